"""Thin Python handle over the C ABI: torch only supplies device memory, streams and (for N>1)
the rendezvous that carries the NCCL unique id.  All arithmetic runs in libdppo_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib as L


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    if isinstance(t, np.ndarray):
        return C.c_void_p(t.ctypes.data)
    raise TypeError(type(t))


def _as_dev(x, dev, dtype=torch.float32):
    """Contiguous tensor of `dtype` on `dev` (no copy if it already is)."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(device=dev, dtype=dtype).contiguous()


def _as_host(x, dtype=np.float32) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=dtype)


class Engine:
    """One `dppo_handle` bound to one GPU (one per rank / process)."""

    def __init__(self, cfg: L.DppoCfg, device: int = 0):
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise L.DppoError("no CUDA device visible: libdppo_b200 has no CPU fallback")
        self.cfg = cfg
        self.device_index = int(device)
        self.dev = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        L.check(self.lib.dppo_create(C.byref(cfg), self.device_index, C.byref(h)), "dppo_create")
        self.h = h
        self.Do = cfg.obs_dim * cfg.cond_steps
        self.A = cfg.action_dim * cfg.horizon_steps
        self.T = cfg.denoising_steps
        self.n_actor = self.lib.dppo_num_params(C.byref(cfg), L.NET_ACTOR)
        self.n_critic = self.lib.dppo_num_params(C.byref(cfg), L.NET_CRITIC)
        self.rank, self.world = 0, 1

    @property
    def K(self):
        return self.cfg.ft_denoising_steps

    def close(self):
        if getattr(self, "h", None):
            self.lib.dppo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    # ---------------------------------------------------------------- weights / optimizer state
    def num_params(self, net: int) -> int:
        return self.n_critic if net == L.NET_CRITIC else self.n_actor

    def set_weights(self, net: int, flat):
        if isinstance(flat, torch.Tensor) and flat.is_cuda:
            flat = flat.to(torch.float32).contiguous().reshape(-1)
            L.check(self.lib.dppo_set_weights(self.h, net, _ptr(flat), flat.numel(), 1, self._stream()), "dppo_set_weights")
        else:
            a = _as_host(flat).reshape(-1)
            L.check(self.lib.dppo_set_weights(self.h, net, _ptr(a), a.size, 0, self._stream()), "dppo_set_weights")

    def get_weights(self, net: int) -> np.ndarray:
        out = np.empty(self.num_params(net), np.float32)
        L.check(self.lib.dppo_get_weights(self.h, net, _ptr(out), out.size, 0, self._stream()), "dppo_get_weights")
        return out

    def get_opt_state(self, opt: int):
        n = self.n_actor if opt == L.OPT_PRETRAIN else self.n_actor + self.n_critic
        m, v = np.empty(n, np.float32), np.empty(n, np.float32)
        step = C.c_int64(0)
        L.check(self.lib.dppo_get_opt_state(self.h, opt, _ptr(m), _ptr(v), n, C.byref(step), 0, self._stream()), "dppo_get_opt_state")
        return m, v, int(step.value)

    def set_opt_state(self, opt: int, m, v, step: int):
        m, v = _as_host(m).reshape(-1), _as_host(v).reshape(-1)
        L.check(self.lib.dppo_set_opt_state(self.h, opt, _ptr(m), _ptr(v), m.size, int(step), 0, self._stream()), "dppo_set_opt_state")

    def set_ft_denoising_steps(self, K: int):
        L.check(self.lib.dppo_set_ft_denoising_steps(self.h, int(K)), "dppo_set_ft_denoising_steps")
        self.cfg.ft_denoising_steps = int(K)

    def set_grad_clip_norm(self, clip_norm: Optional[float]):
        """Per-variable tf.clip_by_norm before AdamW (train_ppo_diffusion_agent.py:349-354); None / <= 0 switches it off."""
        L.check(self.lib.dppo_set_grad_clip_norm(self.h, float(clip_norm) if clip_norm else 0.0), "dppo_set_grad_clip_norm")

    # ---------------------------------------------------------------- forward-only
    def ffma_peak_tflops(self) -> float:
        """Measured sustained CUDA-core FFMA rate of this GPU (bench.py: the strict-fp32 mode's roofline denominator)."""
        v = C.c_double(0.0)
        L.check(self.lib.dppo_debug_ffma_peak(self.h, C.byref(v)), "dppo_debug_ffma_peak")
        return float(v.value)

    def actor_forward(self, net: int, x, t, obs) -> torch.Tensor:
        x = _as_dev(x, self.dev).reshape(-1, self.A)
        N = x.shape[0]
        t = _as_dev(t, self.dev, torch.int32).reshape(N)
        obs = _as_dev(obs, self.dev).reshape(N, self.Do)
        eps = torch.empty(N, self.A, device=self.dev, dtype=torch.float32)
        L.check(self.lib.dppo_actor_forward(self.h, net, _ptr(x), _ptr(t), _ptr(obs), N, _ptr(eps), self._stream()), "dppo_actor_forward")
        return eps

    def value(self, obs) -> torch.Tensor:
        obs = _as_dev(obs, self.dev).reshape(-1, self.Do)
        N = obs.shape[0]
        v = torch.empty(N, device=self.dev, dtype=torch.float32)
        L.check(self.lib.dppo_value(self.h, _ptr(obs), N, _ptr(v), self._stream()), "dppo_value")
        return v

    def sample(self, obs, deterministic=False, use_base_policy=False, min_sampling_std: float = -1.0,
               seed: int = 0, offset: int = 0, row_offset: int = 0, x_T=None, noise=None,
               return_chain=True, actions_out: Optional[torch.Tensor] = None,
               chains_out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """Device-resident call: obs is (moved to) a CUDA tensor; returns CUDA tensors.  `actions_out` [B,A] /
        `chains_out` [B,K+1,A] (contiguous fp32 CUDA views, e.g. one step of a resident rollout buffer) are written in
        place by the sampling kernel when given."""
        obs = _as_dev(obs, self.dev).reshape(-1, self.Do)
        B = obs.shape[0]
        x_T = None if x_T is None else _as_dev(x_T, self.dev).reshape(B, self.A)
        noise = None if noise is None else _as_dev(noise, self.dev).reshape(self.T, B, self.A)
        for name, buf, n in (("actions_out", actions_out, B * self.A), ("chains_out", chains_out, B * (self.K + 1) * self.A)):
            if buf is not None and not (buf.is_cuda and buf.dtype == torch.float32 and buf.is_contiguous() and buf.numel() == n):
                raise ValueError(f"{name} must be a contiguous fp32 CUDA tensor of {n} elements")
        actions = actions_out if actions_out is not None else torch.empty(B, self.A, device=self.dev, dtype=torch.float32)
        chains = chains_out if chains_out is not None else (
            torch.empty(B, self.K + 1, self.A, device=self.dev, dtype=torch.float32) if return_chain else None)
        L.check(self.lib.dppo_sample(self.h, _ptr(obs), B, int(deterministic), int(use_base_policy), float(min_sampling_std),
                                     int(seed), int(offset), int(row_offset), _ptr(x_T), _ptr(noise),
                                     _ptr(actions), _ptr(chains), self._stream()), "dppo_sample")
        return actions, chains

    def set_env_normalization(self, obs_min, obs_max, action_min, action_max):
        """The task's normalization.npz (mujoco_locomotion_lowdim.py:21-25): enables `rollout_step`."""
        arrs = [_as_host(a).reshape(-1) for a in (obs_min, obs_max, action_min, action_max)]
        if arrs[0].size != self.cfg.obs_dim or arrs[1].size != self.cfg.obs_dim or arrs[2].size != self.cfg.action_dim or arrs[3].size != self.cfg.action_dim:
            raise ValueError("normalisation arrays must have obs_dim / action_dim entries")
        L.check(self.lib.dppo_set_env_normalization(self.h, *[_ptr(a) for a in arrs]), "dppo_set_env_normalization")

    def rollout_step(self, raw_obs, obs_out, actions_out, chains_out, raw_actions_out, act_steps, deterministic=False, use_base_policy=False,
                     min_sampling_std: float = -1.0, seed: int = 0, offset: int = 0, row_offset: int = 0, x_T=None, noise=None):
        """One rollout step around RAW env data: `raw_obs` float64 [E, To*obs_dim] and `raw_actions_out` fp32
        [E, act_steps*action_dim] are pinned host tensors (or CUDA tensors) the kernels read / write directly; `obs_out` [E,Do],
        `actions_out` [E,A], `chains_out` [E,K+1,A] are CUDA tensors.  Asynchronous: synchronise the stream before the envs read
        `raw_actions_out`."""
        E = obs_out.shape[0]
        if raw_obs.dtype != torch.float64 or raw_actions_out.dtype != torch.float32:
            raise ValueError("raw_obs must be float64 and raw_actions_out float32")
        for t in (raw_obs, raw_actions_out):
            if not (t.is_cuda or t.is_pinned()) or not t.is_contiguous():
                raise ValueError("host tensors handed to rollout_step must be pinned and contiguous")
        x_T = None if x_T is None else _as_dev(x_T, self.dev).reshape(E, self.A)
        noise = None if noise is None else _as_dev(noise, self.dev).reshape(self.T, E, self.A)
        L.check(self.lib.dppo_rollout_step(self.h, _ptr(raw_obs), E, int(deterministic), int(use_base_policy), float(min_sampling_std),
                                           int(seed), int(offset), int(row_offset), _ptr(x_T), _ptr(noise), _ptr(obs_out), _ptr(actions_out),
                                           _ptr(chains_out), _ptr(raw_actions_out), int(act_steps), self._stream()), "dppo_rollout_step")

    def sample_host(self, obs: np.ndarray, actions_out: np.ndarray, chains_out: Optional[np.ndarray],
                    deterministic=False, use_base_policy=False, min_sampling_std: float = -1.0,
                    seed: int = 0, offset: int = 0, row_offset: int = 0, x_T=None, noise=None):
        """Host-buffer call (H2D + chain + D2H + sync inside the library).  Buffers should be pinned."""
        B = obs.shape[0]
        L.check(self.lib.dppo_sample_host(self.h, _ptr(obs), B, int(deterministic), int(use_base_policy), float(min_sampling_std),
                                          int(seed), int(offset), int(row_offset), _ptr(x_T), _ptr(noise),
                                          _ptr(actions_out), _ptr(chains_out), self._stream()), "dppo_sample_host")

    def logprobs(self, obs, chains, use_base_policy=False) -> torch.Tensor:
        obs = _as_dev(obs, self.dev).reshape(-1, self.Do)
        B = obs.shape[0]
        chains = _as_dev(chains, self.dev).reshape(B, self.K + 1, self.A)
        logp = torch.empty(B * self.K, self.A, device=self.dev, dtype=torch.float32)
        L.check(self.lib.dppo_logprobs(self.h, _ptr(obs), _ptr(chains), B, int(use_base_policy), _ptr(logp), self._stream()), "dppo_logprobs")
        return logp

    def logprobs_subsample(self, obs, prev, nxt, inds, use_base_policy=False) -> torch.Tensor:
        obs = _as_dev(obs, self.dev).reshape(-1, self.Do)
        N = obs.shape[0]
        prev = _as_dev(prev, self.dev).reshape(N, self.A)
        nxt = _as_dev(nxt, self.dev).reshape(N, self.A)
        inds = _as_dev(inds, self.dev, torch.int32).reshape(N)
        logp = torch.empty(N, self.A, device=self.dev, dtype=torch.float32)
        L.check(self.lib.dppo_logprobs_subsample(self.h, _ptr(obs), _ptr(prev), _ptr(nxt), _ptr(inds), N, int(use_base_policy),
                                                 _ptr(logp), self._stream()), "dppo_logprobs_subsample")
        return logp

    # ---------------------------------------------------------------- updates
    def ppo_step(self, obs, prev, nxt, inds, returns, oldvalues, advantages, oldlogp, lr: float, apply=True,
                 n_global: Optional[int] = None, adv_mean: float = 0.0, adv_std: float = -1.0, want_grads=False):
        obs = _as_dev(obs, self.dev).reshape(-1, self.Do)
        N = obs.shape[0]
        prev = _as_dev(prev, self.dev).reshape(N, self.A)
        nxt = _as_dev(nxt, self.dev).reshape(N, self.A)
        inds = _as_dev(inds, self.dev, torch.int32).reshape(N)
        returns = _as_dev(returns, self.dev).reshape(N)
        oldvalues = _as_dev(oldvalues, self.dev).reshape(N)
        advantages = _as_dev(advantages, self.dev).reshape(N)
        oldlogp = _as_dev(oldlogp, self.dev).reshape(N, self.A)
        metrics = torch.empty(8, device=self.dev, dtype=torch.float32)
        grads = torch.empty(self.n_actor + self.n_critic, device=self.dev, dtype=torch.float32) if want_grads else None
        L.check(self.lib.dppo_ppo_step(self.h, _ptr(obs), _ptr(prev), _ptr(nxt), _ptr(inds), _ptr(returns), _ptr(oldvalues),
                                       _ptr(advantages), _ptr(oldlogp), N, int(n_global if n_global is not None else N),
                                       float(adv_mean), float(adv_std), float(lr), int(apply),
                                       _ptr(metrics), _ptr(grads), self._stream()), "dppo_ppo_step")
        return (metrics, grads) if want_grads else metrics

    def ppo_step_host(self, obs, prev, nxt, inds, returns, oldvalues, advantages, oldlogp, metrics_out: np.ndarray,
                      lr: float, apply=True, n_global: Optional[int] = None, adv_mean: float = 0.0, adv_std: float = -1.0):
        """All arguments are host (ideally pinned) contiguous arrays; metrics_out is host float32[8]."""
        N = obs.shape[0]
        L.check(self.lib.dppo_ppo_step_host(self.h, _ptr(obs), _ptr(prev), _ptr(nxt), _ptr(inds), _ptr(returns), _ptr(oldvalues),
                                            _ptr(advantages), _ptr(oldlogp), N, int(n_global if n_global is not None else N),
                                            float(adv_mean), float(adv_std), float(lr), int(apply),
                                            _ptr(metrics_out), self._stream()), "dppo_ppo_step_host")

    def ppo_step_indexed(self, obs_buf, chains_buf, oldlogp_buf, returns_buf, values_buf, adv_buf, inds_k, lr: float, apply=True,
                         n_global: Optional[int] = None, adv_mean: float = 0.0, adv_std: float = -1.0, want_grads=False,
                         metrics_host: Optional[np.ndarray] = None):
        """Index-driven update (train_ppo_diffusion_agent.py:287-312): the rollout buffers are CUDA tensors that stay
        resident; `inds_k` holds flat (b*K + k) indices - a CUDA int32 tensor, or (with `metrics_host`, a pinned float32[8]
        array) a host int32 array, in which case the call copies the indices in, the metrics out and synchronises."""
        obs_buf = _as_dev(obs_buf, self.dev).reshape(-1, self.Do)
        P = obs_buf.shape[0]
        chains_buf = _as_dev(chains_buf, self.dev).reshape(P, self.K + 1, self.A)
        oldlogp_buf = _as_dev(oldlogp_buf, self.dev).reshape(P, self.K, self.A)
        returns_buf = _as_dev(returns_buf, self.dev).reshape(P)
        values_buf = _as_dev(values_buf, self.dev).reshape(P)
        adv_buf = _as_dev(adv_buf, self.dev).reshape(P)
        if metrics_host is not None:
            inds = _as_host(inds_k, np.int32).reshape(-1) if not isinstance(inds_k, np.ndarray) else inds_k
            N = inds.shape[0]
            L.check(self.lib.dppo_ppo_step_indexed_host(self.h, _ptr(obs_buf), _ptr(chains_buf), _ptr(oldlogp_buf), _ptr(returns_buf),
                                                        _ptr(values_buf), _ptr(adv_buf), P, _ptr(inds), N,
                                                        int(n_global if n_global is not None else N), float(adv_mean), float(adv_std),
                                                        float(lr), int(apply), _ptr(metrics_host), self._stream()), "dppo_ppo_step_indexed_host")
            return metrics_host
        inds = _as_dev(inds_k, self.dev, torch.int32).reshape(-1)
        N = inds.shape[0]
        metrics = torch.empty(8, device=self.dev, dtype=torch.float32)
        grads = torch.empty(self.n_actor + self.n_critic, device=self.dev, dtype=torch.float32) if want_grads else None
        L.check(self.lib.dppo_ppo_step_indexed(self.h, _ptr(obs_buf), _ptr(chains_buf), _ptr(oldlogp_buf), _ptr(returns_buf), _ptr(values_buf),
                                               _ptr(adv_buf), P, _ptr(inds), N, int(n_global if n_global is not None else N),
                                               float(adv_mean), float(adv_std), float(lr), int(apply), _ptr(metrics), _ptr(grads),
                                               self._stream()), "dppo_ppo_step_indexed")
        return (metrics, grads) if want_grads else metrics

    def gae(self, rewards, terminated, values, next_values, reward_scale_const=1.0, gamma=0.999, gae_lambda=0.95):
        """train_ppo_diffusion_agent.py:242-263 on the device: rewards [S,E] (float64), terminated [S,E], values [S,E],
        next_values [E] -> (advantages [S,E], returns [S,E]) fp32 CUDA tensors."""
        rewards = _as_dev(rewards, self.dev, torch.float64)
        S, E = rewards.shape
        terminated = _as_dev(terminated, self.dev).reshape(S, E)
        values = _as_dev(values, self.dev).reshape(S, E)
        next_values = _as_dev(next_values, self.dev).reshape(E)
        adv = torch.empty(S, E, device=self.dev, dtype=torch.float32); ret = torch.empty_like(adv)
        L.check(self.lib.dppo_gae(self.h, _ptr(rewards), _ptr(terminated), _ptr(values), _ptr(next_values), S, E,
                                  float(reward_scale_const), float(gamma), float(gae_lambda), _ptr(adv), _ptr(ret), self._stream()), "dppo_gae")
        return adv, ret

    def pretrain_step(self, actions, obs, lr: float, apply=True, t=None, noise=None, seed: int = 0, offset: int = 0,
                      n_global: Optional[int] = None, row_offset: int = 0, want_grads=False):
        actions = _as_dev(actions, self.dev).reshape(-1, self.A)
        N = actions.shape[0]
        obs = _as_dev(obs, self.dev).reshape(N, self.Do)
        t = None if t is None else _as_dev(t, self.dev, torch.int32).reshape(N)
        noise = None if noise is None else _as_dev(noise, self.dev).reshape(N, self.A)
        loss = torch.empty(1, device=self.dev, dtype=torch.float32)
        grads = torch.empty(self.n_actor, device=self.dev, dtype=torch.float32) if want_grads else None
        L.check(self.lib.dppo_pretrain_step(self.h, _ptr(actions), _ptr(obs), N, int(n_global if n_global is not None else N),
                                            int(row_offset), _ptr(t), _ptr(noise), int(seed), int(offset), float(lr), int(apply),
                                            _ptr(loss), _ptr(grads), self._stream()), "dppo_pretrain_step")
        return (loss, grads) if want_grads else loss

    def ema_update(self, decay: float):
        L.check(self.lib.dppo_ema_update(self.h, float(decay), self._stream()), "dppo_ema_update")

    # ---------------------------------------------------------------- multi-GPU
    def init_comm(self, group=None):
        """Attach an NCCL communicator spanning torch.distributed's (default) group.  torch only
        carries the 128-byte unique id; the gradient all-reduce is issued by the library."""
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        buf = (C.c_char * 128)()
        if rank == 0:
            L.check(self.lib.dppo_comm_unique_id(buf), "dppo_comm_unique_id")
        obj = [bytes(buf.raw) if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0, group=group)
        idb = C.create_string_buffer(obj[0], 128)
        L.check(self.lib.dppo_comm_init(self.h, idb, rank, world), "dppo_comm_init")
        self.rank, self.world = rank, world
        # single node, <= 8 ranks: attach the peers' gradient buffers (CUDA IPC) so that the update uses the fused
        # peer-memory all-reduce + AdamW kernel instead of ncclAllReduce + AdamW (DPPO_NO_PEER_ALLREDUCE=1 keeps NCCL)
        import os
        if world <= 8 and os.environ.get("DPPO_NO_PEER_ALLREDUCE") != "1":
            blob = (C.c_char * 192)()
            L.check(self.lib.dppo_comm_ipc_export(self.h, blob), "dppo_comm_ipc_export")
            blobs = [None] * world
            dist.all_gather_object(blobs, bytes(blob.raw), group=group)
            allb = C.create_string_buffer(b"".join(blobs), 192 * world)
            L.check(self.lib.dppo_comm_ipc_attach(self.h, allb, rank, world), "dppo_comm_ipc_attach")
            self.peer_allreduce = True

    # ---------------------------------------------------------------- introspection
    def launch_count(self) -> int:
        return int(self.lib.dppo_launch_count(self.h))

    def tc_launch_count(self) -> int:
        return int(self.lib.dppo_tc_launch_count(self.h))

    def fused_launch_count(self) -> int:
        return int(self.lib.dppo_fused_launch_count(self.h))

    def last_path(self) -> int:
        return int(self.lib.dppo_last_path(self.h))

    def force_path(self, path: int):
        self.lib.dppo_force_path(self.h, int(path))

    def profile_enable(self, on: bool = True):
        L.check(self.lib.dppo_profile_enable(self.h, int(on)), "dppo_profile_enable")

    def profile_read_class(self, cls: int):
        """-> (ms, launches, algorithmic flops) of one kernel class: 0 fused chain, 1 tcgen05 GEMM, 2 FFMA SGEMM."""
        ms, n, fl = C.c_double(0), C.c_int64(0), C.c_double(0)
        L.check(self.lib.dppo_profile_read_class(self.h, int(cls), C.byref(ms), C.byref(n), C.byref(fl)), "dppo_profile_read_class")
        return float(ms.value), int(n.value), float(fl.value)

    def profile_read_exec(self, cls: int) -> float:
        """Tensor-pipe flops the launches of a class issued (call after profile_read_class)."""
        fl = C.c_double(0)
        L.check(self.lib.dppo_profile_read_exec(self.h, int(cls), C.byref(fl)), "dppo_profile_read_exec")
        return float(fl.value)

    def profile_read(self):
        """-> (gemm_ms, gemm_launches, gemm_flops) accumulated since profile_enable(True)."""
        ms, n, fl = C.c_double(0), C.c_int64(0), C.c_double(0)
        L.check(self.lib.dppo_profile_read(self.h, C.byref(ms), C.byref(n), C.byref(fl)), "dppo_profile_read")
        return float(ms.value), int(n.value), float(fl.value)
