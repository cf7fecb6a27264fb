#!/usr/bin/env python
"""bench.py — DPPO hot path on B200 (BASELINE.json).

Workload (N=1 default = BASELINE.json configs[1]): walker2d-v2 ft_ppo_diffusion_mlp shapes
(obs 17, act 6, horizon 4, T=20, K=10).  One "step" = one PPO minibatch update of 50 000
(env-step, k) rows PER GPU: log-prob forward of the stored denoising chains under the current
weights, clipped-ratio + value loss, backward, (N>1: one all-reduce of grads+metrics), AdamW.
`value` = PPO log-prob-update samples/s over all ranks (weak scaling).  The other half of the
metric — denoised action chunks/s of the T=20 chain at 40 env copies — is measured in the same
run and reported under "sampling" (it is dependency-latency-bound, not a throughput kernel).

Precision: the default is the tensor-core mode (tcgen05, bf16 operands, fp32 accumulation / masters / loss / AdamW) that
BASELINE.json's north_star allows with stated looser bounds (tests/test_gpu_bf16.py); `--precision fp32` runs the strict
parity mode (CUDA-core FFMA, 1e-4 / 1e-3 bounds) and its throughput is also reported in every line under "fp32_parity_mode".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

TASK = "walker2d"
N_ROWS = 50_000            # minibatch rows per GPU (cfg train.batch_size, ft_ppo_diffusion_mlp.yaml:69)
N_ENVS = 40                # env copies per rollout step (ft_ppo_diffusion_mlp_run.yaml:26)
METRIC = "ppo_logprob_update_samples_per_sec"
UNIT = "samples/s"
PATHS = {1: "persistent cluster kernel (fp32 FFMA, DSMEM)", 2: "layered fp32", 3: "layered tcgen05", 4: "fused tcgen05 chain (one launch, T steps on chip)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ----------------------------------------------------------------------------- clocks sampler
class Clocks(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- synthetic data
DIMS = dict(obs_dim=17, action_dim=6, horizon_steps=4, cond_steps=1, denoising_steps=20, ft_denoising_steps=10,
            time_dim=16, actor_hidden=512, critic_hidden=256)


class D:
    """Shapes of the workload."""
    def __init__(self):
        self.__dict__.update(DIMS)
        self.A = self.action_dim * self.horizon_steps
        self.Do = self.obs_dim * self.cond_steps
        self.Din = self.A + self.time_dim + self.Do


def make_gpu_engine(precision, device):
    """Engine with Glorot-initialised base policy, fine-tuned copy = base + 5e-3 N(0,1), random critic.
    Uses only the product package (never the oracle)."""
    import diffusionpolicyoptimization_b200 as dp
    from diffusionpolicyoptimization_b200 import _lib as L
    cfg = dp.default_cfg()
    for k in ("obs_dim", "action_dim", "horizon_steps", "cond_steps", "denoising_steps", "ft_denoising_steps", "time_dim",
              "actor_hidden", "critic_hidden"):
        setattr(cfg, k, DIMS[k])
    cfg.precision = precision
    e = dp.Engine(cfg, device)
    actor = dp.DiffusionMLP(action_dim=6, horizon_steps=4, cond_dim=17, mlp_dims=[512, 512, 512], activation_type="ReLU",
                            residual_style=True, seed=0)
    critic = dp.CriticObs(cond_dim=17, mlp_dims=[256, 256, 256], residual_style=True, seed=1)
    w = actor.get_flat_weights()
    rng = np.random.default_rng(2)
    e.set_weights(L.NET_ACTOR, w)
    e.set_weights(L.NET_ACTOR_EMA, w)
    e.set_weights(L.NET_ACTOR_FT, w + 5e-3 * rng.standard_normal(w.size).astype(np.float32))
    e.set_weights(L.NET_CRITIC, critic.get_flat_weights())
    return e


def make_gpu_batches(e, n_rows, n_batches, seed):
    """a pool of stored chains produced by the sampler itself (in-kernel Philox),
    old log-probs under the pre-update (base) weights, old values from the critic; minibatch rows
    are random (env-step, k) pairs.  Everything here runs before any timed region."""
    dev, K = e.dev, e.K
    g = torch.Generator(device=dev); g.manual_seed(seed)
    P = 4096
    obs = torch.rand(P, e.Do, device=dev, generator=g) * 2 - 1
    _, chains = e.sample(obs, seed=seed, offset=1)
    oldlogp = e.logprobs(obs, chains, use_base_policy=True).reshape(P, K, e.A)
    values = e.value(obs)
    out = []
    for _ in range(n_batches):
        flat = torch.randint(0, P * K, (n_rows,), device=dev, generator=g)
        b, k = flat // K, flat % K
        out.append([obs[b].contiguous(), chains[b, k].contiguous(), chains[b, k + 1].contiguous(), k.to(torch.int32).contiguous(),
                    torch.randn(n_rows, device=dev, generator=g), (values[b] + 0.1 * torch.randn(n_rows, device=dev, generator=g)).contiguous(),
                    torch.randn(n_rows, device=dev, generator=g), oldlogp[b, k].contiguous()])
    torch.cuda.synchronize()
    return out


def flops_per_sample(d):
    Fa = 2 * (d.Din * d.actor_hidden + 2 * d.actor_hidden ** 2 + d.actor_hidden * d.A)
    Fc = 2 * (d.Do * d.critic_hidden + 2 * d.critic_hidden ** 2 + d.critic_hidden)
    return (3 * Fa - 2 * d.Din * d.actor_hidden) + (3 * Fc - 2 * d.Do * d.critic_hidden), Fa


# ----------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args):
    """The reference's CPU implementation of the path = the oracle restatement (TensorFlow is not
    installable here), all host threads, same config/metric, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import dppo_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    o = O.make_oracle(TASK, seed=0)
    n_cpu = N_ROWS
    batch = O.make_ppo_batch(o, n_cpu, pool=512, seed=1)
    params = [p.clone() for p in o.actor_ft] + [p.clone() for p in o.critic]
    m = [torch.zeros_like(p) for p in params]; v = [torch.zeros_like(p) for p in params]

    def step(i):
        oo = O.Oracle(o.d, o.h, o.actor, params[:12], params[12:])
        _, ga, gc = oo.ppo_grads(*batch)
        O.adamw_keras(params, ga + gc, m, v, i + 1, 1e-4, o.h.beta1, o.h.beta2, o.h.adam_eps, o.h.weight_decay)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    dt = time.perf_counter() - t0
    val = n_cpu * args.steps / dt
    # sampling half of the metric on the CPU, for the "sampling" block
    obs, x_T, noise = O.make_rollout_inputs(o, N_ENVS, seed=3)
    o.sample(obs, x_T, noise)
    t1 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        o.sample(obs, x_T, noise)
    dts = (time.perf_counter() - t1) / reps
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": f"{TASK}-v2 ft_ppo_diffusion_mlp PPO update, N={n_cpu} rows x 1 host (obs 17, act 6, Ta 4, T 20, K 10)",
                   "note": "oracle = torch-CPU fp32 restatement of the reference TF graph (TF not installable); runs on rank 0 only"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full minibatches of {n_cpu} rows, torch {torch.get_num_threads()} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sampling": {"chunks_per_sec": N_ENVS / dts, "us_per_rollout_step": dts * 1e6, "n_envs": N_ENVS},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm (GPU)
def run_ours(args):
    import torch.distributed as dist
    from diffusionpolicyoptimization_b200 import _lib as L
    from diffusionpolicyoptimization_b200.parallel import advantage_stats, init_process_group_from_env

    rank, world, local = init_process_group_from_env("nccl")
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    prec = L.PREC_BF16 if args.precision == "bf16" else L.PREC_FP32
    d = D()
    e = make_gpu_engine(prec, local)
    if world > 1:
        e.init_comm()
    fps, Fa = flops_per_sample(d)
    pk, pk_src = peaks()

    # ---- inputs: 2 distinct minibatches per rank, pinned on the host and resident on the device
    devb = make_gpu_batches(e, N_ROWS, 2, seed=10 + rank)
    pinned = [[t.cpu().pin_memory() for t in b] for b in devb]
    n_global = N_ROWS * world
    # global advantage statistics (diffusion_ppo.py:74-75 normalises over the whole minibatch)
    stats = []
    for i in range(2):
        adv = devb[i][6]
        if world > 1:
            parts = [torch.empty_like(adv) for _ in range(world)]
            dist.all_gather(parts, adv)
            stats.append(advantage_stats(torch.cat(parts).cpu().numpy()))
        else:
            stats.append(advantage_stats(adv.cpu().numpy()))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    lr = 1e-4

    def dev_step(i):
        b = devb[i & 1]; mean, std = stats[i & 1]
        return e.ppo_step(*b, lr=lr, apply=True, n_global=n_global, adv_mean=mean, adv_std=std)

    metrics_host = torch.empty(8, dtype=torch.float32).pin_memory()

    def e2e_step(i):
        b = pinned[i & 1]; mean, std = stats[i & 1]
        e.ppo_step_host(*[t.numpy() for t in b], metrics_host.numpy(), lr=lr, apply=True, n_global=n_global, adv_mean=mean, adv_std=std)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm):
        for i in range(warm):
            fn(i)
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            flush.zero_()                               # L2 flush between timed steps (outside the event pair)
            evs[i][0].record(); fn(warm + i); evs[i][1].record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    clk = Clocks(local); clk.start()
    l0 = e.launch_count()
    ms_dev = timed(dev_step, args.steps, args.warmup)
    launches = e.launch_count() - l0 - 0
    launches_per_step = launches // (args.steps + args.warmup)
    ms_e2e = timed(e2e_step, args.steps, max(3, args.warmup // 2))
    # the same host buffers copied to the device with nothing else running: the host-link floor of `e2e` on this box
    h2d_dst = [torch.empty_like(t, device=dev) for t in pinned[0]]

    def h2d_only(i):
        for hb, db in zip(pinned[i & 1], h2d_dst):
            db.copy_(hb, non_blocking=True)
    ms_h2d = timed(h2d_only, args.steps, 2)
    del h2d_dst
    # ---- e2e through the index-driven entry point (the reference's own data flow, :266-312): the rollout
    #      of an iteration (P = 500 steps x 40 envs chains, old log-probs, returns/values/advantages) is uploaded from pinned
    #      host memory once per 20 minibatch updates (update_epochs 5 x 4 minibatches) and stays resident; every step copies
    #      only its 50 000 shuffled flat indices in and the 8 metrics out.  The upload is inside the timed region.
    P_roll, K = 20000, d.ft_denoising_steps
    g2 = torch.Generator(device=dev); g2.manual_seed(99 + rank)
    obs_r = torch.rand(P_roll, d.Do, device=dev, generator=g2) * 2 - 1
    _, chains_r = e.sample(obs_r, seed=5, offset=7)
    olp_r = e.logprobs(obs_r, chains_r, use_base_policy=True).reshape(P_roll, K, d.A)
    val_r = e.value(obs_r)
    roll_host = [t.cpu().pin_memory() for t in (obs_r, chains_r, olp_r, torch.randn(P_roll, device=dev, generator=g2), val_r,
                                                torch.randn(P_roll, device=dev, generator=g2))]
    roll_dev = [torch.empty_like(t, device=dev) for t in roll_host]
    inds_host = [torch.randint(0, P_roll * K, (N_ROWS,), dtype=torch.int32).pin_memory() for _ in range(2)]
    UPLOAD_EVERY = 20
    # global advantage statistics of the (uniformly indexed) rollout advantages, identical on every rank
    adv_all = roll_host[5].to(dev)
    if world > 1:
        parts = [torch.empty_like(adv_all) for _ in range(world)]
        dist.all_gather(parts, adv_all)
        adv_all = torch.cat(parts)
    idx_mean, idx_std = advantage_stats(adv_all.cpu().numpy())

    def e2e_indexed_step(i):
        if i % UPLOAD_EVERY == 0:
            for hb, db in zip(roll_host, roll_dev):
                db.copy_(hb, non_blocking=True)
        e.ppo_step_indexed(*roll_dev, inds_host[i & 1].numpy(), lr=lr, apply=True, n_global=n_global, adv_mean=idx_mean, adv_std=idx_std,
                           metrics_host=metrics_host.numpy())

    idx_steps = max(args.steps, UPLOAD_EVERY)
    ms_e2e_idx = timed(e2e_indexed_step, idx_steps, UPLOAD_EVERY)     # warm-up and timed region each start with an upload
    roll_bytes = sum(t.numel() * t.element_size() for t in roll_host)
    # the timed regions last only milliseconds: keep the same step running for ~0.3 s so that nvidia-smi (20 ms period)
    # sees the clocks and throttle reasons of this workload under load
    t_load = time.perf_counter(); n_load = 0
    while time.perf_counter() - t_load < 0.3:
        for i in range(20):
            dev_step(i)
        torch.cuda.synchronize(); n_load += 20
    clocks = clk.finish()
    clocks["load_steps_sampled"] = n_load

    # ---- roofline of the dominant kernel, timed live with CUDA events inside the library (every tensor-class launch is
    #      bracketed on its launching stream; classes: 0 fused tcgen05 layer chain, 1 tcgen05 GEMM (dW), 2 FFMA SGEMM)
    #      While the profile is on the library runs the actor and critic chains back to back (in the timed region above the critic
    #      chain runs on a second stream and fills the idle SMs of the actor chain's last wave), so each duration is that kernel alone.
    e.profile_enable(True)
    for i in range(args.steps):
        flush.zero_(); dev_step(i)
    cls = [e.profile_read_class(c) for c in range(4)]
    e.profile_enable(False)
    tensor = prec == L.PREC_BF16
    step_ms = ms_dev / args.steps
    names = ["fc::chain_kernel<512> (fused tcgen05 actor forward / backward chains, CTA pairs)", "tcp::dw_pair_kernel (tcgen05 grouped split-K weight gradients on CTA pairs)",
             "sgemm_kernel (fp32 FFMA layers + gradients)", "fc::chain_kernel<256> (fused tcgen05 Mish critic forward / backward chains)"]
    dom = max(range(4), key=lambda c: cls[c][0])
    d_ms, d_n, d_fl = cls[dom]
    achieved = d_fl / (d_ms * 1e-3) / 1e12 if d_ms > 0 else 0.0
    all_ms = sum(c[0] for c in cls); all_fl = sum(c[2] for c in cls)
    roof = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
            "frac": achieved / pk["bf16_tflops_sustained"],
            # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this bench
            # (profiles/r01_final_chain_dw_ncu.txt): mean of the step's two chain_kernel<512> launches (actor forward 115 MB,
            # actor backward 112 MB); algorithmic HBM bytes of the same two: 171 / 160 MB (part of the stores is still in L2
            # when the kernel ends)
            "traffic": 111.9e6 if tensor and dom == 0 else None,
            "traffic_source": "profiles/r01_chain_r1s_ncu.txt (static, from the committed ncu --set full capture: mean dram read+write of the step's two chain_kernel<512> launches)" if tensor and dom == 0 else None,
            "kernel": names[dom],
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk_src}); the kernel runs inside a long step",
            "avg_launch_us": d_ms / max(d_n, 1) * 1e3, "launches_timed": d_n,
            "algorithmic_flops_per_launch": d_fl / max(d_n, 1),
            "share_of_step": d_ms / args.steps / step_ms,
            "all_tensor_kernels": {"achieved": all_fl / (all_ms * 1e-3) / 1e12 if all_ms > 0 else 0.0,
                                   "share_of_step": all_ms / args.steps / step_ms,
                                   "classes": {names[c].split(" ")[0]: {"ms_per_step": cls[c][0] / args.steps, "launches_per_step": cls[c][1] // args.steps,
                                                                        "tflops": cls[c][2] / (cls[c][0] * 1e-3) / 1e12 if cls[c][0] > 0 else 0.0}
                                               for c in range(4) if cls[c][1] > 0}},
            "note": "per-kernel durations are timed with the actor and critic chains serialized; in the timed region of `value` the critic chain overlaps the actor chain's last wave on a second stream, so shares are relative to the serialized sum" if tensor else "fp32 parity mode runs on CUDA cores: FFMA peak is ~74.5 TFLOP/s (148 SM x 128 x 2 x 1.965 GHz); frac is still quoted against the bf16 tensor peak"}

    # ---- sampling half of the metric: walker2d, 40 env copies, T=20 chain, in-kernel Philox
    obs40 = torch.rand(N_ENVS, d.Do, device=dev) * 2 - 1
    obs40_h = obs40.cpu().pin_memory()
    act_h = torch.empty(N_ENVS, d.A).pin_memory(); ch_h = torch.empty(N_ENVS, d.ft_denoising_steps + 1, d.A).pin_memory()
    cnt = [0]

    def samp_dev(i):
        cnt[0] += 1
        e.sample(obs40, seed=1, offset=cnt[0])

    def samp_e2e(i):
        cnt[0] += 1
        e.sample_host(obs40_h.numpy(), act_h.numpy(), ch_h.numpy(), seed=1, offset=cnt[0])

    ls0 = e.launch_count()
    ms_s = timed(samp_dev, 50, 5)
    samp_launches = (e.launch_count() - ls0) // 55
    samp_path = e.last_path()
    ms_s_e2e = timed(samp_e2e, 50, 5)
    BL = 148 * 128                                     # one 128-row tile per SM
    obsL = torch.rand(BL, d.Do, device=dev) * 2 - 1

    def samp_large(i):
        cnt[0] += 1
        e.sample(obsL, seed=1, offset=cnt[0], return_chain=True)

    ms_L = timed(samp_large, 5, 3)
    largeB_path = e.last_path()
    sampling = {
        "chunks_per_sec": world * N_ENVS * 50 / (ms_s * 1e-3), "us_per_rollout_step": ms_s / 50 * 1e3,
        "launches_per_rollout_step": samp_launches, "path": PATHS[samp_path],
        "n_envs_per_gpu": N_ENVS, "e2e_chunks_per_sec": world * N_ENVS * 50 / (ms_s_e2e * 1e-3),
        "e2e_us_per_rollout_step": ms_s_e2e / 50 * 1e3,
        "fp32_fma_frac": (N_ENVS * d.denoising_steps * Fa / (ms_s / 50 * 1e-3)) / 74.5e12,
        "large_batch": {"rows_per_gpu": BL, "chunks_per_sec": world * BL * 5 / (ms_L * 1e-3),
                        "tflops": BL * d.denoising_steps * Fa * 5 / (ms_L * 1e-3) / 1e12,
                        "frac_of_bf16_sustained": BL * d.denoising_steps * Fa * 5 / (ms_L * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
                        "launches_per_rollout_step": 1 if largeB_path in (1, 4) else None,
                        "path": PATHS[largeB_path]},
    }

    # ---- BASELINE.json configs[4]: halfcheetah-medium-v2 pre_diffusion_mlp eps-MSE pre-train step, batch 4096 per GPU
    #      (q_sample with in-kernel Philox t / noise, forward, MSE, backward, (all-reduce), AdamW, EMA every step excluded)
    NB = 4096
    x0_p = torch.rand(NB, d.A, device=dev) * 2 - 1
    obs_p = torch.rand(NB, d.Do, device=dev) * 2 - 1
    pcnt = [0]

    def pre_step(i):
        pcnt[0] += 1
        e.pretrain_step(x0_p, obs_p, lr=1e-3, apply=True, seed=3, offset=pcnt[0], n_global=NB * world, row_offset=rank * NB)

    ms_pre = timed(pre_step, 20, 5)
    pre_flops = 3 * Fa - 2 * d.Din * d.actor_hidden
    pretrain = {"samples_per_sec": world * NB * 20 / (ms_pre * 1e-3), "ms_per_step": ms_pre / 20, "batch_per_gpu": NB,
                "tflops": pre_flops * NB * 20 / (ms_pre * 1e-3) / 1e12,
                "note": "latency-bound at this batch (32 row tiles for 148 SMs): 13.7 GFLOP per step"}

    # ---- the strict-parity fp32 mode on the same workload (3 steps), reported beside the headline
    fp32_mode = None
    if tensor:
        e32 = make_gpu_engine(L.PREC_FP32, local)
        if world > 1:
            e32.init_comm()
        def dev_step32(i):
            b = devb[i & 1]; mean, std = stats[i & 1]
            return e32.ppo_step(*b, lr=lr, apply=True, n_global=n_global, adv_mean=mean, adv_std=std)
        ms32 = timed(dev_step32, 3, 3)
        fp32_mode = {"value": n_global * 3 / (ms32 * 1e-3), "unit": UNIT, "ms_per_step": ms32 / 3, "dtype": "fp32",
                     "tflops": fps * N_ROWS / (ms32 / 3 * 1e-3) / 1e12, "frac_of_fp32_ffma_peak": fps * N_ROWS / (ms32 / 3 * 1e-3) / 74.5e12}
        e32.close()

    # ---- CPU baseline (oracle port) on rank 0, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import dppo_oracle as O      # cpu_baseline leg: the only oracle use in this arm
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n_cpu = N_ROWS
        o = O.make_oracle(TASK, seed=0)
        batch = O.make_ppo_batch(o, n_cpu, pool=512, seed=1)
        o.ppo_grads(*batch)
        t0 = time.perf_counter(); reps = 0
        while reps < 3 or (time.perf_counter() - t0 < 10 and reps < 20):
            o.ppo_grads(*batch); reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": n_cpu / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{reps} x loss+gradient of one {n_cpu}-row minibatch (AdamW excluded), torch-CPU fp32, {cores} threads"}

    if rank == 0:
        h2d = sum(t.numel() * t.element_size() for t in pinned[0])
        line = {
            "metric": METRIC, "value": n_global * args.steps / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if tensor else "fp32", "data": "synthetic",
            "config": {"workload": f"{TASK}-v2 ft_ppo_diffusion_mlp PPO update: {N_ROWS} (env-step,k) rows per GPU "
                                   f"(obs 17, act 6, Ta 4, T 20, K 10, actor 512x3 ReLU, critic 256x3 Mish), fused loss+backward+AdamW",
                       "global_rows": n_global, "parallelism": f"dp{world}", "precision": args.precision if not tensor else "bf16 operands (tcgen05), fp32 accumulate/master weights/loss/AdamW",
                       "l2": "256 MB flush buffer written between timed steps; two alternating minibatches",
                       "flops_per_sample": fps},
            "clocks": clocks,
            "e2e": {"value": n_global * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32,
                    "ms_per_step": ms_e2e / args.steps, "api": "Engine.ppo_step_host -> dppo_ppo_step_host (pinned host buffers)",
                    "h2d_alone_ms_per_step": ms_h2d / args.steps,
                    "h2d_alone_gbps": h2d / (ms_h2d / args.steps * 1e-3) / 1e9},
            "e2e_indexed": {"value": n_global * idx_steps / (ms_e2e_idx * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_idx / idx_steps,
                            "steps": idx_steps, "h2d_bytes_per_step": N_ROWS * 4 + roll_bytes * ((idx_steps + UPLOAD_EVERY - 1) // UPLOAD_EVERY) / idx_steps,
                            "d2h_bytes_per_step": 32,
                            "api": "Engine.ppo_step_indexed -> dppo_ppo_step_indexed_host: rollout buffers resident in HBM (re-uploaded from pinned host "
                                   "memory every 20 steps, inside the timed region), per step only the flat minibatch indices go in"},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "roofline": roof,
            "cpu_baseline": cpu,
            "sampling": sampling,
            "pretrain": pretrain,
            "fp32_parity_mode": fp32_mode,
            "step_tflops": fps * N_ROWS / (ms_dev / args.steps * 1e-3) / 1e12,
        }
        print(json.dumps(line), flush=True)
    e.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DPPO_BENCH_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        if args.steps > 5:
            args.steps = 5        # each step is a full 50 000-row minibatch on the CPU (~1 s+)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
