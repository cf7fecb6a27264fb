"""DPPO fine-tuning loop over the B200 library — the caller side of the hot path, with the rollout resident in HBM.

Mirrors `TrainPPODiffusionAgent.run` (agent/finetune/train_ppo_diffusion_agent.py:47-468) and the attributes it inherits
from `TrainPPOAgent` / `TrainAgent` (agent/finetune/train_ppo_agent.py:17-80, train_agent.py:17-125), minus Hydra, WandB,
video rendering and checkpoint files (the reference side keeps those).  What changes relative to the
reference is where the data lives:

* `obs_trajs` / `chains_trajs` are CUDA buffers [n_steps, n_envs, ...]; the sampling kernel writes each step's chains
  straight into `chains_trajs[step]` (reference: NumPy float64 holders filled from TF tensors, :82-97,132);
* the old-value / old-log-prob pass runs once over the whole buffer on the device (reference: `logprob_batch_size`
  splits through NumPy, :191-229) and GAE is the fp64 device scan `dppo_gae` (bit-identical to the NumPy loop :242-263);
* a minibatch is the slice of the shuffled flat index handed to `dppo_ppo_step_indexed` (reference: `tf.gather` /
  `gather_nd` of seven tensors, :287-312), loss + backward + AdamW fused in the same call (:314-356).

Data parallel (absent in the reference): under an initialised `torch.distributed` group every rank owns a
contiguous block of the `n_envs` env copies (its own `venv` with that many envs, its own resident rollout) and a replica of
the weights (`model.engine.init_comm()` is called here).  Every rank draws the SAME minibatch permutation over the global
(step, env, k) pool and keeps the rows of its env block (`parallel.local_minibatch`), so the union over ranks is exactly the
reference's minibatch; the advantage normalisation of a minibatch (diffusion_ppo.py:74-75) uses the global mean / std from one
small all-reduce of (sum, sum of squares); the library's gradient exchange returns global metrics, so the KL early stop
(:366-368) is taken identically everywhere; reward scaling and episode statistics run on the all-gathered (tiny) reward arrays.

Only rewards / terminated / firsts and the reward scaler stay on the host (float64 NumPy, like the reference), because the
environment produces them there.  `venv` needs `reset_arg(options_list) -> {"state": [E,To,Do]}` and
`step(action [E,act_steps,Da]) -> (obs dict, reward [E], terminated [E], truncated [E], info)`, the reference's vector-env API.
"""
import logging
import time as _time

import numpy as np
import torch

from ...parallel import local_minibatch, shard_range, stats_from_moments
from ...util.reward_scaling import RunningRewardScaler

log = logging.getLogger(__name__)


class TrainPPODiffusionAgent:
    def __init__(self, model, venv, *, n_envs, n_steps, act_steps, n_train_itr, batch_size, update_epochs,
                 gamma=0.99, gae_lambda=0.95, target_kl=1, actor_lr=1e-4, val_freq=10, force_train=False,
                 reset_at_iteration=True, reward_scale_running=True, reward_scale_const=1.0, n_critic_warmup_itr=0,
                 reward_horizon=None, max_grad_norm=None, best_reward_threshold_for_success=3.0,
                 furniture_sparse_reward=False, log_freq=1, seed=42, noise_fn=None, shuffle_fn=None, normalization=None):
        """`actor_lr`: float or a schedule called with the optimizer's iteration count (Keras semantics of
        train_ppo_agent.py:35-49).  `noise_fn(itr, step, B) -> (x_T [B,A], noise [T,B,A])` and
        `shuffle_fn(itr, epoch, total) -> int permutation` replace the library's Philox stream / torch.randperm
        (tests inject them to replay the same draws on a CPU restatement of the loop).
        `normalization`: dict with obs_min / obs_max / action_min / action_max (the task's normalization.npz).  When
        given, `venv` is the RAW environment (no MujocoLocomotionLowdimWrapper): it emits raw float64 observations and takes raw
        actions; normalize_obs, the fp32 cast, the `[:, :act_steps]` slice and unnormalize_action (mujoco_locomotion_lowdim.py:57-62,
        train_ppo_diffusion_agent.py:111-123) run on the device around the sampler (`dppo_rollout_step`), with the obs / action exchange
        through two alternating pinned host buffers that the kernels read and write directly."""
        self.model, self.venv, self.engine = model, venv, model.engine
        # data parallel: `n_envs` is the GLOBAL env count; this rank steps env columns [env_lo, env_hi) (its venv has that many)
        import torch.distributed as dist
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.n_envs_global = int(n_envs)
        self.env_lo, self.env_hi = shard_range(int(n_envs), self.rank, self.world)
        if self.world > 1:
            self.engine.init_comm()
        n_envs = self.env_hi - self.env_lo
        # :349-354: any max_grad_norm switches on tf.clip_by_norm(grad, clip_norm=1.0) per variable (the constant is the reference's)
        self.max_grad_norm = max_grad_norm
        self.engine.set_grad_clip_norm(1.0 if max_grad_norm is not None else None)
        self.n_envs, self.n_steps, self.act_steps = int(n_envs), int(n_steps), int(act_steps)
        self.n_train_itr, self.batch_size, self.update_epochs = int(n_train_itr), int(batch_size), int(update_epochs)
        self.gamma, self.gae_lambda, self.target_kl = gamma, gae_lambda, target_kl
        self.actor_lr, self.val_freq, self.force_train = actor_lr, val_freq, force_train
        self.reset_at_iteration = reset_at_iteration
        self.reward_scale_running, self.reward_scale_const = reward_scale_running, reward_scale_const
        self.n_critic_warmup_itr = n_critic_warmup_itr
        self.reward_horizon = self.act_steps if reward_horizon is None else reward_horizon    # :27
        if self.reward_horizon != model.cfg.reward_horizon:
            raise ValueError("reward_horizon must equal the model's cfg.reward_horizon (fixed at construction)")
        self.best_reward_threshold_for_success = best_reward_threshold_for_success
        self.furniture_sparse_reward, self.log_freq = furniture_sparse_reward, log_freq
        self.n_cond_step, self.obs_dim = model.cfg.cond_steps, model.cfg.obs_dim
        self.horizon_steps, self.action_dim = model.horizon_steps, model.action_dim
        self.running_reward_scaler = RunningRewardScaler(self.n_envs_global) if reward_scale_running else None   # train_ppo_agent.py:70-72
        self.noise_fn, self.shuffle_fn = noise_fn, shuffle_fn
        self.normalization = None
        if normalization is not None:
            self.normalization = {k: np.asarray(normalization[k], np.float32).reshape(-1) for k in ("obs_min", "obs_max", "action_min", "action_max")}
            self.engine.set_env_normalization(**self.normalization)
        self._perm_gen = torch.Generator().manual_seed(seed)
        self.itr, self.cnt_train_step, self.opt_iterations = 0, 0, 0
        self.run_results = []
        self._done_venv = np.zeros((1, self.n_envs))
        self._prev_obs_venv = None
        self._alloc()

    # ---------------------------------------------------------------- resident rollout storage
    def _alloc(self):
        dev = self.engine.dev
        S, E, K = self.n_steps, self.n_envs, self.model.ft_denoising_steps
        A, Do = self.horizon_steps * self.action_dim, self.n_cond_step * self.obs_dim
        self._K = K
        self.obs_trajs = torch.zeros(S, E, Do, device=dev)
        self.chains_trajs = torch.zeros(S, E, K + 1, A, device=dev)
        self.actions_dev = torch.zeros(E, A, device=dev)
        self.obs_stage = torch.zeros(E, Do).pin_memory()
        self.actions_stage = torch.zeros(E, A).pin_memory()
        # raw env exchange (normalization given): two alternating pinned buffers per direction
        self.raw_obs_stage = [torch.zeros(E, Do, dtype=torch.float64).pin_memory() for _ in range(2)]
        self.raw_act_stage = [torch.zeros(E, self.act_steps * self.action_dim).pin_memory() for _ in range(2)]
        self.inds_stage = torch.zeros(self.batch_size, dtype=torch.int32).pin_memory()
        self.metrics_stage = torch.zeros(8).pin_memory()

    # ---------------------------------------------------------------- checkpoints (train_agent.py:127-142)
    def save_model(self, checkpoint_dir):
        """`state_{itr}.weights.h5` in the reference's Keras layout (rank 0 only) + `state_{itr}.opt.npz` with what the reference
        does NOT save but a true resume needs: AdamW moments / step, the optimizer iteration count, the reward scaler."""
        import os
        if self.rank != 0:
            return None
        os.makedirs(checkpoint_dir, exist_ok=True)
        path = os.path.join(checkpoint_dir, f"state_{self.itr}.weights.h5")
        self.model.save_weights(path)
        m, v, step = self.engine.get_opt_state(1)
        extra = dict(m=m, v=v, step=step, itr=self.itr, opt_iterations=self.opt_iterations, cnt_train_step=self.cnt_train_step)
        np.savez(os.path.join(checkpoint_dir, f"state_{self.itr}.opt.npz"), **extra)
        log.info("Saved model to %s", path)
        return path

    def load(self, checkpoint_dir, itr, with_optimizer=True):
        import os
        self.model.load_weights(os.path.join(checkpoint_dir, f"state_{itr}.weights.h5"))
        side = os.path.join(checkpoint_dir, f"state_{itr}.opt.npz")
        if with_optimizer and os.path.exists(side):
            z = np.load(side, allow_pickle=False)
            self.engine.set_opt_state(1, z["m"], z["v"], int(z["step"]))
            self.itr, self.opt_iterations, self.cnt_train_step = int(z["itr"]), int(z["opt_iterations"]), int(z["cnt_train_step"])

    def reset_env_all(self, options_venv=None):
        """train_agent.py:144-153."""
        options_venv = options_venv if options_venv is not None else [{} for _ in range(self.n_envs)]
        return self.venv.reset_arg(options_list=options_venv)

    def _lr(self):
        return float(self.actor_lr(self.opt_iterations)) if callable(self.actor_lr) else float(self.actor_lr)

    # ---------------------------------------------------------------- one iteration of :59-468
    def run_iteration(self):
        S, E, K = self.n_steps, self.n_envs, self.model.ft_denoising_steps
        if K != self._K:
            self._alloc()                                           # ft_denoising_steps annealed (diffusion_vpg.py:131-134)
        eval_mode = self.itr % self.val_freq == 0 and not self.force_train       # :70
        t0 = _time.time()

        firsts_trajs = np.zeros((S + 1, E))
        # :75-81 (`last_itr_eval` is assigned from eval_mode just before this test in the reference, so it adds nothing)
        if self.reset_at_iteration or eval_mode or self._prev_obs_venv is None:
            self._prev_obs_venv = self.reset_env_all()
            firsts_trajs[0] = 1
        else:
            firsts_trajs[0] = self._done_venv
        prev_obs_venv = self._prev_obs_venv
        terminated_trajs = np.zeros((S, E))
        reward_trajs = np.zeros((S, E))
        stream = torch.cuda.current_stream(self.engine.dev)

        for step in range(S):                                        # :107-142
            kw = {}
            if self.noise_fn is not None:
                kw["x_T"], kw["noise"] = self.noise_fn(self.itr, step, E)
            if self.normalization is not None:
                # raw observations in, raw actions out: normalisation, slice and un-normalisation run on the device (8f.4)
                ro, ra = self.raw_obs_stage[step & 1], self.raw_act_stage[step & 1]
                ro.copy_(torch.from_numpy(np.ascontiguousarray(prev_obs_venv["state"], np.float64)).reshape(E, -1))
                self.engine.rollout_step(ro, self.obs_trajs[step], self.actions_dev, self.chains_trajs[step], ra, self.act_steps,
                                         deterministic=eval_mode, min_sampling_std=float(self.model.get_min_sampling_denoising_std()),
                                         seed=self.model.seed, offset=self.model._next_offset(), row_offset=self.env_lo, **kw)
                stream.synchronize()
                action_venv = ra.numpy().reshape(E, self.act_steps, self.action_dim)
            else:
                self.obs_stage.copy_(torch.from_numpy(np.ascontiguousarray(prev_obs_venv["state"], np.float32)).reshape(E, -1))
                self.obs_trajs[step].copy_(self.obs_stage, non_blocking=True)
                self.engine.sample(self.obs_trajs[step], deterministic=eval_mode,
                                   min_sampling_std=float(self.model.get_min_sampling_denoising_std()),
                                   seed=self.model.seed, offset=self.model._next_offset(), row_offset=self.env_lo,
                                   actions_out=self.actions_dev, chains_out=self.chains_trajs[step], **kw)
                self.actions_stage.copy_(self.actions_dev, non_blocking=True)
                stream.synchronize()
                output_venv = self.actions_stage.numpy().reshape(E, self.horizon_steps, self.action_dim)
                action_venv = output_venv[:, : self.act_steps]      # :123
            obs_venv, reward_venv, terminated_venv, truncated_venv, _info = self.venv.step(np.array(action_venv))
            done_venv = terminated_venv | truncated_venv
            reward_trajs[step] = reward_venv
            terminated_trajs[step] = terminated_venv
            firsts_trajs[step + 1] = done_venv
            prev_obs_venv = obs_venv
            self.cnt_train_step += self.n_envs_global * self.act_steps if not eval_mode else 0
        self._prev_obs_venv, self._done_venv = prev_obs_venv, done_venv

        result = {"itr": self.itr, "step": self.cnt_train_step, "eval_mode": eval_mode}
        result.update(self._summarize_episodes(firsts_trajs, reward_trajs))
        if not eval_mode:
            result.update(self._update(prev_obs_venv, reward_trajs, terminated_trajs, firsts_trajs))
        self.model.step()                                            # :399
        result["diffusion_min_sampling_std"] = self.model.get_min_sampling_denoising_std()
        result["time"] = _time.time() - t0
        self.run_results.append(result)
        self.itr += 1
        return result

    def run(self):
        while self.itr < self.n_train_itr:
            r = self.run_iteration()
            if r["itr"] % self.log_freq == 0:
                if r["eval_mode"]:
                    log.info("eval: success rate %8.4f | avg episode reward %8.4f | avg best reward %8.4f",
                             r["success_rate"], r["avg_episode_reward"], r["avg_best_reward"])
                else:
                    log.info("%d: step %8d | loss %8.4f | pg loss %8.4f | value loss %8.4f | reward %8.4f | t:%8.4f",
                             r["itr"], r["step"], r["loss"], r["pg_loss"], r["v_loss"], r["avg_episode_reward"], r["time"])
        return self.run_results

    # ---------------------------------------------------------------- data-parallel plumbing (tiny host arrays)
    def _gather_envs(self, x):
        """[..., E_local] host array of every rank -> [..., E_global] (env columns in rank order), identical on every rank."""
        if self.world == 1:
            return x
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, np.ascontiguousarray(x))
        return np.concatenate(parts, axis=-1)

    # ---------------------------------------------------------------- :145-186, host bookkeeping
    def _summarize_episodes(self, firsts_trajs, reward_trajs):
        firsts_trajs, reward_trajs = self._gather_envs(firsts_trajs), self._gather_envs(reward_trajs)
        spans = []
        for env_ind in range(self.n_envs_global):
            starts = np.where(firsts_trajs[:, env_ind] == 1)[0]
            spans += [(env_ind, a, b - 1) for a, b in zip(starts[:-1], starts[1:]) if b - a > 1]
        if not spans:
            log.info("[WARNING] No episode completed within the iteration!")
            return dict(num_episode_finished=0, avg_episode_reward=0, avg_best_reward=0, success_rate=0)
        pieces = [reward_trajs[a: b + 1, e] for e, a, b in spans]
        episode_reward = np.array([np.sum(p) for p in pieces])
        best = episode_reward if self.furniture_sparse_reward else np.array([np.max(p) / self.act_steps for p in pieces])
        return dict(num_episode_finished=len(pieces), avg_episode_reward=np.mean(episode_reward),
                    avg_best_reward=np.mean(best), success_rate=np.mean(best >= self.best_reward_threshold_for_success))

    # ---------------------------------------------------------------- :189-377, everything on the device
    def _update(self, obs_venv, reward_trajs, terminated_trajs, firsts_trajs):
        S, E, K = self.n_steps, self.n_envs, self.model.ft_denoising_steps
        eng = self.engine
        obs_k = self.obs_trajs.reshape(S * E, -1)
        chains_k = self.chains_trajs.reshape(S * E, K + 1, -1)
        values_k = eng.value(obs_k)                                  # :205
        logprobs_k = eng.logprobs(obs_k, chains_k)                   # :223  [S*E*K, A], row = b*K + k
        if self.reward_scale_running:                                # :232-236 (running statistics over ALL env copies: gathered, then this rank's columns)
            scaled = self.running_reward_scaler(reward=self._gather_envs(reward_trajs).T, first=self._gather_envs(firsts_trajs)[:-1].T).T
            reward_trajs = np.ascontiguousarray(scaled[:, self.env_lo:self.env_hi])
        last_obs = obs_venv["state"]
        if self.normalization is not None:                           # the raw env's last observation: normalize_obs on the host (tiny)
            n = self.normalization
            last_obs = 2 * ((np.asarray(last_obs, np.float64) - n["obs_min"]) / (n["obs_max"] - n["obs_min"] + 1e-6) - 0.5)
        next_values = eng.value(np.ascontiguousarray(last_obs, np.float32).reshape(E, -1))   # :252
        advantages_k, returns_k = eng.gae(np.ascontiguousarray(reward_trajs), terminated_trajs, values_k.reshape(S, E),
                                          next_values, self.reward_scale_const, self.gamma, self.gae_lambda)

        total_steps = S * self.n_envs_global * K                     # the GLOBAL (step, env, k) pool
        num_batch = max(1, total_steps // self.batch_size)           # :285, the tail is skipped
        adv_flat = advantages_k.reshape(-1)
        clipfracs, metrics, flag_break = [], None, False
        apply = self.itr >= self.n_critic_warmup_itr                 # :348
        for update_epoch in range(self.update_epochs):
            if self.shuffle_fn is not None:
                inds_k = np.asarray(self.shuffle_fn(self.itr, update_epoch, total_steps), np.int32)
            else:
                inds_k = torch.randperm(total_steps, generator=self._perm_gen).to(torch.int32).numpy()   # :284
            for batch in range(num_batch):
                inds_b = inds_k[batch * self.batch_size: (batch + 1) * self.batch_size]
                kw = {}
                if self.world > 1:
                    # this rank's rows of the global minibatch + the minibatch's global advantage statistics (one small all-reduce)
                    import torch.distributed as dist
                    n_glob = int(inds_b.shape[0])
                    inds_b = local_minibatch(inds_b, self.n_envs_global, K, self.env_lo, self.env_hi)
                    if inds_b.shape[0] < 1:
                        raise RuntimeError("a rank holds no row of this minibatch: use fewer ranks or a larger batch_size")
                    a = adv_flat[torch.from_numpy(inds_b // K).to(adv_flat.device, torch.int64)].to(torch.float64)
                    mom = torch.stack([a.sum(), (a * a).sum()])
                    dist.all_reduce(mom)
                    mean, std = stats_from_moments(float(mom[0]), float(mom[1]), n_glob)
                    kw = dict(n_global=n_glob, adv_mean=mean, adv_std=std)
                stage = self.inds_stage[: inds_b.shape[0]]
                stage.copy_(torch.from_numpy(np.ascontiguousarray(inds_b)))
                eng.ppo_step_indexed(obs_k, chains_k, logprobs_k, returns_k, values_k, advantages_k, stage.numpy(),
                                     lr=self._lr(), apply=apply, metrics_host=self.metrics_stage.numpy(), **kw)
                if apply:
                    self.opt_iterations += 1
                metrics = self.metrics_stage.numpy().copy()
                clipfracs.append(float(metrics[3]))
                log.info("approx_kl: %s, update_epoch: %d, num_batch: %d", metrics[4], update_epoch, num_batch)
                if self.target_kl is not None and metrics[4] > self.target_kl:     # :366-368
                    flag_break = True
                    break
            if flag_break:
                break
        y_pred = self._gather_envs(values_k.reshape(S, -1).cpu().numpy()).reshape(-1)     # :373-377 (over the whole rollout)
        y_true = self._gather_envs(returns_k.reshape(S, -1).cpu().numpy()).reshape(-1)
        var_y = np.var(y_true)
        explained_var = np.nan if var_y == 0 else 1 - np.var(y_true - y_pred) / var_y
        pg_loss, entropy_loss, v_loss, _cf, approx_kl, ratio, bc_loss, eta = [float(m) for m in metrics]
        return dict(loss=pg_loss + v_loss * self.model.vf_coef, pg_loss=pg_loss, v_loss=v_loss, bc_loss=bc_loss, eta=eta,
                    approx_kl=approx_kl, ratio=ratio, clipfrac=float(np.mean(clipfracs)), explained_var=explained_var,
                    entropy_loss=entropy_loss, n_updates=len(clipfracs))
