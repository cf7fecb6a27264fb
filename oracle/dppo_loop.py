"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement of ONE iteration of the reference's DPPO fine-tuning loop
(agent/finetune/train_ppo_diffusion_agent.py:59-400) on top of oracle/dppo_oracle.py: NumPy float64 holders, per-step
sampling, value / log-prob pass, running reward scaling (util/reward_scaling.py), the GAE scan, shuffled minibatches
gathered with fancy indexing, PPO loss + gradients + Keras-3 AdamW, KL early stop.  The Gaussian draws and the
permutations are injected so the CUDA agent can replay them.  Parity pin: tests/golden/ref_loop.npz - two iterations produced by
the reference's own rollout / update blocks exec'd verbatim over the TF shim (tests/golden/make_ref_loop.py); this file is checked
against it by tests/test_ref_golden.py.
"""
import numpy as np
import torch

from . import dppo_oracle as O


class RewardScaler:
    """util/reward_scaling.py:43-100 (global statistics, the agent's default)."""

    def __init__(self, num_envs, cliprew=10.0, gamma=0.99, epsilon=1e-8):
        self.mean, self.var, self.count = 0.0, 1.0, 1e-4
        self.ret = np.zeros(num_envs)
        self.cliprew, self.gamma, self.epsilon = cliprew, gamma, epsilon

    def __call__(self, reward, first):
        rets = np.zeros_like(reward)
        prev = self.ret
        for t in range(reward.shape[1]):                              # :89-100
            prev = rets[:, t] = reward[:, t] + (1 - first[:, t]) * self.gamma * prev
        self.ret = rets[:, -1]
        x = rets.reshape(-1)
        bm, bv, bc = np.mean(x), np.var(x), x.shape[0]                # :24-40
        delta, tot = bm - self.mean, self.count + bc
        self.mean = self.mean + delta * bc / tot
        self.var = (self.var * self.count + bv * bc + delta ** 2 * self.count * bc / tot) / (tot - 1)
        self.count = tot
        return np.clip(reward / np.sqrt(self.var + self.epsilon), -self.cliprew, self.cliprew)   # :68-73


def gather_minibatch(obs_k, chains_k, logprobs_k, returns_k, values_k, advantages_k, inds_b, K):
    """train_ppo_diffusion_agent.py:292-312: flat index -> (env-step row, denoising index) by row-major unravel over
    (n_steps*n_envs, K); prev / next are chains[b, k] / chains[b, k+1].  Returns the argument tuple of PPODiffusion.c_loss.
    (Pinned by tests/golden/ref_agent_blocks.npz, produced by the reference's own lines.)"""
    inds_b = torch.as_tensor(inds_b).long()
    bi, ki = inds_b // K, inds_b % K
    return (obs_k[bi], chains_k[bi, ki], chains_k[bi, ki + 1], ki.to(torch.int32), returns_k[bi], values_k[bi],
            advantages_k[bi], logprobs_k[bi, ki])


def ppo_iteration(o: O.Oracle, opt: dict, venv, itr: int, prev_obs_venv, *, n_steps, act_steps, batch_size,
                  update_epochs, gamma, gae_lambda, target_kl, lr, reward_scaler, reward_scale_const, noise_fn,
                  shuffle_fn, firsts0, reward_horizon=None):
    """Training iteration (eval_mode False).  `opt` = {"m": [...], "v": [...], "step": int} (AdamW state over
    actor_ft ++ critic); `o.actor_ft` / `o.critic` are updated in place.  Returns (metrics dict, last obs, last done)."""
    d = o.d
    E = prev_obs_venv["state"].shape[0]
    K = d.ft_denoising_steps
    reward_horizon = act_steps if reward_horizon is None else reward_horizon
    obs_trajs = np.zeros((n_steps, E, d.cond_steps, d.obs_dim))                    # :85-97
    chains_trajs = np.zeros((n_steps, E, K + 1, d.horizon_steps, d.action_dim))
    terminated_trajs, reward_trajs = np.zeros((n_steps, E)), np.zeros((n_steps, E))
    firsts_trajs = np.zeros((n_steps + 1, E))
    firsts_trajs[0] = firsts0
    for step in range(n_steps):                                                    # :107-142
        x_T, noise = noise_fn(itr, step, E)
        cond = torch.from_numpy(np.asarray(prev_obs_venv["state"], np.float32))
        s = o.sample(cond, torch.as_tensor(x_T).reshape(E, d.horizon_steps, d.action_dim),
                     torch.as_tensor(noise).reshape(d.denoising_steps, E, d.horizon_steps, d.action_dim))
        action_venv = s.trajectories.numpy()[:, :act_steps]
        obs_venv, reward_venv, terminated_venv, truncated_venv, _ = venv.step(np.array(action_venv))
        done_venv = terminated_venv | truncated_venv
        obs_trajs[step] = prev_obs_venv["state"]
        chains_trajs[step] = s.chains.numpy()
        reward_trajs[step], terminated_trajs[step], firsts_trajs[step + 1] = reward_venv, terminated_venv, done_venv
        prev_obs_venv = obs_venv

    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))     # noqa: E731
    obs_k = f32(obs_trajs).reshape(n_steps * E, d.cond_steps, d.obs_dim)            # :191-229
    chains_k = f32(chains_trajs).reshape(n_steps * E, K + 1, d.horizon_steps, d.action_dim)
    with torch.no_grad():
        values_trajs = O.critic_obs(o.critic, obs_k, o.h.critic_act).numpy().reshape(-1, E).astype(np.float64)
        logprobs_k = o.get_logprobs(obs_k, chains_k).reshape(n_steps * E, K, d.horizon_steps, d.action_dim)
        next_values = O.critic_obs(o.critic, f32(obs_venv["state"]), o.h.critic_act).numpy().reshape(1, -1)
    if reward_scaler is not None:                                                   # :232-236
        reward_trajs = reward_scaler(reward=reward_trajs.T, first=firsts_trajs[:-1].T).T
    advantages_trajs, returns_trajs = O.gae(reward_trajs, terminated_trajs, values_trajs.astype(np.float32),
                                            next_values.reshape(-1), reward_scale_const, gamma, gae_lambda)   # :242-263
    returns_k, values_k, advantages_k = f32(returns_trajs).reshape(-1), f32(values_trajs).reshape(-1), f32(advantages_trajs).reshape(-1)

    total_steps = n_steps * E * K
    num_batch = max(1, total_steps // batch_size)
    last, clipfracs, stop = None, [], False
    for epoch in range(update_epochs):                                              # :281-370
        inds_k = torch.as_tensor(np.asarray(shuffle_fn(itr, epoch, total_steps))).long()
        for b in range(num_batch):
            batch = gather_minibatch(obs_k, chains_k, logprobs_k, returns_k, values_k, advantages_k,
                                     inds_k[b * batch_size: (b + 1) * batch_size], K)
            metrics, ga, gc = o.ppo_grads(*batch, reward_horizon=reward_horizon)
            opt["step"] += 1
            params = o.actor_ft + o.critic
            O.adamw_keras(params, ga + gc, opt["m"], opt["v"], opt["step"], lr, o.h.beta1, o.h.beta2, o.h.adam_eps,
                          o.h.weight_decay)
            last = [float(m) for m in metrics]
            clipfracs.append(last[3])
            if target_kl is not None and last[4] > target_kl:
                stop = True
                break
        if stop:
            break
    y_pred, y_true = values_k.numpy(), returns_k.numpy()
    var_y = np.var(y_true)
    out = dict(pg_loss=last[0], v_loss=last[2], approx_kl=last[4], ratio=last[5], clipfrac=float(np.mean(clipfracs)),
               explained_var=np.nan if var_y == 0 else 1 - np.var(y_true - y_pred) / var_y, n_updates=len(clipfracs),
               returns_k=y_true, values_k=y_pred, advantages_k=advantages_k.numpy(), reward_trajs=reward_trajs,
               chains_k=chains_k.numpy(), logprobs_k=logprobs_k.numpy())
    return out, prev_obs_venv, done_venv


def pretrain_epochs(o: O.Oracle, actions, states, *, n_epochs, batch_size, lr_fn, weight_decay, ema_decay, epoch_start_ema,
                    update_ema_freq, draws_fn):
    """agent/pretrain/train_diffusion_agent.py:56-120 + train_agent.py:46-58,139-148: sequential batches (short tail kept), eps-MSE
    loss with injected (t, eps) draws, Keras-3 AdamW under `lr_fn(iteration)`, EMA copy before `epoch_start_ema` and decay after.
    Returns (per-epoch mean losses, network params, EMA params).  Pinned by tests/golden/ref_pretrain_loop.npz."""
    net = [p.clone() for p in o.actor]
    ema = [p.clone() for p in net]                                     # reset_parameters() in PreTrainAgent.__init__
    m = [torch.zeros_like(p) for p in net]; v = [torch.zeros_like(p) for p in net]
    it, losses = 0, []
    M = actions.shape[0]
    for epoch in range(1, n_epochs + 1):
        ep = []
        for nb, r0 in enumerate(range(0, M, batch_size)):
            a = torch.as_tensor(actions[r0:r0 + batch_size]); s = torch.as_tensor(states[r0:r0 + batch_size])
            t, eps = draws_fn(epoch, nb, a.shape[0])
            oo = O.Oracle(o.d, o.h, net, o.actor_ft, o.critic)
            loss, g = oo.pretrain_grads(a, s, torch.as_tensor(t).long(), torch.as_tensor(eps).reshape(a.shape))
            O.adamw_keras(net, g, m, v, it + 1, lr_fn(it), o.h.beta1, o.h.beta2, o.h.adam_eps, weight_decay)
            it += 1
            ep.append(float(loss))
        losses.append(float(np.mean(ep)))
        if epoch % update_ema_freq == 0:
            if epoch < epoch_start_ema:
                ema = [p.clone() for p in net]
            else:
                O.ema_update(ema, net, ema_decay)
    return losses, net, ema
