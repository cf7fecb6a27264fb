"""Generates tests/golden/ref_loop.npz: TWO training iterations of the reference's DPPO fine-tuning loop, produced by the
reference's own code:

  * model classes: model/diffusion/*.py, model/common/{mlp,critic}.py imported unmodified (as in make_ref_golden.py),
  * reward scaling: util/reward_scaling.py imported unmodified,
  * the rollout block (agent/finetune/train_ppo_diffusion_agent.py:106-141) and the whole update block (:187-377: value /
    log-prob pass, reward scaling, GAE, shuffled minibatches, c_loss under the tape, gradients, optimizer step, KL stop,
    explained variance) sliced out of the file at generation time and exec'd verbatim,

over tests/golden/tf_shim/ (torch-CPU TF primitives, GradientTape on torch autograd).  NOT the reference's: the environment
(tests/toy_env.py), the few lines of reset / `firsts_trajs[0]` handling around the blocks (restated below from :73-81), and the
optimizer object — `keras.optimizers.AdamW` is a third-party class, stubbed with the Keras-3 update rule the oracle restates
(oracle.adamw_keras), so the optimizer arithmetic stays unpinned while its call sequence is the reference's.
Gaussian draws and permutations are recorded so that the oracle loop and the CUDA agent can replay them.

    python tests/golden/make_ref_loop.py
"""
import logging
import math
import os
import sys
import tempfile
import textwrap
import types

import einops
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("DPPO_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import tensorflow as tf  # noqa: E402  (the shim)
from model.common.critic import CriticObs  # noqa: E402  (reference)
from model.diffusion.diffusion_ppo import PPODiffusion  # noqa: E402
from model.diffusion.mlp_diffusion import DiffusionMLP  # noqa: E402
from util.reward_scaling import RunningRewardScaler  # noqa: E402  (reference)

from oracle import dppo_oracle as O  # noqa: E402  (weights recipe, AdamW rule)
from toy_env import ToyVecEnv  # noqa: E402

torch.set_grad_enabled(False)          # TF records nothing outside a GradientTape

src = open(os.path.join(REF, "agent", "finetune", "train_ppo_diffusion_agent.py")).read().splitlines()


def block(first, last, last_extra=0):
    a = next(i for i, l in enumerate(src) if l.strip().startswith(first))
    b = next(i for i, l in enumerate(src) if i > a and l.strip().startswith(last)) + last_extra
    return textwrap.dedent("\n".join(src[a:b + 1])), (a + 1, b + 1)


ROLLOUT, rl = block("for step in range(self.n_steps):", "cnt_train_step += self.n_envs * self.act_steps")
UPDATE, ul = block('obs_trajs["state"] = tf.identity(', "np.nan if var_y == 0 else", last_extra=1)
print(f"rollout block: reference lines {rl[0]}-{rl[1]}; update block: {ul[0]}-{ul[1]} ({len(UPDATE.splitlines())} lines)")


class Draws:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed); self.normal, self.perm = [], []

    def __call__(self, kind, shape, **kw):
        if kind == "normal":
            v = torch.from_numpy(self.rng.standard_normal(size=tuple(shape)).astype(np.float32)); self.normal.append(v.clone())
        elif kind == "shuffle":
            v = torch.from_numpy(self.rng.permutation(shape[0])); self.perm.append(v.clone())
        else:
            raise NotImplementedError(kind)
        return v


class AdamWStub:
    """Call surface of keras.optimizers.AdamW as the agent uses it (`apply_gradients(zip(grads, vars))`, lr schedule called with
    the iteration count); the arithmetic is oracle.adamw_keras (Keras-3 semantics as restated in the oracle header)."""

    def __init__(self, lr, h):
        self.lr, self.h, self.iterations, self.m, self.v = lr, h, 0, None, None

    def apply_gradients(self, grads_and_vars):
        grads, vs = zip(*list(grads_and_vars))
        if self.m is None:
            self.m = [torch.zeros_like(p) for p in vs]; self.v = [torch.zeros_like(p) for p in vs]
        lr = float(self.lr(self.iterations)) if callable(self.lr) else float(self.lr)
        self.iterations += 1
        O.adamw_keras([p.data for p in vs], [g.detach() for g in grads], self.m, self.v, self.iterations, lr, self.h.beta1, self.h.beta2,
                      self.h.adam_eps, self.h.weight_decay)


def np_(x):
    return x.detach().numpy().copy() if isinstance(x, torch.Tensor) else np.asarray(x)


def _rearrange(x, pattern, **kw):
    """The two einops patterns the block uses merge the leading (step, env) axes.  (einops itself cannot be called here: with a
    module named `tensorflow` importable it instantiates its TensorFlow backend, which needs the real Keras backend.)"""
    lhs, rhs = (t.strip() for t in pattern.split("->"))
    assert lhs.startswith("s e") and rhs.startswith("(s e)") and lhs[3:].strip() == rhs[5:].strip(), pattern
    return x.reshape((-1,) + tuple(x.shape[2:]))


EINOPS = types.SimpleNamespace(rearrange=_rearrange)


def main():
    seed, E, S, ACT, BATCH, EPOCHS, LR = 21, 8, 6, 4, 160, 2, 1e-4
    o = O.make_oracle("hopper", seed=seed)
    d = o.d
    actor = DiffusionMLP(action_dim=d.action_dim, horizon_steps=d.horizon_steps, cond_dim=d.obs_dim, time_dim=16, mlp_dims=[512] * 3,
                         activation_type="ReLU", residual_style=True)
    critic = CriticObs(cond_dim=d.obs_dim, mlp_dims=[256] * 3, activation_type="Mish", residual_style=True)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "base.npz"); np.savez(path, *[np_(p) for p in o.actor])
        model = PPODiffusion(actor=actor, critic=critic, network_path=path, ft_denoising_steps=d.ft_denoising_steps, horizon_steps=4,
                             obs_dim=11, action_dim=3, denoising_steps=20, device="cpu", gamma_denoising=0.99, clip_ploss_coef=0.01,
                             clip_ploss_coef_base=0.01, clip_ploss_coef_rate=3, randn_clip_value=3, min_sampling_denoising_std=0.1,
                             min_logprob_denoising_std=0.1)
    model.actor_ft.set_weights([np_(p) for p in o.actor_ft]); model.critic.set_weights([np_(p) for p in o.critic])
    draws = Draws(seed + 1); tf.random.source = draws
    venv = ToyVecEnv(E, 11, 3, seed=3)
    self = types.SimpleNamespace(
        n_steps=S, n_envs=E, act_steps=ACT, n_cond_step=1, obs_dim=11, horizon_steps=4, action_dim=3, save_full_observations=False,
        model=model, venv=venv, logprob_batch_size=16, reward_scale_running=True, running_reward_scaler=RunningRewardScaler(E),
        reward_scale_const=1.0, gamma=0.99, gae_lambda=0.95, update_epochs=EPOCHS, batch_size=BATCH, use_bc_loss=False,
        reward_horizon=ACT, vf_coef=0.5, itr=0, n_critic_warmup_itr=0, max_grad_norm=None, target_kl=1,
        actor_optimizer=AdamWStub(LR, o.h), learn_eta=False)
    out = dict(seed=np.array([seed]), cfg=np.array([E, S, ACT, BATCH, EPOCHS]), lr=np.array([LR]))
    prev_obs_venv, done_venv = None, np.zeros((1, E))
    for itr in range(2):
        self.itr = itr
        n0, p0 = len(draws.normal), len(draws.perm)
        # ---- restated from :73-97 (reset at the first iteration only: reset_at_iteration False, no eval iteration)
        firsts_trajs = np.zeros((S + 1, E))
        if prev_obs_venv is None:
            prev_obs_venv = venv.reset_arg(options_list=[{} for _ in range(E)]); firsts_trajs[0] = 1
        else:
            firsts_trajs[0] = done_venv
        ns = dict(np=np, tf=tf, math=math, einops=EINOPS, log=logging.getLogger("ref"), self=self, eval_mode=False, cnt_train_step=0,
                  firsts_trajs=firsts_trajs, prev_obs_venv=prev_obs_venv,
                  obs_trajs={"state": np.zeros((S, E, 1, 11))}, chains_trajs=np.zeros((S, E, d.ft_denoising_steps + 1, 4, 3)),
                  terminated_trajs=np.zeros((S, E)), reward_trajs=np.zeros((S, E)), print=lambda *a, **k: None)
        exec(ROLLOUT, ns)                     # the reference's rollout loop
        exec(UPDATE, ns)                      # the reference's update
        prev_obs_venv, done_venv = ns["prev_obs_venv"], ns["done_venv"]
        k = f"it{itr}_"
        nrm = draws.normal[n0:]
        assert len(nrm) == S * 21
        out[k + "x_T"] = np.stack([np_(nrm[s * 21]) for s in range(S)])                                   # [S,E,4,3]
        out[k + "noise"] = np.stack([np.stack([np_(v) for v in nrm[s * 21 + 1:(s + 1) * 21]]) for s in range(S)])   # [S,20,E,4,3]
        out[k + "perms"] = np.stack([np_(v) for v in draws.perm[p0:]])
        out[k + "chains"] = np.asarray(ns["chains_trajs"], np.float32)
        out[k + "rewards_scaled"] = np.asarray(ns["reward_trajs"])
        for name in ("returns_k", "values_k", "advantages_k", "logprobs_k"):
            out[k + name] = np_(ns[name])
        out[k + "metrics"] = np.array([float(ns[n]) for n in ("pg_loss", "entropy_loss", "v_loss", "clipfrac", "approx_kl", "ratio", "bc_loss", "eta")])
        out[k + "explained_var"] = np.array([float(ns["explained_var"])])
        out[k + "clipfrac_mean"] = np.array([float(np.mean([float(c) for c in ns["clipfracs"]]))])
        out[k + "n_updates"] = np.array([len(ns["clipfracs"])])
        flat = np.concatenate([np_(v).reshape(-1) for v in model.trainable_variables]).astype(np.float32)
        out[k + "weights_fp"] = flat[::97].copy()
        print(k, "metrics", out[k + "metrics"], "explained_var", out[k + "explained_var"], "updates", out[k + "n_updates"], "opt its", self.actor_optimizer.iterations)
    out["base_unchanged"] = np.array([float(np.abs(np.concatenate([np_(v).reshape(-1) for v in model.actor.variables]) - O.flatten_params(o.actor)).max())])
    np.savez_compressed(os.path.join(HERE, "ref_loop.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
