"""Generates tests/golden/ref_*.npz by executing the REFERENCE'S OWN model code.

    python tests/golden/make_ref_golden.py            (needs /root/reference; runs in the build container only)

TensorFlow / Keras / tensorflow-probability are not installable here, so the reference's modules
(/root/reference/model/diffusion/{diffusion,diffusion_vpg,diffusion_ppo,mlp_diffusion,modules,sampling}.py and
model/common/{mlp,critic}.py — imported unmodified) run on `tests/golden/tf_shim/`, a torch-CPU restatement of the
~60 TF primitives they call.  What these fixtures pin is therefore the reference's control flow and arithmetic
composition (schedule constants, extract/gather indexing, clip order, ft/base network switch, chain bookkeeping,
log-prob, PPO loss, metric definitions, which variables receive gradients and in which order) — not TF's kernels.
The random draws of `tf.random.normal` / `tf.random.uniform` are served from a recorded source so the same numbers
can be fed to the oracle and to the CUDA path.

Weights are not stored (0.7 M floats per net): they come from oracle.make_oracle(task, seed) — a pure numpy-rng
recipe — and the fixture keeps the seed plus a strided fingerprint.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("DPPO_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import tensorflow as tf  # noqa: E402  (the shim)
from model.common.critic import CriticObs  # noqa: E402  (reference)
from model.diffusion.diffusion import DiffusionModel  # noqa: E402
from model.diffusion.diffusion_ppo import PPODiffusion  # noqa: E402
from model.diffusion.mlp_diffusion import DiffusionMLP  # noqa: E402

from oracle import dppo_oracle as O  # noqa: E402  (weights recipe and Dims only)

SCHEDULE = ("betas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
            "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "ddpm_logvar_clipped",
            "ddpm_mu_coef1", "ddpm_mu_coef2")


class Draws:
    """Serves tf.random.* from a numpy rng and records every draw in call order."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.log = []

    def __call__(self, kind, shape, minval=0, maxval=None, dtype=None):
        if kind == "normal":
            v = torch.from_numpy(self.rng.standard_normal(size=tuple(shape)).astype(np.float32))
        else:
            assert dtype in (torch.int32, torch.int64)
            v = torch.from_numpy(self.rng.integers(minval, maxval, size=tuple(shape))).to(dtype)
        self.log.append(v.clone())
        return v


def np_(x):
    return x.detach().numpy().copy() if isinstance(x, torch.Tensor) else np.asarray(x)


def grads_summary(grads):
    flat = np.concatenate([np_(g).reshape(-1) for g in grads]).astype(np.float32)
    return dict(fp=flat[::97].copy(), norms=np.array([float(np.linalg.norm(np_(g).astype(np.float64))) for g in grads]),
                tail=np.concatenate([np_(g).reshape(-1) for g in grads[-2:]]))


def build(o, model_kw):
    """Instantiate the reference classes the way cfg/gym/finetune/hopper-v2/ft_ppo_diffusion_mlp.yaml:78-110 does."""
    d = o.d
    actor = DiffusionMLP(action_dim=d.action_dim, horizon_steps=d.horizon_steps, cond_dim=d.obs_dim * d.cond_steps,
                         time_dim=d.time_dim, mlp_dims=[d.actor_hidden] * 3, activation_type="ReLU",
                         residual_style=True)
    critic = CriticObs(cond_dim=d.obs_dim * d.cond_steps, mlp_dims=[d.critic_hidden] * 3, activation_type="Mish",
                       residual_style=True)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "base_policy.npz")       # plays the role of the pre-trained checkpoint
        np.savez(path, *[np_(p) for p in o.actor])
        model = PPODiffusion(actor=actor, critic=critic, network_path=path, ft_denoising_steps=d.ft_denoising_steps,
                             horizon_steps=d.horizon_steps, obs_dim=d.obs_dim, action_dim=d.action_dim,
                             denoising_steps=d.denoising_steps, device="cpu", **model_kw)
    model.actor_ft.set_weights([np_(p) for p in o.actor_ft])     # "fine-tuned for a while"
    model.critic.set_weights([np_(p) for p in o.critic])
    return model


def case(name, seed, B, N, dims_over, model_kw, reward_horizon):
    assert dims_over.get("obs_dim", 11) == 11 and dims_over.get("action_dim", 3) == 3, \
        "VPGDiffusion.__init__ builds its networks on hard-coded hopper dummies (diffusion_vpg.py:83-89)"
    torch.manual_seed(0)
    o = O.make_oracle("hopper", seed=seed, **dims_over)
    d = o.d
    K, T = d.ft_denoising_steps, d.denoising_steps
    model = build(o, model_kw)
    out = dict(seed=np.array([seed]), B=np.array([B]), N=np.array([N]), reward_horizon=np.array([reward_horizon]),
               denoising_steps=np.array([T]), ft_denoising_steps=np.array([K]))
    for k, v in model_kw.items():
        out["kw_" + k] = np.array([np.nan if v is None else float(v)])
    for k in ("actor", "actor_ft", "critic"):
        out[k + "_fp"] = O.flatten_params(getattr(o, k))[::97].copy()
    # the variable order the reference's tape sees (train_ppo_diffusion_agent.py:343): actor_ft then critic
    tv = model.trainable_variables
    assert [tuple(v.shape) for v in tv] == d.actor_shapes() + d.critic_shapes()

    # ---- DDPM constants (diffusion.py:57-72)
    for k in SCHEDULE:
        out["sch_" + k] = np_(getattr(model, k))

    # ---- rollout sampling (diffusion_vpg.py:249-339), three flavours on the same draws
    rng = np.random.default_rng(seed + 1)
    obs = torch.from_numpy(rng.uniform(-1, 1, size=(B, d.cond_steps, d.obs_dim)).astype(np.float32))
    draws = Draws(seed + 2)
    tf.random.source = draws
    with torch.no_grad():
        s = model(cond={"state": obs}, deterministic=False, return_chain=True)
        x_T, noise = draws.log[0], torch.stack(draws.log[1:], 0)
        assert noise.shape[0] == T
        replay = lambda: iter([x_T] + list(noise))      # noqa: E731
        it = replay(); tf.random.source = lambda *a, **k: next(it)
        s_det = model(cond={"state": obs}, deterministic=True, return_chain=True)
        it = replay(); tf.random.source = lambda *a, **k: next(it)
        s_base = model(cond={"state": obs}, deterministic=False, return_chain=True, use_base_policy=True)
        logp = model.get_logprobs({"state": obs}, s.chains)           # mutates its cond dict: pass a fresh one
        value = model.critic({"state": obs})
    out.update(obs=np_(obs), x_T=np_(x_T), noise=np_(noise), actions=np_(s.trajectories), chains=np_(s.chains),
               actions_det=np_(s_det.trajectories), chains_det=np_(s_det.chains), actions_base=np_(s_base.trajectories),
               logp=np_(logp), value=np_(value))

    # ---- PPO minibatch built from the reference's own rollout (train_ppo_diffusion_agent.py:266-312)
    pool = 64
    obs_p = torch.from_numpy(rng.uniform(-1, 1, size=(pool, d.cond_steps, d.obs_dim)).astype(np.float32))
    tf.random.source = Draws(seed + 3)
    with torch.no_grad():
        chains_p = model(cond={"state": obs_p}, deterministic=False, return_chain=True).chains
        # old log-probs from a slightly different policy so that ratios are not all 1: the base net
        oldlogp_p = model.get_logprobs({"state": obs_p}, chains_p, use_base_policy=True).reshape(
            pool, K, d.horizon_steps, d.action_dim)
        values_p = model.critic({"state": obs_p}).reshape(-1)
    flat = torch.from_numpy(rng.integers(0, pool * K, size=N))
    b_inds, k_inds = flat // K, flat % K
    f32 = lambda a: torch.from_numpy(a.astype(np.float32))   # noqa: E731
    batch = dict(
        obs={"state": obs_p[b_inds]}, chains_prev=chains_p[b_inds, k_inds], chains_next=chains_p[b_inds, k_inds + 1],
        denoising_inds=k_inds.to(torch.int32), returns=f32(rng.standard_normal(N)),
        oldvalues=values_p[b_inds] + 0.1 * f32(rng.standard_normal(N)), advantages=f32(rng.standard_normal(N)),
        oldlogprobs=oldlogp_p[b_inds, k_inds])
    metrics = model.c_loss(**batch, use_bc_loss=False, reward_horizon=reward_horizon)
    loss = metrics[0] + metrics[2] * 0.5                      # train_ppo_diffusion_agent.py:340, vf_coef yaml:71
    grads = torch.autograd.grad(loss, tv)
    with torch.no_grad():
        logp_sub = model.get_logprobs_subsample(batch["obs"], batch["chains_prev"], batch["chains_next"],
                                                batch["denoising_inds"])
        value_b = model.critic(batch["obs"])
    g = grads_summary(grads)
    out.update(ppo_obs=np_(batch["obs"]["state"]), ppo_prev=np_(batch["chains_prev"]), ppo_next=np_(batch["chains_next"]),
               ppo_inds=np_(batch["denoising_inds"]), ppo_returns=np_(batch["returns"]),
               ppo_oldvalues=np_(batch["oldvalues"]), ppo_adv=np_(batch["advantages"]),
               ppo_oldlogp=np_(batch["oldlogprobs"]),
               ppo_metrics=np.array([float(m) for m in metrics], np.float32),
               ppo_grads_fp=g["fp"], ppo_grads_norms=g["norms"], ppo_grads_tail=g["tail"],
               ppo_logp_sub=np_(logp_sub), ppo_value=np_(value_b))

    # ---- pre-training loss (diffusion.py:179-202) on the base network
    pre = DiffusionModel(network=model.actor, horizon_steps=d.horizon_steps, obs_dim=d.obs_dim,
                         action_dim=d.action_dim, denoising_steps=T, device="cpu")
    x0 = f32(rng.uniform(-1, 1, size=(N, d.horizon_steps, d.action_dim)))
    draws = Draws(seed + 4)
    tf.random.source = draws
    ploss = pre.c_loss(actions=x0, conditions={"state": batch["obs"]["state"]})
    t_drawn, pn = draws.log
    pgrads = torch.autograd.grad(ploss, model.actor.variables)
    g = grads_summary(pgrads)
    out.update(pre_x0=np_(x0), pre_t=np_(t_drawn).astype(np.int32), pre_noise=np_(pn),
               pre_loss=np.array([float(ploss)], np.float32), pre_grads_fp=g["fp"], pre_grads_norms=g["norms"],
               pre_grads_tail=g["tail"])
    np.savez_compressed(os.path.join(HERE, f"ref_{name}.npz"), **out)
    print(name, "metrics", out["ppo_metrics"], "pre_loss", out["pre_loss"], {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    yaml_kw = dict(gamma_denoising=0.99, clip_ploss_coef=0.01, clip_ploss_coef_base=0.01, clip_ploss_coef_rate=3,
                   randn_clip_value=3, min_sampling_denoising_std=0.1, min_logprob_denoising_std=0.1)
    # the shipped configuration (cfg/gym/finetune/hopper-v2/ft_ppo_diffusion_mlp.yaml)
    case("hopper_yaml", seed=11, B=8, N=96, dims_over={}, model_kw=yaml_kw, reward_horizon=4)
    # every denoising step fine-tuned (chain starts with x_T, diffusion_vpg.py:286), value clipping, final action
    # clip, a wider policy clip that actually triggers, shorter reward horizon
    case("hopper_allft", seed=12, B=6, N=80, dims_over=dict(denoising_steps=8, ft_denoising_steps=8),
         model_kw=dict(yaml_kw, clip_ploss_coef=0.2, clip_ploss_coef_base=0.05, clip_vloss_coef=0.2,
                       final_action_clip_value=1.0, min_sampling_denoising_std=0.08),
         reward_horizon=3)
