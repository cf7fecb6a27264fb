"""Shared test plumbing: build an Engine whose weights are the oracle's."""
import numpy as np
import torch

import diffusionpolicyoptimization_b200 as dp
from diffusionpolicyoptimization_b200 import _lib as L
from oracle import dppo_oracle as O


def make_cfg(o: O.Oracle, precision=L.PREC_FP32):
    d, h = o.d, o.h
    cfg = dp.default_cfg()
    cfg.obs_dim, cfg.action_dim, cfg.horizon_steps, cfg.cond_steps = d.obs_dim, d.action_dim, d.horizon_steps, d.cond_steps
    cfg.denoising_steps, cfg.ft_denoising_steps, cfg.time_dim = d.denoising_steps, d.ft_denoising_steps, d.time_dim
    cfg.actor_hidden, cfg.critic_hidden = d.actor_hidden, d.critic_hidden
    cfg.actor_act = {"ReLU": L.ACT_RELU, "Mish": L.ACT_MISH}[h.actor_act]
    cfg.critic_act = {"ReLU": L.ACT_RELU, "Mish": L.ACT_MISH}[h.critic_act]
    cfg.precision = precision
    cfg.denoised_clip_value = -1.0 if h.denoised_clip_value is None else h.denoised_clip_value
    cfg.randn_clip_value = h.randn_clip_value
    cfg.final_action_clip_value = -1.0 if h.final_action_clip_value is None else h.final_action_clip_value
    cfg.min_sampling_denoising_std, cfg.min_logprob_denoising_std = h.min_sampling_denoising_std, h.min_logprob_denoising_std
    cfg.gamma_denoising = h.gamma_denoising
    cfg.clip_ploss_coef, cfg.clip_ploss_coef_base, cfg.clip_ploss_coef_rate = h.clip_ploss_coef, h.clip_ploss_coef_base, h.clip_ploss_coef_rate
    cfg.clip_vloss_coef = -1.0 if h.clip_vloss_coef is None else h.clip_vloss_coef
    cfg.norm_adv = int(h.norm_adv)
    cfg.reward_horizon = d.horizon_steps
    cfg.vf_coef = h.vf_coef
    cfg.adam_beta1, cfg.adam_beta2, cfg.adam_eps, cfg.weight_decay = h.beta1, h.beta2, h.adam_eps, h.weight_decay
    return cfg


def make_engine(o: O.Oracle, precision=L.PREC_FP32, device=0) -> dp.Engine:
    e = dp.Engine(make_cfg(o, precision), device)
    e.set_weights(L.NET_ACTOR, O.flatten_params(o.actor))
    e.set_weights(L.NET_ACTOR_FT, O.flatten_params(o.actor_ft))
    e.set_weights(L.NET_CRITIC, O.flatten_params(o.critic))
    e.set_weights(L.NET_ACTOR_EMA, O.flatten_params(o.actor))
    return e


def rel_err(a, b):
    """max |a-b| / max |b|  (norm-wise relative error)."""
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def max_abs(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    return float(np.abs(a - b).max())
