"""DiffusionModel with the reference's surface (model/diffusion/diffusion.py:18-202), DDPM branch.

The constructor owns the `Engine` (one dppo_handle on one GPU).  DDIM (`use_ddim=True`) is not
implemented — no reference cfg uses it and its eta module was never ported.
"""
from collections import namedtuple

import numpy as np
import torch

from ... import _lib as L
from ...engine import Engine
from .sampling import ddpm_schedule, extract

Sample = namedtuple("Sample", "trajectories chains")   # diffusion.py:15


def _state(cond):
    return cond["state"] if isinstance(cond, dict) else cond


class DiffusionModel:
    def __init__(self, network, horizon_steps, obs_dim, action_dim, network_path=None, device="cuda:0",
                 denoised_clip_value=1.0, randn_clip_value=10.0, final_action_clip_value=None, eps_clip_value=None,
                 denoising_steps=100, predict_epsilon=True, use_ddim=False, ddim_discretize="uniform", ddim_steps=None,
                 cond_steps=1, precision="fp32", seed=0, _cfg_hook=None, **kwargs):
        if use_ddim or not predict_epsilon:
            raise NotImplementedError("libdppo_b200 implements the DDPM / predict_epsilon path (the only one any reference cfg uses)")
        self.horizon_steps, self.obs_dim, self.action_dim = int(horizon_steps), int(obs_dim), int(action_dim)
        self.denoising_steps = int(denoising_steps)
        self.predict_epsilon, self.use_ddim, self.ddim_steps = predict_epsilon, use_ddim, ddim_steps
        self.denoised_clip_value, self.final_action_clip_value = denoised_clip_value, final_action_clip_value
        self.randn_clip_value, self.eps_clip_value = randn_clip_value, eps_clip_value
        self.network, self.network_path = network, network_path
        self.seed, self._calls = int(seed), 0

        cfg = L.default_cfg()
        cfg.obs_dim, cfg.action_dim, cfg.horizon_steps, cfg.cond_steps = self.obs_dim, self.action_dim, self.horizon_steps, int(cond_steps)
        cfg.denoising_steps = self.denoising_steps
        cfg.ft_denoising_steps = 0
        cfg.time_dim = network.time_dim
        dims = network.mlp_dims
        cfg.actor_hidden = dims[0]
        cfg.actor_act = network.activation_id
        cfg.denoised_clip_value = -1.0 if denoised_clip_value is None else float(denoised_clip_value)
        cfg.randn_clip_value = float(randn_clip_value)
        cfg.final_action_clip_value = -1.0 if final_action_clip_value is None else float(final_action_clip_value)
        cfg.precision = {"fp32": L.PREC_FP32, "bf16": L.PREC_BF16, "bf16x3": L.PREC_BF16X3}[precision]
        if _cfg_hook is not None:
            _cfg_hook(cfg)
        assert network.cond_dim == cfg.obs_dim * cfg.cond_steps
        dev = torch.device(device)
        self.device = dev
        self.engine = Engine(cfg, dev.index or 0)
        self.cfg = cfg

        # DDPM parameters (diffusion.py:58-73), as device tensors with the reference's attribute names
        tabs = ddpm_schedule(self.denoising_steps)
        for k, v in tabs.items():
            setattr(self, k, torch.from_numpy(v).to(dev))
        self.alphas = 1.0 - self.betas

        network._bind(self.engine, L.NET_ACTOR)
        if network_path is not None:
            network.load_weights(network_path)

    # ---- diffusion.py:113-151
    def p_mean_var(self, x, t, cond, index=None, network_override=None):
        net = network_override if network_override is not None else self.network
        x = torch.as_tensor(x, device=self.device, dtype=torch.float32)
        t = torch.as_tensor(t, device=self.device)
        noise = net(x, t, cond=cond)
        x_recon = extract(self.sqrt_recip_alphas_cumprod, t, x.shape) * x \
            - extract(self.sqrt_recipm1_alphas_cumprod, t, x.shape) * noise
        if self.denoised_clip_value is not None:
            x_recon = torch.clamp(x_recon, -self.denoised_clip_value, self.denoised_clip_value)
        mu = extract(self.ddpm_mu_coef1, t, x.shape) * x_recon + extract(self.ddpm_mu_coef2, t, x.shape) * x
        logvar = extract(self.ddpm_logvar_clipped, t, x.shape)
        return mu, logvar

    def _next_offset(self):
        self._calls += 1
        return self._calls

    # ---- diffusion.py:153-177: plain sampler = base network on every step, std clipped at 1e-3, 0 at t=0
    def __call__(self, cond, deterministic=True, x_T=None, noise=None):
        actions, _ = self.engine.sample(_state(cond), deterministic=True, use_base_policy=True, seed=self.seed,
                                        offset=self._next_offset(), x_T=x_T, noise=noise, return_chain=False)
        return Sample(actions.reshape(-1, self.horizon_steps, self.action_dim), None)

    call = __call__

    # ---- diffusion.py:179-202
    def c_loss(self, lr=None, t=None, noise=None, **kwargs):
        """loss = mean((eps_hat - eps)^2).  With `lr` the fused backward + AdamW step is taken
        (the reference's tape.gradient/apply_gradients, train_diffusion_agent.py:63-69)."""
        actions, cond = kwargs.get("actions"), kwargs.get("conditions")
        return self.engine.pretrain_step(actions, _state(cond), lr=0.0 if lr is None else lr, apply=lr is not None,
                                         t=t, noise=noise, seed=self.seed, offset=self._next_offset())[0]

    def p_losses(self, x_start, cond, t, noise=None):
        return self.engine.pretrain_step(x_start, _state(cond), lr=0.0, apply=False, t=t, noise=noise,
                                         seed=self.seed, offset=self._next_offset())[0]

    def q_sample(self, x_start, t, noise=None):
        x_start = torch.as_tensor(x_start, device=self.device, dtype=torch.float32)
        t = torch.as_tensor(t, device=self.device)
        if noise is None:
            noise = torch.randn_like(x_start)
        return extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start \
            + extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise

    def set_weights(self, weights):
        self.network.set_weights(weights)

    def get_weights(self):
        return self.network.get_weights()
