"""Schedule helpers with the reference's names (model/diffusion/sampling.py:7-29).
The schedule itself is computed by the library's host routine `dppo_ddpm_schedule`."""
import numpy as np
import torch

from ... import _lib as L

SCHEDULE_ROWS = ("betas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                 "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "ddpm_logvar_clipped",
                 "ddpm_mu_coef1", "ddpm_mu_coef2")


def ddpm_schedule(timesteps: int) -> dict:
    """All nine fp32 DDPM tables of model/diffusion/diffusion.py:58-73 as numpy arrays."""
    out = np.empty((len(SCHEDULE_ROWS), timesteps), np.float32)
    L.check(L.load().dppo_ddpm_schedule(int(timesteps), out.ctypes.data), "dppo_ddpm_schedule")
    return {k: out[i].copy() for i, k in enumerate(SCHEDULE_ROWS)}


def cosine_beta_schedule(timesteps, s=0.008, dtype=torch.float32):
    """sampling.py:7-17 (s is fixed at 0.008 in the library, as in every reference call site)."""
    assert abs(s - 0.008) < 1e-12, "the library implements the reference's s=0.008 schedule"
    return torch.from_numpy(ddpm_schedule(int(timesteps))["betas"]).to(dtype)


def extract(a, t, x_shape):
    """sampling.py:20-24: a[t] reshaped to broadcast over x."""
    b = t.shape[0]
    out = a.to(t.device)[t.long()]
    return out.reshape([b] + [1] * (len(x_shape) - 1))


def make_timesteps(batch_size, i, device=None, **kwargs):
    """sampling.py:27-29."""
    return torch.full((batch_size,), int(i), dtype=torch.int32, device=device)
