"""Torch-backed stand-in for the one tensorflow_probability name the reference's DPPO path uses
(model/diffusion/diffusion_vpg.py:419,475).  TEST INFRASTRUCTURE ONLY — see ../tensorflow/__init__.py."""
import math as _math
import types as _types

import torch as _torch


class Normal:
    """tfp.distributions.Normal: `_log_prob` restated from TFP's published source
    (log_unnormalized = -0.5 * squared_difference(x / scale, loc / scale);
     log_normalization = 0.5 * log(2 pi) + log(scale))."""

    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale

    def log_prob(self, x):
        log_unnormalized = -0.5 * ((x / self.scale) - (self.loc / self.scale)) ** 2
        log_normalization = 0.5 * _math.log(2.0 * _math.pi) + _torch.log(self.scale)
        return log_unnormalized - log_normalization


distributions = _types.ModuleType("tensorflow_probability.distributions")
distributions.Normal = Normal
distributions.normal = _types.ModuleType("tensorflow_probability.distributions.normal")
distributions.normal.Normal = Normal
