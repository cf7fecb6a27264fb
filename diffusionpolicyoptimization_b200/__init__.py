"""diffusionpolicyoptimization_b200 — B200-native (sm_100a) DPPO hot path.

Python keeps the reference's `model.diffusion` / `model.common` object surface
(`PPODiffusion`, `VPGDiffusion`, `DiffusionModel`, `DiffusionMLP`, `CriticObs`) as thin shims over
libdppo_b200.so (C ABI in include/dppo_b200.h).  No CPU fallback exists.
"""
from . import _lib
from ._lib import DppoCfg, DppoError, default_cfg
from .engine import Engine
from .model.diffusion.diffusion import DiffusionModel, Sample
from .model.diffusion.diffusion_vpg import VPGDiffusion
from .model.diffusion.diffusion_ppo import PPODiffusion
from .model.diffusion.mlp_diffusion import DiffusionMLP
from .model.common.critic import CriticObs

__all__ = ["DppoCfg", "DppoError", "default_cfg", "Engine", "DiffusionModel", "Sample", "VPGDiffusion",
           "PPODiffusion", "DiffusionMLP", "CriticObs"]
