"""CriticObs with the reference's constructor and call (model/common/critic.py:15-54)."""
from typing import Union

import torch

from .mlp import _Net, activation_ids, residual_mlp_shapes


class CriticObs(_Net):
    """State-only critic: ResidualMLP([cond_dim] + mlp_dims + [1])."""

    def __init__(self, cond_dim, mlp_dims, activation_type="Mish", use_layernorm=False, residual_style=False,
                 seed=None, **kwargs):
        if not residual_style or use_layernorm:
            raise ValueError("libdppo_b200 implements the residual_style=True, use_layernorm=False critic of the reference cfgs")
        self.cond_dim, self.mlp_dims = int(cond_dim), list(mlp_dims)
        self.activation_type = activation_type
        self.activation_id = activation_ids[activation_type]
        super().__init__(residual_mlp_shapes([cond_dim] + list(mlp_dims) + [1]), seed=seed)

    def __call__(self, cond: Union[dict, torch.Tensor]):
        """cond: dict with "state" (B, To, Do) or a (B, Do) tensor -> (B, 1)."""
        if self._engine is None:
            raise RuntimeError("critic is not attached to a diffusion model")
        state = cond["state"] if isinstance(cond, dict) else cond
        return self._engine.value(state).reshape(-1, 1)

    call = __call__

    def keras_variable_paths(self, prefix=""):
        from ...util.keras_h5 import keras_paths_critic_obs
        return keras_paths_critic_obs(prefix)
