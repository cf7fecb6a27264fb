"""DiffusionMLP with the reference's constructor and call (model/diffusion/mlp_diffusion.py:12-90)."""
from ...model.common.mlp import _Net, activation_ids, residual_mlp_shapes


class DiffusionMLP(_Net):
    """eps-network: sinusoidal time embedding -> Dense(2td, mish) -> Dense(td); concat [x, t_emb, obs];
    ResidualMLP head."""

    def __init__(self, action_dim, horizon_steps, cond_dim, time_dim=16, mlp_dims=(256, 256), cond_mlp_dims=None,
                 activation_type="Mish", out_activation_type="Identity", use_layernorm=False, residual_style=False,
                 seed=None):
        if cond_mlp_dims is not None or use_layernorm or not residual_style or out_activation_type != "Identity":
            raise ValueError("libdppo_b200 implements the residual_style=True, no-layernorm, no-cond-MLP DiffusionMLP of the reference cfgs")
        self.action_dim, self.horizon_steps, self.cond_dim = int(action_dim), int(horizon_steps), int(cond_dim)
        self.time_dim, self.mlp_dims = int(time_dim), list(mlp_dims)
        self.activation_type = activation_type
        self.activation_id = activation_ids[activation_type]
        td, A = self.time_dim, self.action_dim * self.horizon_steps
        input_dim = td + A + self.cond_dim
        shapes = [(td, 2 * td), (2 * td,), (2 * td, td), (td,)] + residual_mlp_shapes([input_dim] + self.mlp_dims + [A])
        super().__init__(shapes, seed=seed)

    def __call__(self, x, time, cond, **kwargs):
        """x (B,Ta,Da), time (B,), cond {"state": (B,To,Do)} -> eps (B,Ta,Da)."""
        if self._engine is None:
            raise RuntimeError("network is not attached to a diffusion model")
        eps = self._engine.actor_forward(self._net, x, time, cond["state"])
        return eps.reshape(-1, self.horizon_steps, self.action_dim)

    call = __call__

    def keras_variable_paths(self, prefix=""):
        """Dataset paths of this network inside a Keras-3 `.weights.h5` file, in flat variable order (util/keras_h5.py)."""
        from ...util.keras_h5 import keras_paths_diffusion_mlp
        return keras_paths_diffusion_mlp(prefix)
