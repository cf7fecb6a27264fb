"""VPGDiffusion with the reference's surface (model/diffusion/diffusion_vpg.py:27-481), DDPM branch."""
import logging

import numpy as np
import torch

from ... import _lib as L
from .diffusion import DiffusionModel, Sample, _state
from .mlp_diffusion import DiffusionMLP

log = logging.getLogger(__name__)


class VPGDiffusion(DiffusionModel):
    def __init__(self, actor, critic, ft_denoising_steps, ft_denoising_steps_d=0, ft_denoising_steps_t=0,
                 network_path=None, min_sampling_denoising_std=0.1, min_logprob_denoising_std=0.1,
                 eta=None, learn_eta=False, _cfg_hook=None, **kwargs):
        if eta is not None or learn_eta:
            raise NotImplementedError("eta is a DDIM feature (diffusion_vpg.py:52); this library is DDPM only")
        self.ft_denoising_steps = int(ft_denoising_steps)
        self.ft_denoising_steps_d, self.ft_denoising_steps_t = ft_denoising_steps_d, ft_denoising_steps_t
        self.ft_denoising_steps_cnt = 0
        self.min_sampling_denoising_std = min_sampling_denoising_std
        self.min_logprob_denoising_std = float(min_logprob_denoising_std)
        self.learn_eta = False

        def hook(cfg):
            cfg.ft_denoising_steps = self.ft_denoising_steps
            cfg.min_sampling_denoising_std = float(self.get_min_sampling_denoising_std())
            cfg.min_logprob_denoising_std = self.min_logprob_denoising_std
            cfg.critic_hidden = critic.mlp_dims[0]
            cfg.critic_act = critic.activation_id
            if _cfg_hook is not None:
                _cfg_hook(cfg)

        super().__init__(network=actor, network_path=network_path, _cfg_hook=hook, **kwargs)
        assert self.ft_denoising_steps <= self.denoising_steps                      # diffusion_vpg.py:50
        self.actor = self.network                                                    # :76
        # fine-tuned copy starts as a clone of the (loaded) base policy (:94-97)
        self.actor_ft = DiffusionMLP(actor.action_dim, actor.horizon_steps, actor.cond_dim, time_dim=actor.time_dim,
                                     mlp_dims=actor.mlp_dims, activation_type=actor.activation_type, residual_style=True)
        self.actor_ft._bind(self.engine, L.NET_ACTOR_FT)
        self.actor_ft.set_flat_weights(self.actor.get_flat_weights())
        self.critic = critic
        critic._bind(self.engine, L.NET_CRITIC)

    # ---- Keras model.save_weights / load_weights of the whole fine-tuning model (agent/finetune/train_agent.py:127-142): one
    #      `.weights.h5` file with actor/ (= network), actor_ft/ and critic/ (util/keras_h5.py documents the assumed layout)
    def save_weights(self, path):
        from ...util.keras_h5 import write_h5
        data = {}
        for prefix, net in (("actor", self.actor), ("actor_ft", self.actor_ft), ("critic", self.critic)):
            for pth, w in zip(net.keras_variable_paths(prefix), net.get_weights()):
                data[pth] = w
        write_h5(str(path), data)

    def load_weights(self, path):
        from ...util.keras_h5 import load_keras_weights_h5
        for prefix, net in (("actor", self.actor), ("actor_ft", self.actor_ft), ("critic", self.critic)):
            net.set_weights(load_keras_weights_h5(str(path), net.keras_variable_paths(prefix), net.shapes))

    # ---- diffusion_vpg.py:114-148
    def step(self):
        if type(self.min_sampling_denoising_std) is not float:
            self.min_sampling_denoising_std.step()
        self.ft_denoising_steps_cnt += 1
        if (self.ft_denoising_steps_d > 0 and self.ft_denoising_steps_t > 0
                and self.ft_denoising_steps_cnt % self.ft_denoising_steps_t == 0):
            self.ft_denoising_steps = max(0, self.ft_denoising_steps - self.ft_denoising_steps_d)
            self.engine.set_ft_denoising_steps(self.ft_denoising_steps)
            # the fine-tuned actor becomes the new base (:137-138)
            self.actor.set_flat_weights(self.actor_ft.get_flat_weights())
            log.info("Finished annealing fine-tuning denoising steps to %d", self.ft_denoising_steps)

    def get_min_sampling_denoising_std(self):
        if type(self.min_sampling_denoising_std) is float:
            return self.min_sampling_denoising_std
        return self.min_sampling_denoising_std()

    # ---- diffusion_vpg.py:151-245
    def p_mean_var(self, x, t, cond, index=None, use_base_policy=False, deterministic=False):
        t = torch.as_tensor(t, device=self.device)
        # the whole batch switches on the FIRST row's t (:165,172)
        use_ft = bool(t[0] < self.ft_denoising_steps) and not use_base_policy
        mu, logvar = super().p_mean_var(x, t, cond, network_override=self.actor_ft if use_ft else self.actor)
        return mu, logvar, torch.ones_like(mu)

    # ---- diffusion_vpg.py:249-339
    def __call__(self, cond, deterministic=False, return_chain=True, use_base_policy=False, x_T=None, noise=None):
        """cond {"state": (B,To,Do)} -> Sample(trajectories (B,Ta,Da), chains (B,K+1,Ta,Da)).
        Host (NumPy / CPU tensor) observations take the host-buffer entry point and return NumPy,
        like the reference caller's np.array(...) (train_ppo_diffusion_agent.py:111-132)."""
        state = _state(cond)
        min_std = float(self.get_min_sampling_denoising_std())
        off = self._next_offset()
        host = isinstance(state, np.ndarray) or (isinstance(state, torch.Tensor) and not state.is_cuda)
        if host and x_T is None and noise is None:
            obs = np.ascontiguousarray(state.numpy() if isinstance(state, torch.Tensor) else state, np.float32)
            B = obs.shape[0]
            obs = obs.reshape(B, -1)
            actions = np.empty((B, self.horizon_steps, self.action_dim), np.float32)
            chains = np.empty((B, self.ft_denoising_steps + 1, self.horizon_steps, self.action_dim), np.float32) if return_chain else None
            self.engine.sample_host(obs, actions, chains, deterministic=deterministic, use_base_policy=use_base_policy,
                                    min_sampling_std=min_std, seed=self.seed, offset=off)
            return Sample(actions, chains)
        actions, chains = self.engine.sample(state, deterministic=deterministic, use_base_policy=use_base_policy,
                                             min_sampling_std=min_std, seed=self.seed, offset=off, x_T=x_T, noise=noise,
                                             return_chain=return_chain)
        actions = actions.reshape(-1, self.horizon_steps, self.action_dim)
        if chains is not None:
            chains = chains.reshape(-1, self.ft_denoising_steps + 1, self.horizon_steps, self.action_dim)
        return Sample(actions, chains)

    call = __call__

    # ---- diffusion_vpg.py:343-425
    def get_logprobs(self, cond, chains, get_ent: bool = False, use_base_policy: bool = False):
        """-> (B*K, Ta, Da), row = b*K + k.  (The reference tiles cond in place, :374-379; here the
        tiling is index arithmetic inside the kernel and the caller's dict is left alone.)"""
        logp = self.engine.logprobs(_state(cond), chains, use_base_policy=use_base_policy)
        logp = logp.reshape(-1, self.horizon_steps, self.action_dim)
        return (logp, torch.ones_like(logp)) if get_ent else logp

    # ---- diffusion_vpg.py:427-481
    def get_logprobs_subsample(self, cond, chains_prev, chains_next, denoising_inds, get_ent: bool = False,
                               use_base_policy: bool = False):
        logp = self.engine.logprobs_subsample(_state(cond), chains_prev, chains_next, denoising_inds,
                                              use_base_policy=use_base_policy)
        logp = logp.reshape(-1, self.horizon_steps, self.action_dim)
        return (logp, torch.ones_like(logp)) if get_ent else logp
