"""PPODiffusion with the reference's surface (model/diffusion/diffusion_ppo.py:7-132)."""
import torch

from .diffusion import _state
from .diffusion_vpg import VPGDiffusion


class PPODiffusion(VPGDiffusion):
    def __init__(self, gamma_denoising, clip_ploss_coef, clip_ploss_coef_base=1e-3, clip_ploss_coef_rate=3,
                 clip_vloss_coef=None, clip_advantage_lower_quantile=0, clip_advantage_upper_quantile=1,
                 norm_adv=True, vf_coef=0.5, weight_decay=0.004, adam_eps=1e-7, **kwargs):
        self.gamma_denoising, self.clip_ploss_coef = gamma_denoising, clip_ploss_coef
        self.clip_ploss_coef_base, self.clip_ploss_coef_rate = clip_ploss_coef_base, clip_ploss_coef_rate
        self.clip_vloss_coef, self.norm_adv = clip_vloss_coef, norm_adv
        self.clip_advantage_lower_quantile, self.clip_advantage_upper_quantile = clip_advantage_lower_quantile, clip_advantage_upper_quantile
        self.vf_coef = vf_coef
        self.last_gradients = None
        user_hook = kwargs.pop("_cfg_hook", None)

        def hook(cfg):
            cfg.gamma_denoising = float(gamma_denoising)
            cfg.clip_ploss_coef, cfg.clip_ploss_coef_base = float(clip_ploss_coef), float(clip_ploss_coef_base)
            cfg.clip_ploss_coef_rate = float(clip_ploss_coef_rate)
            cfg.clip_vloss_coef = -1.0 if clip_vloss_coef is None else float(clip_vloss_coef)
            cfg.norm_adv = int(bool(norm_adv))
            cfg.vf_coef = float(vf_coef)                 # train_ppo_diffusion_agent.py:340
            cfg.weight_decay = float(weight_decay)       # Keras-3 AdamW default (the `decay=` kwarg is ignored)
            cfg.adam_eps = float(adam_eps)
            if user_hook is not None:
                user_hook(cfg)

        super().__init__(_cfg_hook=hook, **kwargs)

    def _run(self, obs, chains_prev, chains_next, denoising_inds, returns, oldvalues, advantages, oldlogprobs,
             use_bc_loss, reward_horizon, lr, apply, n_global, adv_mean, adv_std):
        if use_bc_loss:
            raise NotImplementedError("use_bc_loss is False in every reference cfg (diffusion_ppo.py:62-71)")
        if reward_horizon != self.cfg.reward_horizon:
            raise ValueError("reward_horizon is fixed at construction (cfg.reward_horizon = act_steps)")
        m, g = self.engine.ppo_step(_state(obs), chains_prev, chains_next, denoising_inds, returns, oldvalues, advantages,
                                    oldlogprobs, lr=lr, apply=apply, n_global=n_global, adv_mean=adv_mean, adv_std=adv_std,
                                    want_grads=True)
        self.last_gradients = g
        return tuple(m[i] for i in range(8))

    # ---- diffusion_ppo.py:32-132: returns (pg_loss, entropy_loss, v_loss, clipfrac, approx_kl, ratio, bc_loss, eta)
    def c_loss(self, obs, chains_prev, chains_next, denoising_inds, returns, oldvalues, advantages, oldlogprobs,
               use_bc_loss=False, reward_horizon=4):
        """Loss scalars; the flat gradient of pg_loss + vf_coef*v_loss wrt [actor_ft, critic] is left in
        `self.last_gradients` (what tape.gradient returns at train_ppo_diffusion_agent.py:345-346)."""
        return self._run(obs, chains_prev, chains_next, denoising_inds, returns, oldvalues, advantages, oldlogprobs,
                         use_bc_loss, reward_horizon, 0.0, False, None, 0.0, -1.0)

    def ppo_update(self, obs, chains_prev, chains_next, denoising_inds, returns, oldvalues, advantages, oldlogprobs,
                   lr, use_bc_loss=False, reward_horizon=4, n_global=None, adv_mean=0.0, adv_std=-1.0):
        """c_loss + tape.gradient + (all-reduce) + actor_optimizer.apply_gradients in one fused call
        (train_ppo_diffusion_agent.py:328-356)."""
        return self._run(obs, chains_prev, chains_next, denoising_inds, returns, oldvalues, advantages, oldlogprobs,
                         use_bc_loss, reward_horizon, lr, True, n_global, adv_mean, adv_std)

    def ppo_update_indexed(self, obs_k, chains_k, logprobs_k, returns_k, values_k, advantages_k, inds_k, lr,
                           n_global=None, adv_mean=0.0, adv_std=-1.0, apply=True):
        """One minibatch of train_ppo_diffusion_agent.py:287-356 from the device-resident rollout of this iteration:
        obs_k [P,1,Do], chains_k [P,K+1,Ta,Da], logprobs_k [P,K,Ta,Da], returns_k / values_k / advantages_k [P] stay on
        the GPU across the update epochs; `inds_k` is the slice of the shuffled flat (step*env, k) index the reference
        unravels at :293-296.  Returns the 8 loss scalars (device tensor)."""
        P = obs_k.shape[0]
        return self.engine.ppo_step_indexed(_state({"state": obs_k}) if not isinstance(obs_k, dict) else _state(obs_k),
                                            chains_k.reshape(P, self.ft_denoising_steps + 1, -1), logprobs_k.reshape(P, self.ft_denoising_steps, -1),
                                            returns_k, values_k, advantages_k, inds_k, lr=lr, apply=apply, n_global=n_global,
                                            adv_mean=adv_mean, adv_std=adv_std)
