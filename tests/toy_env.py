"""A deterministic vector environment with the reference's venv API (reset_arg / step), for agent-loop tests."""
import numpy as np


class ToyVecEnv:
    def __init__(self, n_envs, obs_dim, action_dim, max_episode_steps=4, seed=0, term_threshold=0.93):
        rng = np.random.default_rng(seed)
        self.E, self.Do, self.Da = n_envs, obs_dim, action_dim
        self.W = rng.normal(size=(action_dim, obs_dim)) / np.sqrt(action_dim)
        self.s0 = rng.uniform(-0.8, 0.8, size=(n_envs, obs_dim))
        self.max_episode_steps, self.term_threshold = max_episode_steps, term_threshold
        self.reset_arg()

    def shard(self, lo, hi):
        """The env copies [lo, hi) of this vector env as their own vector env (data-parallel ranks step their own block)."""
        import copy
        o = copy.copy(self)
        o.E = hi - lo
        o.s0 = self.s0[lo:hi].copy()
        o.reset_arg()
        return o

    def _obs(self):
        return {"state": self.s[:, None, :].astype(np.float32)}

    def reset_arg(self, options_list=None):
        self.s = self.s0.copy()
        self.t = np.zeros(self.E, dtype=np.int64)
        return self._obs()

    def step(self, action):
        a = np.asarray(action, np.float64).mean(axis=1)
        self.s = 0.9 * self.s + 0.3 * np.tanh(a @ self.W)
        self.t += 1
        reward = 1.0 - (self.s ** 2).sum(-1) + 0.1 * a.sum(-1)
        terminated = np.abs(self.s).max(-1) > self.term_threshold
        truncated = self.t >= self.max_episode_steps
        done = terminated | truncated
        self.s[done] = self.s0[done]          # auto-reset like the reference's async vector env
        self.t[done] = 0
        return self._obs(), reward, terminated, truncated, [{} for _ in range(self.E)]


class RawToyVecEnv(ToyVecEnv):
    """The same dynamics behind RAW interfaces, like a MuJoCo env before MujocoLocomotionLowdimWrapper: float64 observations in
    physical units, actions in physical units.  `norm` holds the task's obs_min / obs_max / action_min / action_max (fp32)."""

    def __init__(self, n_envs, obs_dim, action_dim, norm, **kw):
        self.norm = norm
        super().__init__(n_envs, obs_dim, action_dim, **kw)

    def _obs(self):
        n = self.norm       # physical units: invert normalize_obs in float64
        raw = (self.s / 2 + 0.5) * (n["obs_max"].astype(np.float64) - n["obs_min"] + 1e-6) + n["obs_min"]
        return {"state": raw[:, None, :]}

    def step(self, raw_action):
        n = self.norm       # back to [-1, 1] for the toy dynamics
        a = (np.asarray(raw_action, np.float64) - n["action_min"]) / (n["action_max"].astype(np.float64) - n["action_min"]) * 2 - 1
        return super().step(a)


class NormalizingVecWrapper:
    """env/gym_utils/wrapper/mujoco_locomotion_lowdim.py:57-62 restated for a vector env (same NumPy expressions per element)."""

    def __init__(self, env, norm):
        self.env = env
        self.obs_min, self.obs_max = norm["obs_min"], norm["obs_max"]
        self.action_min, self.action_max = norm["action_min"], norm["action_max"]

    def normalize_obs(self, obs):
        return 2 * ((obs - self.obs_min) / (self.obs_max - self.obs_min + 1e-6) - 0.5)

    def unnormalize_action(self, action):
        action = (action + 1) / 2  # [-1, 1] -> [0, 1]
        return action * (self.action_max - self.action_min) + self.action_min

    def reset_arg(self, options_list=None):
        return {"state": self.normalize_obs(self.env.reset_arg(options_list)["state"])}

    def step(self, action):
        obs, r, term, trunc, info = self.env.step(self.unnormalize_action(action))
        return {"state": self.normalize_obs(obs["state"])}, r, term, trunc, info
