"""Running reward scaling, host side (float64 NumPy like the reference): util/reward_scaling.py:13-100.

Rewards are divided by the running standard deviation of the discounted return-so-far (no mean subtraction) and clipped.
The arrays involved are [n_envs, n_steps] per PPO iteration (40 x 500 for the shipped configs): host work, not a kernel.
"""
import numpy as np


class RunningMeanStd:
    """Parallel-variance (Chan et al.) accumulator; util/reward_scaling.py:13-40.  Note the reference's `var`
    divides the merged M2 by (count - 1) while the per-batch term uses the population variance — kept as is."""

    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape)
        self.var = np.ones(shape)
        self.count = epsilon

    def update(self, x):
        self.update_from_moments(np.mean(x, axis=0), np.var(x, axis=0), x.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        n_old, n_new = self.count, self.count + batch_count
        shift = batch_mean - self.mean
        merged_m2 = self.var * n_old + batch_var * batch_count + shift ** 2 * n_old * batch_count / n_new
        self.mean = self.mean + shift * batch_count / n_new
        self.var = merged_m2 / (n_new - 1)
        self.count = n_new


def backward_discounted_sum(prevret, reward, first, gamma):
    """ret[:, t] = reward[:, t] + (1 - first[:, t]) * gamma * ret[:, t-1], seeded with `prevret` (reward_scaling.py:89-100)."""
    assert first.ndim == 2
    out = np.zeros_like(reward)
    carry = prevret
    for t in range(reward.shape[1]):
        carry = out[:, t] = reward[:, t] + (1 - first[:, t]) * gamma * carry
    return out


class RunningRewardScaler:
    """util/reward_scaling.py:43-86.  __call__(reward [E,S], first [E,S]) -> scaled reward [E,S]; carries the last
    return of every env and the running variance across iterations."""

    def __init__(self, num_envs, cliprew=10.0, gamma=0.99, epsilon=1e-8, per_env=False):
        self.ret_rms = RunningMeanStd(shape=(num_envs,) if per_env else ())
        self.cliprew, self.gamma, self.epsilon, self.per_env = cliprew, gamma, epsilon, per_env
        self.ret = np.zeros(num_envs)

    def __call__(self, reward, first):
        rets = backward_discounted_sum(self.ret, reward, first, self.gamma)
        self.ret = rets[:, -1]
        self.ret_rms.update(rets if self.per_env else rets.reshape(-1))
        return self.transform(reward)

    def transform(self, reward):
        return np.clip(reward / np.sqrt(self.ret_rms.var + self.epsilon), -self.cliprew, self.cliprew)
