"""Generates tests/golden/ref_reward_scaling.npz by running the reference's util/reward_scaling.py itself
(pure NumPy: imported unmodified from /root/reference, no shim).   python tests/golden/make_ref_reward_scaling.py"""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DPPO_REFERENCE_ROOT", "/root/reference")
spec = importlib.util.spec_from_file_location("ref_reward_scaling", os.path.join(REF, "util", "reward_scaling.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(5)
out = {}
# per_env=True cannot run in the reference for n_steps != n_envs (mean over axis 0 of [E,S] vs shape (E,)); the agent uses the default
for name, E, S, iters, gamma in (("hopper_like", 6, 9, 3, 0.99), ("square", 5, 5, 2, 0.9)):
    scaler = ref.RunningRewardScaler(num_envs=E, gamma=gamma)
    rewards = rng.normal(size=(iters, E, S)) * np.array([1.0, 4.0, 0.3])[:iters, None, None]
    firsts = (rng.uniform(size=(iters, E, S)) < 0.15).astype(np.float64)
    firsts[0, :, 0] = 1
    scaled = np.stack([scaler(reward=rewards[i], first=firsts[i]) for i in range(iters)])
    out.update({f"{name}_rewards": rewards, f"{name}_firsts": firsts, f"{name}_scaled": scaled,
                f"{name}_var": np.asarray(scaler.ret_rms.var), f"{name}_mean": np.asarray(scaler.ret_rms.mean),
                f"{name}_count": np.asarray(scaler.ret_rms.count), f"{name}_ret": scaler.ret, f"{name}_gamma": np.array([gamma])})
np.savez_compressed(os.path.join(HERE, "ref_reward_scaling.npz"), **out)
print({k: np.asarray(v).shape for k, v in out.items()})
