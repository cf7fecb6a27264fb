"""Loader for tests/golden/ref_*.npz — fixtures produced by running the reference's own model code over the
torch-backed TF shim (tests/golden/make_ref_golden.py).  Rebuilds the oracle with the fixture's configuration."""
import glob
import os

import numpy as np

from oracle import dppo_oracle as O

REF_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_hopper_*.npz")))
REF_IDS = [os.path.basename(p)[4:-4] for p in REF_GOLDEN]
BATCH_KEYS = ("ppo_obs", "ppo_prev", "ppo_next", "ppo_inds", "ppo_returns", "ppo_oldvalues", "ppo_adv", "ppo_oldlogp")


def load(path):
    z = np.load(path)

    def kw(name, default):
        key = "kw_" + name
        if key not in z.files:
            return default
        v = float(z[key][0])
        return None if np.isnan(v) else v

    dflt = O.Hyper()
    h = O.Hyper(**{f: kw(f, getattr(dflt, f)) for f in (
        "randn_clip_value", "final_action_clip_value", "min_sampling_denoising_std", "min_logprob_denoising_std",
        "gamma_denoising", "clip_ploss_coef", "clip_ploss_coef_base", "clip_ploss_coef_rate", "clip_vloss_coef")})
    o = O.make_oracle("hopper", seed=int(z["seed"][0]), hyper=h, denoising_steps=int(z["denoising_steps"][0]),
                      ft_denoising_steps=int(z["ft_denoising_steps"][0]))
    for k in ("actor", "actor_ft", "critic"):
        np.testing.assert_array_equal(O.flatten_params(getattr(o, k))[::97], z[k + "_fp"])
    return o, z, int(z["reward_horizon"][0])


def grads_summary(grads):
    """Same digest make_ref_golden.py stores: strided fingerprint, per-tensor L2 norms, last two tensors in full."""
    flat = np.concatenate([np.asarray(g, np.float32).reshape(-1) for g in grads])
    norms = np.array([float(np.linalg.norm(np.asarray(g, np.float64))) for g in grads])
    tail = np.concatenate([np.asarray(g, np.float32).reshape(-1) for g in grads[-2:]])
    return flat[::97], norms, tail
