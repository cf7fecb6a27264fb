"""Generates tests/golden/ref_gae.npz by executing the reference's OWN GAE block
(agent/finetune/train_ppo_diffusion_agent.py: from `advantages_trajs = np.zeros_like(reward_trajs)` to
`returns_trajs = advantages_trajs + values_trajs`), sliced out of the file at generation time and exec'd verbatim on NumPy inputs
with a stub `self` (the block sits inside `TrainPPODiffusionAgent.run`, which needs Hydra / gym / TF and cannot be imported).
The critic call for the bootstrap value goes through the TF shim with a stub critic.   python tests/golden/make_ref_gae.py"""
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DPPO_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
import tensorflow as tf  # noqa: E402  (the shim)

src = open(os.path.join(REF, "agent", "finetune", "train_ppo_diffusion_agent.py")).read().splitlines()
a = next(i for i, l in enumerate(src) if l.strip() == "advantages_trajs = np.zeros_like(reward_trajs)")
b = next(i for i, l in enumerate(src) if l.strip() == "returns_trajs = advantages_trajs + values_trajs")
block = textwrap.dedent("\n".join(src[a:b + 1]))
print(f"reference lines {a + 1}-{b + 1}:\n{block}\n")

out = {}
rng = np.random.default_rng(17)
for name, S, E, gamma, lam, rsc in (("a", 7, 5, 0.99, 0.95, 1.0), ("b", 12, 3, 0.999, 0.9, 0.1)):
    reward_trajs = rng.normal(size=(S, E))
    terminated_trajs = (rng.uniform(size=(S, E)) < 0.2).astype(np.float64)
    values_trajs = rng.normal(size=(S, E)).astype(np.float32).astype(np.float64)      # np.vstack of fp32 critic outputs (:205-209)
    next_values = rng.normal(size=(E,)).astype(np.float32)
    self = types.SimpleNamespace(
        n_steps=S, gamma=gamma, gae_lambda=lam, reward_scale_const=rsc,
        model=types.SimpleNamespace(critic=lambda obs: torch.from_numpy(next_values).reshape(-1, 1)))
    ns = dict(np=np, tf=tf, self=self, reward_trajs=reward_trajs, terminated_trajs=terminated_trajs, values_trajs=values_trajs,
              obs_venv_ts={"state": None})
    exec(block, ns)
    out.update({f"{name}_rewards": reward_trajs, f"{name}_terminated": terminated_trajs, f"{name}_values": values_trajs.astype(np.float32),
                f"{name}_next_values": next_values, f"{name}_hyper": np.array([gamma, lam, rsc]),
                f"{name}_advantages": ns["advantages_trajs"], f"{name}_returns": ns["returns_trajs"]})
np.savez_compressed(os.path.join(HERE, "ref_gae.npz"), **out)
print({k: v.shape for k, v in out.items()})
