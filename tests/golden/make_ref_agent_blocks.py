"""Generates tests/golden/ref_agent_blocks.npz by exec-ing two more blocks of the reference's
agent/finetune/train_ppo_diffusion_agent.py verbatim (sliced out of the file at generation time; `run` itself needs Hydra / gym):

  * the minibatch assembly (`inds_b = inds_k[start:end]` ... `logprobs_b = tf.gather_nd(...)`, :292-312) over the TF shim, and
  * the episode statistics (`episodes_start_end = []` ... the "No episode completed" branch, :144-183) in plain NumPy.

   python tests/golden/make_ref_agent_blocks.py"""
import logging
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


def block(first, last_prefix, extra=0):
    a = next(i for i, l in enumerate(src) if l.strip().startswith(first))
    b = next(i for i, l in enumerate(src) if i > a and l.strip().startswith(last_prefix)) + extra
    return textwrap.dedent("\n".join(src[a:b + 1])), (a + 1, b + 1)


out = {}
rng = np.random.default_rng(23)

# ---- minibatch assembly
code, lines = block("inds_b = inds_k[start:end]", "logprobs_b = tf.gather_nd(", extra=3)
print(f"minibatch assembly, reference lines {lines[0]}-{lines[1]}:\n{code}\n")
S, E, K, Ta, Da, Do = 5, 4, 3, 4, 3, 11
P = S * E
f32 = lambda a: torch.from_numpy(a.astype(np.float32))   # noqa: E731
ns = dict(tf=tf, np=np, start=7, end=7 + 20,
          self=types.SimpleNamespace(n_steps=S, n_envs=E, model=types.SimpleNamespace(ft_denoising_steps=K)),
          inds_k=torch.from_numpy(rng.permutation(P * K)),
          obs_k={"state": f32(rng.normal(size=(P, 1, Do)))}, chains_k=f32(rng.normal(size=(P, K + 1, Ta, Da))),
          returns_k=f32(rng.normal(size=P)), values_k=f32(rng.normal(size=P)), advantages_k=f32(rng.normal(size=P)),
          logprobs_k=f32(rng.normal(size=(P, K, Ta, Da))))
inputs = {k: (v["state"] if isinstance(v, dict) else v) for k, v in ns.items() if k.endswith("_k")}
exec(code, ns)
for k, v in inputs.items():
    out["mb_" + k] = v.numpy()
out["mb_start_end_K"] = np.array([7, 27, K])
for k in ("inds_b", "batch_inds_b", "denoising_inds_b", "chains_prev_b", "chains_next_b", "returns_b", "values_b", "advantages_b", "logprobs_b"):
    out["mb_" + k] = ns[k].numpy()
out["mb_obs_b"] = ns["obs_b"]["state"].numpy()

# ---- episode statistics
code, lines = block("episodes_start_end = []", 'log.info("[WARNING] No episode completed')
print(f"episode statistics, reference lines {lines[0]}-{lines[1]}:\n{code}\n")
for name, S, E, p_first in (("some", 12, 5, 0.25), ("none", 4, 3, 0.0)):
    firsts = (rng.uniform(size=(S + 1, E)) < p_first).astype(np.float64)
    firsts[0] = 1
    rewards = rng.normal(size=(S, E)) + 1.0
    ns = dict(np=np, log=logging.getLogger("ref"), firsts_trajs=firsts, reward_trajs=rewards,
              self=types.SimpleNamespace(n_envs=E, furniture_sparse_reward=False, act_steps=4, best_reward_threshold_for_success=0.3))
    exec(code, ns)
    out.update({f"ep_{name}_firsts": firsts, f"ep_{name}_rewards": rewards,
                f"ep_{name}_stats": np.array([ns["num_episode_finished"], ns["avg_episode_reward"], ns["avg_best_reward"], ns["success_rate"]], np.float64)})
np.savez_compressed(os.path.join(HERE, "ref_agent_blocks.npz"), **out)
print({k: v.shape for k, v in out.items()})
