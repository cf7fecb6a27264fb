"""Host-side data-parallel plumbing (absent in the reference).

One process per GPU.  Rows (env copies for the rollout, minibatch rows for the updates) are split
evenly; weights are replicated.  The rollout needs no collective.  An update needs exactly one:
the library's sum all-reduce of [flat gradient ++ metric partial sums], issued inside
`dppo_ppo_step` / `dppo_pretrain_step`.  What must be global *before* the step is the advantage
normalisation (diffusion_ppo.py:74-75 uses the whole minibatch), computed here once.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of n rows for `rank` (first n % world ranks get one extra)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def advantage_stats(advantages) -> Tuple[float, float]:
    """Population mean / std over the un-sharded minibatch (tf.reduce_mean / tf.math.reduce_std)."""
    a = np.asarray(advantages, dtype=np.float64).reshape(-1)
    return float(a.mean()), float(a.std())


def local_minibatch(inds_global, n_envs: int, K: int, env_lo: int, env_hi: int) -> np.ndarray:
    """Rows of a GLOBAL minibatch that live on the rank owning env columns [env_lo, env_hi), as LOCAL flat indices.

    The reference indexes the flattened (step, env, k) pool of the whole rollout (train_ppo_diffusion_agent.py:287-296):
    global flat f = (step * n_envs + env) * K + k.  A rank stores only its env columns, local flat =
    (step * (env_hi - env_lo) + (env - env_lo)) * K + k.  Every rank draws the same permutation, so the union of the
    ranks' local pieces is exactly the reference's minibatch."""
    f = np.asarray(inds_global, dtype=np.int64).reshape(-1)
    k = f % K
    se = f // K
    env = se % n_envs
    step = se // n_envs
    mine = (env >= env_lo) & (env < env_hi)
    loc = (step[mine] * (env_hi - env_lo) + (env[mine] - env_lo)) * K + k[mine]
    return loc.astype(np.int32)


def stats_from_moments(s1: float, s2: float, n: int) -> Tuple[float, float]:
    """Population mean / std from the all-reduced sum and sum of squares (same formula as the library's adv_stats_kernel)."""
    mean = s1 / n
    var = s2 / n - mean * mean
    return float(np.float32(mean)), float(np.float32(np.sqrt(var if var > 0 else 0.0)))


def init_process_group_from_env(backend: str = "nccl"):
    """torchrun rendezvous (RANK / WORLD_SIZE / MASTER_* from the environment) -> (rank, world, local_rank)."""
    import os
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local
