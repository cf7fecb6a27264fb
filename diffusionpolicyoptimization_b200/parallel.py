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
