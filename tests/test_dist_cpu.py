"""CPU, gloo, world_size 2: the data-parallel contract of SURVEY.md §8(e).  Each rank takes its
shard of a PPO minibatch, uses GLOBAL advantage statistics and 1/N_global scaling, and a single
sum all-reduce of [flat gradient ++ metric partial sums] reproduces the un-sharded result."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusionpolicyoptimization_b200.parallel import advantage_stats, shard_range
from oracle import dppo_oracle as O


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 40, 50000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _sharded_loss(o, batch, lo, hi, n_global, mean, std):
    """What one rank computes: sums over its rows divided by N_global, advantages normalised with
    the global statistics (the library's dppo_ppo_step contract)."""
    sub = tuple(b[lo:hi] for b in batch)
    ft = [p.clone().requires_grad_(True) for p in o.actor_ft]
    cr = [p.clone().requires_grad_(True) for p in o.critic]
    h = o.h
    adv = (sub[6] - mean) / (std + 1e-8)
    o2 = O.Oracle(o.d, O.Hyper(**{**h.__dict__, "norm_adv": False}), o.actor, o.actor_ft, o.critic)
    out = o2.ppo_loss(sub[0], sub[1], sub[2], sub[3], sub[4], sub[5], adv, sub[7], actor_ft=ft, critic=cr)
    frac = (hi - lo) / n_global
    loss = (out[0] + h.vf_coef * out[2]) * frac
    grads = torch.autograd.grad(loss, ft + cr)
    flat = torch.cat([g.reshape(-1) for g in grads])
    metrics = torch.stack([out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7]]).detach() * frac
    return torch.cat([flat, metrics])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    o = O.make_oracle("hopper", seed=0, actor_hidden=64, critic_hidden=32)
    batch = O.make_ppo_batch(o, 101, pool=16, seed=3)      # odd size: ragged shards
    mean, std = advantage_stats(batch[6].numpy())
    lo, hi = shard_range(101, rank, world)
    buf = _sharded_loss(o, batch, lo, hi, 101, mean, std)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)              # the ONE collective
    if rank == 0:
        q.put(buf.numpy())
    dist.destroy_process_group()


def test_two_rank_gradient_equals_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    o = O.make_oracle("hopper", seed=0, actor_hidden=64, critic_hidden=32)
    batch = O.make_ppo_batch(o, 101, pool=16, seed=3)
    metrics, ga, gc = o.ppo_grads(*batch)
    want = np.concatenate([O.flatten_params(ga), O.flatten_params(gc), np.array([float(m) for m in metrics], np.float32)])
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-6)


# ---------------------------------------------------------------------------------------------------------------------
# the data-parallel agent's host logic (agent/finetune/train_ppo_diffusion_agent.py of the package): every rank draws the same
# permutation over the GLOBAL (step, env, k) pool and keeps the rows of its env block; the minibatch's advantage statistics come
# from one all-reduce of (sum, sum of squares)
def test_local_minibatch_partitions_the_global_one():
    from diffusionpolicyoptimization_b200.parallel import local_minibatch
    S, E, K = 5, 7, 3
    rng = np.random.default_rng(0)
    pool = rng.standard_normal((S, E, K))                        # one value per (step, env, k)
    perm = rng.permutation(S * E * K)
    for world in (1, 2, 3):
        for b0 in range(0, S * E * K - 20, 20):
            mb = perm[b0:b0 + 20]
            want = np.sort(pool.reshape(-1)[mb])
            got = []
            for r in range(world):
                lo, hi = shard_range(E, r, world)
                loc = local_minibatch(mb, E, K, lo, hi)
                assert loc.dtype == np.int32 and (loc >= 0).all() and (loc < S * (hi - lo) * K).all()
                got.append(np.ascontiguousarray(pool[:, lo:hi]).reshape(-1)[loc])    # the rank's own resident rollout
            np.testing.assert_array_equal(np.sort(np.concatenate(got)), want)


def _agent_stats_worker(rank, world, port, q):
    from diffusionpolicyoptimization_b200.parallel import local_minibatch, stats_from_moments
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S, E, K = 6, 8, 10
    rng = np.random.default_rng(5)
    adv = rng.standard_normal((S, E)).astype(np.float32)         # advantages per (step, env), identical on both ranks here
    perm = np.random.default_rng(9).permutation(S * E * K)      # the same draw on every rank
    lo, hi = shard_range(E, rank, world)
    mine = torch.from_numpy(np.ascontiguousarray(adv[:, lo:hi]).reshape(-1))
    out = []
    for b0 in range(0, S * E * K, 160):
        mb = perm[b0:b0 + 160]
        loc = local_minibatch(mb, E, K, lo, hi)
        a = mine[torch.from_numpy(loc // K).long()].to(torch.float64)
        mom = torch.stack([a.sum(), (a * a).sum()])
        dist.all_reduce(mom)                                     # the one small collective per minibatch
        out.append(stats_from_moments(float(mom[0]), float(mom[1]), len(mb)))
    if rank == 0:
        q.put((np.array(out), adv, perm))
    dist.destroy_process_group()


def test_two_rank_advantage_statistics_equal_the_unsharded_ones():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_agent_stats_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, adv, perm = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    K = 10
    for i, b0 in enumerate(range(0, adv.size * K, 160)):
        rows = adv.reshape(-1)[perm[b0:b0 + 160] // K].astype(np.float64)      # diffusion_ppo.py:74-75 on the un-sharded minibatch
        np.testing.assert_allclose(got[i], [rows.mean(), rows.std()], rtol=2e-6, atol=1e-7)
