"""GPU (-m gpu), needs >= 2 B200s on the box (skipped otherwise): the data-parallel update.

Two ranks take PPO steps on their own row shards through (a) the fused peer-memory all-reduce + AdamW kernel (one-shot and
two-shot) and (b) ncclAllReduce + AdamW, in the bf16 and the fp32-faithful bf16x3 tensor modes.  Checked on hardware:
  * the REDUCED gradient of every exchange equals the gradient of the un-sharded minibatch on one rank (no communicator)
    to 2e-4 of its largest entry (only fp32 summation order differs), and so do the metrics;
  * the replicas stay bit-identical across ranks after three updates, and the three exchanges agree with each other and with
    the single-rank AdamW step in units of the learning rate;
  * a loss-only step (apply = 0, the critic warm-up iterations of train_ppo_diffusion_agent.py:340-356) still hands out the
    GLOBAL metrics and gradient on every rank, so a KL early stop is taken identically everywhere.
Run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`; the log is kept under profiles/."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LR = 1e-3


def _worker(rank, world, port, prec_name, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from diffusionpolicyoptimization_b200 import _lib as L
    from diffusionpolicyoptimization_b200.parallel import advantage_stats, shard_range
    from oracle import dppo_oracle as O
    from helpers import make_engine
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    prec = {"bf16": L.PREC_BF16, "bf16x3": L.PREC_BF16X3}[prec_name]
    o = O.make_oracle("hopper", seed=0)
    N = 8192
    batch = O.make_ppo_batch(o, N, pool=512, seed=7)
    flat = [batch[0].reshape(N, -1), batch[1].reshape(N, -1), batch[2].reshape(N, -1), batch[3], batch[4], batch[5], batch[6], batch[7].reshape(N, -1)]
    lo, hi = shard_range(N, rank, world)
    shard = [t[lo:hi] for t in flat]
    mean, std = advantage_stats(batch[6].numpy())
    # the un-sharded reference on this rank's GPU: no communicator, same weights, same global statistics
    e1 = make_engine(o, precision=prec, device=rank)
    m_ref, g_ref = e1.ppo_step(*flat, lr=LR, apply=True, n_global=N, adv_mean=mean, adv_std=std, want_grads=True)
    g_ref = g_ref.cpu().numpy(); m_ref = m_ref.cpu().numpy()
    w_ref = np.concatenate([e1.get_weights(L.NET_ACTOR_FT), e1.get_weights(L.NET_CRITIC)])
    e1.close()
    res = {}
    for mode in ("peer", "peer2", "nccl"):       # one-shot, two-shot (reduce-scatter + broadcast of the sum), ncclAllReduce
        os.environ["DPPO_NO_PEER_ALLREDUCE"] = "1" if mode == "nccl" else "0"
        os.environ["DPPO_PEER_TWO_SHOT"] = "1" if mode == "peer2" else "0"
        e = make_engine(o, precision=prec, device=rank)
        e.init_comm()
        assert getattr(e, "peer_allreduce", False) == (mode != "nccl")
        # loss + gradient only: global metrics / gradient on every rank although nothing is applied
        m0, g0 = e.ppo_step(*shard, lr=LR, apply=False, n_global=N, adv_mean=mean, adv_std=std, want_grads=True)
        m, g = e.ppo_step(*shard, lr=LR, apply=True, n_global=N, adv_mean=mean, adv_std=std, want_grads=True)
        torch.cuda.synchronize()
        w_first = np.concatenate([e.get_weights(L.NET_ACTOR_FT), e.get_weights(L.NET_CRITIC)])
        for _ in range(2):
            e.ppo_step(*shard, lr=LR, apply=True, n_global=N, adv_mean=mean, adv_std=std)
        torch.cuda.synchronize()
        L.check(e.lib.dppo_comm_status(e.h), "dppo_comm_status")
        w_last = np.concatenate([e.get_weights(L.NET_ACTOR_FT), e.get_weights(L.NET_CRITIC)])
        scale = np.abs(g_ref).max()
        r = {"grad_err": float(np.abs(g.cpu().numpy() - g_ref).max() / scale),
             "grad_err_noapply": float(np.abs(g0.cpu().numpy() - g_ref).max() / scale),
             "metrics_err": float(np.abs(m.cpu().numpy() - m_ref).max()), "metrics_err_noapply": float(np.abs(m0.cpu().numpy() - m_ref).max()),
             "w_first_err_lr": float(np.abs(w_first - w_ref).max() / LR),
             "w_first_frac_within_0p02lr": float((np.abs(w_first - w_ref) < 0.02 * LR).mean())}
        digests = [None] * world
        dist.all_gather_object(digests, w_last.tobytes())
        r["replicas_identical"] = all(b == digests[0] for b in digests)
        res[mode] = (r, w_last)
        dist.barrier()
        e.close()
    if rank == 0:
        out["res"] = {k: v[0] for k, v in res.items()}
        out["peer_vs_nccl_lr"] = max(float(np.abs(res[m_][1] - res["nccl"][1]).max()) for m_ in ("peer", "peer2")) / LR
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("prec", ["bf16", "bf16x3"])
def test_sharded_update_matches_the_unsharded_one(prec):
    import torch.multiprocessing as mp
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(2, 29577 + (1 if prec == "bf16x3" else 0), prec, out), nprocs=2, join=True)
    print(prec, dict(out["res"]), "peer vs nccl after 3 steps:", out["peer_vs_nccl_lr"], "lr")
    for mode, r in out["res"].items():
        assert r["replicas_identical"], mode
        # reduced gradient == un-sharded gradient up to fp32 summation order (bf16 mode: atomics; bf16x3: another split-K partition)
        assert r["grad_err"] < 2e-4 and r["grad_err_noapply"] < 2e-4, (mode, r)
        assert r["metrics_err"] < 2e-6 and r["metrics_err_noapply"] < 2e-6, (mode, r)
        # AdamW's first step moves an entry by ~lr * g / (|g| + eps): compare in units of lr (entries with |g| ~ eps amplify 1e-9 errors)
        assert r["w_first_frac_within_0p02lr"] > 0.999 and r["w_first_err_lr"] < 1.0, (mode, r)
    assert out["peer_vs_nccl_lr"] < 2.0


# ----------------------------------------------------------------------------------------------------------------------
def _agent_worker(rank, world, port, out):
    """Two ranks, four env copies each, replay the two training iterations of tests/test_gpu_agent.py: the data-parallel agent
    must reproduce the reference loop (oracle/dppo_loop.py on the un-sharded environment) within the single-rank bounds."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from diffusionpolicyoptimization_b200 import _lib as L
    from diffusionpolicyoptimization_b200.agent.finetune.train_ppo_diffusion_agent import TrainPPODiffusionAgent
    from diffusionpolicyoptimization_b200.parallel import shard_range
    from oracle import dppo_loop as OL
    from oracle import dppo_oracle as O
    from helpers import rel_err
    from toy_env import ToyVecEnv
    import test_gpu_agent as T
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    E, S = T.E, T.S
    lo, hi = shard_range(E, rank, world)
    o = O.make_oracle("hopper", seed=21)
    model = T.make_model(o, device=f"cuda:{rank}")

    def noise_local(itr, step, B):
        x_T, nz = T.noise_fn(itr, step, E)                      # the global draws; this rank's env block
        return x_T[lo:hi], nz[:, lo:hi]

    kw = dict(n_steps=S, act_steps=T.ACT_STEPS, batch_size=160, update_epochs=2, gamma=0.99, gae_lambda=0.95, target_kl=1)
    agent = TrainPPODiffusionAgent(model, ToyVecEnv(E, 11, 3, seed=3).shard(lo, hi), n_envs=E, n_train_itr=2, actor_lr=T.LR, force_train=True,
                                   reset_at_iteration=False, reward_scale_running=True, noise_fn=noise_local, shuffle_fn=T.shuffle_fn, **kw)
    assert agent.world == world and agent.n_envs == hi - lo
    got = [agent.run_iteration() for _ in range(2)]
    w = np.concatenate([model.engine.get_weights(L.NET_ACTOR_FT), model.engine.get_weights(L.NET_CRITIC)])
    digests = [None] * world
    dist.all_gather_object(digests, w.tobytes())
    chains = [None] * world
    dist.all_gather_object(chains, agent.chains_trajs.cpu().numpy())
    if rank == 0:
        venv = ToyVecEnv(E, 11, 3, seed=3)
        scaler = OL.RewardScaler(E)
        params = o.actor_ft + o.critic
        opt = dict(m=[torch.zeros_like(p) for p in params], v=[torch.zeros_like(p) for p in params], step=0)
        prev_obs, firsts0 = venv.reset_arg(), 1
        errs = {}
        for itr in range(2):
            want, prev_obs, done = OL.ppo_iteration(o, opt, venv, itr, prev_obs, lr=T.LR, reward_scaler=scaler, reward_scale_const=1.0,
                                                    noise_fn=T.noise_fn, shuffle_fn=T.shuffle_fn, firsts0=firsts0, **kw)
            firsts0 = done
            for k in ("pg_loss", "v_loss", "approx_kl", "ratio", "clipfrac", "explained_var"):
                errs[(itr, k)] = (float(got[itr][k]), float(want[k]))
            errs[(itr, "n_updates")] = (got[itr]["n_updates"], want["n_updates"])
        K = o.d.ft_denoising_steps
        got_chains = np.concatenate(chains, axis=1)             # [S, E, K+1, A] in env order (last iteration)
        out["chains_err"] = rel_err(got_chains.reshape(S * E, K + 1, -1), want["chains_k"].reshape(S * E, K + 1, -1))
        out["errs"] = errs
        out["frac_bad"] = float(np.mean(np.abs(w - O.flatten_params(params)) > 0.25 * T.LR))
        out["opt_steps"] = (agent.opt_iterations, opt["step"])
        out["replicas_identical"] = all(b == digests[0] for b in digests)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_agent_reproduces_the_reference_loop():
    import torch.multiprocessing as mp
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_agent_worker, args=(2, 29591, out), nprocs=2, join=True)
    print(dict(out))
    assert out["replicas_identical"] and out["opt_steps"] == (12, 12)
    assert out["chains_err"] < 5e-4
    for (itr, k), (g, w) in out["errs"].items():
        if k == "n_updates":
            assert g == w == 6
        else:
            assert abs(g - w) < 5e-3 * max(1.0, abs(w)) + 2e-5, (itr, k, g, w)
    assert out["frac_bad"] < 1e-2
