"""Generates tests/golden/*.npz from the oracle (python tests/golden/make_golden.py).

The reference ships no golden vectors and TensorFlow cannot be installed here, so
these fixtures pin the ORACLE: they catch regressions of the restatement and give the GPU tests a
committed, seed-independent target.  Sizes are small (oracle finishes in seconds)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dppo_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def case(task, B, N, seed):
    torch.manual_seed(0)
    o = O.make_oracle(task, seed=seed)
    obs, x_T, noise = O.make_rollout_inputs(o, B, seed=seed + 1)
    s = o.sample(obs, x_T, noise)
    s_det = o.sample(obs, x_T, noise, deterministic=True)
    s_base = o.sample(obs, x_T, noise, use_base_policy=True)
    logp = o.get_logprobs(obs, s.chains)
    batch = O.make_ppo_batch(o, N, pool=64, seed=seed + 2)
    metrics, ga, gc = o.ppo_grads(*batch)
    with torch.no_grad():
        val = O.critic_obs(o.critic, batch[0], o.h.critic_act).reshape(-1)
        logp_sub, _ = o.get_logprobs_subsample(batch[0], batch[1], batch[2], batch[3])
    rng = np.random.default_rng(seed + 3)
    x0 = torch.from_numpy(rng.uniform(-1, 1, size=(N, o.d.horizon_steps, o.d.action_dim)).astype(np.float32))
    t = torch.from_numpy(rng.integers(0, o.d.denoising_steps, size=N))
    pn = torch.from_numpy(rng.standard_normal(size=tuple(x0.shape)).astype(np.float32))
    ploss, pg = o.pretrain_grads(x0, batch[0], t, pn)
    out = dict(
        actor=O.flatten_params(o.actor), actor_ft=O.flatten_params(o.actor_ft), critic=O.flatten_params(o.critic),
        obs=obs.numpy(), x_T=x_T.numpy(), noise=noise.numpy(),
        actions=s.trajectories.numpy(), chains=s.chains.numpy(),
        actions_det=s_det.trajectories.numpy(), actions_base=s_base.trajectories.numpy(),
        logp=logp.numpy(),
        ppo_obs=batch[0].numpy(), ppo_prev=batch[1].numpy(), ppo_next=batch[2].numpy(), ppo_inds=batch[3].numpy(),
        ppo_returns=batch[4].numpy(), ppo_oldvalues=batch[5].numpy(), ppo_adv=batch[6].numpy(), ppo_oldlogp=batch[7].numpy(),
        ppo_metrics=np.array([float(m) for m in metrics], np.float32),
        ppo_grads=np.concatenate([O.flatten_params(ga), O.flatten_params(gc)]),
        value=val.numpy(), logp_sub=logp_sub.numpy(),
        pre_x0=x0.numpy(), pre_t=t.numpy().astype(np.int32), pre_noise=pn.numpy(),
        pre_loss=np.array([float(ploss)], np.float32), pre_grads=O.flatten_params(pg),
    )
    # weights/grads are big (0.55M floats each): keep a strided fingerprint of the gradients only
    out["ppo_grads_fp"] = out.pop("ppo_grads")[::97].copy()
    out["pre_grads_fp"] = out.pop("pre_grads")[::97].copy()
    for k in ("actor", "actor_ft", "critic"):
        out[k + "_fp"] = out.pop(k)[::97].copy()
    np.savez_compressed(os.path.join(HERE, f"{task}_B{B}_N{N}.npz"), **out)
    print(task, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    case("hopper", B=8, N=96, seed=0)
    case("walker2d", B=5, N=64, seed=7)
