"""Dev tool (GPU box): per-tensor error of every precision mode against the oracle at BASELINE.json's full sizes
(walker2d / halfcheetah shapes: 50 000-row PPO minibatch, 18 944-row rollout, 4096-row pre-train step).  The numbers it
prints are what the bounds in tests/test_gpu_fullsize_oracle.py are derived from.

    python tools/measure_parity.py [--modes fp32,bf16,bf16x3] [--out gpurun_out/parity_full.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

from diffusionpolicyoptimization_b200 import _lib as L   # noqa: E402
from oracle import dppo_oracle as O                      # noqa: E402
from helpers import make_engine                          # noqa: E402

MODES = {"fp32": L.PREC_FP32, "bf16": L.PREC_BF16}
if hasattr(L, "PREC_BF16X3"):
    MODES["bf16x3"] = L.PREC_BF16X3


def per_var_err(got_flat, want_list):
    """max |got - want| / max |want| per Keras variable, and over the whole net."""
    out, off = [], 0
    for w in want_list:
        n = w.numel()
        g = got_flat[off:off + n]; ww = w.reshape(-1).numpy()
        out.append(float(np.abs(g - ww).max() / max(np.abs(ww).max(), 1e-30)))
        off += n
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default=",".join(MODES))
    ap.add_argument("--task", default="walker2d")
    ap.add_argument("--rows", type=int, default=50_000)
    ap.add_argument("--sample-rows", type=int, default=148 * 128)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_full.json"))
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    o = O.make_oracle(a.task, seed=0)
    d = o.d
    N = a.rows
    res = {"task": a.task, "rows": N, "modes": {}}

    t0 = time.time()
    batch = O.make_ppo_batch(o, N, pool=4096, seed=5)
    with torch.no_grad():
        want_lp = o.get_logprobs_subsample(batch[0], batch[1], batch[2], batch[3])[0].reshape(N, -1)
        want_v = O.critic_obs(o.critic, batch[0], o.h.critic_act).reshape(-1)
    metrics, ga, gc = o.ppo_grads(*batch)
    res["oracle_ppo_seconds"] = time.time() - t0
    wg = np.concatenate([O.flatten_params(ga), O.flatten_params(gc)])
    # one AdamW step on the oracle's side (Keras-3 rule)
    ft = [p.clone() for p in o.actor_ft]; cr = [p.clone() for p in o.critic]
    m = [torch.zeros_like(p) for p in ft + cr]; v = [torch.zeros_like(p) for p in ft + cr]
    lr = 1e-4
    O.adamw_keras(ft + cr, ga + gc, m, v, 1, lr, o.h.beta1, o.h.beta2, o.h.adam_eps, o.h.weight_decay)
    want_w = np.concatenate([O.flatten_params(ft), O.flatten_params(cr)])
    w0 = np.concatenate([O.flatten_params(o.actor_ft), O.flatten_params(o.critic)])

    # sampler inputs
    Bs = a.sample_rows
    obs_s, xT_s, nz_s = O.make_rollout_inputs(o, Bs, seed=9)
    t1 = time.time()
    want_s = o.sample(obs_s, xT_s, nz_s)
    res["oracle_sample_seconds"] = time.time() - t1
    with torch.no_grad():
        want_slp = o.get_logprobs(obs_s[:4096], want_s.chains[:4096]).reshape(4096 * d.ft_denoising_steps, -1)

    # pre-train inputs
    Np = 4096
    rng = np.random.default_rng(11)
    acts = torch.from_numpy(rng.uniform(-1, 1, (Np, d.horizon_steps, d.action_dim)).astype(np.float32))
    st = torch.from_numpy(rng.uniform(-1, 1, (Np, 1, d.obs_dim)).astype(np.float32))
    tt = torch.from_numpy(rng.integers(0, d.denoising_steps, Np))
    nzp = torch.from_numpy(rng.standard_normal((Np, d.horizon_steps, d.action_dim)).astype(np.float32))
    want_pl, want_pg = o.pretrain_grads(acts, st, tt, nzp)
    wpg = O.flatten_params(want_pg)

    for name in a.modes.split(","):
        if name not in MODES:
            continue
        e = make_engine(o, precision=MODES[name])
        r = {}
        fb = [batch[0].reshape(N, -1), batch[1].reshape(N, -1), batch[2].reshape(N, -1), batch[3], batch[4], batch[5], batch[6],
              batch[7].reshape(N, -1)]
        lp = e.logprobs_subsample(fb[0], fb[1], fb[2], fb[3]).cpu().numpy()
        r["logp_abs_max"] = float(np.abs(lp - want_lp.numpy()).max())
        r["logp_abs_mean"] = float(np.abs(lp - want_lp.numpy()).mean())
        val = e.value(fb[0]).cpu().numpy()
        r["value_rel"] = float(np.abs(val - want_v.numpy()).max() / np.abs(want_v.numpy()).max())
        mt, g = e.ppo_step(*fb, lr=lr, apply=True, want_grads=True)
        torch.cuda.synchronize()
        g = g.cpu().numpy(); mt = mt.cpu().numpy()
        r["tc_launches"] = e.tc_launch_count()
        r["grad_rel_max"] = float(np.abs(g - wg).max() / np.abs(wg).max())
        nA = e.n_actor
        r["grad_rel_actor"] = float(np.abs(g[:nA] - wg[:nA]).max() / np.abs(wg[:nA]).max())
        r["grad_rel_critic"] = float(np.abs(g[nA:] - wg[nA:]).max() / np.abs(wg[nA:]).max())
        r["grad_rel_per_var"] = per_var_err(g, ga + gc)
        r["metrics"] = [float(x) for x in mt]
        r["metrics_oracle"] = [float(x) for x in metrics]
        r["metrics_abs_err"] = [abs(float(x) - float(y)) for x, y in zip(mt, metrics)]
        w1 = np.concatenate([e.get_weights(L.NET_ACTOR_FT), e.get_weights(L.NET_CRITIC)])
        # AdamW's first step moves every entry by ~lr * sign(g): compare in units of lr
        r["weights_err_in_lr"] = float(np.abs(w1 - want_w).max() / lr)
        r["weights_frac_within_0p1lr"] = float((np.abs(w1 - want_w) < 0.1 * lr).mean())
        r["weights_moved_in_lr"] = float(np.abs(want_w - w0).max() / lr)
        e.close()

        # ratio == 1 at unchanged weights: old log-probs from dppo_logprobs, new ones inside the update (apply = 0)
        e = make_engine(o, precision=MODES[name])
        P = 4096
        obs_p = fb[0][:P].contiguous().cuda()
        _, ch = e.sample(obs_p, seed=3, offset=1)
        olp = e.logprobs(obs_p, ch).reshape(P, e.K, e.A)
        vals = e.value(obs_p)
        gen = torch.Generator(device="cuda"); gen.manual_seed(0)
        flat = torch.randint(0, P * e.K, (N,), device="cuda", generator=gen)
        b, k = flat // e.K, flat % e.K
        mt = e.ppo_step(obs_p[b].contiguous(), ch[b, k].contiguous(), ch[b, k + 1].contiguous(), k.to(torch.int32), vals[b].contiguous(),
                        vals[b].contiguous(), torch.randn(N, device="cuda", generator=gen), olp[b, k].contiguous(), lr=0.0, apply=False)
        mt = mt.cpu().numpy()
        r["unchanged_weights"] = {"clipfrac": float(mt[3]), "approx_kl": float(mt[4]), "ratio": float(mt[5]), "v_loss": float(mt[2])}

        # sampler
        act, chn = e.sample(obs_s.reshape(Bs, -1), x_T=xT_s.reshape(Bs, -1), noise=nz_s.reshape(d.denoising_steps, Bs, -1))
        torch.cuda.synchronize()
        r["sample_path"] = e.last_path()
        wa = want_s.trajectories.reshape(Bs, -1).numpy(); ga_ = act.cpu().numpy()
        r["sample_actions_rel"] = float(np.abs(ga_ - wa).max() / np.abs(wa).max())
        r["sample_actions_abs_mean"] = float(np.abs(ga_ - wa).mean())
        r["sample_actions_abs_p999"] = float(np.quantile(np.abs(ga_ - wa), 0.999))
        r["sample_chain_abs_mean"] = float(np.abs(chn.cpu().numpy() - want_s.chains.reshape(Bs, d.ft_denoising_steps + 1, -1).numpy()).mean())
        slp = e.logprobs(obs_s[:4096].reshape(4096, -1), want_s.chains[:4096].reshape(4096, d.ft_denoising_steps + 1, -1)).cpu().numpy()
        r["chain_logp_abs_max"] = float(np.abs(slp - want_slp.numpy()).max())

        # pre-train step
        loss, pg = e.pretrain_step(acts.reshape(Np, -1), st.reshape(Np, -1), lr=1e-3, apply=False, t=tt, noise=nzp.reshape(Np, -1), want_grads=True)
        pg = pg.cpu().numpy()
        r["pretrain_loss_rel"] = abs(float(loss) - float(want_pl)) / abs(float(want_pl))
        r["pretrain_grad_rel_max"] = float(np.abs(pg - wpg).max() / np.abs(wpg).max())
        e.close()
        res["modes"][name] = r
        print(name, json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
