"""Dev probe: is the PPO actor-gradient error of a mode explained by rows that sit on a discontinuity of the loss (ratio on the clip
boundary 1 +- c, log-prob on the [-5, 2] clip, x0 on the +-1 clip)?  Drops the rows the oracle finds within `margin` of the ratio
boundary and compares again."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import dppo_oracle as O
from helpers import make_engine
from diffusionpolicyoptimization_b200 import _lib as L

MODES = {"fp32": L.PREC_FP32, "bf16": L.PREC_BF16, "bf16x3": L.PREC_BF16X3}
task, N = (sys.argv[1] if len(sys.argv) > 1 else "hopper"), int(sys.argv[2]) if len(sys.argv) > 2 else 4099
modes = (sys.argv[3] if len(sys.argv) > 3 else "fp32,bf16x3").split(",")
noclip = len(sys.argv) > 4 and sys.argv[4] == "noclip"
o = O.make_oracle(task, seed=0, hyper=O.Hyper(denoised_clip_value=None) if noclip else None)
seed = int(sys.argv[5]) if len(sys.argv) > 5 else 3
batch = O.make_ppo_batch(o, N, pool=512 if N < 10000 else 4096, seed=seed)


def ratio_of(b):
    with torch.no_grad():
        lp = o.get_logprobs_subsample(b[0], b[1], b[2], b[3])[0]
        new = lp.clamp(-5, 2).mean(dim=(-1, -2)); old = b[7].clamp(-5, 2).mean(dim=(-1, -2))
        return torch.exp(new - old)


def run(b, tag):
    n = b[0].shape[0]
    metrics, ga, gc = o.ppo_grads(*b)
    wg = np.concatenate([O.flatten_params(ga), O.flatten_params(gc)])
    for mode in modes:
        e = make_engine(o, precision=MODES[mode])
        m, g = e.ppo_step(b[0].reshape(n, -1), b[1].reshape(n, -1), b[2].reshape(n, -1), b[3], b[4], b[5], b[6], b[7].reshape(n, -1), lr=0.0,
                          apply=False, want_grads=True)
        g = g.cpu().numpy(); nA = e.n_actor
        pv, off = [], 0
        for w in ga + gc:
            k = w.numel(); pv.append(float(np.abs(g[off:off + k] - w.reshape(-1).numpy()).max() / np.abs(wg[:nA] if off < nA else wg[nA:]).max())); off += k
        print(f"[{tag}] {mode}: actor {np.abs(g[:nA]-wg[:nA]).max()/np.abs(wg[:nA]).max():.2e} critic {np.abs(g[nA:]-wg[nA:]).max()/np.abs(wg[nA:]).max():.2e} "
              f"per-var(of net max) {['%.1e' % x for x in pv[:12]]}  clipfrac {float(m[3]):.5f} vs {float(metrics[3]):.5f}", flush=True)
        e.close()


with torch.no_grad():
    K = o.d.ft_denoising_steps
    t_all = torch.arange(K - 1, -1, -1)[batch[3].long()]
    eps = O.diffusion_mlp(o.actor_ft, batch[1], t_all, batch[0], o.d, o.h.actor_act)
    x0 = o._extract("sqrt_recip_alphas_cumprod", t_all) * batch[1] - o._extract("sqrt_recipm1_alphas_cumprod", t_all) * eps
    d0 = (x0.abs() - 1.0).abs()
    print(f"noclip={noclip}: elements with |x0| within 1e-4 / 3e-5 / 1e-5 of the +-1 clip: {int((d0 < 1e-4).sum())} / {int((d0 < 3e-5).sum())} / {int((d0 < 1e-5).sum())} of {d0.numel()}; clipped fraction {float((x0.abs() >= 1).float().mean()):.3f}")
r = ratio_of(batch)
c = o.h.clip_ploss_coef
dist = torch.minimum((r - (1 - c)).abs(), (r - (1 + c)).abs())
print(f"{task} N={N}: ratio range [{float(r.min()):.4f}, {float(r.max()):.4f}], rows within 1e-4 / 1e-5 of the clip boundary: {int((dist < 1e-4).sum())} / {int((dist < 1e-5).sum())}")
run(batch, "all rows")
if os.environ.get("FLIP_DROP") == "1":
    keep = dist >= 1e-4
    fb = tuple(t[keep] for t in batch)
    run(fb, f"{int(keep.sum())} rows, boundary rows dropped")
