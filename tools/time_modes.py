"""Dev probe: PPO step / log-prob / sampler time per precision mode at the benchmarked size (walker2d shapes), with the
library's per-kernel-class device times (dppo_profile_*).   python tools/time_modes.py [modes, comma separated] [rows]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import dppo_oracle as O
from helpers import make_engine
from diffusionpolicyoptimization_b200 import _lib as L

MODES = {"fp32": L.PREC_FP32, "bf16": L.PREC_BF16, "bf16x3": L.PREC_BF16X3}


def timeit(fn, iters=10, warm=3, blocks=7):
    """median over `blocks` timed blocks of `iters` calls (the min is printed too: clocks move under the power cap)"""
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(blocks):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters): fn()
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / iters)
    ts.sort()
    timeit.last_min = ts[0]
    return ts[len(ts) // 2]


modes = (sys.argv[1] if len(sys.argv) > 1 else "bf16x3").split(",")
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
o = O.make_oracle("walker2d", seed=0)
b = O.make_ppo_batch(o, N, pool=2048, seed=1)
args = [x.cuda() for x in b]
for i in (0, 1, 2, 7): args[i] = args[i].reshape(N, -1).contiguous()
for mode in modes:
    e = make_engine(o, precision=MODES[mode])
    ms = timeit(lambda: e.ppo_step(*args, lr=1e-4, apply=True))
    print(f"[{mode}] ppo N={N}: {ms:.3f} ms  {N/ms*1e3/1e6:.2f} M samples/s  ({4.141568e6*N/ms/1e9:.1f} algorithmic TFLOP/s)  [min {timeit.last_min:.3f} ms]", flush=True)
    e.profile_enable(True)
    l0 = e.launch_count()
    for _ in range(5): e.ppo_step(*args, lr=1e-4, apply=True)
    torch.cuda.synchronize()
    print(f"   launches/step {(e.launch_count()-l0)/5:.0f}")
    for cls, name in ((0, "chain512"), (1, "tcgen05 gemm"), (2, "sgemm"), (3, "chain256")):
        t, n, fl = e.profile_read_class(cls)
        if n: print(f"   class {name}: {t/5:.3f} ms/step over {n//5} launches, {fl/t/1e9:.1f} algorithmic TFLOP/s", flush=True)
    e.profile_enable(False)
    ms = timeit(lambda: e.logprobs_subsample(args[0], args[1], args[2], args[3]))
    print(f"[{mode}] logprobs N={N}: {ms:.3f} ms  {N/ms*1e3/1e6:.2f} M rows/s", flush=True)
    B = 148 * 128
    obs = torch.rand(B, o.d.Do, device="cuda") * 2 - 1
    ms = timeit(lambda: e.sample(obs, seed=1, offset=2), iters=5, warm=2)
    print(f"[{mode}] sample B={B}: {ms:.3f} ms  path={e.last_path()}  {B/ms*1e3/1e6:.2f} M chunks/s", flush=True)
    e.close()
