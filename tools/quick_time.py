"""Dev probe: time the main entry points with CUDA events (not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import dppo_oracle as O
from helpers import make_engine
from diffusionpolicyoptimization_b200 import _lib as L

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

prec = L.PREC_BF16 if "bf16" in sys.argv else L.PREC_FP32
for task in ("hopper", "walker2d"):
    o = O.make_oracle(task, seed=0)
    e = make_engine(o, precision=prec)
    for B in (40, 256, 1024, 4096, 16384):
        obs = torch.rand(B, o.d.Do, device="cuda") * 2 - 1
        ms = timeit(lambda: e.sample(obs, seed=1, offset=2), iters=10 if B > 1000 else 50)
        print(f"{task} sample B={B}: {ms*1e3:.1f} us  path={e.last_path()}  {B/ms*1e3:.0f} chunks/s", flush=True)
    for N in (4096, 50000):
        b = O.make_ppo_batch(o, N, pool=512, seed=1)
        args = [x.cuda() for x in b]
        args[0] = args[0].reshape(N, -1); args[1] = args[1].reshape(N, -1); args[2] = args[2].reshape(N, -1); args[7] = args[7].reshape(N, -1)
        ms = timeit(lambda: e.ppo_step(*args, lr=1e-4, apply=True), iters=5)
        print(f"{task} ppo N={N}: {ms:.3f} ms  {N/ms*1e3:.0f} samples/s", flush=True)
        ms = timeit(lambda: e.logprobs_subsample(args[0], args[1], args[2], args[3]), iters=5)
        print(f"{task} logprobs N={N}: {ms:.3f} ms  {N/ms*1e3:.0f} rows/s", flush=True)
    e.close()
