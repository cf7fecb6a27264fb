import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import dppo_oracle as O
from helpers import make_engine
from diffusionpolicyoptimization_b200 import _lib as L
for task in ("hopper", "walker2d"):
    o = O.make_oracle(task, seed=0)
    e = make_engine(o, precision=L.PREC_BF16X3)
    for B in (40, 16, 8, 64):
        obs = torch.rand(B, o.d.Do, device="cuda") * 2 - 1
        for _ in range(5): e.sample(obs, seed=1, offset=2)
        torch.cuda.synchronize()
        ts = []
        for blk in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20): e.sample(obs, seed=1, offset=2)
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / 20 * 1e3)
        ts.sort()
        print(f"{task} B={B}: median {ts[3]:.1f} us  min {ts[0]:.1f} us  path={e.last_path()}", flush=True)
    e.close()
