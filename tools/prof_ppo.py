"""Profiling target: a few PPO updates + old-log-prob passes at bench shapes (walker2d, 50 000 rows).
Usage: python tools/prof_ppo.py [bf16|fp32] [N]   (run plain first, then under ncu)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from diffusionpolicyoptimization_b200 import _lib as L

prec = L.PREC_FP32 if "fp32" in sys.argv else L.PREC_BF16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
e = bench.make_gpu_engine(prec, 0)
b = bench.make_gpu_batches(e, N, 1, seed=3)[0]
for i in range(3):
    e.ppo_step(*b, lr=1e-4, apply=True)
torch.cuda.synchronize()
a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(3):
    e.ppo_step(*b, lr=1e-4, apply=True)
c.record(); torch.cuda.synchronize()
print(f"ppo_step N={N}: {a.elapsed_time(c) / 3:.3f} ms")
a.record()
for i in range(3):
    e.logprobs_subsample(b[0], b[1], b[2], b[3])
c.record(); torch.cuda.synchronize()
print(f"logprobs N={N}: {a.elapsed_time(c) / 3:.3f} ms")
e.close()
