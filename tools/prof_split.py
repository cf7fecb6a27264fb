"""Profiling target: a few PPO updates in the split-precision mode at bench shapes (run plain first, then under ncu)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from diffusionpolicyoptimization_b200 import _lib as L

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
e = bench.make_gpu_engine({"bf16x3": L.PREC_BF16X3, "bf16": L.PREC_BF16, "fp32": L.PREC_FP32}[mode], 0)
b = bench.make_gpu_batches(e, N, 1, seed=3)[0]
for i in range(2):
    e.ppo_step(*b, lr=1e-4, apply=True)
torch.cuda.synchronize()
a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(steps):
    e.ppo_step(*b, lr=1e-4, apply=True)
c.record(); torch.cuda.synchronize()
print(f"[{mode}] ppo_step N={N}: {a.elapsed_time(c) / steps:.3f} ms")
e.close()
