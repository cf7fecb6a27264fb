"""Profiling target for the fused layer-chain kernel: log-prob forward (N rows) and the T-step sampler (B rows)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from diffusionpolicyoptimization_b200 import _lib as L

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 18944
e = bench.make_gpu_engine(L.PREC_BF16, 0)
b = bench.make_gpu_batches(e, N, 1, seed=3)[0]
obs = torch.rand(B, e.Do, device="cuda") * 2 - 1
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    c.record(); torch.cuda.synchronize()
    return a.elapsed_time(c) / n
print(f"logprobs N={N}: {t(lambda: e.logprobs_subsample(b[0], b[1], b[2], b[3])):.3f} ms")
print(f"sample B={B}: {t(lambda: e.sample(obs, seed=1, offset=2), 3):.3f} ms  path={e.last_path()}")
print(f"ppo N={N}: {t(lambda: e.ppo_step(*b, lr=1e-4, apply=True, adv_mean=0.0, adv_std=1.0)):.3f} ms")
e.close()
