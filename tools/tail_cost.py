"""Dev tool: what the serial tail of the PPO update costs (apply=True vs apply=False: AdamW + table/operand refresh)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from diffusionpolicyoptimization_b200 import _lib as L
e = bench.make_gpu_engine(L.PREC_BF16, 0)
b = bench.make_gpu_batches(e, 50000, 1, seed=3)[0]
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    c.record(); torch.cuda.synchronize()
    return a.elapsed_time(c) / n * 1e3
print("apply=True  us/step", round(t(lambda: e.ppo_step(*b, lr=1e-5, apply=True, adv_mean=0.0, adv_std=1.0)), 1))
print("apply=False us/step", round(t(lambda: e.ppo_step(*b, lr=0.0, apply=False, adv_mean=0.0, adv_std=1.0)), 1))
print("logprobs    us/call", round(t(lambda: e.logprobs_subsample(b[0], b[1], b[2], b[3])), 1))
print("value       us/call", round(t(lambda: e.value(b[0])), 1))
e.close()
