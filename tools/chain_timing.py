"""Dev tool: per-role cycle counters of the fused chain kernel (dppo_debug_chain_timing)."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from diffusionpolicyoptimization_b200 import _lib as L

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 18944
e = bench.make_gpu_engine(L.PREC_BF16, 0)
b = bench.make_gpu_batches(e, N, 1, seed=3)[0]
obs = torch.rand(B, e.Do, device="cuda") * 2 - 1
nsm = C.c_int(0)
L.check(e.lib.dppo_debug_chain_timing(e.h, 1, None, C.byref(nsm)))
buf = np.zeros((nsm.value, 8), np.int64)
names = ["prod_wait_wempty", "mma_wait_xfull", "mma_wait_wfull", "mma_total", "epi_wait_acc", "epi_generic", "epi_final", "-"]
def report(tag):
    L.check(e.lib.dppo_debug_chain_timing(e.h, 1, buf.ctypes.data_as(C.c_void_p), None))
    act = buf[buf[:, 3] > 0]
    print(tag, f"({len(act)} CTAs)", {n: int(act[:, i].mean()) for i, n in enumerate(names[:7])}, "max total", int(act[:, 3].max()))
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    c.record(); torch.cuda.synchronize()
    return a.elapsed_time(c) / n
print(f"logprobs N={N}: {timed(lambda: e.logprobs_subsample(b[0], b[1], b[2], b[3])):.3f} ms"); report("logprobs")
print(f"sample B={B}: {timed(lambda: e.sample(obs, seed=1, offset=2)):.3f} ms"); report("sample")
print(f"value N={N}: {timed(lambda: e.value(b[0])):.3f} ms"); report("critic forward (infer)")
ms = timed(lambda: e.ppo_step(*b, lr=1e-4, apply=False, adv_mean=0.0, adv_std=1.0))
print(f"ppo N={N}: {ms:.3f} ms"); report("ppo(last chain = actor bwd)")
e.close()
