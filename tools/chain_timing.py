"""Dev tool: per-role cycle counters of the fused chain kernel launches (dppo_debug_chain_timing)."""
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
buf = np.zeros((16, nsm.value, 8), np.int64)
names = ["prod_wait_wempty", "mma_wait_xready", "mma_wait_wfull", "mma_total", "epi_wait_acc", "epi_generic", "epi_final"]
def read():
    L.check(e.lib.dppo_debug_chain_timing(e.h, 1, buf.ctypes.data_as(C.c_void_p), None))
def report(tag, slots):
    for sl in slots:
        a = buf[sl]
        lead = a[a[:, 3] > 0]      # leader CTAs carry the MMA counters
        epi = a[a[:, 5] + a[:, 6] > 0]
        if len(epi) == 0: continue
        d = {n: int(lead[:, i].mean()) if i in (1, 2, 3) and len(lead) else int(epi[:, i].mean()) for i, n in enumerate(names)}
        print(f"{tag} slot {sl} ({len(epi)} CTAs, {len(lead)} leaders)", d, flush=True)
def once(fn):
    fn(); torch.cuda.synchronize(); read()       # warm + clear
    fn(); torch.cuda.synchronize(); read()
once(lambda: e.logprobs_subsample(b[0], b[1], b[2], b[3])); report("logprobs", [0])
once(lambda: e.sample(obs, seed=1, offset=2)); report("sample", [0])
once(lambda: e.value(b[0])); report("critic fwd infer", [0])
once(lambda: e.ppo_step(*b, lr=1e-4, apply=False, adv_mean=0.0, adv_std=1.0)); report("ppo [actor fwd, critic fwd, actor bwd, critic bwd]", [0, 1, 2, 3])
e.close()
