"""Dev probe: tcgen05.mma issue cost, TMA round trip and TMA throughput per SM (see fc::mma_probe_kernel)."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from diffusionpolicyoptimization_b200 import _lib as L
e = bench.make_gpu_engine(L.PREC_BF16, 0)
def run(grid, mode, iters, N, depth):
    out = np.zeros(grid * 2, np.int64)
    L.check(e.lib.dppo_debug_mma_probe(e.h, grid, mode, iters, N, depth, out.ctypes.data_as(C.c_void_p)))
    return out[0::2].astype(float), out[1::2].astype(float)
grid = 148
for N in (64, 256):
    for flags in (0, 16, 1, 2, 4, 7, 23, 8, 24, 31):
        a, b = run(grid, 0, 2048, N, flags)
        nm = 4 if flags & 8 else 2
        print(f"issue loop N={N:3d} commit={flags & 1} poll={(flags >> 1) & 1} fence={(flags >> 2) & 1} mma/stage={nm} altD={(flags >> 4) & 1}: issue {a.mean() / 2048:6.1f} cyc/stage, with drain {b.mean() / 2048:6.1f} (ideal exec {nm * 128 * N / 256:.0f})", flush=True)
if "tma" in sys.argv:
    a, b = run(grid, 1, 256, 256, 1)
    print(f"TMA round trip (16 KB, 4 boxes, one in flight): {a.mean() / b.mean():7.1f} cyc", flush=True)
    for depth in (1, 2, 3, 4):
        a, b = run(grid, 2, 1024, 256, depth)
        print(f"TMA throughput depth={depth}: {(b * 16384 / a).mean():6.1f} B/cyc/SM  ({a.mean() / b.mean():6.1f} cyc per 16 KB stage)", flush=True)
e.close()
