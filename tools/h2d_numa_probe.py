"""Dev probe: does pinning the process to the GPU's NUMA-local CPUs change pinned H2D bandwidth?"""
import os, subprocess, sys, torch
def bw(mb=18.6, reps=20):
    n = int(mb * 1e6 / 4)
    h = torch.empty(n, dtype=torch.float32).pin_memory(); d = torch.empty(n, dtype=torch.float32, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): d.copy_(h, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    return mb / (a.elapsed_time(b) / reps)
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
print("cpus allowed:", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8], "...")
torch.cuda.init()
print(f"default affinity: {bw():.1f} GB/s, {bw(4.8):.1f} GB/s (4.8 MB)")
try:
    out = subprocess.run(["nvidia-smi", "topo", "-C", "-i", "0"], capture_output=True, text=True).stdout
    print("topo -C:", out.strip())
except Exception as ex:
    print(ex)
for lo in (0, 8, 16, 32, 48, 64, 96):
    try:
        os.sched_setaffinity(0, set(range(lo, lo + 8)))
        print(f"cpus {lo}-{lo+7}: {bw():.1f} GB/s", flush=True)
    except Exception as ex:
        print(lo, ex)
