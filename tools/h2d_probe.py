"""Dev probe: pinned host -> device copy bandwidth at the bench's minibatch size, 1 vs 2 vs 4 concurrent streams."""
import torch
def run(total_mb, nstreams, reps=10):
    n = int(total_mb * 1e6 / 4 / nstreams)
    hs = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(nstreams)]
    ds = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(nstreams)]
    ss = [torch.cuda.Stream() for _ in range(nstreams)]
    def go():
        for h, d, s in zip(hs, ds, ss):
            with torch.cuda.stream(s):
                d.copy_(h, non_blocking=True)
    for _ in range(3): go()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in ss: s.wait_event(a)
    for _ in range(reps): go()
    for s in ss: b.wait(s) if False else torch.cuda.current_stream().wait_stream(s)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"H2D {total_mb:5.1f} MB over {nstreams} stream(s): {ms * 1e3:7.1f} us  {total_mb / ms:6.1f} GB/s", flush=True)
for mb in (7.0, 18.6, 64.0):
    for ns in (1, 2, 4):
        run(mb, ns)
