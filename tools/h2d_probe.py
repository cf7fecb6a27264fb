"""Dev probe: pinned host -> device copy bandwidth at the bench's minibatch size."""
import torch, time
for mb in (1.2, 4.8, 18.6, 64.0):
    n = int(mb * 1e6 / 4)
    h = torch.empty(n, dtype=torch.float32).pin_memory(); d = torch.empty(n, dtype=torch.float32, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"H2D {mb:5.1f} MB: {ms * 1e3:7.1f} us  {mb / ms:6.1f} GB/s")
