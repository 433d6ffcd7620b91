"""HBM bandwidth by access mix (torch elementwise kernels): pure write, pure read (reduction), copy."""
import torch
n = 1 << 30
x = torch.empty(n, device="cuda", dtype=torch.uint8); y = torch.empty_like(x)
xf = x.view(torch.float32)
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
ms = t(lambda: x.zero_());            print(f"write  1 GiB: {ms:.3f} ms  {n / ms / 1e9:.2f} TB/s")
ms = t(lambda: xf.sum());             print(f"read   1 GiB: {ms:.3f} ms  {n / ms / 1e9:.2f} TB/s")
ms = t(lambda: y.copy_(x));           print(f"copy   1 GiB: {ms:.3f} ms  {2 * n / ms / 1e9:.2f} TB/s (read + write)")
