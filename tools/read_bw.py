"""Read-only HBM bandwidth yard-sticks (torch kernels): what a streaming READ kernel can reach on this GPU."""
import torch
n = 1 << 30  # 4 GiB of fp32
x = torch.ones(n, device="cuda")
y = torch.empty_like(x)
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: x.sum()); print(f"sum (read only)      {4*n/ms/1e6:7.0f} GB/s")
ms = t(lambda: x.max()); print(f"max (read only)      {4*n/ms/1e6:7.0f} GB/s")
ms = t(lambda: y.copy_(x)); print(f"copy (read + write)  {8*n/ms/1e6:7.0f} GB/s")
ms = t(lambda: y.fill_(1.0)); print(f"fill (write only)    {4*n/ms/1e6:7.0f} GB/s")
x2 = x.view(-1, 250)[: (n // 250)]
ms = t(lambda: x2.argmin(dim=1)); print(f"argmin rows of 250   {4*x2.numel()/ms/1e6:7.0f} GB/s")
