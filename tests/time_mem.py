"""Developer tool: raw HBM write / copy rates for the sizes of the layer1 tensors (what the conv epilogues compete with)."""
import torch

def t(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for mb in (68, 273, 1024):
    n = mb * 1024 * 1024 // 2
    a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    b = torch.randn(n // 2, device="cuda").view(torch.bfloat16) if False else torch.ones(n, dtype=torch.bfloat16, device="cuda")
    us = t(lambda: a.fill_(1.0))
    print(f"fill  {mb} MB: {us:.1f} us  {mb * 1.048576 / us * 1e3:.0f} GB/s (write only)")
    us = t(lambda: a.copy_(b))
    print(f"copy  {mb} MB: {us:.1f} us  {2 * mb * 1.048576 / us * 1e3:.0f} GB/s (read+write)")
    us = t(lambda: torch.relu_(a))
    print(f"relu_ {mb} MB: {us:.1f} us  {2 * mb * 1.048576 / us * 1e3:.0f} GB/s (read+write in place)")
    c = torch.ones(n, dtype=torch.bfloat16, device="cuda")
    us = t(lambda: torch.add(b, c, out=a))
    print(f"add   {mb} MB: {us:.1f} us  {3 * mb * 1.048576 / us * 1e3:.0f} GB/s (2 reads + write)")
