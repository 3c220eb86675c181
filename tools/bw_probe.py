#!/usr/bin/env python
"""HBM bandwidth context for K3's roofline (torch kernels, 8 GB buffers, CUDA events): fill, copy,
read-only reduction and a 1:4 read:write expand-copy.  Measured on B200 (round 1): fill 3.9 TB/s,
copy 6.67 TB/s, read 6.87 TB/s, 1r:4w 4.3 TB/s -- write-heavy streams do not exceed the copy figure,
so K3 (81 % writes) at 6.55-6.6 TB/s sits at the practical ceiling."""
import torch, time
torch.cuda.set_device(0)
n = 8 * 1024**3
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
ms = t(lambda: a.zero_()); print("memset   GB/s", n / ms / 1e6)
ms = t(lambda: b.copy_(a)); print("copy r+w GB/s", 2 * n / ms / 1e6)
ms = t(lambda: a.sum(dtype=torch.int64) if False else torch.sum(a.view(torch.int64)[: n // 8])); print("read     GB/s", n / ms / 1e6)
# 4:1 write:read mix like K3 (25+2 MB written, 6.2 MB read per frame): out = 4x expand
src = a[: n // 4].view(torch.int32)
dst = b.view(torch.int32).view(4, -1)
ms = t(lambda: dst.copy_(src.unsqueeze(0).expand(4, -1))); print("1r:4w    GB/s", (n // 4 + n) / ms / 1e6)
