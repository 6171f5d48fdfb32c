#!/usr/bin/env python3
"""What HBM takes for a pure write stream, a copy and a read (torch kernels, 1 GiB buffers): the write-heavy GEMM epilogues
(qkv, fc1: 3-4 bytes written per byte read) are bounded by the first number, not by the copy bandwidth MEASURED_PEAKS quotes."""
import torch

x = torch.empty(2 ** 30, dtype=torch.uint8, device="cuda")
y = torch.empty(2 ** 30, dtype=torch.uint8, device="cuda")
for name, fn, nbytes in (("memset (write only)", lambda: x.zero_(), 2 ** 30), ("copy (read + write)", lambda: y.copy_(x), 2 ** 31),
                         ("int32 sum (read only, torch reduction)", lambda: x.view(torch.int32).sum(), 2 ** 30)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:42s} {nbytes * 10 / e0.elapsed_time(e1) / 1e6:8.0f} GB/s")
