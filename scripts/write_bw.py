#!/usr/bin/env python
"""Write-only HBM stream ceiling on this GPU: cudaMemsetAsync (torch zero_) and a plain copy over buffers far larger
than L2.  Context for the sparse-support (store-bound) regime of the mode sum, whose algorithmic traffic is writes only."""
import json
import torch

x = torch.empty(int(3.2e9) // 8, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
res = {}
for name, fn, nbytes in (("memset_GBps", lambda: x.zero_(), x.numel() * 8), ("copy_GBps", lambda: y.copy_(x), 2 * x.numel() * 8)):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        fn()
    b.record()
    torch.cuda.synchronize()
    res[name] = nbytes * 10 / (a.elapsed_time(b) * 1e-3) / 1e9
print(json.dumps(res))
