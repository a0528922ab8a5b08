"""HBM write/copy ceiling probe (development aid): cudaMemset-style fill and copy of 8 GB."""
import torch
n = 8 * 1024 ** 3
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn, nbytes in (("fill(write-only)", lambda: a.fill_(1), n), ("zero_", lambda: a.zero_(), n),
                         ("copy(read+write)", lambda: b.copy_(a), 2 * n)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%s: %.3f ms  %.0f GB/s" % (name, ms, nbytes / ms / 1e6))
