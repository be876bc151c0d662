"""Pure-write and pure-read HBM rates on this GPU (context for the write-bound PLDA / TDNN epilogues)."""
import torch
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda")   # 4 GiB
for name, fn, nbytes in (("fill (pure write)", lambda: x.fill_(1.0), x.numel() * 4),
                         ("sum (pure read)", lambda: x.sum(), x.numel() * 4),
                         ("copy (read+write)", lambda: x[: 1 << 29].copy_(x[1 << 29:]), x.numel() * 4)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"{name:20s} {nbytes / best / 1e6:8.1f} GB/s")
