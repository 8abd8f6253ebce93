"""What a write-dominated kernel can reach on this B200: write-only (fill), read-only (sum) and copy bandwidth at the
sizes the projections move (CUDA events, L2 flushed between launches). Not a bench.py number.
  gpurun -- python profiles/hbm_rw_probe.py"""
import torch

dev = "cuda"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def t(fn, n=7):
    fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


for mb in (67, 117, 256, 1024):
    n = mb * 1000 * 1000
    a = torch.empty(n, dtype=torch.uint8, device=dev)
    b = torch.empty(n, dtype=torch.uint8, device=dev)
    tf, tc = t(lambda: a.zero_()), t(lambda: b.copy_(a))
    ah = a[: n // 2 * 2].view(torch.float16)
    tr = t(lambda: torch.sum(ah, dtype=torch.float32))
    print(f"{mb:5d} MB: fill {tf:6.1f} us = {n / tf / 1e6:.2f} TB/s (write only); copy {tc:6.1f} us = {2 * n / tc / 1e6:.2f} "
          f"TB/s (read + write); sum {tr:6.1f} us = {n / tr / 1e6:.2f} TB/s (read only)")
