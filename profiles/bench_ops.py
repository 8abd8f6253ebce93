"""Micro-timings of the projection kernel on the rollout's shapes (CUDA events, warm, L2-flushed between launches).
  python profiles/bench_ops.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tbpkg  # noqa: E402,F401
from trafficbotsv1_5_b200 import ops  # noqa: E402

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def time_op(fn, n=10):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


shapes = [(65536, 896, 128, "in-proj self"), (65536, 640, 128, "in-proj cross q|u"), (65536, 128, 640, "out-proj"),
          (65536, 512, 128, "ffn1"), (65536, 128, 512, "ffn2"), (65536, 128, 128, "head 128"),
          (720896, 64, 128, "pointnet"), (720896, 64, 20, "ag input mlp0"), (720896, 64, 64, "ag input mlp"),
          (16384, 256, 128, "kv table")]
for M, N, K, name in shapes:
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    y = torch.empty(M, N, device=dev)
    gb = (M * K + M * N + N * K) * 4 / 1e9
    for prec in (0, 1):
        t = time_op(lambda: ops.linear(x, w, b, out=y, precision=prec))
        print(f"{name:20s} M={M:7d} N={N:4d} K={K:4d} prec={prec}: {t:8.1f} us  {gb / t * 1e6 / 1e3:6.2f} TB/s  "
              f"{2 * M * N * K / t / 1e6:7.1f} TFLOP/s")
