"""Micro-benchmarks of the training path's dominant kernels at the config-4 shapes (CUDA events, L2 flushed between
launches): tcgen05 weight gradient, FFMA weight gradient, attention backward. Not a bench.py number.
  gpurun -- python profiles/train_micro.py > gpurun_out/train_micro.txt
  gpurun -- ncu --set full --clock-control none -k regex:wgrad_tc -c 2 -o gpurun_out/wgrad_tc python profiles/train_micro.py wgrad"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tbpkg  # noqa: E402,F401
from trafficbotsv1_5_b200 import autograd as AG  # noqa: E402
from trafficbotsv1_5_b200 import ops  # noqa: E402

dev = "cuda"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=5):
    fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


what = sys.argv[1] if len(sys.argv) > 1 else "all"
M = 64 * 90 * 128  # agent tokens of one training step at config 4
if what in ("all", "wgrad"):
    for N, K, name in ((1152, 128, "self in-proj"), (128, 640, "out-proj"), (512, 128, "FFN-1"), (128, 512, "FFN-2")):
        dy = torch.randn(M, N, device=dev); x = torch.randn(M, K, device=dev)
        for prec in (1, 0) if what == "all" else (1,):
            ms = timed(lambda: AG.wgrad(dy, x, True, precision=prec))
            fl = 2.0 * M * N * K
            by = 4.0 * (M * N + M * K + N * K)
            print(f"wgrad {name:12s} M={M} N={N} K={K} precision={prec}: {ms:7.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s  "
                  f"{by / ms / 1e6:7.0f} GB/s (algorithmic: both activations read once)")
        del dy, x
if what in ("all", "attn"):
    B, S, D, H = 64 * 90, 128, 128, 4
    for (T0, K0, T1, K1, name) in ((1024, 64, 40, 25, "agent cross (K=89)"), (128, 25, 0, 0, "agent self (K=25)")):
        g = torch.Generator(device=dev).manual_seed(0)
        qu = torch.randn(B * S, D + H * D, device=dev, generator=g) * 0.3
        div0 = 90 if T1 else 1
        kv0 = torch.randn((B // div0) * T0, 2 * D, device=dev, generator=g)
        kv1 = torch.randn(B * T1, 2 * D, device=dev, generator=g) if T1 else None
        K = K0 + K1
        idx = torch.cat([torch.randint(0, T0, (B, S, K0), device=dev, generator=g, dtype=torch.int32)] +
                        ([torch.randint(0, T1, (B, S, K1), device=dev, generator=g, dtype=torch.int32)] if K1 else []), 2).contiguous()
        inv = torch.rand(B, S, K, device=dev, generator=g) < 0.1
        rel = torch.randn(B, S, K, 3, device=dev, generator=g) * torch.tensor([50.0, 50.0, 1.5], device=dev)
        freq = ops.pe_freq_xy(D, 1e3, dev)
        d_out = torch.randn(B * S, D + H * D, device=dev, generator=g)
        ms_f = timed(lambda: ops.knarpe_attn(qu[:, :D], qu[:, D:], kv0, T0, div0, K0, idx, inv, rel, freq, B, S, D, H, kv1=kv1,
                                             T1=T1, div1=1, K1=K1, fast_trig=True), 3)
        ms_b = timed(lambda: ops.knarpe_attn_bwd(qu[:, :D], qu[:, D:], kv0, T0, div0, K0, idx, inv, rel, freq, B, S, D, d_out,
                                                 kv1=kv1, T1=T1, div1=1, K1=K1), 3)
        pairs = float((~inv).sum())
        print(f"attention {name}: {B * S} tokens, {pairs / 1e6:.1f} M valid pairs: forward {ms_f:.2f} ms, backward {ms_b:.2f} ms "
              f"({pairs / ms_b / 1e6:.1f} G pairs/s; K|V gradient atomics {pairs * 1024 / ms_b / 1e6:.0f} GB/s)")
