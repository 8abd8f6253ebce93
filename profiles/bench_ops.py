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


def time_burst(fn, n=20, reps=5):
    """n back-to-back launches per event pair (launch latency amortised; inputs stay L2-warm as inside the step)."""
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / n)
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

# ---- the same projections as the tensor-core mode runs them: fp16 rows in (kind::f16), fp16 rows or fp32 + residual out
if os.environ.get("TB_BENCH_F16", "1") != "0":
    M = 65536
    for N, K, name, out_h, extra in [(896, 128, "in-proj self", True, False), (640, 128, "in-proj cross q|u", True, False),
                                     (128, 640, "out-proj", False, True), (512, 128, "ffn1 (relu)", True, False),
                                     (128, 512, "ffn2", False, True)]:
        x = (torch.randn(M, K, device=dev)).half()
        w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
        b = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev) if extra else None
        nv = (torch.rand(M, device=dev) < 0.01) if extra else None
        yh = torch.empty(M, N, dtype=torch.float16, device=dev) if out_h else None
        y = None if out_h else torch.empty(M, N, device=dev)
        nbytes = M * K * 2 + N * K * 2 + (M * N * 2 if out_h else M * N * 8)
        if out_h:
            fn = lambda: ops.linear(x, w, b, precision=2, out_h=yh, col_h=0, relu="relu" in name)  # noqa: E731
        else:
            fn = lambda: ops.linear(x, w, b, precision=2, out=y, res=res, mask_pre=nv)  # noqa: E731
        t = time_burst(fn)
        print(f"f16 {name:18s} M={M:7d} N={N:4d} K={K:4d}: {t:8.1f} us  {nbytes / t / 1e6:6.2f} TB/s "
              f"(byte bound {nbytes / 6545.3e3:5.1f} us)")

    x = torch.randn(M, 128, device=dev)
    g_, b_ = torch.randn(128, device=dev), torch.randn(128, device=dev)
    t = time_burst(lambda: ops.layernorm(x, g_, b_, out_dtype=torch.float16))
    print(f"f16 layernorm          M={M:7d} D= 128        : {t:8.1f} us  {M * 128 * 6 / t / 1e6:6.2f} TB/s "
          f"(byte bound {M * 128 * 6 / 6545.3e3:5.1f} us)")

# ---- KNARPE core forward / backward at the agent cross-attention shape (fp32 tables; SURVEY 8(f) rank 2 first piece)
B, S, T0, K0, d = 512, 128, 1024, 89, 128
g = torch.Generator(device=dev).manual_seed(0)
M = B * S
q = torch.randn(M, d, device=dev, generator=g) * 0.3
u = torch.randn(M, 4 * d, device=dev, generator=g) * 0.1
kv = torch.randn(16 * T0, 2 * d, device=dev, generator=g)
idx = torch.randint(0, T0, (B, S, K0), device=dev, generator=g, dtype=torch.int32)
inv = torch.rand(B, S, K0, device=dev, generator=g) < 0.1
rel = torch.cat([(torch.rand(B, S, K0, 2, device=dev, generator=g) * 2 - 1) * 100,
                 (torch.rand(B, S, K0, 1, device=dev, generator=g) * 2 - 1) * 3], -1).contiguous()
freq = ops.pe_freq_xy(d, 1e3, dev)
d_out = torch.randn(M, 5 * d, device=dev, generator=g)
args = (q, u, kv, T0, 32, K0, idx, inv, rel, freq, B, S, d)
t_f = time_op(lambda: ops.knarpe_attn(*args), 5)
t_b = time_op(lambda: ops.knarpe_attn_bwd(*args, d_out), 5)
print(f"KNARPE core fp32, {M} tokens x {K0} neighbours: forward {t_f:.0f} us, backward {t_b:.0f} us "
      f"(incl. zero-fill of the gradient table)")
