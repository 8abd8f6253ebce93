"""Decomposition of the fused-chain kernel's time per tile: synthetic programs at M = 65,536 rows (512 tiles)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tbpkg  # noqa: E402,F401
from trafficbotsv1_5_b200 import ops  # noqa: E402

DEV = "cuda"


def time_us(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    big = torch.empty(64 * 2**20, device=DEV)
    torch.cuda.synchronize()
    for _ in range(4):
        big.fill_(1.0)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


def main():
    M, d = 65536, 128
    g = torch.Generator().manual_seed(0)
    x = torch.randn(M, d, generator=g).half().to(DEV)
    w128 = [(torch.randn(d, d, generator=g) * 0.1).half().to(DEV) for _ in range(8)]
    w512 = (torch.randn(d, 4 * d, generator=g) * 0.05).half().to(DEV)
    y = torch.zeros(M, d, device=DEV)

    def prog(build, n_buf=4):
        p = ops.ChainProgram(DEV, n_buf=n_buf)
        p.load(0, d, 0, f16=True)
        build(p)
        return p.finish()

    cases = {
        "load only + 1 gemm K=128 -> global": lambda p: p.gemm(w128[0], [0], out_g=1, ldg=d),
        "1 gemm K=512 (same buffer x4, 8 boxes) -> global": lambda p: p.gemm(w512, [0, 0, 0, 0], out_g=1, ldg=d),
        "8 independent gemms K=128 -> buffers": lambda p: [p.gemm(w128[i], [0], out_buf=1 + i % 3) for i in range(8)],
        "8 dependent gemms K=128": lambda p: [p.gemm(w128[i], [i % 2], out_buf=(i + 1) % 2) for i in range(8)],
        "8 dependent gemms, same weight": lambda p: [p.gemm(w128[0], [i % 2], out_buf=(i + 1) % 2) for i in range(8)],
        "16 dependent gemms K=128": lambda p: [p.gemm(w128[i % 8], [i % 2], out_buf=(i + 1) % 2) for i in range(16)],
    }
    for name, b in cases.items():
        p = prog(b)
        t = time_us(lambda: p.run([x, y], M))
        print(f"{name:55s} {t:8.1f} us  = {t / (512 / 148):6.2f} us per tile-round")


if __name__ == "__main__":
    main()
