"""Fused chain programs (tb_chain_run) vs the one-launch-per-layer path: accuracy against the fp32 path and time per
launch at config-3 row counts (65,536 agent tokens). python profiles/chain_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tbpkg  # noqa: E402,F401
from trafficbotsv1_5_b200 import config, ops, params  # noqa: E402
from trafficbotsv1_5_b200.model import HotPathModel  # noqa: E402

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
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    cfg = config.default_model_cfg()
    P = params.init_params(cfg, 0)
    sz = config.derived_sizes(cfg)
    g = torch.Generator().manual_seed(0)
    d = 128
    m32 = HotPathModel(P, cfg, sz, DEV, 0)
    m16 = HotPathModel(P, cfg, sz, DEV, 1)
    x = torch.randn(M, d, generator=g).to(DEV)
    st = dict(pose=(torch.randn(M, 3, generator=g) * 30).to(DEV), navi_invalid=(torch.rand(M, generator=g) < 0.2).to(DEV),
              latent=torch.randn(M, 16, generator=g).to(DEV), latent_invalid=(torch.rand(M, generator=g) < 0.1).to(DEV))
    navi0 = dict(feat=torch.randn(M, d, generator=g).to(DEV) * 0.3, pose=(torch.randn(M, 3, generator=g) * 30).to(DEV))

    def run_heads(m):
        navi = dict(navi0)
        m.latent_static(navi, st)
        x_cat = torch.zeros(M, 2 * d, device=DEV)
        x_cat[:, :d] = x
        return m.heads(x_cat, st, navi), navi, x_cat

    ref, _, _ = run_heads(m32)
    scale = float(ref.abs().max())
    for chain in ("1", "0"):
        m16.opt["chain"] = chain == "1"
        out, navi, x_cat = run_heads(m16)
        err = float((out - ref).abs().max())
        t = time_us(lambda: m16.heads(x_cat, st, navi))
        print(f"heads  chain={chain}: max |act - fp32| = {err:.3e} (scale {scale:.2f}), {t:.1f} us per call (M={M})")
    # ---- FFN of agent layer 0
    p = "ag_encoder.tf_ag2agmptl.layers.0"
    src = torch.randn(M, d, generator=g).to(DEV)
    inv = (torch.rand(M, generator=g) < 0.1).to(DEV)
    x2_32 = m32.ln(src, f"{p}.norm2")
    h = m32.lin(x2_32, f"{p}.linear1", relu=True)
    ref = m32.lin(h, f"{p}.linear2", res=src, mask_post=inv)
    nxt = "ag_encoder.tf_ag2agmptl.layers.1.norm_src"
    ref_ln = m32.ln(ref, nxt)
    x2 = m16.ln(src, f"{p}.norm2", half=True)

    def ffn(m):
        if m.chain_fused:
            y = torch.empty(M, d, device=DEV)
            ln_rows = torch.empty(M, d, dtype=torch.float16, device=DEV)
            m._ffn_program(p, d, nxt).run([x2, src, ops._u8(inv), y, ln_rows], M)
            return y, ln_rows
        hh = torch.empty(M, 4 * d, dtype=torch.float16, device=DEV)
        m._proj(x2, f"{p}.linear1", m.P[f"{p}.linear1.weight"], m.P[f"{p}.linear1.bias"], relu=True, out_h=hh, col_h=0)
        return ops.linear_ln(hh, m._half(f"{p}.linear2", m.P[f"{p}.linear2.weight"]), m.P[f"{p}.linear2.bias"],
                             m.P[f"{nxt}.weight"], m.P[f"{nxt}.bias"], res=src, mask_post=inv, precision=2)

    for chain in ("1", "0"):
        m16.opt["chain"] = chain == "1"
        y, ln_rows = ffn(m16)
        e1 = float((y - ref).abs().max())
        e2 = float((ln_rows.float() - ref_ln)[~inv].abs().max())
        t = time_us(lambda: ffn(m16))
        print(f"FFN    chain={chain}: max |y - fp32| = {e1:.3e} (scale {float(ref.abs().max()):.2f}), LN rows {e2:.3e}, "
              f"{t:.1f} us per call")


if __name__ == "__main__":
    with torch.no_grad():
        main()
