"""Does running a layer on two half-batches (working set inside the 126 MB L2) beat one full batch (working set through
HBM)? One agent-layer front half of the 16-bit mode - LayerNorm, self in-projection, self-attention, out-projection with
fused LayerNorm - on the config-3 token count, full batch vs two halves that reuse the same activation buffers.
  gpurun -- python profiles/l2_chunk_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tbpkg  # noqa: E402,F401
from trafficbotsv1_5_b200 import config, params, ops  # noqa: E402
from trafficbotsv1_5_b200.model import HotPathModel  # noqa: E402

dev = "cuda"
cfg = config.default_model_cfg()
sz = config.derived_sizes(cfg)
m = HotPathModel(params.init_params(cfg, 0), cfg, sz, dev, precision=1)
B, A, d, K = 512, 120, 128, sz["k_ag2ag"]
g = torch.Generator(device=dev).manual_seed(0)
src = torch.randn(B * A, d, device=dev, generator=g)
inv = torch.rand(B * A, device=dev, generator=g) < 0.1
idx = torch.randint(0, A, (B, A, K), device=dev, generator=g, dtype=torch.int32)
kinv = torch.rand(B, A, K, device=dev, generator=g) < 0.15
rel = torch.randn(B, A, K, 3, device=dev, generator=g) * torch.tensor([40.0, 40.0, 1.5], device=dev)
p = "ag_encoder.tf_ag2agmptl.layers.1"
f = m.fa[f"{p}.attn_src"]


def half_layer(b0, b1):
    r0, r1 = b0 * A, b1 * A
    knn = dict(idx=idx[b0:b1], inv=kinv[b0:b1], rel=rel[b0:b1])
    x0 = m.ln(src[r0:r1], f"{p}.norm_src", half=True)
    proj, kv = m._in_self(f, x0, K, f"{p}.attn_src")
    o, nv = m._attend(f, proj, b1 - b0, A, kv, A, 1, K, knn)
    return m._out_proj(f"{p}.attn_src", f, o, nv, src[r0:r1], ln_next=f"{p}.norm1")


def timed(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for parts in (1, 2, 4):
    step = B // parts
    t = timed(lambda: [half_layer(i * step, (i + 1) * step) for i in range(parts)])
    print(f"{parts} part(s) of {step} rollout-scenes ({step * A} tokens): {t:.1f} us for LN + in-proj + self-attention + out-proj/LN")
