"""Host-side logic that needs no GPU: weight re-association and layout permutations of the tensor-core mode."""
import torch

from oracle import tb_oracle as O
from trafficbotsv1_5_b200 import params
from trafficbotsv1_5_b200.model import H, fuse_attention, head_interleave_perm


def test_head_interleave_perm_layout():
    """tb_knarpe_attn flags bit 4: position 64*(h>>1) + 16*(c>>3) + 8*(h&1) + (c&7) holds channel c of head h — a
    permutation whose 32-byte pieces (16 halves) hold 8 channels of head 2i and 8 of head 2i+1."""
    perm = head_interleave_perm(128)
    assert sorted(perm.tolist()) == list(range(128))
    for h in range(4):
        for c in range(32):
            pos = 64 * (h >> 1) + 16 * (c >> 3) + 8 * (h & 1) + (c & 7)
            assert int(perm[pos]) == 32 * h + c
    pieces = perm.view(8, 16) // 32  # head of every channel, per 32-byte piece
    for i in range(8):
        assert pieces[i, :8].unique().tolist() == [2 * (i // 4)] and pieces[i, 8:].unique().tolist() == [2 * (i // 4) + 1]


def test_fused_attention_weights_reproduce_the_reference_formulation():
    """The re-associated projections (u = W_rk^T q, out-proj over [ov | z]) evaluated densely in torch
    equal the oracle's AttentionRPE (attention_rpe.py:137-190), including under the head-interleaved permutation of the
    q / k / v output features (a per-head permutation applied to q and k alike leaves q.k unchanged; v is un-permuted
    by the kernel's epilogue, modelled here by the inverse permutation)."""
    d, B, S, K = 128, 2, 5, 7
    g = torch.Generator().manual_seed(3)
    shapes = {"in_proj_weight": (3 * d, d), "in_proj_bias": (3 * d,), "out_proj_weight": (d, d), "out_proj_bias": (d,),
              "linear_rpe.weight": (2 * d, d), "linear_rpe.bias": (2 * d,)}
    sd = params.rand_like_state_dict(shapes, 5)
    P = {f"a.{k}": v for k, v in sd.items()}
    src, tgt = torch.randn(B, S, d, generator=g), torch.randn(B, S, K, d, generator=g)
    mask = torch.rand(B, S, K, generator=g) < 0.3
    mask[0, 0] = True
    rel = torch.cat([torch.randn(B, S, K, 2, generator=g) * 20, torch.randn(B, S, K, 1, generator=g)], -1)
    emb = O.pose_emb_xy_yaw(rel[..., :2], rel[..., 2], d)
    ref = O.attention_rpe(P, "a", src, tgt, mask, emb, H).double()

    f = {k: v.double() for k, v in fuse_attention(P, "a", d).items()}
    perm = head_interleave_perm(d)
    inv = torch.argsort(perm)
    for interleaved in (False, True):
        qu = src.double() @ f["w_in_q"].T + f["b_in_q"]
        kv = tgt.double() @ f["w_kv"].T + f["b_kv"]
        q, u, k, v = qu[..., :d], qu[..., d:].view(B, S, H, d), kv[..., :d], kv[..., d:]
        if interleaved:
            q, k, v = q[..., perm], k[..., perm], v[..., perm]
            head_of = (2 * (torch.arange(d) >> 6) + ((torch.arange(d) >> 3) & 1))
        else:
            head_of = torch.arange(d) // (d // H)
        sel = torch.nn.functional.one_hot(head_of, H).double()                       # [d, H]
        logit = torch.einsum("bsc,bskc,ch->bskh", q, k, sel) + torch.einsum("bshc,bskc->bskh", u, emb.double())
        logit = logit.masked_fill(mask[..., None], float("-inf"))
        a = torch.softmax(logit * 0.6931471805599453, dim=2)                         # q, u carry log2(e)
        a = torch.nan_to_num(a)
        ov = torch.einsum("bskh,bskc,ch->bsc", a, v, sel)
        if interleaved:
            ov = ov[..., inv]
        z = torch.einsum("bskh,bskc->bshc", a, emb.double()).reshape(B, S, H * d)
        out = torch.cat([ov, z], -1) @ f["w_out"].T + f["b_out"]
        out = out.masked_fill(mask.all(-1)[..., None], 0.0)
        assert float((out - ref).abs().max()) < 1e-5 * float(ref.abs().max()), interleaved
