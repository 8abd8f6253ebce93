"""Differentiable forms of the CUDA ops (training path, SURVEY.md 8(f) rank 2 / BASELINE config 4).

torch.autograd is the TAPE only (it orders the backward calls and sums fan-out gradients); every node's forward and
backward is a launch of this library through the C ABI: tb_linear (forward and data gradient, with the transposed
weight), tb_linear_wgrad, tb_grad_mask, tb_group_sum, tb_layernorm / tb_layernorm_bwd, tb_pointnet_pool /
tb_pointnet_pool_bwd, tb_knarpe_attn / tb_knarpe_attn_bwd, tb_il_loss_fwd / _bwd, tb_tl_nll. `ops.linear`,
`ops.layernorm` and `ops.pointnet_pool` route here when an input requires grad, so `model.HotPathModel`'s kernel
sequences are the training forward as they are. fp32 activations (precision 0 = FFMA, 1 = tf32 tcgen05 GEMMs).

Reference: autograd through modules/mlp.py:69, transformer_rpe.py:175-245, attention_rpe.py:58-198,
polyline_encoder.py:50-53, utils/dynamics.py:66-141, utils/rewards.py:35-85, models/metrics/training.py:76-160.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor
from torch.autograd import Function

from . import lib as L
from . import ops


def needs_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


def _rows(t: Tensor) -> Tensor:
    """2-D row-strided fp32 view the kernels accept (unit column stride)."""
    return t if t.stride(1) == 1 else t.contiguous()


def _wt(w: Tensor) -> Tensor:
    """W^T (contiguous) for the data-gradient GEMM, cached on the weight object until it is modified in place."""
    c = getattr(w, "_tb_wt", None)
    if c is None or c[0] != w._version:
        c = (w._version, w.detach().t().contiguous())
        w._tb_wt = c
    return c[1]


def grad_mask(dy: Tensor, y: Optional[Tensor], mask_a: Optional[Tensor], mask_b: Optional[Tensor]) -> Tensor:
    M, N = dy.shape
    out = torch.empty(M, N, dtype=torch.float32, device=dy.device)
    L.check(L.load().tb_grad_mask(L.ptr(dy), dy.stride(0), L.ptr(y), y.stride(0) if y is not None else 0,
                                  L.ptr(ops._u8(mask_a)), L.ptr(ops._u8(mask_b)), M, N, L.ptr(out), N, L.stream()),
            "tb_grad_mask")
    ops._count()
    return out


def wgrad(dv: Tensor, x: Tensor, want_bias: bool, precision: int = 0) -> Tuple[Tensor, Optional[Tensor]]:
    M, N = dv.shape
    K = x.shape[1]
    dw = torch.zeros(N, K, dtype=torch.float32, device=dv.device)
    db = torch.zeros(N, dtype=torch.float32, device=dv.device) if want_bias else None
    L.check(L.load().tb_linear_wgrad(L.ptr(dv), dv.stride(0), L.ptr(x), x.stride(0), M, N, K, L.ptr(dw), K, L.ptr(db),
                                     precision, L.stream()), "tb_linear_wgrad")
    ops._count()
    return dw, db


class _Linear(Function):
    """tb_linear with its epilogue (bias / grouped bias, ReLU, row masks, residual). ReLU together with a residual is
    split by the caller (the ReLU mask is read off the output)."""

    @staticmethod
    def forward(ctx, x, w, b, res, relu, mask_pre, mask_post, precision, bias_group):
        y = ops.linear(x, w, b, relu=relu, mask_pre=mask_pre, res=res, mask_post=mask_post, precision=precision,
                       bias_group=bias_group)
        ctx.cfg = (relu, precision, bias_group, b is not None, res is not None)
        ctx.masks = (mask_pre, mask_post)
        ctx.wt = _wt(w) if ctx.needs_input_grad[0] else None
        ctx.save_for_backward(x, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        relu, precision, bias_group, has_b, has_res = ctx.cfg
        mask_pre, mask_post = ctx.masks
        x, y = ctx.saved_tensors
        dy = _rows(dy)
        need = ctx.needs_input_grad
        dv = grad_mask(dy, y, mask_pre, mask_post) if (relu or mask_pre is not None or mask_post is not None) else dy
        d_res = None
        if has_res and need[3]:
            d_res = grad_mask(dy, None, mask_post, None) if mask_post is not None else dy
        dx = ops.linear(dv, ctx.wt, None, precision=precision) if need[0] else None
        dw = db = None
        if need[1] or (has_b and need[2]):
            plain_bias = has_b and need[2] and not bias_group
            dw, db = wgrad(dv, x, plain_bias, precision)
            if has_b and need[2] and bias_group:
                M, N = dv.shape
                assert M % bias_group == 0
                db = torch.empty(M // bias_group, N, dtype=torch.float32, device=dv.device)
                L.check(L.load().tb_group_sum(L.ptr(dv), dv.stride(0), M // bias_group, bias_group, N, L.ptr(db), N,
                                              L.stream()), "tb_group_sum")
                ops._count()
        return dx, dw, db, d_res, None, None, None, None, None


def linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None, relu: bool = False, mask_pre: Optional[Tensor] = None,
           res: Optional[Tensor] = None, mask_post: Optional[Tensor] = None, out: Optional[Tensor] = None,
           precision: int = 0, bias_group: int = 0) -> Tensor:
    assert precision in (0, 1) and x.dtype == torch.float32, "the training path keeps fp32 activations"
    if relu and res is not None:  # y = mask_post(res + mask_pre(relu(.))): the ReLU mask must be read before the add
        y = _Linear.apply(x, w, b, None, True, mask_pre, None, precision, bias_group) + res
        if mask_post is not None:
            y = y.masked_fill(mask_post.bool().unsqueeze(-1), 0.0)
    else:
        y = _Linear.apply(x, w, b, res, relu, mask_pre, mask_post, precision, bias_group)
    if out is not None:  # column slice of a wider buffer (cat inputs): autograd tracks the copy into the view
        out.copy_(y)
        return out
    return y


class _LayerNorm(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, relu):
        y = ops.layernorm(x, gamma, beta, relu=relu)
        ctx.relu = relu
        ctx.save_for_backward(x, gamma, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, y = ctx.saved_tensors
        dy = _rows(dy)
        if ctx.relu:
            dy = grad_mask(dy, y, None, None)
        M, D = x.shape
        dx = torch.empty(M, D, dtype=torch.float32, device=x.device)
        dg = torch.zeros(D, dtype=torch.float32, device=x.device)
        db = torch.zeros(D, dtype=torch.float32, device=x.device)
        L.check(L.load().tb_layernorm_bwd(L.ptr(x), x.stride(0), L.ptr(gamma), L.ptr(dy), dy.stride(0), L.ptr(dx), D,
                                          L.ptr(dg), L.ptr(db), M, D, L.stream()), "tb_layernorm_bwd")
        ops._count()
        return dx, dg, db, None


def layernorm(x: Tensor, gamma: Tensor, beta: Tensor, relu: bool = False, out: Optional[Tensor] = None) -> Tensor:
    y = _LayerNorm.apply(x, gamma.contiguous(), beta.contiguous(), relu)
    if out is not None and out.data_ptr() != x.data_ptr():
        out.copy_(y)
        return out
    return y


class _Pool(Function):
    @staticmethod
    def forward(ctx, x, invalid, G, Lg, mode):
        assert mode in (1, 2)
        out = ops.pointnet_pool(x, invalid, G, Lg, mode)
        ctx.cfg = (G, Lg, mode)
        ctx.inv = invalid
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (x,) = ctx.saved_tensors
        G, Lg, mode = ctx.cfg
        d_out = _rows(d_out)
        C = x.shape[1]
        dx = torch.empty(G * Lg, C, dtype=torch.float32, device=x.device)
        L.check(L.load().tb_pointnet_pool_bwd(L.ptr(x), x.stride(0), L.ptr(ops._u8(ctx.inv)), G, Lg, C, mode,
                                              L.ptr(d_out), d_out.stride(0), L.ptr(dx), C, L.stream()),
                "tb_pointnet_pool_bwd")
        ops._count()
        return dx, None, None, None, None


def pointnet_pool(x: Tensor, invalid: Tensor, G: int, Lg: int, mode: int) -> Tensor:
    return _Pool.apply(x, invalid, G, Lg, mode)


class _Attn(Function):
    """KNARPE core on fp32 rows: qu [M, D + H*D] = [q | u], K|V tables kv0 (/ kv1) [rows, 2D] (row-strided views)."""

    @staticmethod
    def forward(ctx, qu, kv0, kv1, T0, div0, K0, T1, div1, K1, idx, inv, rel, freq, B, S, D, H, fast_trig):
        out, nv = ops.knarpe_attn(qu[:, :D], qu[:, D:], kv0, T0, div0, K0, idx, inv, rel, freq, B, S, D, H, kv1=kv1,
                                  T1=T1, div1=div1, K1=K1, fast_trig=fast_trig)
        ctx.cfg = (T0, div0, K0, T1, div1, K1, B, S, D, H)
        ctx.lists = (idx, inv, rel, freq)
        ctx.save_for_backward(qu, kv0, kv1)
        ctx.mark_non_differentiable(nv)
        return out, nv

    @staticmethod
    def backward(ctx, d_out, _d_nv):
        qu, kv0, kv1 = ctx.saved_tensors
        T0, div0, K0, T1, div1, K1, B, S, D, H = ctx.cfg
        idx, inv, rel, freq = ctx.lists
        d_out = _rows(d_out)
        M = B * S
        d_qu = torch.empty(M, D + H * D, dtype=torch.float32, device=qu.device)

        def table_grad(kv):  # same leading dimension as the table (the kernel indexes both with it)
            if kv is None:
                return None
            return torch.zeros(kv.shape[0], kv.stride(0), dtype=torch.float32, device=kv.device)[:, :kv.shape[1]]

        d_kv0, d_kv1 = table_grad(kv0), table_grad(kv1)
        L.check(L.load().tb_knarpe_attn_bwd(
            L.ptr(qu), qu.stride(0), L.ptr(qu[:, D:]), qu.stride(0), L.ptr(kv0), kv0.stride(0), T0, div0, K0,
            L.ptr(kv1), kv1.stride(0) if kv1 is not None else 0, T1, div1, K1, L.ptr(idx), L.ptr(ops._u8(inv)),
            L.ptr(rel), L.ptr(freq), B, S, D, H, L.ptr(d_out), L.ptr(d_out[:, D:]), d_out.stride(0), L.ptr(d_qu),
            d_qu.stride(0), L.ptr(d_kv0), L.ptr(d_kv1), L.stream()), "tb_knarpe_attn_bwd")
        ops._count()
        return (d_qu, d_kv0, d_kv1) + (None,) * 15


def knarpe_attn(qu: Tensor, kv0: Tensor, T0: int, div0: int, K0: int, idx: Tensor, inv: Tensor, rel: Tensor,
                freq: Tensor, B: int, S: int, D: int, H: int, kv1: Optional[Tensor] = None, T1: int = 0, div1: int = 1,
                K1: int = 0, fast_trig: bool = False) -> Tuple[Tensor, Tensor]:
    assert qu.dtype == torch.float32 and kv0.dtype == torch.float32 and rel is not None
    return _Attn.apply(qu, kv0, kv1, T0, div0, K0, T1, div1, K1, idx.contiguous(), inv.contiguous(),
                       rel.contiguous(), freq, B, S, D, H, fast_trig)


# ---------------------------------------------------------------------------------------------------- closed loop
class _IlLoss(Function):
    """(sum of weighted imitation errors, count) of a recorded rollout as a function of the stacked action-head
    outputs [T, M, 6] (tb_il_loss_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, act, rec, dyn, weights, step_start):
        T, M, _ = act.shape
        act = act.contiguous()
        order = ("veh", "ped", "cyc")
        ma = ops.host_f3([dyn[k]["max_acc"] for k in order])
        my = ops.host_f3([dyn[k]["max_yaw_rate"] for k in order])
        state_in = torch.empty(T, M, 4, dtype=torch.float32, device=act.device)
        out = torch.zeros(2, dtype=torch.float32, device=act.device)
        B, A = rec["B"], rec["A"]
        assert B * A == M
        common = (L.ptr(rec["gt_valid"]), L.ptr(rec["gt_pose"]), L.ptr(rec["gt_motion"]), L.ptr(rec["tf_mask"]),
                  L.ptr(ops._u8(rec.get("loss_mask"))), rec["n_gt"], rec["sc_div"], B, A, T, step_start,
                  float(weights[0]), float(weights[1]), float(weights[2]))
        L.check(L.load().tb_il_loss_fwd(L.ptr(act), L.ptr(rec["ag_type"]), ma, my, float(dyn["dt"]),
                                        L.ptr(ops._u8(rec["pred_valid"])), L.ptr(rec["pose0"]), L.ptr(rec["motion0"]),
                                        *common, L.ptr(state_in), L.ptr(out), L.stream()), "tb_il_loss_fwd")
        ops._count()
        ctx.args = (rec, ma, my, float(dyn["dt"]), common)
        ctx.save_for_backward(act, state_in)
        ctx.mark_non_differentiable(state_in)
        return out, state_in

    @staticmethod
    def backward(ctx, g_out, _g_state):
        act, state_in = ctx.saved_tensors
        rec, ma, my, dt, common = ctx.args
        g = g_out[:1].contiguous()  # only the sum is differentiable; the count is a constant
        d_act = torch.empty_like(act)
        L.check(L.load().tb_il_loss_bwd(L.ptr(act), L.ptr(rec["ag_type"]), ma, my, dt, L.ptr(ops._u8(rec["pred_valid"])),
                                        *common, L.ptr(state_in), L.ptr(g), L.ptr(d_act), L.stream()), "tb_il_loss_bwd")
        ops._count()
        return d_act, None, None, None, None


def il_loss(act: Tensor, rec: dict, dyn: dict, weights=(0.1, 10.0, 0.1), step_start: int = 10) -> Tuple[Tensor, Tensor]:
    """act [T, M, 6]; rec: recorded rollout (engine state tensors). Returns (out [2] = (sum, count), state_in)."""
    return _IlLoss.apply(act, rec, dyn, weights, step_start)


class _TlNll(Function):
    @staticmethod
    def forward(ctx, logits, tl_invalid, gt_tl, n_gt):
        T, n, _ = logits.shape
        logits = logits.contiguous()
        out = torch.zeros(2, dtype=torch.float32, device=logits.device)
        L.check(L.load().tb_tl_nll(L.ptr(logits), L.ptr(ops._u8(tl_invalid)), L.ptr(gt_tl), n_gt, n, T, L.ptr(out), None,
                                   None, L.stream()), "tb_tl_nll")
        ops._count()
        ctx.args = (tl_invalid, gt_tl, n_gt)
        ctx.save_for_backward(logits)
        return out

    @staticmethod
    def backward(ctx, g_out):
        (logits,) = ctx.saved_tensors
        tl_invalid, gt_tl, n_gt = ctx.args
        T, n, _ = logits.shape
        g = g_out[:1].contiguous()
        d = torch.empty_like(logits)
        L.check(L.load().tb_tl_nll(L.ptr(logits), L.ptr(ops._u8(tl_invalid)), L.ptr(gt_tl), n_gt, n, T, None, L.ptr(g),
                                   L.ptr(d), L.stream()), "tb_tl_nll")
        ops._count()
        return d, None, None, None


def tl_nll(logits: Tensor, tl_invalid: Tensor, gt_tl: Tensor, n_gt: int) -> Tensor:
    """logits [T, n, 5] pre-clamp; gt_tl u8 [n, n_gt, 5]. Returns out [2] = (sum of NLL, count)."""
    return _TlNll.apply(logits, tl_invalid.contiguous(), gt_tl.contiguous(), n_gt)


class _SoftmaxNll(Function):
    @staticmethod
    def forward(ctx, logits, target, row_valid):
        R, Cn = logits.shape
        logits = _rows(logits)
        out = torch.zeros(2, dtype=torch.float32, device=logits.device)
        L.check(L.load().tb_softmax_nll(L.ptr(logits), logits.stride(0), L.ptr(target), L.ptr(ops._u8(row_valid)), R, Cn,
                                        L.ptr(out), None, None, 0, L.stream()), "tb_softmax_nll")
        ops._count()
        ctx.args = (target, row_valid)
        ctx.save_for_backward(logits)
        return out

    @staticmethod
    def backward(ctx, g_out):
        (logits,) = ctx.saved_tensors
        target, row_valid = ctx.args
        R, Cn = logits.shape
        g = g_out[:1].contiguous()
        d = torch.empty(R, Cn, dtype=torch.float32, device=logits.device)
        L.check(L.load().tb_softmax_nll(L.ptr(logits), logits.stride(0), L.ptr(target), L.ptr(ops._u8(row_valid)), R, Cn,
                                        None, L.ptr(g), L.ptr(d), Cn, L.stream()), "tb_softmax_nll")
        ops._count()
        return d, None, None


def softmax_nll(logits: Tensor, target: Tensor, row_valid: Tensor) -> Tensor:
    """logits [R, C] (-inf entries allowed), target int64 [R], row_valid bool [R] -> out [2] = (sum of NLL, count)."""
    assert target.dtype == torch.int64
    return _SoftmaxNll.apply(logits, target.contiguous(), row_valid.contiguous())
