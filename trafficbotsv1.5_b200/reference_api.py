"""Drop-in mirror of the reference's Python call surface for the hot path: same names, constructor
arguments, `forward` signatures, `state_dict` keys and error behaviour as

  utils.rpe.get_rel_pose / get_tgt_knn_idx                         src/utils/rpe.py:9-90
  utils.pose_emb.PoseEmb (mode pe_xy_yaw)                          src/utils/pose_emb.py:7-56
  models.modules.attention_rpe.AttentionRPE                        src/models/modules/attention_rpe.py:10-198
  models.modules.transformer_rpe.TransformerRPE / TransformerBlockRPE   src/models/modules/transformer_rpe.py:19-245

but executing on the CUDA kernels of libtbknarpe.so. Only the inference path the rollout uses is implemented
(eval mode, RPE branch, no attn_mask / need_weights); anything else raises NotImplementedError — there is no silent
PyTorch fallback. A reference checkpoint loads unchanged (`load_state_dict`), fused projection weights are rebuilt
lazily whenever a parameter changes.

The reference API hands attention a PRE-GATHERED target tensor [B,S,K,d]; these classes honour that (projecting the
gathered rows, K-times redundant exactly like the reference). The rollout engine (engine.py) uses the efficient
project-once-then-gather form instead.
"""
from typing import Optional, Tuple, Union

import math

import torch
from torch import Tensor, nn

from . import ops
from .model import H, HotPathModel


# ---------------------------------------------------------------------------------------------------- utils.rpe
class _LazyRelPose:
    """Handle returned by get_rel_pose: the [B,S,T,3] / [B,S,T] tensors are never materialised; get_tgt_knn_idx runs
    the fused rel-pose + top-K kernel on the stored poses."""

    def __init__(self, pose, invalid, pose2, invalid2):
        self.pose, self.invalid, self.pose2, self.invalid2 = pose, invalid, pose2, invalid2

    @property
    def shape(self):
        return (self.pose.shape[0], self.pose.shape[1], self.pose2.shape[1])


@torch.no_grad()
def get_rel_pose(pose: Tensor, invalid: Tensor, pose2: Optional[Tensor] = None, invalid2: Optional[Tensor] = None):
    """utils/rpe.py:9-37. Returns (rel_pose, rel_dist) handles to be passed to get_tgt_knn_idx."""
    if pose2 is None:
        pose2, invalid2 = pose, invalid
    h = _LazyRelPose(pose, invalid, pose2, invalid2)
    return h, h


@torch.no_grad()
def get_tgt_knn_idx(tgt_invalid: Tensor, rel_pose, rel_dist, n_tgt_knn: int, dist_limit: Union[float, Tensor]
                    ) -> Tuple[Tensor, Tensor, Optional[Tensor]]:
    """utils/rpe.py:62-90. idx int64 [B,S,K] (ascending target index; the reference's order is unspecified),
    tgt_invalid_knn bool [B,S,K], rpe [B,S,K,3]."""
    h = rel_dist
    if not isinstance(h, _LazyRelPose):
        raise NotImplementedError("pass the handles returned by get_rel_pose (materialised rel_dist is not supported)")
    if isinstance(dist_limit, Tensor):
        raise NotImplementedError("per-target dist_limit tensors are not on the rollout path")
    n_tgt = h.pose2.shape[1]
    assert 0 < n_tgt_knn < n_tgt  # utils/rpe.py:79
    idx, inv, rel = ops.knn_select(h.pose, h.invalid, h.pose2, h.invalid2, n_tgt_knn, float(dist_limit))
    return idx.long(), inv, (rel if rel_pose is not None else None)


def knn_rel_pose(pose, invalid, pose2, invalid2, n_tgt_knn: int, dist_limit: float, tgt_div: int = 1):
    """Fused form of the two calls above (int32 indices, shared target tables via tgt_div)."""
    return ops.knn_select(pose, invalid, pose if pose2 is None else pose2, invalid if invalid2 is None else invalid2,
                          n_tgt_knn, dist_limit, tgt_div=tgt_div)


# ---------------------------------------------------------------------------------------------------- PoseEmb
class PoseEmb(nn.Module):
    """utils/pose_emb.py:7-56, mode "pe_xy_yaw" only (the mode of pose_rpe and the agent pose embedding)."""

    def __init__(self, mode: str, pe_dim: int = 256, theta_xy: float = 1e3, theta_cs: float = 1e1):
        super().__init__()
        if mode != "pe_xy_yaw":
            raise NotImplementedError(f"PoseEmb mode {mode!r} is not on the rollout path")
        self.mode, self.out_dim, self.theta_xy = mode, pe_dim, theta_xy
        # same persistent buffers (and names) as the reference: pe_xy.freqs / pe_yaw.freqs (positional_emb.py:11-14,40-42)
        dim = pe_dim // 4
        fxy = 1.0 / (theta_xy ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        self.pe_xy, self.pe_yaw = nn.Module(), nn.Module()
        self.pe_xy.register_buffer("freqs", fxy.repeat_interleave(2, 0))
        self.pe_yaw.register_buffer("freqs", (torch.arange(0, pe_dim // 4) + 1.0).repeat_interleave(2, 0))

    def forward(self, xy: Tensor, dir: Tensor) -> Tensor:
        if dir.shape[-1] != 1:
            dir = torch.atan2(dir[..., 1:2], dir[..., 0:1])
        pose = torch.cat([xy, dir], -1)
        out = ops.pose_emb(pose.reshape(-1, 3).contiguous(), self.pe_xy.freqs[::2].contiguous(), self.out_dim)
        return out.view(*xy.shape[:-1], self.out_dim)


# ---------------------------------------------------------------------------------------------------- attention
class _FusedMixin:
    """Rebuilds the fused projection weights when any parameter / buffer was modified in place or re-loaded."""

    def _runner(self, d_model: int) -> HotPathModel:
        sd = self.state_dict()
        ver = tuple((k, v._version, v.data_ptr()) for k, v in sd.items())
        if getattr(self, "_tb_ver", None) != ver:
            dev = next(iter(sd.values())).device
            if dev.type != "cuda":
                raise RuntimeError("the B200 drop-in modules run on CUDA only (no CPU fallback): call .cuda() first")
            self._tb_runner = HotPathModel.from_state_dict(sd, d_model, device=dev,
                                                           precision=getattr(self, "precision", 0))
            self._tb_ver = ver
        return self._tb_runner


def _knn_dict(idx: Tensor, mask: Tensor, rpe: Tensor, d_rpe: int) -> dict:
    out = dict(idx=idx.to(torch.int32).contiguous(), inv=mask.contiguous())
    if rpe.shape[-1] == 3 and d_rpe != 3:
        out["rel"] = rpe.float().contiguous()          # raw relative pose: embedding evaluated in-kernel
    else:
        out["emb"] = rpe.float().contiguous()          # reference form: materialised PoseEmb tensor
    return out


class _KnarpeCore(torch.autograd.Function):
    """Differentiable KNARPE core: forward tb_knarpe_attn, backward tb_knarpe_attn_bwd (fp32; gradients w.r.t. q, u and
    the K|V table, none into the relative pose). Used by AttentionRPE when gradients are required."""

    @staticmethod
    def forward(ctx, q, u, kv, idx, inv, rel, freq, B, S, K):
        d = q.shape[1]
        out, nv = ops.knarpe_attn(q, u, kv, S * K, 1, K, idx, inv, rel, freq, B, S, d, H)
        ctx.save_for_backward(q, u, kv, idx, inv, rel, freq)
        ctx.dims = (B, S, K, d)
        ctx.mark_non_differentiable(nv)
        return out, nv

    @staticmethod
    def backward(ctx, d_out, _d_nv):
        q, u, kv, idx, inv, rel, freq = ctx.saved_tensors
        B, S, K, d = ctx.dims
        d_qu, d_kv, _ = ops.knarpe_attn_bwd(q, u, kv, S * K, 1, K, idx, inv, rel, freq, B, S, d, d_out.contiguous())
        return d_qu[:, :d], d_qu[:, d:], d_kv, None, None, None, None, None, None, None


class AttentionRPE(nn.Module, _FusedMixin):
    """KNARPE attention, src/models/modules/attention_rpe.py:10-198 (same parameters: in_proj_weight, in_proj_bias,
    out_proj_weight, out_proj_bias, linear_rpe.*)."""

    def __init__(self, d_model: int, n_head: int, dropout_p: float = 0.1, bias: bool = True, d_rpe: int = -1,
                 apply_q_rpe: bool = False, precision: int = 0) -> None:
        super().__init__()
        self.precision = precision  # 0: fp32 everywhere (1e-4 parity); 1: 16-bit mode (extra argument of the drop-in)
        self.d_model, self.n_head, self.d_head = d_model, n_head, d_model // n_head
        self.apply_q_rpe, self.d_rpe = apply_q_rpe, d_rpe
        assert self.d_head * n_head == d_model, "d_model must be divisible by n_head"  # attention_rpe.py:31
        if apply_q_rpe or not bias or n_head != H or d_rpe != d_model:
            raise NotImplementedError("only apply_q_rpe=False, bias=True, n_head=4, d_rpe=d_model (sim_agent.yaml)")
        self.linear_rpe = nn.Linear(d_rpe, 2 * d_model, bias=bias)
        self.in_proj_weight = nn.Parameter(torch.empty((3 * d_model, d_model)))
        self.out_proj_weight = nn.Parameter(torch.empty((d_model, d_model)))
        self.in_proj_bias = nn.Parameter(torch.empty(3 * d_model))
        self.out_proj_bias = nn.Parameter(torch.empty(d_model))
        self.dropout = nn.Dropout(p=dropout_p) if dropout_p > 0 else None
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.xavier_uniform_(self.out_proj_weight)
        nn.init.constant_(self.in_proj_bias, 0.0)
        nn.init.constant_(self.out_proj_bias, 0.0)

    def forward(self, src: Tensor, tgt: Optional[Tensor] = None, tgt_padding_mask: Optional[Tensor] = None,
                attn_mask: Optional[Tensor] = None, rpe: Optional[Tensor] = None, need_weights=False
                ) -> Tuple[Tensor, Optional[Tensor]]:
        if self.training and self.dropout is not None:
            raise NotImplementedError("training-mode dropout is outside the rollout hot path: call .eval()")
        if rpe is None or tgt is None or tgt.dim() != 4 or attn_mask is not None or need_weights:
            raise NotImplementedError("only the KNN + RPE branch (attention_rpe.py:137-164) is implemented")
        assert self.d_rpe > 0  # attention_rpe.py:139
        # gradients are produced when autograd is on and either an input carries a graph or the module is in train()
        # mode (eval-mode calls on plain tensors take the fused inference path, like the reference under no_grad)
        needs_grad = torch.is_grad_enabled() and (src.requires_grad or tgt.requires_grad or self.training)
        if needs_grad:
            return self._forward_autograd(src, tgt, tgt_padding_mask, rpe), None
        with torch.no_grad():
            return self._forward_inference(src, tgt, tgt_padding_mask, rpe), None

    def _forward_autograd(self, src: Tensor, tgt: Tensor, tgt_padding_mask: Optional[Tensor], rpe: Tensor) -> Tensor:
        """Differentiable path (fp32): the projections are torch GEMMs (autograd), the attention core is the CUDA
        kernel pair tb_knarpe_attn / tb_knarpe_attn_bwd. Same re-association as DESIGN.md 3, written with the
        module's own parameters so that gradients reach in_proj / linear_rpe / out_proj (attention_rpe.py:92-97,
        147-161, 180-186)."""
        B, S, K, d = tgt.shape
        if d != 128 or rpe.shape[-1] != 3:
            raise NotImplementedError("autograd path: d_model 128 and raw relative poses [B,S,K,3] only")
        M, dh = B * S, self.d_head
        w_q, w_kv = self.in_proj_weight[:d], self.in_proj_weight[d:]
        b_q, b_kv = self.in_proj_bias[:d], self.in_proj_bias[d:]
        w_rk, w_rv = self.linear_rpe.weight[:d], self.linear_rpe.weight[d:]
        b_rv = self.linear_rpe.bias[d:]
        scale = math.log2(math.e) / math.sqrt(dh)
        q = torch.nn.functional.linear(src.reshape(M, d).float(), w_q, b_q) * scale
        u = torch.einsum("mhc,hcr->mhr", q.view(M, H, dh), w_rk.view(H, dh, self.d_rpe)).reshape(M, H * self.d_rpe)
        kv = torch.nn.functional.linear(tgt.reshape(M * K, d).float(), w_kv, b_kv)
        idx = torch.arange(S * K, dtype=torch.int32, device=src.device).view(1, S, K).expand(B, -1, -1).contiguous()
        mask = (tgt_padding_mask if tgt_padding_mask is not None
                else torch.zeros(B, S, K, dtype=torch.bool, device=src.device)).contiguous()
        freq = ops.pe_freq_xy(d, 1e3, src.device)
        o, nv = _KnarpeCore.apply(q.contiguous(), u.contiguous(), kv.contiguous(), idx, mask,
                                  rpe.float().contiguous(), freq, B, S, K)
        ov, z = o[:, :d], o[:, d:].view(M, H, self.d_rpe)
        rv = torch.einsum("mhr,hcr->mhc", z, w_rv.view(H, dh, self.d_rpe)).reshape(M, d) + b_rv  # sum_j a_j = 1
        out = torch.nn.functional.linear(ov + rv, self.out_proj_weight, self.out_proj_bias)
        return out.masked_fill(nv[:, None], 0.0).view(B, S, d)                                   # :188-190

    def _forward_inference(self, src: Tensor, tgt: Tensor, tgt_padding_mask: Optional[Tensor], rpe: Tensor) -> Tensor:
        B, S, K, d = tgt.shape
        m = self._runner(d)
        f = m.fa[""]
        idx = torch.arange(S * K, dtype=torch.int32, device=src.device).view(1, S, K).expand(B, -1, -1)
        mask = tgt_padding_mask if tgt_padding_mask is not None else torch.zeros(B, S, K, dtype=torch.bool,
                                                                                  device=src.device)
        knn = _knn_dict(idx, mask, rpe, self.d_rpe)
        x, t = src.reshape(B * S, d).float().contiguous(), tgt.reshape(B * S * K, d).float().contiguous()
        if m.precision == 1 and "rel" in knn:
            # 16-bit mode (`precision = 1` on the module): tcgen05 projections that write fp16 [q|u] rows and the fp16
            # K|V table, the 16-bit attention core (mma.sync kernel for d_model 128 and K <= 128, the SIMT kernel on
            # fp16 tables otherwise: d_model 256 = BASELINE config 2), fp16 [ov|z] rows into a kind::f16 out-projection.
            qu = torch.empty(B * S, d + H * d, dtype=torch.float16, device=src.device)
            ops.linear(x, f["w_in_q"], f["b_in_q"], precision=1, out_h=qu, col_h=0)
            kv = torch.empty(B * S * K, 2 * d, dtype=torch.float16, device=src.device)
            ops.linear(t, f["w_kv"], f["b_kv"], precision=1, out_h=kv, col_h=0)
            o, nv = ops.knarpe_attn(qu[:, :d], qu[:, d:], kv, S * K, 1, K, knn["idx"], knn["inv"], knn["rel"], m.freq_rpe,
                                    B, S, d, H, fast_trig=True, out_dtype=torch.float16)
            out = ops.linear(o, m._half("w_out", f["w_out"]), f["b_out"], mask_pre=nv, precision=2)
            return out.view(B, S, d)
        proj = ops.linear(x, f["w_in_q"], f["b_in_q"], precision=m.precision)
        kv = ops.linear(t, f["w_kv"], f["b_kv"], precision=m.precision)
        o, nv = ops.knarpe_attn(proj[:, :d], proj[:, d:], kv, S * K, 1, K, knn["idx"], knn["inv"], knn.get("rel"),
                                m.freq_rpe, B, S, d, H, emb=knn.get("emb"))
        out = ops.linear(o, f["w_out"], f["b_out"], mask_pre=nv, precision=m.precision)
        return out.view(B, S, d)


class TransformerRPE(nn.Module):
    """One pre-LN layer, src/models/modules/transformer_rpe.py:138-245 — parameter container with the reference's
    sub-module names; executed by TransformerBlockRPE."""

    def __init__(self, d_model: int, n_head: int, k_feedforward: int, dropout_p: float, bias: bool, activation: str,
                 mode: str, d_rpe: int = -1, apply_q_rpe: bool = False) -> None:
        super().__init__()
        if activation != "relu":
            raise NotImplementedError("activation relu only (sim_agent.yaml:39)")
        self.mode = mode
        self.norm1 = nn.LayerNorm(d_model)
        self.norm_tgt = nn.LayerNorm(d_model)
        if mode == "dec_cross_attn":
            self.attn_src = AttentionRPE(d_model, n_head, dropout_p, bias, d_rpe, apply_q_rpe)
            self.norm_src = nn.LayerNorm(d_model)
        self.attn = AttentionRPE(d_model, n_head, dropout_p, bias, d_rpe, apply_q_rpe)
        self.linear1 = nn.Linear(d_model, k_feedforward * d_model)
        self.linear2 = nn.Linear(k_feedforward * d_model, d_model)
        self.norm2 = nn.LayerNorm(d_model)


class TransformerBlockRPE(nn.Module, _FusedMixin):
    """src/models/modules/transformer_rpe.py:19-135. Modes enc_self_attn (tgt = int64 KNN indices) and dec_cross_attn
    (tgt = pre-gathered [B,S,K,d] features, decoder_tgt = int64 KNN indices), as used by the map / TL / agent encoders."""

    def __init__(self, d_model: int, n_head: int = 4, k_feedforward: int = 4, dropout_p: float = 0.1, bias: bool = True,
                 activation: str = "relu", out_layernorm: bool = False, apply_q_rpe: bool = False, n_layer: int = 1,
                 mode: str = "enc_self_attn", d_rpe: int = -1, precision: int = 0) -> None:
        super().__init__()
        assert mode in ("enc_self_attn", "enc_cross_attn", "dec_cross_attn")  # transformer_rpe.py:36
        if mode == "enc_cross_attn" or out_layernorm:
            raise NotImplementedError("enc_cross_attn / out_layernorm are unused by the HPTR rollout configuration")
        self.mode, self.d_model, self.d_rpe, self.precision = mode, d_model, d_rpe, precision
        self.layers = nn.ModuleList([TransformerRPE(d_model, n_head, k_feedforward, dropout_p, bias, activation, mode,
                                                    d_rpe, apply_q_rpe) for _ in range(n_layer)])
        self.out_layernorm = None

    @torch.no_grad()
    def forward(self, src: Tensor, src_padding_mask: Optional[Tensor] = None, tgt: Optional[Tensor] = None,
                tgt_padding_mask: Optional[Tensor] = None, rpe: Optional[Tensor] = None,
                decoder_tgt: Optional[Tensor] = None, decoder_tgt_padding_mask: Optional[Tensor] = None,
                decoder_rpe: Optional[Tensor] = None, attn_mask: Optional[Tensor] = None, need_weights: bool = False
                ) -> Tuple[Tensor, Optional[Tensor]]:
        if self.training:
            raise NotImplementedError("training mode (dropout) is outside the rollout hot path: call .eval()")
        if attn_mask is not None or need_weights or rpe is None:
            raise NotImplementedError("only the KNN + RPE inference path is implemented")
        B, S, d = src.shape
        m = self._runner(d)
        x = src.reshape(B * S, d).float().contiguous()
        inv = (src_padding_mask if src_padding_mask is not None
               else torch.zeros(B, S, dtype=torch.bool, device=src.device)).reshape(-1).contiguous()
        if self.mode == "enc_self_attn":
            if tgt is None or tgt.dtype != torch.int64:
                raise NotImplementedError("enc_self_attn expects int64 KNN indices as tgt (transformer_rpe.py:87)")
            knn_self, cross = _knn_dict(tgt, tgt_padding_mask, rpe, self.d_rpe), None
        else:
            if decoder_tgt is None or decoder_tgt.dtype != torch.int64 or tgt is None or tgt.dim() != 4:
                raise NotImplementedError("dec_cross_attn expects int64 decoder_tgt and a gathered tgt [B,S,K,d]")
            knn_self = _knn_dict(decoder_tgt, decoder_tgt_padding_mask, decoder_rpe, self.d_rpe)
            K2 = tgt.shape[2]
            idx2 = torch.arange(S * K2, dtype=torch.int32, device=src.device).view(1, S, K2).expand(B, -1, -1)
            cross = _knn_dict(idx2, tgt_padding_mask, rpe, self.d_rpe)
            table = tgt.reshape(B * S * K2, d).float().contiguous()
        # fp16 K|V tables + tensor-core attention need the in-kernel embedding and <= 128 neighbours per token
        m.kv_half = (m.precision == 1 and d == 128 and "rel" in knn_self and knn_self["idx"].shape[-1] <= 128 and
                     (cross is None or ("rel" in cross and cross["idx"].shape[-1] <= 128)))
        for i in range(len(self.layers)):
            p = f"layers.{i}"
            c = None
            if cross is not None:
                c = dict(cross, kv0=m.kv_table(table, p, "norm_tgt"), T0=S * K2, div0=1, K0=K2)
            x = m.tf_layer(p, self.mode, x, inv, B, S, knn_self, c)
        return x.view(B, S, d), None
