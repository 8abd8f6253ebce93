"""CPU ORACLE (test infrastructure, NOT product code) for the TrafficBots-V1.5 hot path.

A functional, torch-CPU fp32 restatement of the reference's KNARPE attention + closed-loop rollout,
written from the reference's behaviour; every function cites the reference file:line it follows
(paths relative to /root/reference/src). Weights are a flat dict `P` keyed by the reference's own
`state_dict` names, so a reference checkpoint drives the oracle unchanged.

Pinning: the reference ships no tests/golden vectors ("parity unpinned by the reference"), so the
pins are golden vectors produced by importing the real reference modules in the build container
(tests/golden/make_golden.py -> tests/golden/*.pt) and checked in tests/test_oracle_golden.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module. The CUDA product path never does.
"""
import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor
import torch.nn.functional as F

INF = float("inf")


# --------------------------------------------------------------------------------------------------
# geometry: utils/transform_utils.py:121-213
# --------------------------------------------------------------------------------------------------
def to_local_xy(xy: Tensor, origin_xy: Tensor, origin_yaw: Tensor) -> Tensor:
    """(p - p0) @ [[c,-s],[s,c]] -> x' = dx c + dy s ; y' = -dx s + dy c
    (transform_utils.py:121-131 torch_rad2rot, :146-157 torch_pos2local).
    xy [..., M, 2], origin_xy [..., 1, 2], origin_yaw [...]."""
    c, s = torch.cos(origin_yaw), torch.sin(origin_yaw)
    rot = torch.stack([torch.stack([c, -s], -1), torch.stack([s, c], -1)], -2)
    return torch.matmul(xy - origin_xy, rot)


def get_rel_pose(pose: Tensor, invalid: Tensor, pose2: Optional[Tensor] = None,
                 invalid2: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """utils/rpe.py:9-37. Returns rel_pose [B,S,T,3] (yaw difference NOT wrapped, :31) and
    rel_dist [B,S,T] with +inf where src or tgt is invalid (:36)."""
    if pose2 is None:
        pose2, invalid2 = pose, invalid
    xy_l = to_local_xy(pose2[:, None, :, :2], pose[:, :, None, :2], pose[:, :, 2])
    dyaw = pose2[:, None, :, 2] - pose[:, :, None, 2]
    rel_pose = torch.cat([xy_l, dyaw.unsqueeze(-1)], -1)
    rel_dist = torch.norm(rel_pose[..., :2], dim=-1)
    rel_dist = rel_dist.masked_fill(invalid[:, :, None] | invalid2[:, None, :], INF)
    return rel_pose, rel_dist


def knn_select(tgt_invalid: Tensor, rel_pose: Tensor, rel_dist: Tensor, k: int, dist_limit: float):
    """utils/rpe.py:62-90. idx [B,S,k] int64, invalid [B,S,k] bool, rpe [B,S,k,3].
    The reference's order within the k winners is unspecified (topk sorted=False); the oracle sorts
    ascending by (dist, idx) — compare as SETS (SURVEY.md §7 hard part 2)."""
    B, S, T = rel_dist.shape
    assert 0 < k < T  # rpe.py:79
    # stable sort on distance == tie-break by lower index
    order = torch.sort(rel_dist, dim=-1, stable=True)[1][..., :k]
    d_k = torch.gather(rel_dist, 2, order)
    inv = torch.gather(tgt_invalid[:, None, :].expand(-1, S, -1), 2, order) | (d_k > dist_limit)
    rpe = torch.gather(rel_pose, 2, order[..., None].expand(-1, -1, -1, 3))
    return order, inv, rpe


# --------------------------------------------------------------------------------------------------
# embeddings: utils/positional_emb.py:6-54, utils/pose_emb.py:26-89
# --------------------------------------------------------------------------------------------------
def pe_freqs_xy(dim: int, theta: float) -> Tensor:
    """positional_emb.py:11 — one frequency per (cos,sin) pair: theta^(-2i/dim), i < dim/2."""
    return 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))


def pe_freqs_yaw(dim: int) -> Tensor:
    """positional_emb.py:40 — integer frequencies 1..dim/2."""
    return torch.arange(0, dim // 2) + 1.0


def _pe(x: Tensor, freqs: Tensor) -> Tensor:
    """positional_emb.py:24-25: [cos(x f_0..) | sin(x f_0..)]."""
    a = x.unsqueeze(-1) * freqs
    return torch.cat([a.cos(), a.sin()], -1)


def pose_emb_xy_yaw(xy: Tensor, yaw: Tensor, pe_dim: int, theta_xy: float = 1e3) -> Tensor:
    """PoseEmb mode pe_xy_yaw (pose_emb.py:21-22,50-55): [PE(x) | PE(y) | PERad(yaw)], dims pe/4, pe/4, pe/2.
    xy [...,2], yaw [...]."""
    fxy = pe_freqs_xy(pe_dim // 4, theta_xy)
    return torch.cat([_pe(xy[..., 0], fxy), _pe(xy[..., 1], fxy), _pe(yaw, pe_freqs_yaw(pe_dim // 2))], -1)


def encode_polyline(pos: Tensor, dirv: Tensor) -> Tensor:
    """PoseEmb mode mpa_pl (pose_emb.py:59-89): 7 geometric features of a segment seen from the origin."""
    eps = torch.finfo(pos.dtype).eps
    proj = (-pos * dirv).sum(-1) / ((dirv * dirv).sum(-1) + eps)
    closest = pos + proj.clamp(0, 1).unsqueeze(-1) * dirv
    r = torch.norm(closest, dim=-1, keepdim=True)
    dn = torch.norm(dirv, dim=-1, keepdim=True)
    return torch.cat([r, closest / (r + eps), dirv / (dn + eps), dn,
                      torch.norm(pos + dirv - closest, dim=-1, keepdim=True)], -1)


# --------------------------------------------------------------------------------------------------
# small networks: modules/mlp.py:20-72, polyline_encoder.py:36-63, pooling.py:7-38, input_encoder.py:41-61
# --------------------------------------------------------------------------------------------------
def mlp(P: Dict[str, Tensor], prefix: str, x: Tensor, idxs, end_act: bool) -> Tensor:
    """Sequential of Linear(+ReLU) at `fc_layers.<i>` for i in idxs (eval mode: dropout = identity)."""
    for n, i in enumerate(idxs):
        x = F.linear(x, P[f"{prefix}.fc_layers.{i}.weight"], P[f"{prefix}.fc_layers.{i}.bias"])
        if n < len(idxs) - 1 or end_act:
            x = F.relu(x)
    return x


def pointnet(P: Dict[str, Tensor], prefix: str, x: Tensor, invalid: Tensor, n_layer: int = 3) -> Tensor:
    """PolylineEncoder (use_pointnet, pooling max_valid): polyline_encoder.py:50-53 + pooling.py:18-19,38.
    x [B,N,L,d], invalid [B,N,L] -> [B,N,d]."""
    L = invalid.shape[-1]
    m = invalid.unsqueeze(-1)
    for i in range(n_layer):
        h = F.relu(F.linear(x, P[f"{prefix}.mlp_layers.{i}.fc_layers.0.weight"],
                            P[f"{prefix}.mlp_layers.{i}.fc_layers.0.bias"]))
        h = h.masked_fill(m, -INF)
        x = torch.cat([h, h.amax(2, keepdim=True).expand(-1, -1, L, -1)], -1)
        x = x.masked_fill(m, 0.0)
    pooled = x.masked_fill(m, -INF).amax(2)
    return pooled.masked_fill(invalid.all(-1, keepdim=True), 0.0)


def last_valid(x: Tensor, valid: Tensor) -> Tensor:
    """seq_pooling mode last_valid (pooling.py:24-29,38). x [B,N,L,c], valid [B,N,L]."""
    L = valid.shape[-1]
    idx = L - 1 - torch.max(valid.flip(2).to(torch.uint8), dim=2)[1]
    out = torch.gather(x, 2, idx[:, :, None, None].expand(-1, -1, 1, x.shape[-1])).squeeze(2)
    return out.masked_fill(~valid.any(-1, keepdim=True), 0.0)


# --------------------------------------------------------------------------------------------------
# KNARPE: modules/attention_rpe.py:58-198 (RPE branch), modules/transformer_rpe.py:48-245
# --------------------------------------------------------------------------------------------------
def attention_rpe(P, prefix: str, src: Tensor, tgt: Tensor, mask: Tensor, rpe: Tensor, n_head: int) -> Tensor:
    """src [B,S,d]; tgt [B,S,K,d] (already gathered + normed); mask [B,S,K] True=invalid; rpe [B,S,K,d_rpe]."""
    B, S, d = src.shape
    K = tgt.shape[2]
    dh = d // n_head
    w, b = P[f"{prefix}.in_proj_weight"], P[f"{prefix}.in_proj_bias"]
    q = F.linear(src, w[:d], b[:d])                                   # :96
    kv = F.linear(tgt, w[d:], b[d:])                                  # :97
    k, v = kv.chunk(2, -1)
    r = F.linear(rpe, P[f"{prefix}.linear_rpe.weight"], P[f"{prefix}.linear_rpe.bias"])  # :147
    rk, rv = r.chunk(2, -1)                                           # :152 (apply_q_rpe False)
    hv = lambda t: t.view(B, S, K, n_head, dh).movedim(3, 1)          # noqa: E731  [B,H,S,K,dh]
    qh = q.view(B, S, n_head, dh).transpose(1, 2).unsqueeze(3)        # [B,H,S,1,dh]
    logits = (qh * (hv(k) + hv(rk))).sum(-1)                          # :161
    none_valid = mask.all(-1)                                         # :112-118
    m = mask & ~none_valid.unsqueeze(-1)
    logits = logits.masked_fill(m.unsqueeze(1), -INF)                 # :168
    a = torch.softmax(logits / math.sqrt(dh), -1)                     # :170
    o = ((hv(v) + hv(rv)) * a.unsqueeze(-1)).sum(3)                   # :182
    o = o.transpose(1, 2).flatten(2, 3)
    o = F.linear(o, P[f"{prefix}.out_proj_weight"], P[f"{prefix}.out_proj_bias"])  # :186
    return o.masked_fill(none_valid.unsqueeze(-1), 0.0)               # :188-190


def _ln(P, prefix, x):
    return F.layer_norm(x, (x.shape[-1],), P[f"{prefix}.weight"], P[f"{prefix}.bias"], 1e-5)


def _gather_rows(x: Tensor, idx: Tensor) -> Tensor:
    """x [B,T,d], idx [B,S,K] -> [B,S,K,d] (transformer_rpe.py:88,117)."""
    B, S, K = idx.shape
    return torch.gather(x[:, None].expand(-1, S, -1, -1), 2, idx[..., None].expand(-1, -1, -1, x.shape[-1]))


def transformer_layer(P, prefix: str, mode: str, n_head: int, src, src_invalid, tgt, tgt_mask, rpe,
                      dec_tgt=None, dec_mask=None, dec_rpe=None) -> Tensor:
    """TransformerRPE.forward, transformer_rpe.py:175-245 (eval mode)."""
    if mode == "dec_cross_attn":                                      # :206-214
        s = attention_rpe(P, f"{prefix}.attn_src", _ln(P, f"{prefix}.norm_src", src),
                          _ln(P, f"{prefix}.norm_src", dec_tgt), dec_mask, dec_rpe, n_head)
        src = src + s
    tnorm = "norm1" if mode == "enc_self_attn" else "norm_tgt"        # :219-223
    s = attention_rpe(P, f"{prefix}.attn", _ln(P, f"{prefix}.norm1", src), _ln(P, f"{prefix}.{tnorm}", tgt),
                      tgt_mask, rpe, n_head)
    src = src + s                                                     # :235
    h = _ln(P, f"{prefix}.norm2", src)
    h = F.relu(F.linear(h, P[f"{prefix}.linear1.weight"], P[f"{prefix}.linear1.bias"]))
    src = src + F.linear(h, P[f"{prefix}.linear2.weight"], P[f"{prefix}.linear2.bias"])  # :236-239
    return src.masked_fill(src_invalid.unsqueeze(-1), 0.0)            # :241-242


def transformer_block(P, prefix: str, mode: str, n_layer: int, n_head: int, src, src_invalid, tgt, tgt_mask, rpe,
                      dec_idx=None, dec_mask=None, dec_rpe=None) -> Tensor:
    """TransformerBlockRPE.forward, transformer_rpe.py:48-135. For enc_self_attn `tgt` is an int64 index
    tensor [B,S,K] re-gathered from the current src every layer (:87-88); for dec_cross_attn `tgt` is the
    pre-gathered feature tensor and `dec_idx` the self-attention neighbour indices (:116-117)."""
    for i in range(n_layer):
        p = f"{prefix}.layers.{i}"
        if mode == "enc_self_attn":
            src = transformer_layer(P, p, mode, n_head, src, src_invalid, _gather_rows(src, tgt), tgt_mask, rpe)
        else:
            src = transformer_layer(P, p, mode, n_head, src, src_invalid, tgt, tgt_mask, rpe,
                                    _gather_rows(src, dec_idx), dec_mask, dec_rpe)
    return src


# --------------------------------------------------------------------------------------------------
# encoders
# --------------------------------------------------------------------------------------------------
def map_encoder(P, cfg, sz, mp_valid, mp_attr, mp_pose) -> Dict[str, Tensor]:
    """MapEncoder.forward, models/map_encoder.py:50-113."""
    d = cfg["hidden_dim"]
    n_sc, n_mp, L = mp_valid.shape
    tok_pose, tok_invalid = mp_pose[:, :, 0], ~mp_valid[:, :, 0]
    xy = to_local_xy(mp_pose[..., :2], tok_pose[:, :, None, :2], tok_pose[..., 2])       # :69-71
    yaw = mp_pose[..., 2] - tok_pose[..., 2:3]                                         # :72 (cast=False)
    pe = encode_polyline(xy, torch.stack([yaw.cos(), yaw.sin()], -1))                  # :73, pose_emb.py:38-41
    attr = torch.cat([mp_attr[:, :, None, :].expand(-1, -1, L, -1),
                      torch.eye(L)[None, None].expand(n_sc, n_mp, -1, -1)], -1)         # :75-77
    feat = torch.cat([mlp(P, "mp_encoder.input_encoder.mlp", attr, (0, 2, 4), False), pe], -1)  # input_encoder.py:57
    tok = pointnet(P, "mp_encoder.pl_encoder", feat, ~mp_valid)                        # :80
    rel_pose, rel_dist = get_rel_pose(tok_pose, tok_invalid)                           # :83
    idx, inv, rpe3 = knn_select(tok_invalid, rel_pose, rel_dist, sz["k_mp2mp"], sz["dl_mp"])  # :88-94
    rpe = pose_emb_xy_yaw(rpe3[..., :2], rpe3[..., 2], d)                              # :97
    tok = transformer_block(P, "mp_encoder.tf_mp2mp", "enc_self_attn", cfg["mp_encoder"]["n_layer_tf"],
                            cfg["tf_cfg"]["n_head"], tok, tok_invalid, idx, inv, rpe)  # :99-105
    return dict(mp_token_invalid=tok_invalid, mp_token_feature=tok, mp_token_pose=tok_pose,
                knn_idx_mp2mp=idx, knn_invalid_mp2mp=inv, rpe3_mp2mp=rpe3)


def tl_pre_compute(P, cfg, sz, tl_valid, tl_attr, tl_pose, mp) -> Dict[str, Tensor]:
    """TrafficLightEncoder.pre_compute, models/traffic_light.py:76-154 (tl_mode lane, HPTR)."""
    d = cfg["hidden_dim"]
    n_sc = tl_valid.shape[0]
    inv = ~tl_valid
    out = dict(tl_token_valid=tl_valid, tl_token_invalid=inv, tl_token_pose=tl_pose)
    out["tl_token_attr"] = mp["mp_token_feature"][torch.arange(n_sc)[:, None], tl_attr]                 # :115
    rp_tt, rd_tt = get_rel_pose(tl_pose, inv)                                                          # :119
    rp_tm, rd_tm = get_rel_pose(tl_pose, inv, mp["mp_token_pose"], mp["mp_token_invalid"])              # :120-122
    out["knn_idx_tl2tl"], out["knn_invalid_tl2tl"], r_tt = knn_select(inv, rp_tt, rd_tt, sz["k_tl2tl"], sz["dl_tl"])
    idx_tm, out["knn_invalid_tl2mp"], r_tm = knn_select(mp["mp_token_invalid"], rp_tm, rd_tm, sz["k_tl2mp"], sz["dl_tl"])
    out["knn_idx_tl2mp"] = idx_tm
    out["knn_tgt_tl2mp"] = _gather_rows(mp["mp_token_feature"], idx_tm)                                 # :146-148
    out["rpe3_tl2tl"], out["rpe3_tl2mp"] = r_tt, r_tm
    out["rpe_tl2tl"] = pose_emb_xy_yaw(r_tt[..., :2], r_tt[..., 2], d)                                  # :150-152
    out["rpe_tl2mp"] = pose_emb_xy_yaw(r_tm[..., :2], r_tm[..., 2], d)
    return out


def tl_forward(P, cfg, tl_state_hist: Tensor, tl: Dict[str, Tensor]) -> Tensor:
    """TrafficLightEncoder.forward, traffic_light.py:184-246. tl_state_hist [B,n_tl,n_step<=11,5] bool."""
    B, n_tl, n_step, _ = tl_state_hist.shape
    W = cfg["temp_window_size"]
    assert n_step <= W                                                                                  # :212
    x = torch.cat([tl_state_hist.float(), torch.eye(W)[None, None, -n_step:].expand(B, n_tl, -1, -1)], -1)  # :223-225
    x = mlp(P, "tl_encoder.input_encoder.mlp", x, (0, 2, 4), False) + tl["tl_token_attr"][:, :, None, :]  # :176-180
    inv = tl["tl_token_invalid"]
    tok = pointnet(P, "tl_encoder.temp_encoder", x, inv[:, :, None].expand(-1, -1, n_step))             # :228
    return transformer_block(P, "tl_encoder.tf_tl2tlmp", "dec_cross_attn", cfg["tl_encoder"]["n_layer_tf"],
                             cfg["tf_cfg"]["n_head"], tok, inv, tl["knn_tgt_tl2mp"], tl["knn_invalid_tl2mp"],
                             tl["rpe_tl2mp"], tl["knn_idx_tl2tl"], tl["knn_invalid_tl2tl"], tl["rpe_tl2tl"])  # :231-240


def tl_state_predictor(P, tl_feat: Tensor, tl_invalid: Tensor) -> Tensor:
    """TrafficLightStatePredictor.forward, traffic_light.py:270-286."""
    x = mlp(P, "tl_state_predictor.mlp", tl_feat, (0, 2, 4), False).masked_fill(tl_invalid.unsqueeze(-1), 0.0)
    return x.clamp(-3, 3)


def ag_encoder(P, cfg, sz, hv, hp, hm, ag_attr, mp, tl, tl_feat, return_knn: bool = False):
    """AgentEncoder._forward_hptr + _get_knn_for_ag, models/agent_encoder.py:114-178, 321-387.
    hv [B,n_ag,n_step] bool, hp/hm [B,n_ag,n_step,3]."""
    d = cfg["hidden_dim"]
    B, n_ag, n_step = hv.shape
    W = cfg["temp_window_size"]
    tok_invalid = ~hv.any(-1)
    tok_pose = last_valid(hp, hv)                                                                       # :132
    mp_inv, tl_inv = mp["mp_token_invalid"], tl["tl_token_invalid"]
    rp_aa, rd_aa = get_rel_pose(tok_pose, tok_invalid)                                                  # :339-345
    rp_am, rd_am = get_rel_pose(tok_pose, tok_invalid, mp["mp_token_pose"], mp_inv)
    rp_at, rd_at = get_rel_pose(tok_pose, tok_invalid, tl["tl_token_pose"], tl_inv)
    i_aa, m_aa, r_aa = knn_select(tok_invalid, rp_aa, rd_aa, sz["k_ag2ag"], sz["dl_ag"])                # :356-379
    i_am, m_am, r_am = knn_select(mp_inv, rp_am, rd_am, sz["k_ag2mp"], sz["dl_ag"])
    i_at, m_at, r_at = knn_select(tl_inv, rp_at, rd_at, sz["k_ag2tl"], sz["dl_ag"])
    tgt = torch.cat([_gather_rows(mp["mp_token_feature"], i_am), _gather_rows(tl_feat, i_at)], 2)       # :371,380,165
    e = lambda r: pose_emb_xy_yaw(r[..., :2], r[..., 2], d)                                             # noqa: E731
    xy = to_local_xy(hp[..., :2], tok_pose[:, :, None, :2], tok_pose[..., 2])                           # :147
    yaw = hp[..., 2] - tok_pose[..., 2:3]                                                               # :148
    attr = torch.cat([ag_attr[:, :, None, :].expand(-1, -1, n_step, -1), hm,
                      torch.eye(W)[None, None, -n_step:].expand(B, n_ag, -1, -1)], -1)                  # :150-157
    feat = torch.cat([mlp(P, "ag_encoder.input_encoder.mlp", attr, (0, 2, 4), False),
                      pose_emb_xy_yaw(xy, yaw, d // 2)], -1)                                            # :159
    tok = pointnet(P, "ag_encoder.temp_encoder", feat, ~hv)                                             # :162
    out = transformer_block(P, "ag_encoder.tf_ag2agmptl", "dec_cross_attn", cfg["ag_encoder"]["n_layer_tf"],
                            cfg["tf_cfg"]["n_head"], tok, tok_invalid, tgt, torch.cat([m_am, m_at], 2),
                            torch.cat([e(r_am), e(r_at)], 2), i_aa, m_aa, e(r_aa))                      # :168-177
    if return_knn:
        return out, dict(idx_aa=i_aa, inv_aa=m_aa, rpe_aa=r_aa, idx_am=i_am, inv_am=m_am, rpe_am=r_am,
                         idx_at=i_at, inv_at=m_at, rpe_at=r_at, tok_pose=tok_pose, tok0=tok)
    return out


def navi_encoder(P, cfg, ag_navi: Tensor, ag_pose: Tensor, mp) -> Tensor:
    """NaviEncoder.forward (dest mode, pairwise_relative), models/navigation.py:65-79."""
    B = ag_navi.shape[0]
    ib = torch.arange(B)[:, None]
    f = mlp(P, "navi_encoder.mlp_mp", mp["mp_token_feature"][ib, ag_navi], (0,), False)
    gp = mp["mp_token_pose"][ib, ag_navi]
    xy = to_local_xy(gp[:, :, None, :2], ag_pose[:, :, None, :2], ag_pose[..., 2]).squeeze(2)
    yaw = gp[..., 2] - ag_pose[..., 2]
    return f + mlp(P, "navi_encoder.mlp_pe", pose_emb_xy_yaw(xy, yaw, cfg["hidden_dim"]), (0,), False)


def navi_predictor(P, cfg, ag_valid: Tensor, ag_attr: Tensor, ag_motion: Tensor, ag_pose: Tensor, mp,
                   ag_type: Tensor, mp_type: Tensor) -> Tensor:
    """NaviPredictor.forward, "dest" mode, HPTR + pairwise_relative (models/navigation.py:175-278): destination
    logits [n_sc, n_ag, n_mp] over the map polylines; the once-per-scene step before the rollout loop
    (waymo_motion.py:469-495). ag_valid [n_sc,n_ag,n_step]; ag_type [n_sc,n_ag,3] / mp_type [n_sc,n_mp,11] bool."""
    d, W = cfg["hidden_dim"], cfg["temp_window_size"]
    n_sc, n_ag, n_step = ag_valid.shape
    tok_valid = ag_valid.any(-1)
    tok_pose = last_valid(ag_pose, ag_valid)                                                            # :204
    if n_step > W:                                                                                      # :210-214
        ag_pose, ag_motion, ag_valid, n_step = ag_pose[:, :, -W:], ag_motion[:, :, -W:], ag_valid[:, :, -W:], W
    xy = to_local_xy(ag_pose[..., :2], tok_pose[:, :, None, :2], tok_pose[..., 2])                      # :218
    yaw = ag_pose[..., 2] - tok_pose[..., 2:3]                                                          # :219
    attr = torch.cat([ag_attr[:, :, None, :].expand(-1, -1, n_step, -1), ag_motion,
                      torch.eye(W)[None, None, -n_step:].expand(n_sc, n_ag, -1, -1)], -1)               # :221-228
    feat = torch.cat([mlp(P, "navi_predictor.input_encoder.mlp", attr, (0, 2, 4), False),
                      pose_emb_xy_yaw(xy, yaw, d // 2)], -1)                                            # :230
    tok = pointnet(P, "navi_predictor.temp_encoder", feat, ~ag_valid)                                   # :232
    n_mp = mp["mp_token_invalid"].shape[1]
    rel, _ = get_rel_pose(tok_pose, ~tok_valid, mp["mp_token_pose"], mp["mp_token_invalid"])            # :259
    x = torch.cat([tok[:, :, None, :].expand(-1, -1, n_mp, -1),
                   mp["mp_token_feature"][:, None].expand(-1, n_ag, -1, -1),
                   pose_emb_xy_yaw(rel[..., :2], rel[..., 2], d)], -1)                                  # :248-261
    for i in (0, 3):                                                                                    # mlp.py:47-51
        x = F.linear(x, P[f"navi_predictor.mlp.fc_layers.{i}.weight"], P[f"navi_predictor.mlp.fc_layers.{i}.bias"])
        x = F.relu(F.layer_norm(x, (d,), P[f"navi_predictor.mlp.fc_layers.{i + 1}.weight"],
                                P[f"navi_predictor.mlp.fc_layers.{i + 1}.bias"]))
    logits = F.linear(x, P["navi_predictor.mlp.fc_layers.6.weight"], P["navi_predictor.mlp.fc_layers.6.bias"]).squeeze(-1)
    return navi_logit_mask(logits, tok_valid, mp["mp_token_invalid"], ag_type, mp_type)


def navi_logit_mask(logits: Tensor, tok_valid: Tensor, mp_invalid: Tensor, ag_type: Tensor, mp_type: Tensor) -> Tensor:
    """Type masks of the destination classifier (navigation.py:265-278): polyline types FREEWAY 0, SURFACE_STREET 1,
    STOP_SIGN 2, BIKE_LANE 3, ROAD_EDGE_BOUNDARY 4 are the only candidates; vehicles exclude 3, pedestrians 0-3,
    cyclists 0-2. Masked logits are -inf; rows of invalid agents or without any candidate are all 0."""
    mp_mask = mp_invalid | ~(mp_type[:, :, :5].any(-1))
    veh = ag_type[:, :, [0]] & mp_type[:, :, 3].unsqueeze(1)
    ped = ag_type[:, :, [1]] & mp_type[:, :, :4].any(-1).unsqueeze(1)
    cyc = ag_type[:, :, [2]] & mp_type[:, :, :3].any(-1).unsqueeze(1)
    inv = mp_mask.unsqueeze(1) | veh | ped | cyc
    logits = logits.masked_fill(inv, -INF)
    return logits.masked_fill((~tok_valid).unsqueeze(-1) | inv.all(-1, keepdim=True), 0.0)


def add_navi_latent(P, prefix: str, x: Tensor, z: Tensor, z_valid: Tensor) -> Tensor:
    """AddNaviLatent.forward (mode cat, res_add True), modules/add_navi_latent.py:33-65."""
    zi = ~z_valid
    z = mlp(P, f"{prefix}.mlp_in", z, (0, 3, 6), True).masked_fill(zi.unsqueeze(-1), 0.0)
    h = mlp(P, f"{prefix}.mlp", torch.cat([x, z], -1), (0, 3, 6), True).masked_fill(zi.unsqueeze(-1), 0.0)
    return h + x


def action_head(P, x: Tensor, valid: Tensor, ag_type: Tensor) -> Tensor:
    """ActionHead.forward (branch_type), modules/action_head.py:64-100 -> action mean [B,n_ag,2]."""
    mean = 0
    for i in range(3):
        m = ~(ag_type[:, :, i] & valid)
        mean = mean + mlp(P, f"action_head.mlp_mean.{i}", x, (0, 2, 4), False).masked_fill(m.unsqueeze(-1), 0.0)
    return mean


class PolicyOracle:
    """TrafficBots.forward with its history ring, models/traffic_bots.py:123-221 (HPTR, eval mode)."""

    def __init__(self, P, cfg, sz):
        self.P, self.cfg, self.sz = P, cfg, sz
        self.init()

    def init(self):                                                                                     # :145-149
        self.hv = self.hp = self.hm = self.ht = None

    def _append(self, v, p, m, t):                                                                      # :123-143
        W = self.cfg["temp_window_size"]
        cat = lambda h, x: x.unsqueeze(2) if h is None else torch.cat([h, x.unsqueeze(2)], 2)[:, :, -W:]  # noqa: E731
        self.hv, self.hp, self.hm, self.ht = cat(self.hv, v), cat(self.hp, p), cat(self.hm, m), cat(self.ht, t)

    def step(self, ag_valid, ag_pose, ag_motion, ag_attr, ag_type, ag_latent, ag_latent_valid, ag_navi,
             ag_navi_valid, tl_state, tl, mp, return_aux: bool = False):
        P, cfg = self.P, self.cfg
        self._append(ag_valid, ag_pose, ag_motion, tl_state)                                            # :188
        navi = navi_encoder(P, cfg, ag_navi, ag_pose, mp)                                               # :191-194
        tl_feat = tl_forward(P, cfg, self.ht, tl)                                                       # :197
        ag_feat = ag_encoder(P, cfg, self.sz, self.hv, self.hp, self.hm, ag_attr, mp, tl, tl_feat)      # :200
        x = add_navi_latent(P, "add_navi", ag_feat, navi, ag_navi_valid)                                # :213
        x = add_navi_latent(P, "add_latent", x, ag_latent, ag_latent_valid)                             # :214
        mean = action_head(P, x, ag_valid, ag_type)                                                     # :217
        logits = tl_state_predictor(P, tl_feat, tl["tl_token_invalid"])                                 # :220
        if return_aux:
            return mean, logits, dict(tl_feat=tl_feat, ag_feat=ag_feat, navi=navi, x=x)
        return mean, logits


# --------------------------------------------------------------------------------------------------
# closed loop: utils/dynamics.py, utils/teacher_forcing.py, utils/traffic_rule_checker.py, pl_modules/waymo_motion.py
# --------------------------------------------------------------------------------------------------
def dynamics_update(pose, motion, valid, ag_type, mean, dyn) -> Tuple[Tensor, Tensor]:
    """Dynamics.update_ag + MultiPathPP.process_action/update, utils/dynamics.py:66-120, 237-274
    (deterministic action, no player override). ag_type bool [B,n_ag,3] order (veh, ped, cyc)."""
    dt = dyn["dt"]
    th = torch.tanh(mean)
    action = 0
    for i, k in enumerate(("veh", "ped", "cyc")):
        a = torch.stack([th[..., 0] * dyn[k]["max_acc"], th[..., 1] * dyn[k]["max_yaw_rate"]], -1)
        action = action + a.masked_fill(~ag_type[:, :, [i]], 0.0)
    action = action.masked_fill(~valid.unsqueeze(-1), 0.0)
    acc, yr = action[..., 0], action[..., 1]
    v_t = motion[..., 0] + 0.5 * dt * acc
    th_t = pose[..., 2] + 0.5 * dt * yr
    new_pose = pose + dt * torch.stack([v_t * th_t.cos(), v_t * th_t.sin(), yr], -1)
    new_motion = torch.stack([motion[..., 0] + dt * acc, acc, yr], -1)
    has_type = ag_type.any(-1, keepdim=True)            # masked sum over the 3 one-hot types (:108-112)
    keep = valid.unsqueeze(-1) & has_type
    return new_pose.masked_fill(~keep, 0.0), new_motion.masked_fill(~keep, 0.0)


def teacher_forcing_mask(gt_valid: Tensor, step_spawn: int, step_warm: int) -> Tensor:
    """TeacherForcing.init, utils/teacher_forcing.py:51-82 (schedules/thresholds off at test time)."""
    tf = torch.zeros_like(gt_valid)
    tf[:, :, 0] |= gt_valid[:, :, 0]
    if step_spawn > 0:
        spawn = (~gt_valid[:, :, :-1]) & gt_valid[:, :, 1:]
        spawn[:, :, step_spawn:] = False
        tf[:, :, 1:] |= spawn
    if step_warm >= 0:
        tf[:, :, : step_warm + 1] |= gt_valid[:, :, : step_warm + 1]
    return tf


def dest_tables(mp_valid, mp_type, mp_pos, mp_dir, ag_dest) -> Dict[str, Tensor]:
    """TrafficRuleChecker._get_dest, utils/traffic_rule_checker.py:86-105."""
    ib = torch.arange(mp_valid.shape[0])[:, None]
    dtype = mp_type[ib, ag_dest]
    ddir = mp_dir[ib, ag_dest]
    ddir = ddir / torch.norm(ddir, dim=-1, keepdim=True)
    thresh = torch.ones(ag_dest.shape) * 50 * (1 - dtype[:, :, 4].float() * 0.8)
    return dict(dest_invalid=~mp_valid[ib, ag_dest], dest_type=dtype, dest_pos=mp_pos[ib, ag_dest], dest_dir=ddir,
                dest_thresh_pos=thresh)


def check_outside_map(valid, pose, boundary) -> Tensor:
    """traffic_rule_checker.py:107-116."""
    x, y = pose[..., 0], pose[..., 1]
    b = boundary
    return ((x > b[:, [1]]) | (x < b[:, [0]]) | (y > b[:, [3]]) | (y < b[:, [2]])) & valid


def check_dest_reached(valid, pose, dest, dest_reached) -> Tensor:
    """traffic_rule_checker.py:291-319."""
    dist = torch.norm(pose[..., None, :2] - dest["dest_pos"], dim=-1).masked_fill(dest["dest_invalid"], INF)
    pos_ok = (dist < dest["dest_thresh_pos"].unsqueeze(-1)).any(-1)
    hf = torch.stack([pose[..., 2].cos(), pose[..., 2].sin()], -1)
    rot = (hf.unsqueeze(2) * dest["dest_dir"]).sum(-1).masked_fill(dest["dest_invalid"], 0.0)
    rot_ok = (rot > math.cos(math.radians(30))).any(-1)
    lane, edge = dest["dest_type"][:, :, :4].any(-1), dest["dest_type"][:, :, 4]
    return (~dest_reached) & valid & ((lane & pos_ok & rot_ok) | (edge & pos_ok))


# --------------------------------------------------------------------------------------------------
# logging-only traffic-rule checks: utils/traffic_rule_checker.py:119-274, utils/wosac_collision.py
# --------------------------------------------------------------------------------------------------
def ag_bbox(pose: Tensor, size: Tensor) -> Tensor:
    """get_ag_bbox, wosac_collision.py:22-48. pose [B,A,3], size [B,A,2] (length, width) -> corners [B,A,4,2] (CCW)."""
    c, s = pose[..., 2].cos(), pose[..., 2].sin()
    f, r = torch.stack([c, s], -1), torch.stack([s, -c], -1)
    of, orr = 0.5 * size[..., [0]] * f, 0.5 * size[..., [1]] * r
    off = torch.stack([of - orr, -of - orr, -of + orr, of + orr], 2)
    return pose[:, :, None, :2] + off


def check_collided(valid, bbox, invalid_pair_mask) -> Tensor:
    """_check_collided, traffic_rule_checker.py:119-149 (separating-axis test on the box edge lines)."""
    nxt = bbox.roll(-1, dims=2)
    line = torch.cat([nxt[..., [1]] - bbox[..., [1]], bbox[..., [0]] - nxt[..., [0]],
                      nxt[..., [0]] * bbox[..., [1]] - nxt[..., [1]] * bbox[..., [0]]], -1)      # [B,A,4,3]
    pt = torch.cat([bbox, torch.ones_like(bbox[..., [0]])], -1)                                  # [B,A,4,3]
    outside = (line[:, :, None, :, None, :] * pt[:, None, :, None, :, :]).sum(-1) > 0            # [B,A,A,4,4]
    no_col = outside.all(-1).any(-1)
    no_col = no_col | no_col.transpose(1, 2)
    no_col = no_col | invalid_pair_mask | ~(valid[:, :, None] & valid[:, None, :])
    return ~no_col.all(-1)


def _signed_dist_origin_to_polygon(poly: Tensor) -> Tensor:
    """_signed_distance_from_point_to_convex_polygon with query = origin, wosac_collision.py:51-113. poly [..., n, 2]."""
    nxt = poly.roll(-1, dims=-2)
    ev = nxt - poly
    el = torch.norm(ev, dim=-1)
    tan = ev / el.unsqueeze(-1)
    nrm = torch.stack([-tan[..., 1], tan[..., 0]], -1)
    v2q = -poly
    vd = torch.norm(v2q, dim=-1)
    sperp = (-nrm * v2q).sum(-1)
    inside = (sperp <= 0).all(-1)
    prop = (tan * v2q).sum(-1) / el
    on_edge = (prop >= 0.0) & (prop <= 1.0)
    ed = torch.where(on_edge, sperp.abs(), torch.zeros_like(sperp) + 1e10)
    md = torch.cat([ed, vd], -1).amin(-1)
    return torch.where(inside, -md, md)


def _minkowski_sum(box1: Tensor, box2: Tensor) -> Tensor:
    """_minkowski_sum_of_box_and_box_points + _get_downmost_edge_in_box, wosac_collision.py:140-192. [..., 4, 2]."""
    def downmost(box):
        i0 = torch.argmin(box[..., 1], dim=-1, keepdim=True)
        st = torch.gather(box, -2, i0[..., None].expand(*i0.shape, 2))
        en = torch.gather(box, -2, ((i0 + 1) % 4)[..., None].expand(*i0.shape, 2))
        e = en - st
        return i0, e / torch.norm(e, dim=-1, keepdim=True)
    o1 = torch.tensor([0, 0, 1, 1, 2, 2, 3, 3])
    o2 = torch.tensor([0, 1, 1, 2, 2, 3, 3, 0])
    s1, d1 = downmost(box1)
    s2, d2 = downmost(box2)
    cond = ((d1[..., 0] * d2[..., 1] - d1[..., 1] * d2[..., 0]) >= 0.0).expand(*s1.shape[:-1], 8)
    i1 = (torch.where(cond, o2, o1) + s1) % 4
    i2 = (torch.where(cond, o1, o2) + s2) % 4
    g = lambda b, i: torch.gather(b, -2, i[..., None].expand(*i.shape, 2))  # noqa: E731
    return g(box1, i1) + g(box2, i2)


def check_collided_wosac(pose, size, valid) -> Tensor:
    """check_collided_wosac, wosac_collision.py:196-239 (Minkowski-difference signed distance of rounded boxes)."""
    B, A, _ = pose.shape
    shrink = torch.minimum(size[:, :, 0], size[:, :, 1]) * 0.7 / 2.0
    corners = ag_bbox(pose, size[:, :, :2] - 2.0 * shrink.unsqueeze(-1))
    ev = corners[:, :, None].expand(-1, -1, A, -1, -1)
    al = corners[:, None].expand(-1, A, -1, -1, -1)
    sd = _signed_dist_origin_to_polygon(_minkowski_sum(ev, -1.0 * al))
    sd = sd - shrink[:, None, :] - shrink[:, :, None]
    bad = ~(valid[:, None, :] & valid[:, :, None]) | torch.eye(A, dtype=torch.bool)[None]
    return sd.masked_fill(bad, 1e10).amin(2) < 0.0


def _ccw(a, b, c):
    return (c[..., 1] - a[..., 1]) * (b[..., 0] - a[..., 0]) > (b[..., 1] - a[..., 1]) * (c[..., 0] - a[..., 0])


def check_run_road_edge(valid, bbox, veh, edge, edge_valid) -> Tensor:
    """_check_run_road_edge, traffic_rule_checker.py:152-173. edge [B,E,2,2], edge_valid [B,E]."""
    nxt = bbox.roll(-1, dims=2)
    a, b = bbox[:, :, None], nxt[:, :, None]                       # [B,A,1,4,2]
    c, d = edge[:, None, :, None, 0], edge[:, None, :, None, 1]   # [B,1,E,1,2]
    hit = (_ccw(a, c, d) != _ccw(b, c, d)) & (_ccw(a, b, c) != _ccw(a, b, d))
    return (hit.any(-1) & edge_valid[:, None]).any(-1) & valid & veh


def check_run_red_light(valid, pose, motion, tl_valid, tl_pose, tl_state, size, veh) -> Tensor:
    """_check_run_red_light, traffic_rule_checker.py:176-218. size = UNscaled ag_size [B,A,>=2]."""
    hc, hs = pose[..., 2].cos(), pose[..., 2].sin()
    hf, hr = torch.stack([hc, hs], -1)[:, :, None], torch.stack([hs, -hc], -1)[:, :, None]
    ln, wd = size[:, :, [0]] * 0.5 * 0.6, size[:, :, [1]] * 0.5 * 1.8
    x0 = pose[:, :, None, :2]
    x1 = x0 + 0.1 * motion[:, :, None, [0]] * hf
    tp = tl_pose[:, None, :, :2]
    ins = lambda x: (((tp - x) * hf).sum(-1).abs() < ln) & (((tp - x) * hr).sum(-1).abs() < wd)  # noqa: E731
    m = (valid & veh)[:, :, None] & (tl_valid & tl_state[:, :, 1])[:, None]
    return (ins(x0) & ~ins(x1) & m).any(-1)


def check_passive(valid, pose, motion, tl_valid, tl_pose, tl_state, lane, lane_valid, veh, counter):
    """_check_passive, traffic_rule_checker.py:221-274. Returns (passive_this_step, new counter)."""
    A = pose.shape[1]
    close = ((torch.norm(pose[:, :, None, :2] - lane[:, None], dim=-1) < 2) & lane_valid[:, None]).any(-1)
    slow = motion[:, :, 0] < 5
    hf = torch.stack([pose[..., 2].cos(), pose[..., 2].sin()], -1)[:, :, None]
    mtl = (tl_valid & tl_state[:, :, [0, 1, 2, 4]].any(-1))[:, None]
    tv = tl_pose[:, None, :, :2] - pose[:, :, None, :2]
    tn = torch.norm(tv, dim=-1)
    red = ((tn < 10) & (((hf * tv).sum(-1) / tn) > 0.95) & mtl).any(-1)
    av = pose[:, None, :, :2] - pose[:, :, None, :2]
    an = torch.norm(av, dim=-1)
    ahead = ((an < 10) & (((hf * av).sum(-1) / an) > 0.95) & valid[:, None] & valid[:, :, None]
             & ~torch.eye(A, dtype=torch.bool)[None]).any(-1)
    p = valid & veh & close & slow & ~red & ~ahead
    counter = (counter + p) * p
    return counter > 20, counter


class RuleCheckOracle:
    """State + per-step evaluation of the five logging-only checks (TrafficRuleChecker.__init__ / check,
    traffic_rule_checker.py:10-84, 343-451). Inputs are the rollout-replicated tensors."""

    def __init__(self, mp_valid, mp_type, mp_pos, mp_dir, ag_type, ag_size, tl_valid, tl_pose, size_scale=1.1):
        self.size_raw = ag_size
        self.size = ag_size[..., :2] * size_scale                                                   # :27
        self.veh = ag_type[:, :, 0]
        A = ag_type.shape[1]
        ped = ag_type[:, :, 1]
        self.pair_invalid = torch.eye(A, dtype=torch.bool)[None] | (ped[:, None] & ped[:, :, None])  # :48-51
        self.edge_valid = (mp_valid & mp_type[:, :, [4, 5, 7]].any(-1, keepdim=True)).flatten(1, 2)  # :476-478
        self.edge = torch.stack([mp_pos, mp_pos + mp_dir], -2).flatten(1, 2)
        self.lane_valid = (mp_valid & mp_type[:, :, :3].any(-1, keepdim=True)).flatten(1, 2)         # :493-495
        self.lane = mp_pos.flatten(1, 2)
        self.tl_valid, self.tl_pose = tl_valid, tl_pose
        self.counter = torch.zeros(ag_type.shape[:2])

    def check(self, valid, pose, motion, tl_state) -> Dict[str, Tensor]:
        bbox = ag_bbox(pose, self.size)
        out = dict(collided=check_collided(valid, bbox, self.pair_invalid),
                   collided_wosac=check_collided_wosac(pose, self.size, valid),
                   run_road_edge=check_run_road_edge(valid, bbox, self.veh, self.edge, self.edge_valid),
                   run_red_light=check_run_red_light(valid, pose, motion, self.tl_valid, self.tl_pose, tl_state,
                                                     self.size_raw, self.veh))
        out["passive"], self.counter = check_passive(valid, pose, motion, self.tl_valid, self.tl_pose, tl_state,
                                                     self.lane, self.lane_valid, self.veh, self.counter)
        return out


def rollout(P, cfg, sz, dyn, rcfg, batch: Dict[str, Tensor], n_rollout: int, step_end: Optional[int] = None,
            mp: Optional[dict] = None, tl: Optional[dict] = None, record=None, rule_checks: bool = False,
            policy=None) -> Dict[str, Tensor]:
    """Restated WOSAC driver: test_step -> joint_future_pred -> rollout -> forward
    (pl_modules/waymo_motion.py:843-876, 439-524, 206-311, 118-204) with the feedback-relevant subset of
    TrafficRuleChecker.check (outside_map, dest_reached; traffic_rule_checker.py:343-451) and fixed
    latent / destination samples. Returns pred_pose [n_sc*R, n_ag, T, 3] etc. (RolloutBuffer.finish)."""
    R = n_rollout
    step_end = rcfg["time_step_end"] if step_end is None else step_end
    if mp is None:
        mp = map_encoder(P, cfg, sz, batch["sc/mp_valid"], batch["sc/mp_attr"], batch["sc/mp_pose"])     # :847
    if tl is None:
        tl = tl_pre_compute(P, cfg, sz, batch["sc/tl_valid"], batch["sc/tl_attr"], batch["sc/tl_pose"], mp)  # :851
    rep = lambda t: t.repeat_interleave(R, 0)                                                           # noqa: E731
    mpR = {k: rep(v) for k, v in mp.items()}                                                            # :458-462
    tlR = {k: rep(v) for k, v in tl.items()}
    gt_valid, gt_pose, gt_motion = rep(batch["sc/ag_valid"]), rep(batch["sc/ag_pose"]), rep(batch["sc/ag_motion"])
    ag_attr, ag_type = rep(batch["sc/ag_attr"]), rep(batch["ref/ag_type"])
    tl_gt = rep(batch["sc/tl_state"])
    n_sc = batch["sc/ag_valid"].shape[0]
    latent = batch["ag_latent"][:, :R].reshape(n_sc * R, *batch["ag_latent"].shape[2:])
    latent_valid = rep(batch["ag_latent_valid"])
    navi, navi_valid = rep(batch["agent/dest"]), rep(batch["ag_navi_valid"]).clone()
    boundary = rep(batch["map/boundary"])
    dest = dest_tables(rep(batch["map/valid"]), rep(batch["map/type"]), rep(batch["map/pos"][..., :2]),
                       rep(batch["map/dir"][..., :2]), navi)
    dest_reached = torch.zeros_like(navi_valid)
    tf = teacher_forcing_mask(gt_valid, rcfg["step_spawn_agent"], rcfg["step_warm_start"])
    n_gt = gt_valid.shape[-1]
    # Dynamics.init, dynamics.py:29-64
    valid, pose, motion = gt_valid[:, :, 0], gt_pose[:, :, 0], gt_motion[:, :, 0]
    disabled = torch.zeros_like(valid)
    tl_state = tl_gt[:, :, 0]
    policy = policy or PolicyOracle(P, cfg, sz)  # tests may plug the CUDA drop-in module into the reference loop
    out = dict(pred_valid=[], pred_pose=[], pred_motion=[], tl_state=[], action_mean=[])
    checker = None
    if rule_checks:
        checker = RuleCheckOracle(rep(batch["map/valid"]), rep(batch["map/type"]), rep(batch["map/pos"][..., :2]),
                                  rep(batch["map/dir"][..., :2]), ag_type, rep(batch["ref/ag_size"]),
                                  rep(batch["sc/tl_valid"]), rep(batch["sc/tl_pose"]))
        out.update({k: [] for k in ("collided", "collided_wosac", "run_road_edge", "run_red_light", "passive")})
    for step in range(1, step_end + 1):                                                                 # :233
        mean, logits = policy.step(valid, pose, motion, ag_attr, ag_type, latent, latent_valid, navi, navi_valid,
                                   tl_state, tlR, mpR)                                                  # :163-177
        if record is not None:
            record(step, dict(valid=valid, pose=pose, motion=motion, tl_state=tl_state, mean=mean, logits=logits,
                              navi_valid=navi_valid))
        pred_valid = valid                                                                              # :183
        pose, motion = dynamics_update(pose, motion, valid, ag_type, mean, dyn)                         # :179
        pred_pose, pred_motion = pose, motion
        if step < n_gt:                                                                                 # teacher_forcing.py:126-147
            ov = tf[:, :, step] & ~disabled                                                             # dynamics.py:135
            valid = valid | ov
            pose = torch.where(ov.unsqueeze(-1), gt_pose[:, :, step], pose)
            motion = torch.where(ov.unsqueeze(-1), gt_motion[:, :, step], motion)
        tl_new = F.one_hot(torch.softmax(logits, -1).argmax(-1), logits.shape[-1]).bool()               # dynamics.py:154-159
        tl_state = tl_gt[:, :, step] if step < n_gt else tl_new                                         # :161-163, tf.py:65,159
        if checker is not None:                                                                         # :250, checker :343-451
            for k, v in checker.check(pred_valid, pred_pose, pred_motion, tl_state).items():
                out[k].append(v)
        outside = check_outside_map(pred_valid, pred_pose, boundary)                                    # :250
        reached = check_dest_reached(pred_valid, pred_pose, dest, dest_reached)
        dest_reached = dest_reached | reached
        out["pred_valid"].append(pred_valid); out["pred_pose"].append(pred_pose)
        out["pred_motion"].append(pred_motion); out["tl_state"].append(tl_state); out["action_mean"].append(mean)
        mask_dis = outside & ~gt_valid[:, :, step] if step < n_gt else outside                          # dynamics.py:176-181
        disabled = disabled | mask_dis
        valid = valid & ~mask_dis
        navi_valid = navi_valid & ~reached                                                              # dynamics.py:195-197
    res = {k: torch.stack(v, 2) for k, v in out.items()}
    res["final_valid"], res["final_navi_valid"] = valid, navi_valid
    return res


# --------------------------------------------------------------------------------------------------
# WOSAC post-processing: data_modules/wosac_post_processing.py:31-75
# --------------------------------------------------------------------------------------------------
def wosac_future_scores(collided: Tensor, run_road_edge: Tensor, ag_role: Tensor, t0: int, w_road_edge: float) -> Tensor:
    """`_filter_futures` violation score (:48-58). Flags [n_sc,K,n_ag,n_step] bool, ag_role [n_sc,n_ag,3] -> [n_sc,K]."""
    role = (ag_role.any(-1) * 1.0).unsqueeze(1)
    col = (collided[..., t0:].any(-1) * role).sum(-1)
    edge = (run_road_edge[..., t0:].any(-1) * role).sum(-1)
    return col + edge * w_road_edge


def wosac_select_futures(scores: Tensor, n_keep: int) -> Tensor:
    """The n_keep smallest scores per scene (:61), made deterministic: ascending (score, index)."""
    K = scores.shape[1]
    key = scores.double() * (K + 1) + torch.arange(K, dtype=torch.float64)  # scores are small non-negative numbers
    return torch.argsort(key, dim=-1, stable=True)[:, :n_keep]


def wosac_to_global(trajs: Tensor, center: Tensor, yaw: Tensor) -> Tuple[Tensor, Tensor]:
    """`forward` (:69-74) with transform_utils.py:160-171 (torch_pos2global), :215-225 (torch_rad2global), :9-11
    (cast_rad). trajs [n_sc,K,n_ag,T,3] scene-centric, center [n_sc,2], yaw [n_sc] -> pos [...,2], yaw [...,1]."""
    c, s = torch.cos(yaw), torch.sin(yaw)
    rot = torch.stack([torch.stack([c, -s], -1), torch.stack([s, c], -1)], -2)  # [n_sc,2,2]
    pos = trajs[..., :2]
    out = torch.matmul(pos.flatten(1, 3), rot.transpose(-1, -2)) + center.unsqueeze(1)
    yy = trajs[..., 2:3]
    oy = (yy.flatten(1, 4) + yaw.unsqueeze(-1) + math.pi) % (2 * math.pi) - math.pi
    return out.view(pos.shape), oy.view(yy.shape)


# --------------------------------------------------------------------------------------------------
# WOMD post-processing: data_modules/womd_post_processing.py:36-106, the configured path of
# configs/model/sim_agent.yaml:170-177 (k_pred 6, use_ade, mpa_nms_thresh [2,2,2]; mtr_nms / traj_aggr off)
# --------------------------------------------------------------------------------------------------
def womd_post_processing(ag_type: Tensor, trajs: Tensor, scores: Optional[Tensor], k_pred: int, use_ade: bool,
                         mpa_nms_thresh, score_temperature: float, track_future_samples: int
                         ) -> Tuple[Tensor, Tensor, Tensor]:
    """`WOMDPostProcessing.forward` (:36-72): softmax of the joint-future log-probs per agent (:48-53), `traj_topk`
    (:170-190) when there are more futures than k_pred, `mpa_nms` (:75-106), optional temperature (:66-67), 2 Hz
    down-sampling (:69). trajs [n_sc,K,n_ag,T,3], scores [n_sc,K,n_ag] or None, ag_type [n_sc,n_ag,3].
    Returns (trajs [n_sc,n_ag,k,n_out,3], scores [n_sc,n_ag,k], mode index [n_sc,n_ag,k]). The reference's top-k is
    `sorted=False` (order unspecified); here modes come in descending score, ties by lower index."""
    tr = trajs.transpose(1, 2)                                       # [n_sc,n_ag,K,T,3]
    n_sc, n_ag, K = tr.shape[:3]
    sc = torch.zeros(n_sc, n_ag, K) if scores is None else scores.transpose(1, 2)
    sc = sc.float().softmax(-1)
    mode = torch.arange(K).expand(n_sc, n_ag, K)
    if K > k_pred:                                                   # traj_topk :170-190
        key = -sc.double() * 4.0 * K + torch.arange(K, dtype=torch.float64) * 1e-12
        mode = torch.argsort(key, dim=-1, stable=True)[..., :k_pred]
        tr = torch.gather(tr, 2, mode[..., None, None].expand(-1, -1, -1, tr.shape[3], 3))
        sc = torch.gather(sc, 2, mode)
        sc = sc / sc.sum(-1, keepdim=True)
    if len(mpa_nms_thresh) > 0:                                      # mpa_nms :75-106
        thresh = sum(ag_type[:, :, i].float() * float(mpa_nms_thresh[i]) for i in range(len(mpa_nms_thresh)))
        xy = tr[..., :2]
        if use_ade:
            dist = torch.norm(xy.unsqueeze(2) - xy.unsqueeze(3), dim=-1).mean(-1)
        else:
            dist = torch.norm(xy[:, :, :, -1].unsqueeze(2) - xy[:, :, :, -1].unsqueeze(3), dim=-1)
        within = dist < thresh[:, :, None, None]
        sc = sc.clone()
        for i in range(n_sc):
            for j in range(n_ag):
                order = sorted(range(sc.shape[-1]), key=lambda k: (-float(sc[i, j, k]), k))
                for k in order:
                    if bool((within[i, j, k] & (sc[i, j] > sc[i, j, k])).any()):
                        sc[i, j, k] = 1e-3
        sc = sc / sc.sum(-1, keepdim=True)
    if score_temperature > 0:
        sc = torch.softmax(torch.log(sc) / score_temperature, dim=-1)
    return tr[:, :, :, 4:track_future_samples:5], sc, mode
