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


def rollout(P, cfg, sz, dyn, rcfg, batch: Dict[str, Tensor], n_rollout: int, step_end: Optional[int] = None,
            mp: Optional[dict] = None, tl: Optional[dict] = None, record=None) -> Dict[str, Tensor]:
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
    policy = PolicyOracle(P, cfg, sz)
    out = dict(pred_valid=[], pred_pose=[], pred_motion=[], tl_state=[], action_mean=[])
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
