"""CPU ORACLE (test infrastructure, NOT product code) of the training step: a torch-CPU fp32 restatement of
`WaymoMotion.training_step` (src/pl_modules/waymo_motion.py:313-385) in the reference's operation order, built on the
forward oracle (`oracle/tb_oracle.py`); gradients come from torch autograd THROUGH this restatement and are what the
CUDA training path (trafficbotsv1.5_b200/training.py) is compared with.

  latent posterior / prior        models/latent_encoder.py:56-122, 222-233
  reactive_replay -> rollout      pl_modules/waymo_motion.py:386-437, 206-311 (model inputs detached :158-161)
  DifferentiableReward.get        utils/rewards.py:35-85 (imitation terms; w_collision = 0)
  TrainingMetrics.update/compute  models/metrics/training.py:76-189, BalancedKL models/metrics/loss.py:40-76
  TeacherForcing.init (training)  utils/teacher_forcing.py:51-92

Parity pin: the reference ships no test for this path; the pin is `tests/golden/train_small.pt`, produced by
`tests/golden/make_golden.py train`, which evaluates the same training_step body with the REAL reference modules
(TrafficBots incl. LatentEncoder / NaviPredictor, Dynamics, TeacherForcing, RolloutBuffer, DifferentiableReward,
TrainingMetrics / BalancedKL; only the Lightning wrapper is restated) and torch autograd: every loss term and two
statistics of every one of the 656 gradient tensors (tests/test_oracle_golden.py, `-m "not gpu"`).
Dropout is off (p = 0 configuration); the Bernoulli draws of a step are inputs. Only tests/ and bench.py's CPU /
eager baseline legs may import this module.
"""
import copy
from typing import Dict, Optional

import torch
from torch import Tensor
import torch.nn.functional as F

from . import tb_oracle as O


def remap(P: Dict[str, Tensor], mapping: Dict[str, str]) -> Dict[str, Tensor]:
    out = {}
    for k, v in P.items():
        for src, dst in mapping.items():
            if k.startswith(src):
                out[dst + k[len(src):]] = v
    return out


def latent_posterior(P, cfg, sz, batch, mp):
    """LatentEncoder.forward(posterior=True), latent_encoder.py:56-122 + DistEncoder diag_gaus :222-233.
    Returns (mean [n_sc, n_ag, latent_dim], valid [n_sc, n_ag])."""
    rate = cfg["latent_encoder"]["temporal_down_sample_rate"]
    Pp = remap(P, {"latent_encoder.tl_encoder_post.": "tl_encoder.", "latent_encoder.ag_encoder_post.": "ag_encoder.",
                   "latent_encoder.latent_dist_post.": "latent_dist."})
    cfg_l = copy.deepcopy(cfg)
    n = cfg["time_step_gt"] + 1
    cfg_l["temp_window_size"] = n // rate + 1 if rate > 1 else n                                         # :33-37
    v, po, mo = (batch[k][:, :, ::rate] for k in ("gt/ag_valid", "gt/ag_pose", "gt/ag_motion"))          # :93-99
    tls = batch["gt/tl_state"][:, :, ::rate]
    mp_det = dict(mp, mp_token_feature=mp["mp_token_feature"].detach())                                  # traffic_light.py:113-115
    tl = O.tl_pre_compute(Pp, cfg_l, sz, batch["sc/tl_valid"], batch["sc/tl_attr"], batch["sc/tl_pose"], mp_det)
    tl_feat = O.tl_forward(Pp, cfg_l, tls, tl)                                                           # :107
    x = O.ag_encoder(Pp, cfg_l, sz, v, po, mo, batch["sc/ag_attr"], mp, tl, tl_feat)                     # :108-119
    valid = v.any(-1)
    mean = O.mlp(Pp, "latent_dist.mlp_mean", x, (0, 2, 4), False).masked_fill(~valid.unsqueeze(-1), 0.0)  # :231-232
    return mean, valid


def teacher_forcing_mask_training(gt_valid: Tensor, step_spawn: int, step_warm: int,
                                  forcing_agent: Optional[Tensor]) -> Tensor:
    """TeacherForcing.init with the training schedule (teacher_forcing.py:51-92): the inference mask plus whole
    trajectories of the agents drawn with prob_forcing_agent."""
    tf = O.teacher_forcing_mask(gt_valid, step_spawn, step_warm)
    if forcing_agent is not None:
        tf = tf | (forcing_agent.unsqueeze(-1) & gt_valid)
    return tf


def training_step(P: Dict[str, Tensor], cfg: dict, sz: dict, dyn: dict, tc: dict, batch: Dict[str, Tensor],
                  n_steps: Optional[int] = None) -> Dict[str, Tensor]:
    """Loss terms of one training step, differentiable w.r.t. the tensors of P."""
    T = n_steps or tc["time_step_end"]
    mp = O.map_encoder(P, cfg, sz, batch["sc/mp_valid"], batch["sc/mp_attr"], batch["sc/mp_pose"])       # :317-319
    mp_det = dict(mp, mp_token_feature=mp["mp_token_feature"].detach())
    tl = O.tl_pre_compute(P, cfg, sz, batch["sc/tl_valid"], batch["sc/tl_attr"], batch["sc/tl_pose"], mp_det)  # :321-324
    gt_valid, gt_pose, gt_motion = batch["gt/ag_valid"], batch["gt/ag_pose"], batch["gt/ag_motion"]
    n_sc, n_ag, n_gt = gt_valid.shape
    eps = batch["ag_latent_eps"]
    has_post = any(k.startswith("latent_encoder.ag_encoder_post.") for k in P)
    mu = log_std = None
    latent, latent_valid = eps, gt_valid.any(-1)
    post_valid = latent_valid
    if has_post:                                                                                        # :326-350
        mu, post_valid = latent_posterior(P, cfg, sz, batch, mp)
        log_std = P["latent_encoder.latent_dist_post.log_std"]
        if bool(batch.get("rollout_prior", False)):
            latent, latent_valid = eps, batch["sc/ag_valid"].any(-1)
        else:
            latent, latent_valid = mu + log_std.exp() * eps, post_valid                                  # rsample
    navi_logits = None
    if any(k.startswith("navi_predictor.") for k in P):                                                 # :352-359
        navi_logits = O.navi_predictor(P, cfg, batch["sc/ag_valid"], batch["sc/ag_attr"], batch["sc/ag_motion"],
                                       batch["sc/ag_pose"], mp_det, batch["ref/ag_type"], batch["ref/mp_type"])
    # ---- rollout (:206-311) with the training teacher forcing
    ag_attr, ag_type = batch["sc/ag_attr"], batch["ref/ag_type"]
    tl_gt = batch["gt/tl_state"]
    navi, navi_valid = batch["gt/ag_navi"], gt_valid.any(-1).clone()
    dest = O.dest_tables(batch["map/valid"], batch["map/type"], batch["map/pos"][..., :2], batch["map/dir"][..., :2], navi)
    dest_reached = torch.zeros_like(navi_valid)
    tf = teacher_forcing_mask_training(gt_valid, tc["step_spawn_agent"], tc["step_warm_start"],
                                       batch.get("tf/forcing_agent"))
    valid, pose, motion = gt_valid[:, :, 0], gt_pose[:, :, 0], gt_motion[:, :, 0]
    disabled = torch.zeros_like(valid)
    tl_state = tl_gt[:, :, 0]
    pol = O.PolicyOracle(P, cfg, sz)
    ib = torch.arange(n_sc)[:, None]
    navi_feat = O.mlp(P, "navi_encoder.mlp_mp", mp_det["mp_token_feature"][ib, navi], (0,), False)       # navigation.py:66-71
    navi_pose = mp["mp_token_pose"][ib, navi]
    t0 = tc["step_training_start"]
    rew_sum, rew_cnt = 0.0, 0.0
    nll_sum, nll_cnt = 0.0, 0.0
    pred_valid_l, pred_pose_l = [], []
    for step in range(1, T + 1):
        p_in, m_in = pose.detach(), motion.detach()                                                      # :158-161
        pol._append(valid, p_in, m_in, tl_state)
        xy = O.to_local_xy(navi_pose[:, :, None, :2], p_in[:, :, None, :2], p_in[..., 2]).squeeze(2)
        nv = navi_feat + O.mlp(P, "navi_encoder.mlp_pe", O.pose_emb_xy_yaw(xy, navi_pose[..., 2] - p_in[..., 2],
                                                                          cfg["hidden_dim"]), (0,), False)
        tl_feat = O.tl_forward(P, cfg, pol.ht, tl)
        ag_feat = O.ag_encoder(P, cfg, sz, pol.hv, pol.hp, pol.hm, ag_attr, mp, tl, tl_feat)
        x = O.add_navi_latent(P, "add_navi", ag_feat, nv, navi_valid)
        x = O.add_navi_latent(P, "add_latent", x, latent, latent_valid)
        mean = O.action_head(P, x, valid, ag_type)
        logits = O.tl_state_predictor(P, tl_feat.detach(), tl["tl_token_invalid"])                       # traffic_light.py:279-286
        pred_valid = valid
        pose, motion = O.dynamics_update(pose, motion, valid, ag_type, mean, dyn)                        # gradient path
        pred_pose, pred_motion = pose, motion
        # ! diffbar reward (rewards.py:62-76) and its masks (training.py:93-138)
        if step - 1 >= t0:
            if step < n_gt:
                ok = pred_valid & gt_valid[:, :, step]
                e_pos = F.smooth_l1_loss(gt_pose[:, :, step, :2], pred_pose[..., :2], reduction="none").sum(-1)
                e_rot = 0.5 * (1 - torch.cos(gt_pose[:, :, step, 2] - pred_pose[..., 2]))
                e_spd = F.smooth_l1_loss(gt_motion[:, :, step, 0], pred_motion[..., 0], reduction="none")
                r = -(tc["w_pos"] * e_pos + tc["w_rot"] * e_rot + tc["w_spd"] * e_spd)
                rew_sum = rew_sum + r.masked_fill(~ok, 0.0).sum()
                rew_cnt = rew_cnt + ok.sum()
            else:
                rew_cnt = rew_cnt + pred_valid.sum()
        if step < n_gt:                                                                                  # :270-277
            gt_idx = tl_gt[:, :, step].max(-1)[1]
            nll = -torch.log_softmax(logits, -1).gather(-1, gt_idx.unsqueeze(-1)).squeeze(-1)
            tv = ~tl["tl_token_invalid"]
            nll_sum = nll_sum + nll.masked_fill(~tv, 0.0).sum()
            nll_cnt = nll_cnt + tv.sum()
        if step < n_gt:                                                                                  # teacher forcing
            ov = tf[:, :, step] & ~disabled
            valid = valid | ov
            pose = torch.where(ov.unsqueeze(-1), gt_pose[:, :, step], pose)
            motion = torch.where(ov.unsqueeze(-1), gt_motion[:, :, step], motion)
        tl_new = F.one_hot(torch.softmax(logits, -1).argmax(-1), logits.shape[-1]).bool()
        tl_state = tl_gt[:, :, step] if step < n_gt else tl_new
        outside = O.check_outside_map(pred_valid, pred_pose.detach(), batch["map/boundary"])
        reached = O.check_dest_reached(pred_valid, pred_pose.detach(), dest, dest_reached)
        dest_reached = dest_reached | reached
        pred_valid_l.append(pred_valid)
        pred_pose_l.append(pred_pose.detach())
        mask_dis = outside & ~gt_valid[:, :, step] if step < n_gt else outside
        disabled = disabled | mask_dis
        valid = valid & ~mask_dis
        navi_valid = navi_valid & ~reached
    out = {}
    loss = torch.zeros(())
    if float(rew_cnt) > 0:
        out["diffbar_reward"] = tc["w_diffbar_reward"] * rew_sum / rew_cnt
        loss = loss - out["diffbar_reward"]
    if float(nll_cnt) > 0:
        out["tl_state_loss"] = tc["w_tl_state"] * nll_sum / nll_cnt
        loss = loss + out["tl_state_loss"]
    pv = torch.stack(pred_valid_l, 2)
    loss_any = pv[:, :, t0:].any(-1)
    if mu is not None:                                                                                   # training.py:108-121
        from torch.distributions import Independent, Normal, kl_divergence
        post = Independent(Normal(mu, log_std.exp().expand_as(mu)), 1)
        prior = Independent(Normal(torch.zeros_like(mu), torch.ones_like(mu)), 1)
        d_post = Independent(Normal(mu.detach(), log_std.exp().expand_as(mu).detach()), 1)
        e0 = kl_divergence(d_post, prior).clamp(min=tc["kl_free_nats"])                                  # loss.py:67-72
        e1 = kl_divergence(post, prior).clamp(min=tc["kl_free_nats"])
        err = e0 + tc["kl_balance_scale"] * e1
        kv = post_valid & loss_any
        if int(kv.sum()) > 0:
            out["vae_kl"] = tc["w_vae_kl"] * err.masked_fill(~kv, 0.0).sum() / kv.sum()
            loss = loss + out["vae_kl"]
    if navi_logits is not None:                                                                          # training.py:146-153
        nvv = batch["sc/ag_valid"].any(-1) & loss_any
        nl = -torch.distributions.Categorical(logits=navi_logits).log_prob(navi)
        if int(nvv.sum()) > 0:
            out["navi_loss"] = tc["w_navi"] * nl.masked_fill(~nvv, 0.0).sum() / nvv.sum()
            loss = loss + out["navi_loss"]
    out["loss"] = loss
    out["pred_valid"], out["pred_pose"] = pv, torch.stack(pred_pose_l, 2)
    return out
