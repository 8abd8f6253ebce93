"""Golden-vector generator (run ONLY in the build container, where /root/reference exists):

    python tests/golden/make_golden.py

Imports the UNMODIFIED reference modules (src/models/*, src/utils/*) with three stub packages standing in
for omegaconf / hydra / transforms3d (tests/golden/_stubs), loads the seeded weights of
`params.init_params` into the reference `TrafficBots` (checking that the key set and shapes match the
reference's own state_dict), runs the reference on the seeded synthetic inputs of `synth.make_scene_batch`
and stores ONLY the reference's outputs (+ weight/input checksums) as small fixtures in tests/golden/*.pt.
The PL module (pl_modules/waymo_motion.py) is not importable here (needs pytorch_lightning, tensorflow,
waymo_open_dataset), so its loop body (:118-204, :206-311, :439-524) is restated around the real
reference Dynamics / TeacherForcing / TrafficRuleChecker / TrafficBots objects.
"""
import hashlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, "/root/reference/src")

import tbpkg  # noqa: E402,F401
from trafficbotsv1_5_b200 import config, params, synth  # noqa: E402
from omegaconf import DictConfig  # noqa: E402  (stub)

from models.traffic_bots import TrafficBots  # noqa: E402  (reference)
from models.modules.attention_rpe import AttentionRPE  # noqa: E402
from models.modules.transformer_rpe import TransformerBlockRPE  # noqa: E402
from utils.rpe import get_rel_pose, get_tgt_knn_idx  # noqa: E402
from utils.pose_emb import PoseEmb  # noqa: E402
from utils.dynamics import Dynamics  # noqa: E402
from utils.teacher_forcing import TeacherForcing  # noqa: E402
from utils.traffic_rule_checker import TrafficRuleChecker  # noqa: E402


def checksum(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().cpu().numpy().tobytes()).hexdigest()[:16]


def build_reference_model(cfg, P):
    model = TrafficBots(**DictConfig(cfg))
    sd = model.state_dict()
    hot = {k: v for k, v in sd.items() if not k.startswith(("latent_encoder.", "navi_predictor."))}
    bufs = {k for k in sd if k.endswith((".freqs", "pl_node_ohe", "hist_ohe"))}  # persistent buffers (App. B)
    hot_params = {k: tuple(v.shape) for k, v in hot.items() if k not in bufs}
    mine = {k: tuple(v) for k, v in params.param_shapes(cfg).items()}
    assert hot_params == mine, (set(hot_params) ^ set(mine))
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected, unexpected
    assert all(m.startswith(("latent_encoder.", "navi_predictor.")) or m in bufs for m in missing), missing
    model.eval()
    return model


def tie_free_poses(g, B, N, lo, hi):
    """positions on a jittered lattice; callers assert the K/K+1 distance gap."""
    xy = torch.rand(B, N, 2, generator=g) * (hi - lo) + lo
    yaw = (torch.rand(B, N, 1, generator=g) * 2 - 1) * 3.0
    return torch.cat([xy, yaw], -1)


@torch.no_grad()
def golden_ops():
    out = {}
    g = torch.Generator().manual_seed(7)
    # ---- rel-pose + knn (utils/rpe.py)
    for name, (B, S, T, K, lim) in dict(small=(2, 24, 50, 12, 60.0), self_=(2, 40, 40, 24, 250.0),
                                        big=(1, 16, 1024, 64, 500.0)).items():
        pose = tie_free_poses(g, B, S, -80, 80)
        inv = torch.rand(B, S, generator=g) < 0.15
        if name == "self_":
            pose2, inv2 = None, None
        else:
            pose2 = tie_free_poses(g, B, T, -120, 120)
            inv2 = torch.rand(B, T, generator=g) < 0.2
        rel_pose, rel_dist = get_rel_pose(pose, inv, pose2, inv2)
        idx, knn_inv, rpe = get_tgt_knn_idx(inv if inv2 is None else inv2, rel_pose, rel_dist, K, lim)
        # tie-free check: gap between K-th and (K+1)-th finite distance must be comfortable
        sd = torch.sort(rel_dist, -1)[0]
        gap = (sd[..., K] - sd[..., K - 1])
        fin = torch.isfinite(sd[..., K])
        assert (gap[fin] > 1e-3).all(), gap[fin].min()
        order = torch.argsort(torch.gather(rel_dist, 2, idx), dim=-1, stable=True)
        srt = lambda t: torch.gather(t, 2, order if t.dim() == 3 else order[..., None].expand(-1, -1, -1, 3))  # noqa
        out[f"knn_{name}"] = dict(pose=pose, inv=inv, pose2=pose2, inv2=inv2, K=K, lim=lim,
                                  idx=srt(idx), knn_inv=srt(knn_inv), rpe=srt(rpe),
                                  dist=srt(torch.gather(rel_dist, 2, idx)))
    # ---- pose embedding (utils/pose_emb.py)
    for pe in (64, 128, 256):
        emb = PoseEmb("pe_xy_yaw", pe_dim=pe, theta_xy=1e3)
        xy = (torch.rand(5, 7, 2, generator=g) * 2 - 1) * 300
        yaw = (torch.rand(5, 7, 1, generator=g) * 2 - 1) * 9
        out[f"pe_{pe}"] = dict(xy=xy, yaw=yaw, emb=emb(xy, yaw))
    # ---- AttentionRPE / TransformerBlockRPE (modules/attention_rpe.py, transformer_rpe.py)
    for name, (d, H, B, S, K) in dict(d128=(128, 4, 2, 16, 12), d256=(256, 4, 1, 8, 36)).items():
        att = AttentionRPE(d, H, dropout_p=0.1, bias=True, d_rpe=d).eval()
        att.load_state_dict(params.rand_like_state_dict(att.state_dict(), seed=11))
        src = torch.randn(B, S, d, generator=g)
        tgt = torch.randn(B, S, K, d, generator=g)
        mask = torch.rand(B, S, K, generator=g) < 0.2
        mask[0, 0] = True  # one all-masked row
        rel = torch.cat([(torch.rand(B, S, K, 2, generator=g) * 2 - 1) * 100,
                         (torch.rand(B, S, K, 1, generator=g) * 2 - 1) * 4], -1)
        rpe = PoseEmb("pe_xy_yaw", pe_dim=d, theta_xy=1e3)(rel[..., :2], rel[..., 2:3])
        o, _ = att(src, tgt, tgt_padding_mask=mask, rpe=rpe)
        out[f"attn_{name}"] = dict(sd_shapes={k: tuple(v.shape) for k, v in att.state_dict().items()}, sd_seed=11,
                                   src=src, tgt=tgt, mask=mask, rel=rel, out=o, n_head=H)
    for mode in ("enc_self_attn", "dec_cross_attn"):
        d, H, B, S, K, T2, K2 = 128, 4, 2, 20, 8, 30, 10
        blk = TransformerBlockRPE(d_model=d, n_head=H, k_feedforward=4, dropout_p=0.1, bias=True, activation="relu",
                                  out_layernorm=False, apply_q_rpe=False, n_layer=2, mode=mode, d_rpe=d).eval()
        blk.load_state_dict(params.rand_like_state_dict(blk.state_dict(), seed=13))
        pe = PoseEmb("pe_xy_yaw", pe_dim=d, theta_xy=1e3)
        src = torch.randn(B, S, d, generator=g)
        src_inv = torch.rand(B, S, generator=g) < 0.15
        idx = torch.stack([torch.stack([torch.randperm(S, generator=g)[:K] for _ in range(S)]) for _ in range(B)])
        m1 = torch.rand(B, S, K, generator=g) < 0.2
        rel1 = torch.cat([(torch.rand(B, S, K, 2, generator=g) * 2 - 1) * 100,
                          (torch.rand(B, S, K, 1, generator=g) * 2 - 1) * 4], -1)
        rec = dict(sd_shapes={k: tuple(v.shape) for k, v in blk.state_dict().items()}, sd_seed=13, src=src, src_inv=src_inv, idx=idx, m1=m1,
                   rel1=rel1, n_head=H, n_layer=2)
        if mode == "enc_self_attn":
            o, _ = blk(src=src, src_padding_mask=src_inv, tgt=idx, tgt_padding_mask=m1,
                       rpe=pe(rel1[..., :2], rel1[..., 2:3]))
        else:
            tgt_tab = torch.randn(B, T2, d, generator=g)
            idx2 = torch.stack([torch.stack([torch.randperm(T2, generator=g)[:K2] for _ in range(S)]) for _ in range(B)])
            m2 = torch.rand(B, S, K2, generator=g) < 0.2
            rel2 = torch.cat([(torch.rand(B, S, K2, 2, generator=g) * 2 - 1) * 100,
                              (torch.rand(B, S, K2, 1, generator=g) * 2 - 1) * 4], -1)
            tgt = tgt_tab[torch.arange(B)[:, None, None], idx2]
            o, _ = blk(src=src, src_padding_mask=src_inv, tgt=tgt, tgt_padding_mask=m2,
                       rpe=pe(rel2[..., :2], rel2[..., 2:3]), decoder_tgt=idx, decoder_tgt_padding_mask=m1,
                       decoder_rpe=pe(rel1[..., :2], rel1[..., 2:3]))
            rec.update(tgt_tab=tgt_tab, idx2=idx2, m2=m2, rel2=rel2)
        rec["out"] = o
        out[f"block_{mode}"] = rec
    return out


def make_dynamics():
    dc = config.DYNAMICS_CFG
    mk = lambda k: DictConfig(dict(_target_="utils.dynamics.MultiPathPP", **dc[k]))  # noqa: E731
    return Dynamics(veh=mk("veh"), ped=mk("ped"), cyc=mk("cyc"), navi_mode="dest")


@torch.no_grad()
def reference_rollout(model, batch, R, step_end, record_steps=(), disable_check=True):
    """waymo_motion.py:439-524 (joint_future_pred) + :206-311 (rollout) + :118-204 (forward), restated."""
    rc = config.ROLLOUT_CFG
    mp_tokens = model.mp_encoder(batch["sc/mp_valid"], batch["sc/mp_attr"], batch["sc/mp_pose"], batch["ref/mp_type"])
    tl_tokens = model.tl_encoder.pre_compute(tl_valid=batch["sc/tl_valid"], tl_attr=batch["sc/tl_attr"],
                                             tl_pose=batch["sc/tl_pose"], **mp_tokens)
    static = dict(mp={k: v.clone() for k, v in mp_tokens.items()},
                  tl={k: v.clone() for k, v in tl_tokens.items() if v is not None})
    rep = lambda t: t.repeat_interleave(R, 0)  # noqa: E731
    n_sc = batch["sc/ag_valid"].shape[0]
    ag_tokens = dict(ag_type=rep(batch["ref/ag_type"]), ag_size=rep(batch["ref/ag_size"]), ag_attr=rep(batch["sc/ag_attr"]),
                     gt_valid=rep(batch["sc/ag_valid"]), gt_pose=rep(batch["sc/ag_pose"]), gt_motion=rep(batch["sc/ag_motion"]))
    mp_tokens = {k: rep(v) for k, v in mp_tokens.items()}
    tl_tokens = {k: (rep(v) if v is not None else None) for k, v in tl_tokens.items()}
    ag_tokens["ag_latent"] = batch["ag_latent"][:, :R].reshape(n_sc * R, *batch["ag_latent"].shape[2:])
    ag_tokens["ag_latent_valid"] = rep(batch["ag_latent_valid"])
    ag_tokens["ag_navi"] = rep(batch["agent/dest"])
    ag_tokens["ag_navi_valid"] = rep(batch["ag_navi_valid"])
    rule_checker = TrafficRuleChecker(
        mp_boundary=rep(batch["map/boundary"]), mp_valid=rep(batch["map/valid"]), mp_type=rep(batch["map/type"]),
        mp_pos=rep(batch["map/pos"]), mp_dir=rep(batch["map/dir"]), ag_type=ag_tokens["ag_type"],
        ag_size=ag_tokens["ag_size"], ag_goal=None, ag_dest=ag_tokens["ag_navi"], tl_valid=tl_tokens["tl_token_valid"],
        tl_pose=tl_tokens["tl_token_pose"], disable_check=disable_check)  # True: logging-only checks off
    tl_state_gt = rep(batch["sc/tl_state"])
    tf = TeacherForcing(step_spawn_agent=rc["step_spawn_agent"], step_warm_start=rc["step_warm_start"])
    tf.init(ag_valid=ag_tokens["gt_valid"], ag_pose=ag_tokens["gt_pose"], ag_motion=ag_tokens["gt_motion"],
            tl_state=tl_state_gt, current_epoch=0)
    dyn = make_dynamics()
    dyn.init(tl_state=tl_state_gt, **ag_tokens)
    model.init()
    out = dict(pred_valid=[], pred_pose=[], pred_motion=[], tl_state=[], action_mean=[])
    vio_keys = ("collided", "collided_wosac", "run_road_edge", "run_red_light", "passive")
    if not disable_check:
        out.update({k: [] for k in vio_keys})
    rec = {}
    for step in range(1, step_end + 1):
        ag_override, tl_override = tf.get(step, dyn.ag_valid, dyn.ag_pose, dyn.ag_motion)
        ag_valid = dyn.ag_valid
        action_dist, tl_dist = model(
            ag_valid=ag_valid, ag_pose=dyn.ag_pose, ag_motion=dyn.ag_motion, ag_attr=dyn.ag_attr, ag_type=dyn.ag_type,
            ag_latent=dyn.ag_latent, ag_latent_valid=dyn.ag_latent_valid, ag_navi=dyn.ag_navi,
            ag_navi_valid=dyn.ag_navi_valid, ag_navi_updated=dyn.ag_navi_updated, tl_state=dyn.tl_state,
            tl_tokens=tl_tokens, mp_tokens=mp_tokens)
        if step in record_steps:
            rec[step] = dict(mean=action_dist.mean.clone(), logits=tl_dist.logits.clone(),
                             valid=ag_valid.clone(), pose=dyn.ag_pose.clone())
        dyn.update_ag(action_dist, True, None)
        pred_pose, pred_motion = dyn.ag_pose, dyn.ag_motion
        dyn.override_ag(ag_override)
        dyn.override_tl(tl_dist, tl_override)
        violation = rule_checker.check(ag_valid, pred_pose, pred_motion, dyn.tl_state)
        gt_valid = ag_tokens["gt_valid"][:, :, step] if step < ag_tokens["gt_valid"].shape[-1] else None
        out["pred_valid"].append(ag_valid); out["pred_pose"].append(pred_pose); out["pred_motion"].append(pred_motion)
        out["tl_state"].append(dyn.tl_state); out["action_mean"].append(action_dist.mean)
        if not disable_check:
            for k in vio_keys:
                out[k].append(violation[f"{k}_this_step"])
        dyn.disable_ag(violation, gt_valid)
        dyn.disable_navi(violation)
    res = {k: torch.stack(v, 2) for k, v in out.items()}
    res["final_valid"], res["final_navi_valid"] = dyn.ag_valid, dyn.ag_navi_valid
    return res, static, rec


@torch.no_grad()
def golden_navi_predictor():
    """the reference NaviPredictor ("dest" mode, navigation.py:175-278) on the reference's own map
    tokens -> destination probabilities (DestCategorical.probs, distributions.py:123-137)."""
    cfg = config.default_model_cfg()
    P = params.init_params(cfg, seed=0, with_navi_predictor=True)
    model = build_reference_model(cfg, P)
    sd = model.state_dict()
    mine = {k: tuple(v) for k, v in params.navi_predictor_shapes(cfg).items()}
    ref = {k: tuple(v.shape) for k, v in sd.items() if k.startswith("navi_predictor.")
           and not k.endswith((".freqs", "hist_ohe"))}
    assert ref == mine, set(ref) ^ set(mine)
    assert all(torch.equal(sd[k], P[k]) for k in mine)
    shape = dict(n_sc=2, n_ag=32, n_mp=96, n_tl=30, seed=5000, boundary=105.0)
    batch = synth.make_scene_batch(**shape)
    mp = model.mp_encoder(batch["sc/mp_valid"], batch["sc/mp_attr"], batch["sc/mp_pose"], batch["ref/mp_type"])
    dist = model.navi_predictor(ag_valid=batch["sc/ag_valid"], ag_attr=batch["sc/ag_attr"], ag_motion=batch["sc/ag_motion"],
                                ag_pose=batch["sc/ag_pose"], mp_token_invalid=mp["mp_token_invalid"],
                                mp_token_feature=mp["mp_token_feature"], mp_token_pose=mp["mp_token_pose"],
                                ag_type=batch["ref/ag_type"], mp_token_type=mp["mp_token_type"])
    fix = dict(shape=shape, param_seed=0, probs=dist.probs, valid=dist.valid, dest_argmax=dist.sample(True),
               input_checksum=checksum(torch.cat([batch[k].float().flatten() for k in sorted(batch)])))
    torch.save(fix, os.path.join(HERE, "navi_pred.pt"))
    print("navi_pred.pt", os.path.getsize(os.path.join(HERE, "navi_pred.pt")) // 1024, "KiB; candidates per agent:",
          float((dist.probs > 0).sum(-1).float().mean()))


@torch.no_grad()
def golden_wosac_post():
    """the real WOSACPostProcessing (`_filter_futures` + the local -> global part of `forward`,
    wosac_post_processing.py:31-75) on seeded joint futures; the Waymo proto imports are stubbed (never executed)."""
    from types import SimpleNamespace
    from data_modules.wosac_post_processing import WOSACPostProcessing  # reference
    shape = dict(seed=7000, n_sc=2, K=40, A=40, T=12)
    t0, w = 4, 0.01
    inp = synth.make_wosac_post_inputs(**shape)
    n_sc, K, A, T = shape["n_sc"], shape["K"], shape["A"], shape["T"]
    pp = WOSACPostProcessing(step_gt=90, step_current=10, const_vel_z_sim=True, const_vel_no_sim=True, w_road_edge=w,
                             use_wosac_col=True)
    buffer = SimpleNamespace(pred_pose=inp["pose"].view(n_sc, K, A, T, 3), step_future_start=t0,
                             violation=dict(collided_wosac=inp["collided"], run_road_edge=inp["run_road_edge"]))
    z = torch.zeros
    batch = {"ref/ag_role": inp["role"], "scenario_center": inp["center"], "scenario_yaw": inp["yaw"],
             "history/agent_no_sim/pos": z(n_sc, 3, 11, 3), "history/agent_no_sim/yaw_bbox": z(n_sc, 3, 11, 1),
             "scenario_id": ["abc"] * n_sc, "history/agent/valid": z(n_sc, A, 11, dtype=torch.bool),
             "history/agent/pos": z(n_sc, A, 11, 3), "history/agent/object_id": z(n_sc, A),
             "history/agent_no_sim/valid": z(n_sc, 3, 11, dtype=torch.bool),
             "history/agent_no_sim/object_id": z(n_sc, 3)}
    out = pp(batch, buffer)
    # recover which futures the reference kept by matching its (untransformed) selection against the inputs
    trajs = pp._filter_futures(buffer, inp["role"])
    d = (trajs[:, :, None] - buffer.pred_pose[:, None, :, :, t0:]).abs().flatten(3).amax(-1)  # [n_sc, 32, K]
    idx = d.argmin(-1)
    assert float(d.min(-1)[0].max()) == 0.0 and all(len(set(r.tolist())) == 32 for r in idx)
    fix = dict(shape=shape, t0=t0, w_road_edge=w, n_keep=32, sel=idx, pos_sim=out["pos_sim"], yaw_sim=out["yaw_sim"])
    torch.save(fix, os.path.join(HERE, "wosac_post.pt"))
    print("wosac_post.pt", os.path.getsize(os.path.join(HERE, "wosac_post.pt")) // 1024, "KiB")


def golden_womd_post():
    """the real WOMDPostProcessing (womd_post_processing.py:36-106) with the configured
    parameters of configs/model/sim_agent.yaml:170-177 on seeded joint futures + log-prob scores."""
    from data_modules.womd_post_processing import WOMDPostProcessing  # reference
    fix = {}
    for name, shape, temp in [("k12", dict(seed=8000, n_sc=2, K=12, A=24, T=80), -1.0),
                              ("k32_temp", dict(seed=8001, n_sc=1, K=32, A=16, T=80), 0.7),
                              ("k6", dict(seed=8002, n_sc=1, K=6, A=8, T=80), -1.0)]:
        inp = synth.make_womd_post_inputs(**shape)
        pp = WOMDPostProcessing(k_pred=6, score_temperature=temp, mpa_nms_thresh=[2.0, 1.0, 3.0], mtr_nms_thresh=[],
                                aggr_thresh=[], n_iter_em=3, use_ade=True, step_gt=90, step_current=10)
        out = pp(ag_type=inp["ag_type"], trajs=inp["trajs"].clone(), scores=inp["scores"].clone())
        fix[name] = dict(shape=shape, score_temperature=temp, mpa_nms_thresh=[2.0, 1.0, 3.0], k_pred=6,
                         trajs=out["trajs"].clone(), scores=out["scores"].clone())
    torch.save(fix, os.path.join(HERE, "womd_post.pt"))
    print("womd_post.pt", os.path.getsize(os.path.join(HERE, "womd_post.pt")) // 1024, "KiB")


def golden_checks_redlight(model=None):
    """Second TrafficRuleChecker fixture (`run_red_light` had 0 positives in checks_dense.pt): the crafted
    scene of synth.make_rule_scene_batch through the real reference with all checks on."""
    if model is None:
        cfg = config.default_model_cfg()
        model = build_reference_model(cfg, params.init_params(cfg, seed=0))
    shape = dict(n_sc=3, n_ag=48, n_mp=96, n_tl=40, seed=5000, boundary=80.0, scale=0.4)
    batch = synth.make_rule_scene_batch(**shape)
    R, T = 2, 36
    res, _, _ = reference_rollout(model, batch, R, T, disable_check=False)
    keep = ("pred_valid", "pred_pose", "pred_motion", "tl_state", "collided", "collided_wosac", "run_road_edge",
            "run_red_light", "passive")
    fix = dict(shape=shape, maker="make_rule_scene_batch", R=R, T=T, param_seed=0, **{k: res[k] for k in keep})
    torch.save(fix, os.path.join(HERE, "checks_redlight.pt"))
    print("checks_redlight.pt", os.path.getsize(os.path.join(HERE, "checks_redlight.pt")) // 1024, "KiB",
          {k: int(res[k].sum()) for k in keep[4:]})


TRAIN_STATS_SEED = 12345


def grad_stats(grads: dict) -> dict:
    """Two numbers per gradient tensor (norm, projection on a seeded random direction): pins every gradient of the
    training step with a few KB (the 656 tensors themselves are 32 MB)."""
    out = {}
    for i, k in enumerate(sorted(grads)):
        g = grads[k]
        if g is None:
            out[k] = None
            continue
        r = torch.randn(g.shape, generator=torch.Generator().manual_seed(TRAIN_STATS_SEED + i), dtype=torch.float64)
        out[k] = (float(g.double().norm()), float((g.double() * r).sum()))
    return out


def golden_train():
    """the body of WaymoMotion.training_step (waymo_motion.py:313-385) restated around the REAL
    reference modules - TrafficBots (map / TL / agent encoders, LatentEncoder posterior + prior, NaviPredictor, heads),
    Dynamics, TeacherForcing, TrafficRuleChecker, RolloutBuffer, DifferentiableReward, TrainingMetrics (BalancedKL) -
    with autograd. The step's random draws are the inputs of synth.make_train_batch (forced agents are OR-ed into the
    TeacherForcing mask, the latent sample is mean + std * eps = Normal.rsample). Stores the loss terms and two
    statistics per gradient tensor. Dropout is off (model.eval(): the p = 0 configuration; the hot path has no
    batch-norm, so eval only switches dropout)."""
    from utils.buffer import RolloutBuffer
    from utils.rewards import DifferentiableReward
    from models.metrics.training import TrainingMetrics
    from trafficbotsv1_5_b200.training import TRAIN_CFG
    cfg = config.default_model_cfg()
    P = params.init_params(cfg, 0, with_navi_predictor=True, with_latent_post=True)
    model = TrafficBots(**DictConfig(cfg))
    sd = model.state_dict()
    mine = {k: tuple(v) for k, v in params.latent_post_shapes(cfg).items()}
    ref = {k: tuple(v.shape) for k, v in sd.items() if k.startswith(("latent_encoder.tl_encoder_post.",
           "latent_encoder.ag_encoder_post.", "latent_encoder.latent_dist_post.")) and not k.endswith((".freqs", "hist_ohe"))}
    assert ref == mine, set(ref) ^ set(mine)
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected, unexpected
    assert all(m.startswith(("latent_encoder.tl_encoder_prior.", "latent_encoder.ag_encoder_prior.",
                             "latent_encoder.latent_dist_prior.")) or m.endswith((".freqs", "pl_node_ohe", "hist_ohe"))
               for m in missing), missing
    model.eval()
    model.double()  # float64 end to end: the fixture then pins the oracle to ~1e-9 instead of fp32 summation noise
    tc = dict(TRAIN_CFG)
    shape = dict(n_sc=2, n_ag=28, n_mp=70, n_tl=27, seed=3000, boundary=120.0)
    n_steps = 14
    batch = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v)
             for k, v in synth.make_train_batch(**shape).items()}  # drawn in fp32 (the seeded scene), then widened
    torch.set_default_dtype(torch.float64)
    out = {}
    for variant in ("posterior", "prior_kl"):
        if variant == "prior_kl":
            batch["rollout_prior"] = True
            tc["kl_free_nats"] = 0.01
        model.zero_grad()
        mp_tokens = model.mp_encoder(batch["sc/mp_valid"], batch["sc/mp_attr"], batch["sc/mp_pose"], batch["ref/mp_type"])
        tl_tokens = model.tl_encoder.pre_compute(tl_valid=batch["sc/tl_valid"], tl_attr=batch["sc/tl_attr"],
                                                 tl_pose=batch["sc/tl_pose"], **mp_tokens)                       # :317-324
        lat = dict(ag_attr=batch["sc/ag_attr"], ag_type=batch["ref/ag_type"], mp_tokens=mp_tokens, tl_tokens=tl_tokens)
        latent_post = model.latent_encoder(ag_valid=batch["gt/ag_valid"], ag_motion=batch["gt/ag_motion"],
                                           ag_pose=batch["gt/ag_pose"], tl_state=batch["gt/tl_state"], posterior=True, **lat)
        latent_prior = model.latent_encoder(ag_valid=batch["sc/ag_valid"], ag_motion=batch["sc/ag_motion"],
                                            ag_pose=batch["sc/ag_pose"], tl_state=batch["sc/tl_state"], posterior=False, **lat)
        ag_latent = latent_prior if batch.get("rollout_prior", False) else latent_post                          # :348
        ag_latent_valid = ag_latent.valid
        ag_latent = ag_latent.mean + ag_latent.stddev * batch["ag_latent_eps"]                                  # rsample
        navi_pred = model.navi_predictor(ag_valid=batch["sc/ag_valid"], ag_attr=batch["sc/ag_attr"],
                                         ag_motion=batch["sc/ag_motion"], ag_pose=batch["sc/ag_pose"],
                                         ag_type=batch["ref/ag_type"], **mp_tokens)                              # :352-359
        # ---- reactive_replay (:386-437) -> rollout (:206-311) -> forward (:118-204)
        rule_checker = TrafficRuleChecker(
            mp_boundary=batch["map/boundary"], mp_valid=batch["map/valid"], mp_type=batch["map/type"],
            mp_pos=batch["map/pos"], mp_dir=batch["map/dir"], ag_type=batch["ref/ag_type"], ag_size=batch["ref/ag_size"],
            ag_goal=None, ag_dest=batch["gt/ag_navi"], tl_valid=tl_tokens["tl_token_valid"],
            tl_pose=tl_tokens["tl_token_pose"], disable_check=True)
        ag_tokens = dict(ag_type=batch["ref/ag_type"], ag_size=batch["ref/ag_size"], ag_attr=batch["sc/ag_attr"],
                         gt_valid=batch["gt/ag_valid"], gt_pose=batch["gt/ag_pose"], gt_motion=batch["gt/ag_motion"],
                         ag_latent=ag_latent, ag_latent_valid=ag_latent_valid, ag_navi=batch["gt/ag_navi"],
                         ag_navi_valid=batch["gt/ag_valid"].any(-1))
        tl_state_gt = batch["gt/tl_state"]
        tf = TeacherForcing(step_spawn_agent=tc["step_spawn_agent"], step_warm_start=tc["step_warm_start"],
                            prob_forcing_agent=0.0)
        tf.init(ag_valid=ag_tokens["gt_valid"], ag_pose=ag_tokens["gt_pose"], ag_motion=ag_tokens["gt_motion"],
                tl_state=tl_state_gt, current_epoch=0)
        tf.ag_teacher_forcing |= batch["tf/forcing_agent"].unsqueeze(-1) & ag_tokens["gt_valid"]              # teacher_forcing.py:87-92
        dyn = make_dynamics()
        dyn.init(tl_state=tl_state_gt, **ag_tokens)
        model.init()
        reward_fn = DifferentiableReward(
            l_pos=DictConfig(dict(weight=tc["w_pos"], criterion="SmoothL1Loss")),
            l_rot=DictConfig(dict(weight=tc["w_rot"], criterion="SmoothL1Loss", angular_type="cosine")),
            l_spd=DictConfig(dict(weight=tc["w_spd"], criterion="SmoothL1Loss")),
            w_collision=0, use_il_loss=True, reduce_collsion_with_max=True, is_enabled=True)
        buf = RolloutBuffer(n_steps, 10)
        buf.add_navi_log_prob(torch.zeros_like(batch["sc/ag_attr"][:, :, 0]), ag_tokens["ag_navi_valid"])
        for step in range(1, n_steps + 1):
            ag_override, tl_override = tf.get(step, dyn.ag_valid, dyn.ag_pose, dyn.ag_motion)
            ag_valid = dyn.ag_valid
            action_dist, tl_dist = model(
                ag_valid=ag_valid, ag_pose=dyn.ag_pose.detach(), ag_motion=dyn.ag_motion.detach(), ag_attr=dyn.ag_attr,
                ag_type=dyn.ag_type, ag_latent=dyn.ag_latent, ag_latent_valid=dyn.ag_latent_valid, ag_navi=dyn.ag_navi,
                ag_navi_valid=dyn.ag_navi_valid, ag_navi_updated=dyn.ag_navi_updated, tl_state=dyn.tl_state.detach(),
                tl_tokens=tl_tokens, mp_tokens=mp_tokens)                                                        # :158-177
            _, action_log_prob = dyn.update_ag(action_dist, True, None)
            pred = dict(action_log_prob=action_log_prob, pred_valid=ag_valid, pred_pose=dyn.ag_pose,
                        pred_motion=dyn.ag_motion)
            dyn.override_ag(ag_override)
            dyn.override_tl(tl_dist, tl_override)
            violation = rule_checker.check(pred["pred_valid"], pred["pred_pose"], pred["pred_motion"], dyn.tl_state)
            gv, gp, gm = (ag_tokens[k][:, :, step] for k in ("gt_valid", "gt_pose", "gt_motion"))
            reward = reward_fn.get(pred_valid=pred["pred_valid"], pred_pose=pred["pred_pose"],
                                   pred_motion=pred["pred_motion"], gt_valid=gv, gt_pose=gp, gt_motion=gm,
                                   ag_size=ag_tokens["ag_size"])
            nll = -1.0 * tl_dist.log_prob(tl_state_gt[:, :, step].max(-1)[1])                                    # :270-277
            buf.add(violation=violation, diffbar_reward=reward, tl_state_nll=nll,
                    tl_state_nll_invalid=tl_tokens["tl_token_invalid"], vis_dict={}, ag_override=ag_override, **pred)
            dyn.disable_ag(violation, gv)
            dyn.disable_navi(violation)
        buf.finish()
        buf.flatten_joint_future(1)
        metrics = TrainingMetrics(prefix="t", train_navi=True, train_latent=True, w_vae_kl=tc["w_vae_kl"],
                                  kl_balance_scale=tc["kl_balance_scale"], kl_free_nats=tc["kl_free_nats"],
                                  kl_for_unseen_agent=True, w_diffbar_reward=tc["w_diffbar_reward"], w_navi=tc["w_navi"],
                                  w_tl_state=tc["w_tl_state"], w_relevant_agent=0, p_loss_for_irrelevant=1.0,
                                  step_training_start=tc["step_training_start"], temporal_discount=-1.0,
                                  loss_for_teacher_forcing=True)
        md = metrics(buffer=buf, ag_role=batch["ref/ag_role"], navi_pred=navi_pred, navi_gt=batch["gt/ag_navi"],
                     latent_post=latent_post, latent_prior=latent_prior)
        md["t/loss"].backward()
        grads = {k: p.grad for k, p in model.named_parameters() if k in P}
        out[variant] = dict(terms={k[2:]: float(v.detach()) for k, v in md.items() if torch.is_tensor(v) and v.dim() == 0},
                            grad_stats=grad_stats(grads), pred_valid=buf.pred_valid.squeeze(1).clone(),
                            pred_pose=buf.pred_pose.squeeze(1).detach().clone())
        print(variant, out[variant]["terms"], "tensors with gradient:", sum(g is not None for g in grads.values()))
    torch.set_default_dtype(torch.float32)
    fix = dict(shape=shape, n_steps=n_steps, param_seed=0, stats_seed=TRAIN_STATS_SEED, dtype="float64", **out)
    torch.save(fix, os.path.join(HERE, "train_small.pt"))
    print("train_small.pt", os.path.getsize(os.path.join(HERE, "train_small.pt")) // 1024, "KiB")


@torch.no_grad()
def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "train":  # only (re)generate train_small.pt
        with torch.enable_grad():
            return golden_train()
    if len(sys.argv) > 1 and sys.argv[1] == "navi":  # only (re)generate navi_pred.pt
        return golden_navi_predictor()
    if len(sys.argv) > 1 and sys.argv[1] == "wosac":  # only (re)generate wosac_post.pt
        return golden_wosac_post()
    if len(sys.argv) > 1 and sys.argv[1] == "checks2":  # only (re)generate checks_redlight.pt
        return golden_checks_redlight()
    if len(sys.argv) > 1 and sys.argv[1] == "womd":  # only (re)generate womd_post.pt
        return golden_womd_post()
    ops = golden_ops()
    torch.save(ops, os.path.join(HERE, "ops.pt"))
    print("ops.pt", os.path.getsize(os.path.join(HERE, "ops.pt")) // 1024, "KiB")

    cfg = config.default_model_cfg()
    P = params.init_params(cfg, seed=0)
    model = build_reference_model(cfg, P)
    shape = dict(n_sc=2, n_ag=32, n_mp=96, n_tl=30, seed=1000, boundary=105.0)
    batch = synth.make_scene_batch(**shape)
    R, T = 2, 24
    res, static, rec = reference_rollout(model, batch, R, T, record_steps=(1, 5, 12, 20))
    fix = dict(shape=shape, R=R, T=T, param_seed=0,
               param_checksum=checksum(torch.cat([P[k].flatten() for k in sorted(P)])),
               input_checksum=checksum(torch.cat([batch[k].float().flatten() for k in sorted(batch)])),
               mp_token_feature=static["mp"]["mp_token_feature"], mp_token_invalid=static["mp"]["mp_token_invalid"],
               tl_token_attr=static["tl"]["tl_token_attr"],
               knn_idx_tl2tl=static["tl"]["knn_idx_tl2tl"].masked_fill(static["tl"]["knn_invalid_tl2tl"], -1).sort(-1)[0],
               rec=rec, **res)
    torch.save(fix, os.path.join(HERE, "rollout_small.pt"))
    print("rollout_small.pt", os.path.getsize(os.path.join(HERE, "rollout_small.pt")) // 1024, "KiB")
    print("outside/disabled agents:", int((~res["final_valid"]).sum()), "navi reached:", int((~res["final_navi_valid"]).sum()))
    print("max |action mean|", float(res["action_mean"].abs().max()))

    # ---- dense scene with ALL TrafficRuleChecker checks on (traffic_rule_checker.py:343-451): SURVEY 8(f) rank 1
    shape = dict(n_sc=2, n_ag=48, n_mp=96, n_tl=30, seed=3000, boundary=60.0, scale=0.2)
    batch = synth.make_scene_batch(**shape)
    R, T = 2, 36
    res, _, _ = reference_rollout(model, batch, R, T, disable_check=False)
    keep = ("pred_valid", "pred_pose", "pred_motion", "tl_state", "collided", "collided_wosac", "run_road_edge",
            "run_red_light", "passive")
    fix = dict(shape=shape, R=R, T=T, param_seed=0, **{k: res[k] for k in keep})
    torch.save(fix, os.path.join(HERE, "checks_dense.pt"))
    print("checks_dense.pt", os.path.getsize(os.path.join(HERE, "checks_dense.pt")) // 1024, "KiB",
          {k: int(res[k].sum()) for k in keep[4:]})
    golden_checks_redlight(model)
    golden_navi_predictor()
    golden_wosac_post()
    golden_womd_post()


if __name__ == "__main__":
    main()
