"""Rollout-level and dynamics-level drop-ins: `WaymoMotionRollout.rollout` with the reference's
argument list and reference-style (repeat_interleave'd) token dicts, returning a `RolloutBuffer`; `Dynamics` /
`MultiPathPP` with the reference's constructor and methods."""
import math

import pytest
import torch
from torch.distributions import Categorical, Independent, Normal

from oracle import tb_oracle as O
from trafficbotsv1_5_b200 import config, params, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"

if torch.cuda.is_available():
    from trafficbotsv1_5_b200.traffic_bots import TrafficBots
    from trafficbotsv1_5_b200.waymo_motion import (Dynamics, MultiPathPP, TeacherForcing, TrafficRuleChecker,
                                                   WaymoMotionRollout)


def _model(precision=0):
    cfg = config.default_model_cfg()
    P = params.init_params(cfg, seed=0)
    model = TrafficBots(cfg, precision=precision)
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected
    return model.eval().to(DEV), P, cfg


def _reference_style_inputs(model, batch, R):
    """What WaymoMotion.joint_future_pred builds before calling self.rollout (waymo_motion.py:449-510)."""
    gb = {k: v.to(DEV) for k, v in batch.items()}
    mp_tokens = model.mp_encoder(gb["sc/mp_valid"], gb["sc/mp_attr"], gb["sc/mp_pose"], gb["ref/mp_type"])
    tl_tokens = model.tl_encoder.pre_compute(tl_valid=gb["sc/tl_valid"], tl_attr=gb["sc/tl_attr"], tl_pose=gb["sc/tl_pose"],
                                             **mp_tokens)
    rep = lambda t: t.repeat_interleave(R, 0)  # noqa: E731
    n_sc, A = gb["sc/ag_valid"].shape[:2]
    ag_tokens = dict(ag_type=rep(gb["ref/ag_type"]), ag_size=rep(gb["ref/ag_size"]), ag_attr=rep(gb["sc/ag_attr"]),
                     gt_valid=rep(gb["sc/ag_valid"]), gt_pose=rep(gb["sc/ag_pose"]), gt_motion=rep(gb["sc/ag_motion"]),
                     ag_latent=gb["ag_latent"][:, :R].reshape(n_sc * R, A, -1), ag_latent_valid=rep(gb["ag_latent_valid"]),
                     ag_navi=rep(gb["agent/dest"]), ag_navi_valid=rep(gb["ag_navi_valid"]),
                     ag_navi_log_prob=torch.full((n_sc * R, A), -1.5, device=DEV))
    mpR = {k: rep(v) for k, v in mp_tokens.items()}
    tlR = {k: rep(v) for k, v in tl_tokens.items()}
    return gb, ag_tokens, mpR, tlR


@pytest.mark.parametrize("disable_check", [True, False])
def test_rollout_dropin_vs_reference_golden(golden_rollout, disable_check):
    """The reference's own rollout() call, argument for argument, against the golden rollout of the real reference."""
    g = golden_rollout
    model, P, cfg = _model()
    batch = synth.make_scene_batch(**g["shape"])
    R, T = g["R"], g["T"]
    gb, ag_tokens, mpR, tlR = _reference_style_inputs(model, batch, R)
    rep = lambda t: t.repeat_interleave(R, 0)  # noqa: E731
    checker = TrafficRuleChecker(mp_boundary=rep(gb["map/boundary"]), mp_valid=rep(gb["map/valid"]), mp_type=rep(gb["map/type"]),
                                 mp_pos=rep(gb["map/pos"]), mp_dir=rep(gb["map/dir"]), ag_type=ag_tokens["ag_type"],
                                 ag_size=ag_tokens["ag_size"], ag_goal=None, ag_dest=ag_tokens["ag_navi"],
                                 tl_valid=tlR["tl_token_valid"], tl_pose=tlR["tl_token_pose"], disable_check=disable_check)
    wm = WaymoMotionRollout(model, Dynamics.default(), n_joint_future=R)
    buf = wm.rollout(ag_tokens=ag_tokens, mp_tokens=mpR, tl_tokens=tlR, tl_state_gt=rep(gb["sc/tl_state"]),
                     teacher_forcing=TeacherForcing(step_spawn_agent=10, step_warm_start=10), rule_checker=checker,
                     step_end=T, deterministic_action=True)
    assert buf.step_start == 1 and buf.step_end == T and buf.step_future_start == 10
    assert torch.equal(buf.pred_valid.cpu(), g["pred_valid"])
    assert float((buf.pred_pose.cpu() - g["pred_pose"]).abs().max()) < 5e-3
    assert float((buf.pred_motion.cpu() - g["pred_motion"]).abs().max()) < 5e-3
    assert torch.equal(buf.vis_dict["tl_state"].cpu(), g["tl_state"])
    B, A = g["pred_valid"].shape[:2]
    assert buf.action_log_prob.shape == (B, A, T) and buf.mask_teacher_forcing.shape == (B, A, T)
    # log-density of Normal(mean, exp(-2)) at its mean, 2 dims (action_head.py:48-50, dynamics.py:90); 0 when invalid
    lp = 2 * 2.0 - math.log(2 * math.pi)
    assert torch.allclose(buf.action_log_prob[buf.pred_valid], torch.full((1,), lp, device=DEV).expand(int(buf.pred_valid.sum())))
    assert float(buf.action_log_prob[~buf.pred_valid].abs().max()) == 0.0
    # feedback flags reproduce the reference's end state (dynamics.py:166-204)
    reached = buf.violation["dest_reached"][:, :, -1].cpu()
    assert torch.equal(rep(batch["ag_navi_valid"]) & ~reached, g["final_navi_valid"])
    assert set(buf.violation) == {f"{k}{s}" for k in ("outside_map", "collided", "collided_wosac", "run_road_edge",
                                                     "run_red_light", "passive", "goal_reached", "dest_reached")
                                  for s in ("", "_this_step")}
    assert bool((buf.violation["outside_map"][:, :, 1:] >= buf.violation["outside_map"][:, :, :-1]).all())  # cumulative
    if disable_check:
        assert not bool(buf.violation["collided_this_step"].any())
    # teacher forcing mask: all ground-truth-valid agents during the warm start, nothing afterwards
    assert torch.equal(buf.mask_teacher_forcing[:, :, :10].cpu(), rep(batch["sc/ag_valid"])[:, :, 1:11])
    assert not bool(buf.mask_teacher_forcing[:, :, 10:].any())
    # tl_state_nll: finite and positive while ground truth exists, masked afterwards (waymo_motion.py:270-277)
    assert buf.tl_state_nll.shape == (B, tlR["tl_token_pose"].shape[1], T)
    assert bool((buf.tl_state_nll[:, :, :10] > 0).all()) and float(buf.tl_state_nll[:, :, 10:].abs().max()) == 0.0
    assert bool(buf.tl_state_nll_invalid[:, :, 10:].all())
    buf.flatten_joint_future(R)
    buf.compute_log_prob(None)
    n_sc = B // R
    assert buf.pred_pose.shape == (n_sc, R, A, T, 3) and buf.violation["collided"].shape == (n_sc, R, A, T)
    assert buf.log_prob.shape == (n_sc, R, A)
    nv = rep(batch["ag_navi_valid"]).view(n_sc, R, A)
    assert torch.allclose(buf.log_prob.cpu()[nv], torch.full((int(nv.sum()),), -1.5))


@pytest.mark.parametrize("R,precision", [(32, 0), (32, 1), (128, 1)])
def test_rollout_dropin_wosac_widths_vs_oracle(R, precision):
    """R = 32 (WOSAC) and R = 128 (configs/resume/submission.yaml:5) joint futures through joint_future_pred, with all
    rule checks on, against the oracle's closed loop; then the 128 -> 32 future filter on the buffer's flags."""
    model, P, cfg = _model(precision)
    shape = dict(n_sc=1, n_ag=40, n_mp=96, n_tl=30, seed=77, boundary=60.0, scale=0.2, n_rollout=R)
    T = 20
    batch = synth.make_scene_batch(**shape)
    gb = {k: v.to(DEV) for k, v in batch.items()}
    mp_tokens = model.mp_encoder(gb["sc/mp_valid"], gb["sc/mp_attr"], gb["sc/mp_pose"], gb["ref/mp_type"])
    tl_tokens = model.tl_encoder.pre_compute(tl_valid=gb["sc/tl_valid"], tl_attr=gb["sc/tl_attr"], tl_pose=gb["sc/tl_pose"],
                                             **mp_tokens)
    wm = WaymoMotionRollout(model, time_step_end=T)
    buf = wm.joint_future_pred(gb, mp_tokens, tl_tokens, gb["ag_latent"], gb["ag_latent_valid"], gb["agent/dest"],
                               gb["ag_navi_valid"], TeacherForcing(10, 10), n_joint_future=R)
    ref = O.rollout(P, cfg, config.derived_sizes(cfg), config.DYNAMICS_CFG, config.ROLLOUT_CFG, batch, R, T, rule_checks=True)
    A = shape["n_ag"]
    assert buf.pred_pose.shape == (1, R, A, T, 3)
    assert torch.equal(buf.pred_valid.view(R, A, T).cpu(), ref["pred_valid"])
    tol = 2e-2 if precision else 5e-3
    assert float((buf.pred_pose.view(R, A, T, 3).cpu() - ref["pred_pose"]).abs().max()) < tol
    tot = 0
    for k in ("collided", "collided_wosac", "run_road_edge", "run_red_light", "passive"):
        mine = buf.violation[f"{k}_this_step"].view(R, A, T).cpu()
        n_diff, n_pos = int((mine != ref[k]).sum()), int(ref[k].sum())
        tot += n_pos
        assert n_diff <= max(2, int(0.01 * n_pos)), (k, n_diff, n_pos)
    assert tot > 100
    # rollouts differ through their latent samples only
    assert float((buf.pred_pose[0, 0] - buf.pred_pose[0, R - 1]).abs().max()) > 0


def test_dynamics_dropin_matches_reference_semantics():
    """Dynamics / MultiPathPP: constructor, attributes and every method of utils/dynamics.py against the oracle's
    restatement of update_ag and plain tensor algebra for the overrides."""
    g = torch.Generator().manual_seed(3)
    B, A, n_tl = 3, 17, 5
    d = config.DYNAMICS_CFG
    dyn = Dynamics(veh=dict(_target_="utils.dynamics.MultiPathPP", **d["veh"]), ped=d["ped"], cyc=d["cyc"], navi_mode="dest")
    assert dyn.dt == 0.1 and dyn.action_dim == 2 and isinstance(dyn.ag_dynamics[0], MultiPathPP)
    gt_valid = torch.rand(B, A, 4, generator=g) < 0.8
    gt_pose = torch.randn(B, A, 4, 3, generator=g) * 20
    gt_motion = torch.randn(B, A, 4, 3, generator=g)
    ag_type = torch.nn.functional.one_hot(torch.randint(0, 3, (B, A), generator=g), 3).bool()
    tl_state = torch.nn.functional.one_hot(torch.randint(0, 5, (B, n_tl, 4), generator=g), 5).bool()
    c = lambda t: t.to(DEV)  # noqa: E731
    dyn.init(tl_state=c(tl_state), gt_valid=c(gt_valid), gt_pose=c(gt_pose), gt_motion=c(gt_motion), ag_type=c(ag_type),
             ag_attr=c(torch.zeros(B, A, 6)), ag_latent=None, ag_latent_valid=None, ag_navi=c(torch.zeros(B, A, dtype=torch.long)),
             ag_navi_valid=c(gt_valid[:, :, 0]))
    assert torch.equal(dyn.ag_valid.cpu(), gt_valid[:, :, 0]) and dyn.ag_navi_updated and not bool(dyn.ag_disabled.any())
    mean = torch.randn(B, A, 2, generator=g)
    valid = gt_valid[:, :, 0]
    scale = torch.full((B, A, 2), math.exp(-2.0))
    dist = Independent(Normal(c(mean), c(scale)), 1)
    action, log_prob = dyn.update_ag(dist, True, None)
    ref_pose, ref_motion = O.dynamics_update(gt_pose[:, :, 0], gt_motion[:, :, 0], valid, ag_type, mean, d)
    assert float((dyn.ag_pose.cpu() - ref_pose).abs().max()) < 1e-5 and float((dyn.ag_motion.cpu() - ref_motion).abs().max()) < 1e-5
    assert torch.allclose(log_prob.cpu(), dist.log_prob(c(mean)).cpu().masked_fill(~valid, 0))
    assert float(action.cpu()[~valid].abs().max()) == 0.0
    # player override replaces the physical action of valid agents (dynamics.py:96-99)
    dyn.ag_pose, dyn.ag_motion = c(gt_pose[:, :, 0]), c(gt_motion[:, :, 0])
    pv = torch.rand(B, A, generator=g) < 0.5
    pa = torch.randn(B, A, 2, generator=g)
    action, _ = dyn.update_ag(dist, True, dict(valid=c(pv), action=c(pa)))
    assert torch.allclose(action.cpu()[pv & valid], pa[pv & valid])
    # MultiPathPP stand-alone (dynamics.py:237-274)
    mp = MultiPathPP(0.1, max_acc=5.0, max_yaw_rate=1.5)
    a = mp.process_action(c(mean))
    assert torch.allclose(a.cpu(), torch.tanh(mean) * torch.tensor([5.0, 1.5]), atol=1e-6)
    p2, m2 = mp.update(c(gt_pose[:, :, 0]), c(gt_motion[:, :, 0]), a)
    acc, yr = a.cpu()[..., 0], a.cpu()[..., 1]
    v_t, th_t = gt_motion[:, :, 0, 0] + 0.05 * acc, gt_pose[:, :, 0, 2] + 0.05 * yr
    exp_pose = gt_pose[:, :, 0] + 0.1 * torch.stack([v_t * th_t.cos(), v_t * th_t.sin(), yr], -1)
    assert float((p2.cpu() - exp_pose).abs().max()) < 1e-5
    assert torch.allclose(m2.cpu(), torch.stack([gt_motion[:, :, 0, 0] + 0.1 * acc, acc, yr], -1), atol=1e-6)
    # overrides and disabling (dynamics.py:122-204)
    dyn.ag_valid = c(valid)
    ov = dict(valid=c(gt_valid[:, :, 1]), pose=c(gt_pose[:, :, 1]), motion=c(gt_motion[:, :, 1]))
    dyn.override_ag(ov)
    assert torch.equal(dyn.ag_valid.cpu(), valid | gt_valid[:, :, 1])
    assert torch.equal(dyn.ag_pose.cpu()[gt_valid[:, :, 1]], gt_pose[:, :, 1][gt_valid[:, :, 1]])
    logits = torch.randn(B, n_tl, 5, generator=g)
    tlv = torch.rand(B, n_tl, generator=g) < 0.5
    dyn.override_tl(Categorical(logits=c(logits)), dict(valid=c(tlv), state=c(tl_state[:, :, 1])))
    exp_tl = torch.where(tlv[..., None], tl_state[:, :, 1], torch.nn.functional.one_hot(logits.argmax(-1), 5).bool())
    assert torch.equal(dyn.tl_state.cpu(), exp_tl)
    out = torch.rand(B, A, generator=g) < 0.3
    before = dyn.ag_valid.cpu().clone()
    dyn.disable_ag(dict(outside_map_this_step=c(out)), c(gt_valid[:, :, 2]))
    assert torch.equal(dyn.ag_valid.cpu(), before & ~(out & ~gt_valid[:, :, 2]))
    dyn.override_ag(dict(valid=c(torch.ones(B, A, dtype=torch.bool)), pose=c(gt_pose[:, :, 2]), motion=c(gt_motion[:, :, 2])))
    assert not bool((dyn.ag_valid & dyn.ag_disabled).any())  # disabled agents are never re-spawned
    reached = torch.rand(B, A, generator=g) < 0.3
    dyn.disable_navi(dict(dest_reached_this_step=c(reached)))
    assert torch.equal(dyn.ag_navi_valid.cpu(), gt_valid[:, :, 0] & ~reached)
    dyn.ag_navi_updated = False
    dyn.override_navi(c(torch.full((B, A), 7)))
    assert dyn.ag_navi_updated and torch.equal(dyn.ag_navi.cpu()[reached], torch.full((int(reached.sum()),), 7))


def test_rollout_dropin_refuses_training_configurations():
    model, P, cfg = _model()
    wm = WaymoMotionRollout(model)
    with pytest.raises(NotImplementedError):
        wm.rollout({}, {}, {}, None, TeacherForcing(), None, 10, deterministic_action=False)
    from trafficbotsv1_5_b200.waymo_motion import _check_teacher_forcing
    with pytest.raises(NotImplementedError):
        _check_teacher_forcing(TeacherForcing(prob_scheduled_sampling=0.5))
    with pytest.raises(NotImplementedError):
        Dynamics(veh=dict(_target_="utils.dynamics.StateIntegrator"), ped={}, cyc={}, navi_mode="dest")


def test_training_step_dropin_trains_the_module():
    """`WaymoMotionRollout.training_step`: the module's own parameters are the leaves of the CUDA training step, so a
    torch optimiser over `model.parameters()` trains it. Same loss as a stand-alone `TrainStep` on the same weights; a
    few AdamW steps on one batch lower the loss."""
    from trafficbotsv1_5_b200.traffic_bots import TrafficBots
    from trafficbotsv1_5_b200.training import TrainStep
    from trafficbotsv1_5_b200.waymo_motion import WaymoMotionRollout
    model = TrafficBots(seed=0, training_modules=True).cuda()
    wm = WaymoMotionRollout(model, time_step_end=14)
    batch = synth.make_train_batch(2, n_ag=28, n_mp=70, n_tl=27, seed=3000, boundary=120.0)
    ref = TrainStep({k: v.detach().clone() for k, v in model.named_parameters()}, model.cfg, "cuda",
                    train_cfg=dict(time_step_end=14))
    loss_ref = float(ref.step(batch, backward=False)["loss"].detach())
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4, weight_decay=1e-1, betas=(0.9, 0.95))
    losses = []
    for _ in range(4):
        model.zero_grad()
        out = wm.training_step(batch)
        losses.append(float(out["loss"].detach()))
        assert sum(p.grad is not None for p in model.parameters()) > 600
        opt.step()
    assert abs(losses[0] - loss_ref) < 1e-4 * abs(loss_ref), (losses[0], loss_ref)
    assert losses[-1] < losses[0] - 1e-3, losses
    # the updated weights drive the inference path of the same module (fused weights are rebuilt on a version change)
    assert model._runner() is not None
