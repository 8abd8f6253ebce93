"""GPU parity of the encoders and the closed-loop rollout against (a) golden vectors of the real reference and
(b) the CPU oracle on other seeds, plus size-independent properties at the full BASELINE shape."""
import pytest
import torch

from oracle import tb_oracle as O
from trafficbotsv1_5_b200 import config, params, synth

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from trafficbotsv1_5_b200.engine import RolloutEngine

DEV = "cuda"
# Stated closed-loop tolerance (fp32 path): per-step position 5e-3 m, yaw 1e-3 rad, speed 5e-3 m/s over the whole
# rollout; masks (valid / TL state / navigation reached) bit-exact on these tie-free fixtures.
TOL_XY, TOL_YAW, TOL_SPD = 5e-3, 1e-3, 5e-3


def maxerr(a, b):
    return float((a.detach().float().cpu() - b.detach().float().cpu()).abs().max())


def _engine(g_shape, R, T, **kw):
    cfg = config.default_model_cfg()
    P = params.init_params(cfg, seed=0)
    eng = RolloutEngine(P, cfg, DEV, n_rollout=R, step_end=T, **kw)
    batch = synth.make_scene_batch(**g_shape)
    return eng, batch, P, cfg


def test_scene_encoders_golden(golden_rollout):
    g = golden_rollout
    eng, batch, P, cfg = _engine(g["shape"], g["R"], g["T"])
    static = eng.encode_scenes(batch)
    assert torch.equal(static["mp"]["mp_token_invalid"].cpu(), g["mp_token_invalid"])
    scale = float(g["mp_token_feature"].abs().max())
    err = maxerr(static["mp"]["mp_token_feature"], g["mp_token_feature"])
    assert err < 1e-4 * scale, f"map encoder max abs err {err:.3e} (scale {scale:.3e})"  # 1e-4 relative to scale
    err = maxerr(static["tl"]["tl_token_attr"].view(g["tl_token_attr"].shape), g["tl_token_attr"])
    assert err < 1e-4 * scale
    ks = static["tl"]["knn_self"]
    mine = ks["idx"].long().masked_fill(ks["inv"], -1).sort(-1)[0].cpu()
    assert torch.equal(mine, g["knn_idx_tl2tl"])


@pytest.mark.parametrize("use_graph", [False, True])
def test_rollout_golden(golden_rollout, use_graph):
    g = golden_rollout
    eng, batch, P, cfg = _engine(g["shape"], g["R"], g["T"], use_graph=use_graph)
    res = eng.rollout(batch)
    torch.cuda.synchronize()
    assert torch.equal(res["pred_valid"].cpu(), g["pred_valid"]), "pred_valid differs"
    assert torch.equal(res["tl_state"].cpu(), g["tl_state"]), "tl_state differs"
    assert torch.equal(res["final_valid"].cpu(), g["final_valid"])
    assert torch.equal(res["final_navi_valid"].cpu(), g["final_navi_valid"])
    e_xy = maxerr(res["pred_pose"][..., :2], g["pred_pose"][..., :2])
    e_yaw = maxerr(res["pred_pose"][..., 2], g["pred_pose"][..., 2])
    e_spd = maxerr(res["pred_motion"], g["pred_motion"])
    print(f"rollout vs reference golden: xy {e_xy:.3e} m, yaw {e_yaw:.3e} rad, motion {e_spd:.3e}")
    assert e_xy < TOL_XY and e_yaw < TOL_YAW and e_spd < TOL_SPD


def test_policy_step_vs_oracle_intermediates(golden_rollout):
    """step-by-step (no graph): action-head outputs and TL logits of selected steps vs the reference recordings."""
    g = golden_rollout
    eng, batch, P, cfg = _engine(g["shape"], g["R"], g["T"], use_graph=False)
    eng.prepare(batch)
    rec = {}

    def record(s, aux):
        if s in g["rec"]:
            mean = torch.zeros(aux["act"].shape[0], 2, device=DEV)
            from trafficbotsv1_5_b200 import lib as L, ops
            st = eng._st
            # valid at the time of the policy call = pred_valid[:, :, s-1]
            v = st["pred_valid"][:, :, s - 1].contiguous()
            L.check(L.load().tb_action_mean(L.ptr(aux["act"]), L.ptr(st["ag_type"]), L.ptr(v), mean.shape[0],
                                            L.ptr(mean), L.stream()), "tb_action_mean")
            rec[s] = dict(mean=mean.cpu(), logits=aux["logits"].cpu())

    eng.run(max(g["rec"]), record=record)
    for s, r in g["rec"].items():
        B, A, _ = r["mean"].shape
        e = maxerr(rec[s]["mean"].view(B, A, 2), r["mean"])
        assert e < 2e-4, f"action mean @ step {s}: {e:.3e}"  # outputs O(0.1..1): 1e-4 relative class
        tl_inv = ~batch["sc/tl_valid"]
        lg = rec[s]["logits"].view(tl_inv.shape[0], -1, 5).masked_fill(tl_inv[..., None], 0.0).clamp(-3, 3)
        e = maxerr(torch.log_softmax(lg, -1).repeat_interleave(g["R"], 0), r["logits"])
        assert e < 2e-4, f"tl logits @ step {s}: {e:.3e}"


def test_rollout_vs_oracle_other_seed():
    shape = dict(n_sc=1, n_ag=40, n_mp=128, n_tl=32, seed=4242, boundary=110.0)
    R, T = 3, 30
    eng, batch, P, cfg = _engine(shape, R, T)
    res = eng.rollout(batch)
    ref = O.rollout(P, cfg, config.derived_sizes(cfg), config.DYNAMICS_CFG, config.ROLLOUT_CFG, batch, R, T)
    assert torch.equal(res["pred_valid"].cpu(), ref["pred_valid"])
    assert torch.equal(res["tl_state"].cpu(), ref["tl_state"])
    e_xy = maxerr(res["pred_pose"][..., :2], ref["pred_pose"][..., :2])
    e_yaw = maxerr(res["pred_pose"][..., 2], ref["pred_pose"][..., 2])
    print(f"rollout vs oracle: xy {e_xy:.3e} m, yaw {e_yaw:.3e} rad")
    assert e_xy < TOL_XY and e_yaw < TOL_YAW
    # rollouts of one scene differ only through their latent sample; rollout r of the engine == oracle row r
    assert float((res["joint_pose"][0, 0] - res["joint_pose"][0, 1]).abs().max()) > 0


def test_full_shape_properties():
    """BASELINE shape (128 agents, 1024 polylines, 40 TL, 32 rollouts) at reduced scene/step count:
    determinism across graph replays, finite outputs, invalid agents exactly zero, warm-start steps equal GT."""
    shape = dict(n_sc=2, n_ag=128, n_mp=1024, n_tl=40, seed=7)
    eng, batch, P, cfg = _engine(shape, 32, 14)
    r1 = {k: v.clone() for k, v in eng.rollout(batch).items()}
    r2 = eng.run()
    for k in ("pred_pose", "pred_motion", "pred_valid", "tl_state"):
        assert torch.equal(r1[k], r2[k]), f"{k} not deterministic across replays"
    assert bool(torch.isfinite(r1["pred_pose"]).all())
    inv = ~r1["pred_valid"]
    assert float(r1["pred_pose"][inv].abs().max()) == 0.0
    # agents valid at t and t+1 in the history are teacher-forced: state after step s equals GT[s] for s <= 10
    # => the prediction of step s+1 starts from GT[s]; constant-velocity GT => prediction error stays small
    gt = batch["sc/ag_pose"].repeat_interleave(32, 0).to(DEV)
    gv = batch["sc/ag_valid"].repeat_interleave(32, 0).to(DEV)
    both = gv[:, :, 1:11] & gv[:, :, 0:10]
    err = (r1["pred_pose"][:, :, 0:10, :2] - gt[:, :, 1:11, :2]).norm(dim=-1)[both]
    assert float(err.max()) < 1.0  # |a|<=7 m/s^2 over 0.1 s on top of constant velocity
    # all 32 rollouts share the warm-start; they must diverge afterwards only through the latent
    jp = r1["joint_pose"]
    assert float((jp[:, 0, :, 12] - jp[:, 1, :, 12]).abs().max()) > 0


# Tensor-core mode (precision=1: tcgen05 kind::tf32 projections, everything else fp32). Stated closed-loop tolerance vs
# the fp32 reference over the 24-step golden rollout: position 2e-2 m, yaw 5e-3 rad, speed 2e-2 m/s per step;
# validity / traffic-light masks identical on this fixture.
TC_TOL_XY, TC_TOL_YAW, TC_TOL_SPD = 2e-2, 5e-3, 2e-2


def test_rollout_golden_tensor_core(golden_rollout):
    g = golden_rollout
    eng, batch, P, cfg = _engine(g["shape"], g["R"], g["T"], precision=1)
    res = eng.rollout(batch)
    torch.cuda.synchronize()
    e_xy = maxerr(res["pred_pose"][..., :2], g["pred_pose"][..., :2])
    e_yaw = maxerr(res["pred_pose"][..., 2], g["pred_pose"][..., 2])
    e_spd = maxerr(res["pred_motion"], g["pred_motion"])
    n_valid_diff = int((res["pred_valid"].cpu() != g["pred_valid"]).sum())
    n_tl_diff = int((res["tl_state"].cpu() != g["tl_state"]).sum())
    print(f"tf32 rollout vs reference golden: xy {e_xy:.3e} m, yaw {e_yaw:.3e} rad, motion {e_spd:.3e}, "
          f"valid mismatches {n_valid_diff}, tl mismatches {n_tl_diff}")
    assert n_valid_diff == 0 and n_tl_diff == 0
    assert e_xy < TC_TOL_XY and e_yaw < TC_TOL_YAW and e_spd < TC_TOL_SPD


# Stated closed-loop tolerance of the BENCHMARKED mode (precision=1: tcgen05 tf32 / fp16 projections, fp16 K|V, q|u,
# ov|z, FFN-hidden and LayerNorm rows, mma.sync attention, fused history encoder) against the fp32 oracle over ALL 90
# policy iterations (80 counted WOSAC steps): at policy iteration t the position error stays below
# 2e-3 + 1.2e-4 t^2 m (0.05 m at t = 20, 0.30 m at t = 50, 0.97 m at t = 90) and the heading error below
# 1e-3 + 2e-5 t^2 rad; validity, traffic-light and navigation masks identical. Measured on four scenes and two builds
# (different rounding patterns): 0.21-0.54 m / 0.011-0.05 rad at t = 90 while the agents travel 50-90 m. The envelope
# is quadratic because the error is not noise:
# 10-bit-mantissa WEIGHTS (tf32 / fp16 operands, as in the reference's own AMP-fp16 runs) are a fixed ~5e-4 relative
# perturbation of the policy, i.e. a nearly constant acceleration error that integrates twice. The oracle itself with
# nothing but its nn.Linear weights rounded to fp16 drifts 0.29 m by t = 90 on the config-1 scene (activations
# rounded instead: 0.05 m) - profiles/precision_floor.py, profiles/r2/precision_floor.txt. The fp32 mode
# (precision=0) keeps 1e-2 m / 2e-3 rad over the whole horizon.
def _tc90_envelope(T):
    t = torch.arange(1, T + 1, dtype=torch.float32)
    return 2e-3 + 1.2e-4 * t * t, 1e-3 + 2e-5 * t * t
_ORACLE_CACHE = {}


def _oracle_rollout(shape, R, T, P, cfg):
    key = (tuple(sorted(shape.items())), R, T)
    if key not in _ORACLE_CACHE:
        batch = synth.make_scene_batch(**shape)
        _ORACLE_CACHE[key] = O.rollout(P, cfg, config.derived_sizes(cfg), config.DYNAMICS_CFG, config.ROLLOUT_CFG, batch,
                                       R, T)
    return _ORACLE_CACHE[key]


def _compare_90(res, ref, precision, what):
    n_bad = {k: int((res[k].cpu() != ref[k]).sum()) for k in ("pred_valid", "tl_state", "final_valid", "final_navi_valid")}
    err = (res["pred_pose"].cpu() - ref["pred_pose"]).abs()
    xy, yaw = err[..., :2].amax(dim=(0, 1, 3)), err[..., 2].amax(dim=(0, 1))
    print(f"{what} precision={precision}: mask mismatches {n_bad}")
    print("  max xy err per 10th step [m]:  ", " ".join(f"{float(v):.1e}" for v in xy[9::10]))
    print("  max yaw err per 10th step [rad]:", " ".join(f"{float(v):.1e}" for v in yaw[9::10]))
    assert not any(n_bad.values()), n_bad
    if precision:
        tol_xy, tol_yaw = _tc90_envelope(xy.numel())
        print("  tolerance envelope at those steps [m]:", " ".join(f"{float(v):.1e}" for v in tol_xy[9::10]))
        assert bool((xy < tol_xy).all()) and bool((yaw < tol_yaw).all()), (float((xy / tol_xy).max()),
                                                                             float((yaw / tol_yaw).max()))
    else:
        assert float(xy.max()) < 1e-2 and float(yaw.max()) < 2e-3, (float(xy.max()), float(yaw.max()))


@pytest.mark.parametrize("precision", [0, 1])
def test_rollout_full_90_steps_vs_oracle(precision):
    """BASELINE config-1-sized scene (64 agents, 256 polylines x 20, 40 TL, 11-step history), all 90 policy iterations
    (80 counted WOSAC steps) x 2 rollouts against the CPU oracle, in the fp32 mode AND in the benchmarked tensor-core
    mode. fp32: stated per-step position tolerance over the whole horizon 1e-2 m / 2e-3 rad. This is the fp32 noise
    floor of the closed loop, not slack: the fp32 oracle (= the reference's arithmetic) differs from ITS OWN float64
    evaluation by 4.2e-3 m at step 90 on this scene (PoseEmb evaluates cos(x * 1 rad/m) on coordinates up to ~150 m,
    so fp32 rounding of x is amplified); the float64 oracle is therefore checked too. Masks must be identical."""
    shape = dict(n_sc=1, n_ag=64, n_mp=256, n_tl=40, seed=31, boundary=120.0)
    R, T = 2, 90
    eng, batch, P, cfg = _engine(shape, R, T, precision=precision)
    res = eng.rollout(batch)
    sz = config.derived_sizes(cfg)
    ref = _oracle_rollout(shape, R, T, P, cfg)
    _compare_90(res, ref, precision, "config-1 scene, 90 iterations")
    if precision == 0:
        # float64 evaluation of the same algorithm
        torch.set_default_dtype(torch.float64)
        try:
            P64 = {k: v.double() for k, v in P.items()}
            b64 = {k: (v.double() if v.dtype == torch.float32 else v) for k, v in batch.items()}
            ref64 = O.rollout(P64, cfg, sz, config.DYNAMICS_CFG, config.ROLLOUT_CFG, b64, R, T)
        finally:
            torch.set_default_dtype(torch.float32)
        e64 = (res["pred_pose"].cpu().double() - ref64["pred_pose"]).abs()[..., :2].amax(dim=(0, 1, 3))
        o64 = (ref["pred_pose"].double() - ref64["pred_pose"]).abs()[..., :2].amax(dim=(0, 1, 3))
        print("max xy err vs fp64 oracle:", [f"{float(v):.1e}" for v in e64[9::10]], "| fp32 oracle vs fp64 oracle:",
              [f"{float(v):.1e}" for v in o64[9::10]])
        assert float(e64.max()) < 1e-2
    # WOSAC slice = the 80 steps after the 10-step warm start (wosac_post_processing / waymo_motion.py:888-902)
    assert res["pred_pose"][:, :, 10:].shape[2] == 80


@pytest.mark.parametrize("precision", [0, 1])
def test_rollout_config3_shape_90_steps_vs_oracle(precision):
    """The BENCHMARKED scene shape and horizon (BASELINE config 3: 128 agents, 1024 polylines x 20, 40 TL; 1 scene x 4
    rollouts, all 90 policy iterations) in the parity mode and in the benchmarked mode, against the CPU oracle."""
    shape = dict(n_sc=1, n_ag=128, n_mp=1024, n_tl=40, seed=1000, n_rollout=4)
    R, T = 4, 90
    eng, batch, P, cfg = _engine(shape, R, T, precision=precision)
    res = eng.rollout(batch)
    ref = _oracle_rollout(shape, R, T, P, cfg)
    _compare_90(res, ref, precision, "config-3 scene shape, 90 iterations")


def test_run_refuses_more_steps_than_step_end():
    eng, batch, P, cfg = _engine(dict(n_sc=1, n_ag=30, n_mp=80, n_tl=28, seed=9, boundary=110.0), 2, 12)
    eng.prepare(batch)
    with pytest.raises(ValueError):
        eng.run(13)
    res = eng.run(12)
    assert bool(torch.isfinite(res["pred_pose"]).all())


def test_fp16_range_guard_trips_on_large_activations():
    """the tensor-core mode keeps [q|u], [k|v] and the FFN hidden rows in fp16. With weights scaled so that
    the ReLU hidden rows exceed 65,504 the conversions saturate (no inf / NaN) and the engine raises instead of
    returning silently clamped trajectories; the same weights run cleanly in the fp32 mode."""
    shape = dict(n_sc=1, n_ag=30, n_mp=80, n_tl=28, seed=9, boundary=110.0)
    cfg = config.default_model_cfg()
    P = dict(params.init_params(cfg, seed=0))
    k = "ag_encoder.tf_ag2agmptl.layers.1.linear1"
    P[f"{k}.weight"] = P[f"{k}.weight"] * 3e5
    P[f"{k}.bias"] = P[f"{k}.bias"] + 1e5
    k2 = "ag_encoder.tf_ag2agmptl.layers.1.linear2"
    P[f"{k2}.weight"] = P[f"{k2}.weight"] / 3e5
    batch = synth.make_scene_batch(**shape)
    eng = RolloutEngine(P, cfg, DEV, precision=1, n_rollout=2, step_end=12)
    with pytest.raises(FloatingPointError):
        eng.rollout(batch)
    assert bool(torch.isfinite(eng.results()["pred_pose"]).all())  # saturated, not inf / NaN
    eng0 = RolloutEngine(P, cfg, DEV, precision=0, n_rollout=2, step_end=12)
    assert bool(torch.isfinite(eng0.rollout(batch)["pred_pose"]).all())
    # and the unscaled weights do not trip it
    eng1 = RolloutEngine(params.init_params(cfg, seed=0), cfg, DEV, precision=1, n_rollout=2, step_end=12)
    eng1.rollout(batch)


# ---------------------------------------------------------------------------------------------- rule checks (8(f) rank 1)
VIO = ("collided", "collided_wosac", "run_road_edge", "run_red_light", "passive")


def test_rule_check_kernel_on_reference_predictions(golden_checks_any):
    """tb_rule_check replayed step by step on the reference's own per-step predictions of a dense scene (thousands of
    collision / road-edge events): flags must equal the reference's, up to knife-edge fp32 cases (<= 0.5 % of the
    positives may flip; cos/sin of the heading differ by 1 ulp between libm and CUDA)."""
    from trafficbotsv1_5_b200 import lib as L, ops
    from trafficbotsv1_5_b200.engine import rule_tables
    g = golden_checks_any
    batch = {k: v.to(DEV) for k, v in getattr(synth, g.get("maker", "make_scene_batch"))(**g["shape"]).items()}
    R, T, W = g["R"], g["T"], 11
    B, A = g["pred_valid"].shape[:2]
    n_sc, n_tl = batch["sc/tl_valid"].shape
    tab = rule_tables(batch["map/valid"], batch["map/type"], batch["map/pos"][..., :2].contiguous(),
                      batch["map/dir"][..., :2].contiguous())
    pv = g["pred_valid"].to(DEV).to(torch.uint8).contiguous()
    pp, pm = g["pred_pose"].to(DEV).contiguous(), g["pred_motion"].to(DEV).contiguous()
    ag_type = batch["ref/ag_type"].repeat_interleave(R, 0).to(torch.uint8).contiguous()
    hist_tl = torch.zeros(B, n_tl, W, 5, dtype=torch.uint8, device=DEV)
    tl_inv = (~batch["sc/tl_valid"]).to(torch.uint8).contiguous()
    counter = torch.zeros(B, A, device=DEV)
    outs = {k: torch.zeros(B, A, T, dtype=torch.uint8, device=DEV) for k in VIO}
    d_step = torch.zeros(1, dtype=torch.int32, device=DEV)
    for s in range(1, T + 1):
        d_step.fill_(s)
        hist_tl[:, :, s % W] = g["tl_state"][:, :, s - 1].to(DEV).to(torch.uint8)
        L.check(L.load().tb_rule_check(
            L.ptr(pv), L.ptr(pp), L.ptr(pm), L.ptr(ag_type), L.ptr(batch["ref/ag_size"].contiguous()), L.ptr(hist_tl),
            L.ptr(tl_inv), L.ptr(batch["sc/tl_pose"].contiguous()), L.ptr(tab["seg"]), L.ptr(tab["node_invalid"]),
            L.ptr(tab["poly_circle"]), L.ptr(tab["poly_kind"]), tab["seg"].shape[1], tab["seg"].shape[2], L.ptr(counter),
            *[L.ptr(outs[k]) for k in VIO], L.ptr(d_step), B, A, T, W, n_tl, R, 1, 1.1, L.stream()), "tb_rule_check")
    torch.cuda.synchronize()
    for k in VIO:
        mine, ref = outs[k].bool().cpu(), g[k]
        n_diff, n_pos = int((mine != ref).sum()), int(ref.sum())
        print(f"{k}: reference positives {n_pos}, mismatches {n_diff}")
        assert n_diff <= max(1, int(0.005 * n_pos)), f"{k}: {n_diff} mismatches of {n_pos} positives"
    if "maker" in g:
        assert int(g["run_red_light"].sum()) > 100 and int(g["passive"].sum()) > 100


def test_rule_checks_in_rollout_vs_oracle():
    """engine(rule_checks=True) vs the oracle's closed loop with the same checks on a dense scene (other seed)."""
    shape = dict(n_sc=1, n_ag=40, n_mp=96, n_tl=30, seed=77, boundary=60.0, scale=0.2)
    R, T = 2, 30
    eng, batch, P, cfg = _engine(shape, R, T, rule_checks=True)
    res = eng.rollout(batch)
    ref = O.rollout(P, cfg, config.derived_sizes(cfg), config.DYNAMICS_CFG, config.ROLLOUT_CFG, batch, R, T,
                    rule_checks=True)
    assert torch.equal(res["pred_valid"].cpu(), ref["pred_valid"])
    tot = 0
    for k in VIO:
        n_diff, n_pos = int((res[k].cpu() != ref[k]).sum()), int(ref[k].sum())
        tot += n_pos
        print(f"{k}: oracle positives {n_pos}, mismatches {n_diff}")
        assert n_diff <= max(2, int(0.01 * n_pos))
    assert tot > 100


@pytest.mark.parametrize("shape,R,T", [
    (dict(n_sc=1, n_ag=37, n_mp=70, n_tl=27, seed=5, boundary=103.0), 3, 15),      # ragged sizes just above the K's
    (dict(n_sc=3, n_ag=26, n_mp=65, n_tl=26, seed=6, boundary=101.0), 1, 13),      # minimum sizes (K < T), 1 rollout
])
@pytest.mark.parametrize("precision", [0, 1])
def test_rollout_ragged_shapes_vs_oracle(shape, R, T, precision):
    """Row counts that are not multiples of any tile (128-row GEMM tiles, 8-token attention CTAs, 64-row select CTAs,
    16-neighbour MMA groups, 8 agents per CTA of the fused history encoder) and target counts one above the KNN sizes
    (K_ag2ag = 25 < 26, K_ag2mp = 64 < 65, K_tl2* = 24), in the fp32 and in the tensor-core mode."""
    eng, batch, P, cfg = _engine(shape, R, T, precision=precision)
    res = eng.rollout(batch)
    ref = O.rollout(P, cfg, config.derived_sizes(cfg), config.DYNAMICS_CFG, config.ROLLOUT_CFG, batch, R, T)
    assert torch.equal(res["pred_valid"].cpu(), ref["pred_valid"])
    assert torch.equal(res["tl_state"].cpu(), ref["tl_state"])
    assert maxerr(res["pred_pose"][..., :2], ref["pred_pose"][..., :2]) < (TC_TOL_XY if precision else TOL_XY)
    assert maxerr(res["pred_pose"][..., 2], ref["pred_pose"][..., 2]) < (TC_TOL_YAW if precision else TOL_YAW)


@pytest.mark.parametrize("precision", [0, 1])
def test_rollout_degenerate_scenes(precision):
    """All agents invalid / all traffic lights invalid / all map polylines invalid: every attention row is fully
    masked -> exact zeros, no NaN; agents that are never valid stay exactly zero (fp32 and tensor-core mode)."""
    shape = dict(n_sc=2, n_ag=30, n_mp=80, n_tl=28, seed=9, boundary=110.0)
    R, T = 2, 12
    eng, batch, P, cfg = _engine(shape, R, T, precision=precision)
    batch["sc/ag_valid"][0] = False          # scene 0: no agent at all
    batch["ag_navi_valid"][0] = False
    batch["ag_latent_valid"][0] = False
    batch["sc/tl_valid"][1] = False          # scene 1: no traffic light
    batch["sc/mp_valid"][1] = False          # ... and no valid map polyline
    batch["map/valid"][1] = False
    res = eng.rollout(batch)
    ref = O.rollout(P, cfg, config.derived_sizes(cfg), config.DYNAMICS_CFG, config.ROLLOUT_CFG, batch, R, T)
    assert bool(torch.isfinite(res["pred_pose"]).all()) and bool(torch.isfinite(res["pred_motion"]).all())
    assert torch.equal(res["pred_valid"].cpu(), ref["pred_valid"])
    assert float(res["pred_pose"][: R].abs().max()) == 0.0
    assert maxerr(res["pred_pose"], ref["pred_pose"]) < (TC_TOL_XY if precision else TOL_XY)


def test_module_api_in_reference_loop(golden_rollout):
    """The `TrafficBots` drop-in (init / forward -> distributions, mp_encoder, tl_encoder.pre_compute) driven by the
    reference's host-side loop (restated in the oracle: Dynamics, TeacherForcing, feedback checks on the CPU) with the
    reference's repeat_interleave of every token tensor — vs the golden rollout of the real reference."""
    from trafficbotsv1_5_b200.traffic_bots import TrafficBots
    g = golden_rollout
    cfg = config.default_model_cfg()
    P = params.init_params(cfg, seed=0)
    model = TrafficBots(cfg)
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected and all(k.endswith(("freqs", "_ohe")) for k in missing)
    model = model.eval().to(DEV)
    batch = synth.make_scene_batch(**g["shape"])
    R, T = g["R"], g["T"]
    gb = {k: v.to(DEV) for k, v in batch.items()}
    mp_tokens = model.mp_encoder(gb["sc/mp_valid"], gb["sc/mp_attr"], gb["sc/mp_pose"], gb["ref/mp_type"])
    tl_tokens = model.tl_encoder.pre_compute(tl_valid=gb["sc/tl_valid"], tl_attr=gb["sc/tl_attr"], tl_pose=gb["sc/tl_pose"],
                                             **mp_tokens)
    mpR = {k: v.repeat_interleave(R, 0) for k, v in mp_tokens.items()}          # waymo_motion.py:458-462
    tlR = {k: v.repeat_interleave(R, 0) for k, v in tl_tokens.items()}

    class Adapter:
        def step(self, valid, pose, motion, ag_attr, ag_type, latent, latent_valid, navi, navi_valid, tl_state, _tl, _mp):
            c = lambda t: t.to(DEV)  # noqa: E731
            act, tl_dist = model(ag_valid=c(valid), ag_pose=c(pose), ag_motion=c(motion), ag_attr=c(ag_attr),
                                 ag_type=c(ag_type), ag_latent=c(latent), ag_latent_valid=c(latent_valid),
                                 ag_navi=c(navi), ag_navi_valid=c(navi_valid), ag_navi_updated=False,
                                 tl_state=c(tl_state), tl_tokens=tlR, mp_tokens=mpR)
            assert abs(float(act.base_dist.scale[0, 0, 0]) - (torch.tensor(-2.0).exp() if bool(valid[0, 0]) else 1.0)) < 1e-6
            return act.mean.cpu(), tl_dist.logits.cpu()

    model.init()
    sz = config.derived_sizes(cfg)
    mp_cpu = dict(mp_token_invalid=mp_tokens["mp_token_invalid"].cpu(), mp_token_feature=mp_tokens["mp_token_feature"].cpu(),
                  mp_token_pose=mp_tokens["mp_token_pose"].cpu())
    tl_cpu = dict(tl_token_invalid=tl_tokens["tl_token_invalid"].cpu(), tl_token_pose=tl_tokens["tl_token_pose"].cpu())
    res = O.rollout(P, cfg, sz, config.DYNAMICS_CFG, config.ROLLOUT_CFG, batch, R, T, mp=mp_cpu, tl=tl_cpu, policy=Adapter())
    assert torch.equal(res["pred_valid"], g["pred_valid"])
    assert torch.equal(res["tl_state"], g["tl_state"])
    assert maxerr(res["pred_pose"][..., :2], g["pred_pose"][..., :2]) < TOL_XY
    assert maxerr(res["action_mean"], g["action_mean"]) < 2e-4


def test_navi_predictor_vs_reference_golden(golden_navi):
    """destination classifier on the GPU (HotPathModel.navi_predictor through tb_linear /
    tb_pose_emb / tb_layernorm / tb_pointnet_pool) vs the real reference's probabilities, fp32 and tensor-core mode;
    then destinations sampled by the engine feed a rollout."""
    g = golden_navi
    cfg = config.default_model_cfg()
    P = params.init_params(cfg, seed=g["param_seed"], with_navi_predictor=True)
    batch = synth.make_scene_batch(**g["shape"])
    for prec, tol in ((0, 1e-5), (1, 2e-3)):  # measured 1e-7 / 8e-5
        eng = RolloutEngine(P, cfg, "cuda", precision=prec, n_rollout=4, step_end=20, use_graph=False)
        out = eng.predict_destinations(batch, deterministic_k0=True)
        probs = out["probs"].cpu()
        assert torch.equal(probs > 0, g["probs"] > 0)
        err = float((probs - g["probs"]).abs().max())
        print(f"navi_predictor precision={prec}: max |dprob| {err:.3e}")
        assert err < tol * max(float(g["probs"].max()), 1e-3)
        assert torch.equal(out["navi_valid"].cpu(), g["valid"])
        top2 = g["probs"].topk(2, -1)[0]
        clear = g["valid"] & (top2[..., 0] - top2[..., 1] > 10 * tol)
        assert torch.equal(out["dest"][:, 0].cpu()[clear], g["dest_argmax"][clear])
        # sampled destinations are candidates of their agent
        d = out["dest"].cpu()
        assert bool((torch.gather(g["probs"][:, None].expand(-1, d.shape[1], -1, -1), 3, d[..., None]) > 0)[
            g["valid"][:, None].expand(-1, d.shape[1], -1)].all())
    b2 = dict(batch)
    b2["agent/dest"], b2["ag_navi_valid"] = out["dest"].cpu(), out["navi_valid"].cpu()
    b2["ag_latent"] = b2["ag_latent"][:, :4]
    res = eng.rollout(b2)
    assert bool(torch.isfinite(res["pred_pose"]).all())


def test_wosac_post_processing_kernels_vs_reference(golden_wosac):
    """SURVEY 8(f) rank 4 on the GPU (tb_future_filter + tb_traj_global through the C ABI) vs the real
    WOSACPostProcessing: identical kept set, scores bit-exact vs the oracle, global trajectories within 1 ulp-level
    tolerance of the reference's (positions ~5e3 m: 1e-3 m; yaw 1e-6 rad)."""
    from trafficbotsv1_5_b200 import ops
    g = golden_wosac
    sh = g["shape"]
    inp = synth.make_wosac_post_inputs(**sh)
    n_sc, K, A, T = sh["n_sc"], sh["K"], sh["A"], sh["T"]
    d = lambda t: t.to(DEV)  # noqa: E731
    score, sel = ops.future_filter(d(inp["collided"]).view(n_sc * K, A, T).contiguous(),
                                   d(inp["run_road_edge"]).view(n_sc * K, A, T).contiguous(),
                                   d(inp["role"].any(-1)).contiguous(), n_sc, K, g["t0"], g["w_road_edge"], g["n_keep"])
    ref_sc = O.wosac_future_scores(inp["collided"], inp["run_road_edge"], inp["role"], g["t0"], g["w_road_edge"])
    assert torch.equal(score.cpu(), ref_sc)
    assert torch.equal(sel.cpu().long(), O.wosac_select_futures(ref_sc, g["n_keep"]))
    assert torch.equal(sel.cpu().long().sort(-1)[0], g["sel"].sort(-1)[0])
    pos, yaw = ops.traj_global(d(inp["pose"]), d(g["sel"].to(torch.int32)).contiguous(), d(inp["center"]), d(inp["yaw"]),
                               n_sc, K, g["t0"])
    assert maxerr(pos, g["pos_sim"]) < 1e-3 and maxerr(yaw, g["yaw_sim"]) < 2e-6
    # ties: equal scores are resolved by the lower future index, exactly like the oracle's definition
    col = torch.zeros(1 * 6, 3, 5, dtype=torch.bool)
    col[1, 0, 3] = col[4, 1, 4] = True
    score, sel = ops.future_filter(d(col), d(torch.zeros_like(col)), d(torch.ones(1, 3, dtype=torch.bool)), 1, 6, 2, 0.0, 4)
    assert score.cpu().tolist() == [[0.0, 1.0, 0.0, 0.0, 1.0, 0.0]] and sel.cpu().tolist() == [[0, 2, 3, 5]]
    # all futures, no selection (R <= n_keep)
    pos_all, _ = ops.traj_global(d(inp["pose"]), None, d(inp["center"]), d(inp["yaw"]), n_sc, K, g["t0"])
    assert torch.equal(pos_all[torch.arange(n_sc)[:, None], d(g["sel"])], pos)


def test_post_process_wosac_after_rollout():
    """Engine-level: 6 rollouts with all rule checks, keep the best 4, global frame; vs the oracle on the engine's own
    flags and trajectories."""
    eng, batch, P, cfg = _engine(dict(n_sc=2, n_ag=48, n_mp=96, n_tl=30, seed=3000, boundary=60.0, scale=0.2), 6, 30,
                                 rule_checks=True)
    res = eng.rollout(batch)
    g = torch.Generator().manual_seed(1)
    batch["ref/ag_role"] = torch.rand(2, 48, 3, generator=g) < 0.3
    batch["scenario_center"] = torch.randn(2, 2, generator=g) * 1000
    batch["scenario_yaw"] = torch.randn(2, generator=g)
    out = eng.post_process_wosac(res, batch, n_keep=4, w_road_edge=0.5)
    c = lambda k: res[k].cpu().view(2, 6, 48, 30)  # noqa: E731
    sc = O.wosac_future_scores(c("collided_wosac"), c("run_road_edge"), batch["ref/ag_role"], 10, 0.5)
    assert torch.equal(out["score"].cpu(), sc) and float(sc.max()) > 0
    sel = O.wosac_select_futures(sc, 4)
    assert torch.equal(out["sel"].cpu().long(), sel)
    trajs = res["pred_pose"].cpu().view(2, 6, 48, 30, 3)[torch.arange(2)[:, None], sel][:, :, :, 10:]
    pos, yaw = O.wosac_to_global(trajs, batch["scenario_center"], batch["scenario_yaw"])
    assert maxerr(out["pos_sim"], pos) < 5e-4 and maxerr(out["yaw_sim"], yaw) < 2e-6
    with pytest.raises(RuntimeError):
        RolloutEngine(P, cfg, DEV, n_rollout=6, step_end=30).post_process_wosac(res, batch, n_keep=4)


def test_rollout_golden_tf32_fp32_intermediates(golden_rollout):
    """precision=1 with `kv_half = False`: tf32 tcgen05 projections, fp32 intermediates, SIMT attention, unfused history
    encoder — the fall-back for checkpoints whose activations exceed the fp16 range."""
    g = golden_rollout
    eng, batch, P, cfg = _engine(g["shape"], g["R"], g["T"], precision=1)
    eng.model.kv_half = False
    res = eng.rollout(batch)
    assert int((res["pred_valid"].cpu() != g["pred_valid"]).sum()) == 0
    assert int((res["tl_state"].cpu() != g["tl_state"]).sum()) == 0
    assert maxerr(res["pred_pose"][..., :2], g["pred_pose"][..., :2]) < TC_TOL_XY
    assert maxerr(res["pred_pose"][..., 2], g["pred_pose"][..., 2]) < TC_TOL_YAW


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["k12", "k32_temp", "k6"])
def test_womd_post_processing_kernel_vs_reference(golden_womd, match_womd_modes, case):
    """tb_womd_post vs the real WOMDPostProcessing (golden) and vs the oracle mode by mode."""
    from trafficbotsv1_5_b200 import ops
    g = golden_womd[case]
    inp = synth.make_womd_post_inputs(**g["shape"])
    trajs, scores, mode = ops.womd_post(inp["trajs"].to(DEV), inp["scores"].to(DEV), inp["ag_type"].to(DEV), g["k_pred"],
                                        True, g["mpa_nms_thresh"], g["score_temperature"], 4, 5, 80)
    match_womd_modes(trajs.cpu(), scores.cpu(), g["trajs"], g["scores"])
    o_trajs, o_scores, o_mode = O.womd_post_processing(inp["ag_type"], inp["trajs"], inp["scores"], g["k_pred"], True,
                                                       g["mpa_nms_thresh"], g["score_temperature"], 80)
    assert torch.equal(mode.cpu().long(), o_mode) and torch.equal(trajs.cpu(), o_trajs)  # same modes, same order
    assert torch.allclose(scores.cpu(), o_scores, rtol=1e-5, atol=1e-7)


@pytest.mark.gpu
def test_womd_post_processing_variants_vs_oracle():
    """Uniform scores (reactive replay: scores=None), FDE instead of ADE, NMS off, and the argument checks."""
    from trafficbotsv1_5_b200 import ops
    inp = synth.make_womd_post_inputs(seed=8100, n_sc=2, K=9, A=13, T=37)
    tr, ty = inp["trajs"].to(DEV), inp["ag_type"].to(DEV)
    for scores, use_ade, thr in [(inp["scores"], False, [2.0, 2.0, 2.0]), (inp["scores"], True, []),
                                 (None, True, [1.0, 2.0, 3.0])]:
        t, s, m = ops.womd_post(tr, None if scores is None else scores.to(DEV), ty, 6, use_ade, thr, -1.0, 4, 5, 80)
        ot, os_, om = O.womd_post_processing(inp["ag_type"], inp["trajs"], scores, 6, use_ade, thr, -1.0, 80)
        assert t.shape == (2, 13, 6, 7, 3)  # steps 4, 9, ..., 34 of the 37 available
        assert torch.equal(m.cpu().long(), om) and torch.equal(t.cpu(), ot)
        assert torch.allclose(s.cpu(), os_, rtol=1e-5, atol=1e-7)
    with pytest.raises(RuntimeError):  # more joint futures than the kernel's 128
        ops.womd_post(torch.zeros(1, 129, 2, 10, 3, device=DEV), None, torch.zeros(1, 2, 3, dtype=torch.bool, device=DEV))
    with pytest.raises(RuntimeError):  # k_pred > 8
        ops.womd_post(tr, None, ty, k_pred=9)


@pytest.mark.gpu
def test_post_process_womd_after_rollout():
    """RolloutEngine.post_process_womd on a real rollout: 8 joint futures -> 6 modes per agent at 2 Hz."""
    eng, batch, P, cfg = _engine(dict(n_sc=2, n_ag=48, n_mp=96, n_tl=30, seed=3100, boundary=60.0), 8, 90)
    res = eng.rollout(batch)
    g = torch.Generator().manual_seed(5)
    scores = torch.randn(2, 8, 48, generator=g)
    out = eng.post_process_womd(res, batch, scores)
    assert out["trajs"].shape == (2, 48, 6, 16, 3) and out["scores"].shape == (2, 48, 6)
    fut = res["pred_pose"].view(2, 8, 48, -1, 3)[:, :, :, 10:].cpu()
    ot, os_, om = O.womd_post_processing(batch["ref/ag_type"], fut, scores, 6, True, [2.0, 2.0, 2.0], -1.0, 80)
    assert torch.equal(out["mode"].cpu().long(), om) and torch.equal(out["trajs"].cpu(), ot)
    assert torch.allclose(out["scores"].cpu(), os_, rtol=1e-5, atol=1e-7)
    assert torch.allclose(out["scores"].sum(-1).cpu(), torch.ones(2, 48), atol=1e-5)


@pytest.mark.gpu
def test_rollout_head_interleaved_layout_is_bit_identical(golden_rollout):
    """The head-interleaved q / K / V rows (tb_knarpe_attn flags bit 4, the tensor-core mode's default) are a
    permutation of projection output features: the whole closed-loop rollout is bit-identical to the natural layout
    (`model.opt["attn_il"] = False`, 128-bit gathers)."""
    g = golden_rollout
    eng, batch, P, cfg = _engine(g["shape"], g["R"], g["T"], precision=1)
    assert eng.model.kv_il
    res = {k: v.clone() for k, v in eng.rollout(batch).items() if torch.is_tensor(v)}
    eng2, _, _, _ = _engine(g["shape"], g["R"], g["T"], precision=1)
    eng2.model.opt["attn_il"] = False
    assert not eng2.model.kv_il
    res2 = eng2.rollout(batch)
    for k in ("pred_pose", "pred_motion", "pred_valid", "tl_state"):
        assert torch.equal(res[k], res2[k]), k


@pytest.mark.gpu
def test_rollout_fused_layernorm_vs_separate(golden_rollout):
    """tb_linear_ln (LayerNorm inside the producing projection's epilogue, one-pass variance) against the separate
    tb_layernorm launches (two-pass): same closed-loop rollout within the tensor-core mode's rounding."""
    g = golden_rollout
    eng, batch, P, cfg = _engine(g["shape"], g["R"], g["T"], precision=1)
    assert eng.model.ln_fused
    res = {k: v.clone() for k, v in eng.rollout(batch).items() if torch.is_tensor(v)}
    n_fused = eng.launches_per_step
    eng2, _, _, _ = _engine(g["shape"], g["R"], g["T"], precision=1)
    eng2.model.opt["ln_fused"] = False
    assert not eng2.model.ln_fused
    res2 = eng2.rollout(batch)
    assert eng2.launches_per_step > n_fused
    assert maxerr(res["pred_pose"][..., :2], res2["pred_pose"][..., :2]) < 5e-3
    assert maxerr(res["pred_pose"][..., 2], res2["pred_pose"][..., 2]) < 2e-3
    assert torch.equal(res["pred_valid"], res2["pred_valid"]) and torch.equal(res["tl_state"], res2["tl_state"])


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("gap", [False, True])
def test_warm_start_dedup_matches_per_rollout_encoding(precision, gap):
    """Warm-start de-duplication (engine._warm_steps): while every valid agent is teacher-forced, the encoder inputs of
    a scene's rollouts are identical, so the agent / TL encoders of steps 1 .. S0 run once per scene as one batch. The
    result must equal the plain per-rollout, per-step path: masks identical, poses within 1e-4 m (fp32) / 2e-3 m
    (16-bit mode: the neighbour order of the batched full-scan select and the step-wise temporal select differ, which
    re-orders the attention sums). With a track gap inside the warm start (an agent valid but NOT forced at step 4)
    only the steps before it are rollout-invariant: S0 shrinks and the rest runs on the regular path."""
    shape = dict(n_sc=2, n_ag=40, n_mp=96, n_tl=30, seed=77, boundary=110.0)
    eng, batch, P, cfg = _engine(shape, 3, 24, precision=precision)
    if gap:
        batch["sc/ag_valid"][0, 5, 4] = False   # valid at t = 3, ground truth missing at t = 4, back at t = 5
        batch["sc/ag_valid"][0, 5, 3] = True
        batch["sc/ag_valid"][0, 5, 5] = True
    eng.prepare(batch)
    assert eng._s0 == (4 if gap else 11), eng._s0
    a = {k: v.clone() for k, v in eng.run().items()}
    eng.warm_dedup = False
    eng.prepare(batch)
    assert eng._s0 == 0
    b = eng.run()
    for k in ("pred_valid", "tl_state", "final_valid", "final_navi_valid"):
        assert torch.equal(a[k], b[k]), k
    tol = 1e-4 if precision == 0 else 2e-3
    assert maxerr(a["pred_pose"], b["pred_pose"]) < tol, maxerr(a["pred_pose"], b["pred_pose"])
    assert maxerr(a["pred_motion"], b["pred_motion"]) < 5 * tol  # accelerations / yaw rates of up to 7 per second


@pytest.mark.parametrize("precision", [0, 1])
def test_agent_compaction_matches_padded_run(precision):
    """Agent compaction (engine._compact): slots that are invalid at every ground-truth step can never become valid, so
    they are dropped for the whole rollout and results() scatters back to the caller's agent order. Against the same
    engine with the padding kept: masks identical, dropped slots exactly zero, poses within 1e-5 m (fp32; the KNN lists
    are the same sets in the same order, only the row tiling of the GEMMs changes) / 2e-3 m (16-bit mode)."""
    shape = dict(n_sc=2, n_ag=64, n_mp=96, n_tl=30, seed=91, boundary=110.0)
    eng, batch, P, cfg = _engine(shape, 3, 20, precision=precision, record_feedback=True)
    batch["sc/ag_valid"][:, 40:] = False            # 24 padded slots at the end ...
    batch["sc/ag_valid"][0, 3] = False              # ... and holes in the middle (per scene different)
    batch["sc/ag_valid"][1, 7] = False
    batch["sc/ag_valid"][1, 11] = False
    batch["ag_navi_valid"] = batch["sc/ag_valid"].any(-1)
    batch["ag_latent_valid"] = batch["sc/ag_valid"].any(-1)
    eng.prepare(batch)
    kept = (max(int(batch["sc/ag_valid"].any(-1).sum(1).max()), 26) + 3) // 4 * 4   # largest valid count, rounded up to 4
    assert eng._perm is not None and eng._st["A"] == kept < 44 and eng._A_full == 64
    a = {k: v.clone() for k, v in eng.run().items()}
    eng.compact_agents = False
    eng.prepare(batch)
    assert eng._perm is None and eng._st["A"] == 64
    b = eng.run()
    for k in ("pred_valid", "tl_state", "final_valid", "final_navi_valid", "outside_map", "dest_reached"):
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k
    never = ~batch["sc/ag_valid"].any(-1).repeat_interleave(3, 0).to(DEV)
    assert float(a["pred_pose"][never].abs().max()) == 0.0
    tol = 1e-5 if precision == 0 else 2e-3
    assert maxerr(a["pred_pose"], b["pred_pose"]) < tol, maxerr(a["pred_pose"], b["pred_pose"])
    assert a["joint_pose"].shape == b["joint_pose"].shape


def test_engine_keeps_state_and_graphs_per_batch_shape():
    """Batches with different numbers of kept agent slots alternate (agent compaction): the engine keeps state buffers
    and captured step graphs per shape, so coming back to a shape neither re-captures nor changes the results."""
    shape = dict(n_sc=2, n_ag=64, n_mp=96, n_tl=30, seed=93, boundary=110.0)
    eng, b1, P, cfg = _engine(shape, 2, 14, precision=1)
    b2 = {k: v.clone() for k, v in b1.items()}
    b2["sc/ag_valid"][:, 36:] = False
    b2["ag_navi_valid"] = b2["ag_latent_valid"] = b2["sc/ag_valid"].any(-1)
    keep = ("pred_valid", "pred_pose", "tl_state")
    r1 = {k: eng.rollout(b1)[k].clone() for k in keep}
    g1, a1 = eng._graph, eng._st["A"]
    r2 = {k: eng.rollout(b2)[k].clone() for k in keep}
    g2, a2 = eng._graph, eng._st["A"]
    assert a1 != a2 and g1 is not g2
    for _ in range(2):
        out = eng.rollout(b1)
        assert eng._graph is g1 and eng._st["A"] == a1
        assert all(torch.equal(out[k], r1[k]) for k in keep)
        out = eng.rollout(b2)
        assert eng._graph is g2 and eng._st["A"] == a2
        assert all(torch.equal(out[k], r2[k]) for k in keep)
