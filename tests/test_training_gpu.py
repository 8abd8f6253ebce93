"""GPU parity tests of the training path (BASELINE config 4): every backward kernel against
torch autograd of the same op, and the whole training step (loss terms + gradients of every parameter tensor) against
autograd THROUGH the CPU oracle restatement of `WaymoMotion.training_step` (oracle/tb_oracle_train.py).
Tolerances are written next to each check (fp32 arithmetic in a different operation order)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import tb_oracle as O
from oracle import tb_oracle_train as OT
from trafficbotsv1_5_b200 import config, params, synth

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from trafficbotsv1_5_b200 import autograd as AG
    from trafficbotsv1_5_b200 import lib as L
    from trafficbotsv1_5_b200 import ops
    from trafficbotsv1_5_b200.training import TRAIN_CFG, TrainStep

DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.mark.parametrize("M,N,K,strided", [(1000, 128, 128, False), (777, 896, 128, True), (4096, 6, 384, False),
                                           (513, 5, 128, False), (300, 64, 20, False), (9000, 128, 640, True)])
def test_linear_wgrad_vs_torch(M, N, K, strided):
    """dW = dY^T X, db = colsum(dY) (tb_linear_wgrad, split over M with fp32 atomics): 2e-5 of the largest entry."""
    g = torch.Generator().manual_seed(M + N)
    dy_full = torch.randn(M, N + (8 if strided else 0), generator=g).to(DEV)
    x_full = torch.randn(M, K + (4 if strided else 0), generator=g).to(DEV)
    dy, x = dy_full[:, :N], x_full[:, :K]
    dw, db = AG.wgrad(dy, x, True)
    assert rel(dw, dy.double().t() @ x.double()) < 2e-5
    assert rel(db, dy.double().sum(0)) < 2e-5
    dw2, _ = AG.wgrad(dy, x, False)  # accumulation semantics: a second call into the same buffer doubles it
    L.check(L.load().tb_linear_wgrad(L.ptr(dy), dy.stride(0), L.ptr(x), x.stride(0), M, N, K, L.ptr(dw2), K, None, 0,
                                     L.stream()), "tb_linear_wgrad")
    assert rel(dw2, 2 * (dy.double().t() @ x.double())) < 2e-5
    # tcgen05 path (precision 1: operands read as tf32 - 10-bit mantissas - fp32 accumulate): 2e-3 of the largest entry
    dw3, db3 = AG.wgrad(dy, x, True, precision=1)
    assert rel(dw3, dy.double().t() @ x.double()) < 2e-3
    assert rel(db3, dy.double().sum(0)) < 2e-5


@pytest.mark.parametrize("relu,mask_pre,res,mask_post,group", [(False, False, False, False, 0), (True, False, False, False, 0),
                                                               (True, True, False, False, 0), (False, True, True, True, 0),
                                                               (True, True, True, False, 0), (True, False, False, True, 4),
                                                               (True, False, True, True, 0)])
@pytest.mark.parametrize("precision", [0, 1])
def test_linear_autograd_vs_torch(relu, mask_pre, res, mask_post, group, precision):
    """ops.linear routed through autograd._Linear: gradients w.r.t. x, W, bias (plain / grouped), residual.
    fp32 FFMA 2e-5; tf32 tensor-core GEMMs (precision 1) 3e-3 (10-bit mantissa operands)."""
    M, N, K = 512, 128, 256
    g = torch.Generator().manual_seed(7)
    x = torch.randn(M, K, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV).requires_grad_(True)
    b = torch.randn(M // group if group else 1, N, generator=g).to(DEV)
    b = (b if group else b[0]).contiguous().requires_grad_(True)
    r = torch.randn(M, N, generator=g).to(DEV).requires_grad_(True) if res else None
    mp = (torch.rand(M, generator=g) < 0.3).to(DEV) if mask_pre else None
    mq = (torch.rand(M, generator=g) < 0.3).to(DEV) if mask_post else None
    go = torch.randn(M, N, generator=g).to(DEV)
    y = ops.linear(x, w, b, relu=relu, mask_pre=mp, res=r, mask_post=mq, precision=precision, bias_group=group)
    y.backward(go)
    got = [t.grad.clone() for t in (x, w, b) + ((r,) if res else ())]
    for t in (x, w, b) + ((r,) if res else ()):
        t.grad = None
    bias = b.repeat_interleave(group, 0) if group else b
    v = F.linear(x, w) + bias
    if relu:
        v = v.relu()
    if mp is not None:
        v = v.masked_fill(mp[:, None], 0.0)
    if r is not None:
        v = v + r
    if mq is not None:
        v = v.masked_fill(mq[:, None], 0.0)
    tol = 2e-5 if precision == 0 else 3e-3
    assert rel(y, v) < tol
    v.backward(go)
    # precision 1: a tf32 pre-activation next to zero can land on the other side of the ReLU than the fp32 reference's,
    # which changes single entries of the gradients by O(1) - compare in the L2 norm there (2e-2), entrywise in fp32
    def err(a, t):
        return rel(a, t) if precision == 0 else float((a - t).norm() / t.norm())
    for a, t, nm in zip(got, (x, w, b) + ((r,) if res else ()), ("dx", "dw", "db", "dres")):
        assert err(a, t.grad) < (tol if precision == 0 else 2e-2), (nm, err(a, t.grad))


@pytest.mark.parametrize("D,relu", [(128, False), (256, False), (128, True)])
def test_layernorm_backward_vs_torch(D, relu):
    M = 3001
    g = torch.Generator().manual_seed(D)
    x = (torch.randn(M, D, generator=g) * 2 + 0.5).to(DEV).requires_grad_(True)
    ga = (1 + 0.1 * torch.randn(D, generator=g)).to(DEV).requires_grad_(True)
    be = (0.1 * torch.randn(D, generator=g)).to(DEV).requires_grad_(True)
    go = torch.randn(M, D, generator=g).to(DEV)
    y = ops.layernorm(x, ga, be, relu=relu)
    y.backward(go)
    got = [t.grad.clone() for t in (x, ga, be)]
    for t in (x, ga, be):
        t.grad = None
    ref = F.layer_norm(x, (D,), ga, be)
    if relu:
        ref = ref.relu()
    assert rel(y, ref) < 1e-5
    ref.backward(go)
    for a, t in zip(got, (x, ga, be)):
        assert rel(a, t.grad) < 2e-5  # fp32 atomics over 3001 rows for gamma / beta


@pytest.mark.parametrize("mode", [1, 2])
def test_pointnet_pool_backward_vs_torch(mode):
    G, Lg, C = 257, 11, 64
    g = torch.Generator().manual_seed(mode)
    x = torch.randn(G * Lg, C, generator=g).relu().to(DEV).requires_grad_(True)
    inv = (torch.rand(G * Lg, generator=g) < 0.3)
    inv.view(G, Lg)[5] = True  # a group without valid rows
    inv = inv.to(DEV)
    out = ops.pointnet_pool(x, inv, G, Lg, mode)
    go = torch.randn_like(out)
    out.backward(go)
    got, x.grad = x.grad.clone(), None
    xm = x.view(G, Lg, C).masked_fill(inv.view(G, Lg, 1), float("-inf")).amax(1)
    xm = torch.where(torch.isfinite(xm), xm, torch.zeros_like(xm))
    ref = torch.cat([xm] * mode, 1)
    assert rel(out, ref) < 1e-6
    # positive maxima are unique (continuous values): compare where the pooled value is > 0; at 0 (ReLU ties) torch
    # spreads the gradient evenly while the kernel gives it to the first row - both are dropped by the ReLU backward
    ref.backward(go)
    pos = (xm > 0).repeat_interleave(Lg, 0)
    assert rel(got[pos], x.grad[pos]) < 1e-6


def _il_case(n_sc=2, A=40, T=30, n_gt=31, seed=3):
    g = torch.Generator().manual_seed(seed)
    b = synth.make_train_batch(n_sc, n_ag=A, n_mp=70, n_tl=27, n_step=n_gt, seed=4000 + seed)
    gt_valid, gt_pose, gt_motion = b["gt/ag_valid"], b["gt/ag_pose"], b["gt/ag_motion"]
    tf = OT.teacher_forcing_mask_training(gt_valid, 10, 10, b["tf/forcing_agent"])
    act = torch.randn(T, n_sc * A, 6, generator=g)
    return b, gt_valid, gt_pose, gt_motion, tf, act


def test_il_loss_forward_backward_vs_autograd():
    """tb_il_loss_fwd / _bwd (state recurrence + imitation loss + reverse scan) against torch autograd through the
    oracle's dynamics_update with the same masks: loss 1e-5, d/d(action head output) 1e-4 of its largest entry."""
    n_sc, A, T, n_gt = 2, 40, 30, 31
    b, gt_valid, gt_pose, gt_motion, tf, act = _il_case(n_sc, A, T, n_gt)
    ag_type = b["ref/ag_type"]
    dyn, tc = config.DYNAMICS_CFG, dict(w_pos=0.1, w_rot=10.0, w_spd=0.1)
    # reference: recurrence with validity following teacher forcing (no outside-map disabling here)
    a_ref = act.clone().requires_grad_(True)
    valid, pose, motion = gt_valid[:, :, 0], gt_pose[:, :, 0], gt_motion[:, :, 0]
    tot, cnt, pv = 0.0, 0, []
    for s in range(1, T + 1):
        br = a_ref[s - 1].view(n_sc, A, 3, 2)
        mean = (br * ag_type[..., None]).sum(2)
        pv.append(valid)
        pose, motion = O.dynamics_update(pose, motion, valid, ag_type, mean, dyn)
        if s - 1 >= 10:
            ok = valid & gt_valid[:, :, s]
            e = tc["w_pos"] * F.smooth_l1_loss(gt_pose[:, :, s, :2], pose[..., :2], reduction="none").sum(-1) \
                + tc["w_rot"] * 0.5 * (1 - torch.cos(gt_pose[:, :, s, 2] - pose[..., 2])) \
                + tc["w_spd"] * F.smooth_l1_loss(gt_motion[:, :, s, 0], motion[..., 0], reduction="none")
            tot = tot + e.masked_fill(~ok, 0.0).sum()
            cnt += int(ok.sum())
        ov = tf[:, :, s]
        valid = valid | ov
        pose = torch.where(ov[..., None], gt_pose[:, :, s], pose)
        motion = torch.where(ov[..., None], gt_motion[:, :, s], motion)
    (2.5 * tot).backward()
    pred_valid = torch.stack(pv, 2)
    to = lambda t: t.to(DEV).contiguous()  # noqa: E731
    rec = dict(B=n_sc, A=A, ag_type=to(ag_type.to(torch.uint8)), pred_valid=to(pred_valid.reshape(n_sc * A, T)),
               pose0=to(gt_pose[:, :, 0]), motion0=to(gt_motion[:, :, 0]), gt_valid=to(gt_valid.to(torch.uint8)),
               gt_pose=to(gt_pose), gt_motion=to(gt_motion), tf_mask=to(tf.to(torch.uint8)), n_gt=n_gt, sc_div=1)
    a_gpu = act.to(DEV).requires_grad_(True)
    out, _ = AG.il_loss(a_gpu, rec, dyn, (tc["w_pos"], tc["w_rot"], tc["w_spd"]), 10)
    (2.5 * out[0]).backward()
    assert abs(float(out[0]) - float(tot)) < 1e-5 * abs(float(tot)) and int(out[1]) == cnt
    assert rel(a_gpu.grad, a_ref.grad) < 1e-4


def test_tl_nll_vs_torch():
    T, n, n_gt = 20, 70, 15
    g = torch.Generator().manual_seed(1)
    logits = (torch.randn(T, n, 5, generator=g) * 3).requires_grad_(True)  # some entries beyond the +-3 clamp
    inv = torch.rand(n, generator=g) < 0.2
    gt = F.one_hot(torch.randint(0, 5, (n, n_gt), generator=g), 5).bool()
    tot, cnt = 0.0, 0
    for s in range(1, T + 1):
        if s < n_gt:
            lp = torch.log_softmax(logits[s - 1].clamp(-3, 3), -1).gather(-1, gt[:, s].max(-1)[1][:, None]).squeeze(-1)
            tot = tot - lp.masked_fill(inv, 0.0).sum()
            cnt += int((~inv).sum())
    tot.backward()
    lg = logits.detach().to(DEV).requires_grad_(True)
    out = AG.tl_nll(lg, inv.to(DEV), gt.to(torch.uint8).to(DEV), n_gt)
    out[0].backward()
    assert abs(float(out[0]) - float(tot)) < 1e-5 * float(tot) and int(out[1]) == cnt
    assert rel(lg.grad, logits.grad) < 1e-5


def _grad_report(ts, Pg, tol, floor):
    bad, worst = [], ("", 0.0)
    for k, ref in Pg.items():
        got = ts.params[k].grad
        if ref.grad is None:
            assert got is None or float(got.abs().max()) < floor, f"{k}: gradient where the oracle has none"
            continue
        assert got is not None, f"{k}: no gradient"
        gr, gg = ref.grad.double(), got.double().cpu()
        scale = float(gr.abs().max())
        err = float((gg - gr).abs().max())
        r = err / max(scale, floor)
        if r > worst[1]:
            worst = (k, r)
        if r > tol:
            bad.append((k, r, scale))
    print(f"worst relative gradient error {worst[1]:.2e} at {worst[0]}")
    assert not bad, bad[:10]
    return worst


def _oracle_f64(P, cfg, sz, tc, b, n_steps):
    """The oracle evaluated in float64 (the fp32 oracle's own gradients are 2-3e-3 off it on the traffic-light encoder,
    whose gradient is a small sum with cancellations through the agents' cross-attention)."""
    Pg = {k: v.clone().double().requires_grad_(True) for k, v in P.items()}
    bb = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in b.items()}
    torch.set_default_dtype(torch.float64)
    try:
        ref = OT.training_step(Pg, cfg, sz, config.DYNAMICS_CFG, tc, bb, n_steps=n_steps)
        ref["loss"].backward()
    finally:
        torch.set_default_dtype(torch.float32)
    return ref, Pg


@pytest.mark.parametrize("variant", ["full", "no_latent_encoder", "prior_rollout_kl", "padded_agents"])
def test_training_step_vs_oracle_autograd(variant):
    """The whole training_step body (map / TL / latent posterior / destination predictor / 14-step teacher-forced closed
    loop / TrainingMetrics) on the CUDA path against the oracle evaluated in float64: every loss term within 2e-4
    relative, the gradient of EVERY parameter tensor within 5e-3 of that tensor's largest gradient entry (absolute floor
    1e-4 for the tensors whose whole gradient is smaller - the posterior's TL encoder sits at 1e-5; the fp32 oracle itself is 3e-3 off the float64 one)."""
    cfg = config.default_model_cfg()
    sz = config.derived_sizes(cfg)
    P = params.init_params(cfg, 0, with_navi_predictor=True, with_latent_post=variant != "no_latent_encoder")
    b = synth.make_train_batch(2, n_ag=40 if variant == "padded_agents" else 28, n_mp=70, n_tl=27, seed=3000,
                               boundary=120.0)
    if variant == "padded_agents":  # never-valid slots (scattered): the CUDA path drops them (TrainStep._compact), the
        for k in ("gt/ag_valid", "sc/ag_valid"):  # oracle carries them as masked rows - same loss, same gradients
            b[k][0, 2::3] = False
            b[k][1, 1::3] = False
        b["ag_navi_valid"] = b["ag_latent_valid"] = b["gt/ag_valid"].any(-1)
    tc = dict(TRAIN_CFG)
    if variant == "prior_rollout_kl":  # rollout on the prior sample; free nats below the KL so its gradient is exercised
        b["rollout_prior"] = True
        tc["kl_free_nats"] = 0.01
    n_steps = 14
    ref, Pg = _oracle_f64(P, cfg, sz, tc, b, n_steps)
    ts = TrainStep(P, cfg, DEV, precision=0, train_cfg=tc)
    out = ts.step(b, n_steps=n_steps)
    torch.cuda.synchronize()
    if variant == "padded_agents":
        assert ts.eng._st["A"] == 28  # 40 slots, at most 27 ever valid per scene -> 28 kept
    assert torch.equal(out["pred_valid"].cpu().view_as(ref["pred_valid"]), ref["pred_valid"])
    assert float((out["pred_pose"].cpu().double().view_as(ref["pred_pose"]) - ref["pred_pose"]).abs().max()) < 1e-3
    for k in ("diffbar_reward", "tl_state_loss", "vae_kl", "navi_loss", "loss"):
        if k in ref:
            assert abs(float(out[k]) - float(ref[k])) < 2e-4 * max(1.0, abs(float(ref[k]))), (k, float(out[k]), float(ref[k]))
    _grad_report(ts, Pg, 5e-3, 1e-4)
    # a second step on the same object (fresh graph, re-packed weights) reproduces the gradients
    g1 = {k: v.grad.clone() for k, v in ts.params.items() if v.grad is not None}
    ts.zero_grad()
    ts.step(b, n_steps=n_steps)
    for k, v in g1.items():  # fp32 atomics: the summation order differs from run to run; the floor covers the scalar
        # bias whose gradient cancels to exactly zero (softmax shift invariance) and comes out as +-2e-8 of rounding
        assert float((ts.params[k].grad - v).abs().max()) < 1e-3 * max(float(v.abs().max()), 1e-4), k


def test_training_step_tf32_mode():
    """The benchmarked training mode (precision 1: tf32 tcgen05 GEMMs for forward and data gradients, fp32 rows, fp32
    FFMA weight gradients) against the float64 oracle: loss terms within 1e-2 relative; the gradient of every module
    (all its parameter tensors concatenated) has cosine similarity > 0.995 with the oracle's and a norm within 5 %
    (10-bit-mantissa operands flip ReLU masks next to zero, so single entries are not compared)."""
    cfg = config.default_model_cfg()
    sz = config.derived_sizes(cfg)
    P = params.init_params(cfg, 0, with_navi_predictor=True, with_latent_post=True)
    b = synth.make_train_batch(2, n_ag=28, n_mp=70, n_tl=27, seed=3000, boundary=120.0)
    tc = dict(TRAIN_CFG)
    ref, Pg = _oracle_f64(P, cfg, sz, tc, b, 14)
    ts = TrainStep(P, cfg, DEV, precision=1, train_cfg=tc)
    out = ts.step(b, n_steps=14)
    assert torch.equal(out["pred_valid"].cpu().view_as(ref["pred_valid"]), ref["pred_valid"])
    for k in ("diffbar_reward", "tl_state_loss", "vae_kl", "navi_loss", "loss"):
        assert abs(float(out[k]) - float(ref[k])) < 1e-2 * max(1.0, abs(float(ref[k]))), (k, float(out[k]), float(ref[k]))
    groups = {}
    for k, v in Pg.items():
        if v.grad is not None:
            groups.setdefault(".".join(k.split(".")[:2]), []).append(k)
    for gname, keys in groups.items():
        a = torch.cat([ts.params[k].grad.double().cpu().reshape(-1) for k in keys])
        r = torch.cat([Pg[k].grad.reshape(-1) for k in keys])
        cos = float((a * r).sum() / (a.norm() * r.norm() + 1e-300))
        ratio = float(a.norm() / (r.norm() + 1e-300))
        assert cos > 0.995 and abs(ratio - 1) < 0.05, (gname, cos, ratio)
