"""Training step of the hot path (BASELINE config 4): the body of
`WaymoMotion.training_step` (src/pl_modules/waymo_motion.py:313-385) — map encoder, traffic-light pre-compute, latent
posterior / prior, destination predictor, the 90-step teacher-forced closed loop (`reactive_replay` -> `rollout`
:386-437, 206-311) and `TrainingMetrics` (models/metrics/training.py:76-189) — forward AND backward on the CUDA
kernels of this library.

How it maps onto the B200 path:
  * forward = the same kernel sequences as inference (`model.HotPathModel`, fp32 rows; precision 0 FFMA or 1 tf32
    tcgen05 GEMMs) with `ops.*` routed to their differentiable forms (`autograd.py`), activations of all 90 steps
    kept in HBM (config 4: ~60 GB of 180 GB) instead of being recomputed;
  * the policy inputs are detached in training (`training_detach_model_input`, waymo_motion.py:158-161), so the rollout
    STATE (history rings, validity, teacher forcing, outside-map / destination feedback) is advanced by the inference
    kernels `tb_dyn_step` / `tb_tl_step` in place, and the only gradient path through time — the unicycle recurrence —
    is one reverse-scan kernel (`tb_il_loss_bwd`) that turns the imitation loss into dL/d(action head output) of every
    step; each step's network backward is then independent;
  * gradients w.r.t. the re-associated attention matrices are mapped back onto the reference's
    parameters by autograd through `model.fuse_attention(differentiable=True)`.
Dropout (tf_cfg.dropout_p etc.) is not implemented: this is the p = 0 configuration of BASELINE config 4.
The Bernoulli draws of the step (teacher-forced agents, latent noise, prior-vs-posterior rollout) are inputs.
"""
import copy
from typing import Dict, Optional

import torch
from torch import Tensor

from . import autograd as AG
from . import config as C
from . import lib as L
from . import ops
from .engine import RolloutEngine, teacher_forcing_mask
from .model import HotPathModel, fuse_attention

# configs/model/sim_agent.yaml:140-152 (teacher_forcing_training), :180-216 (differentiable_reward, training_metrics)
TRAIN_CFG = dict(step_spawn_agent=10, step_warm_start=10, w_pos=0.1, w_rot=10.0, w_spd=0.1, w_vae_kl=1.0,
                 kl_balance_scale=0.2, kl_free_nats=1.0, w_diffbar_reward=1.0, w_navi=1.0, w_tl_state=1.0,
                 step_training_start=10, time_step_end=90)


class TrainModel(HotPathModel):
    """HotPathModel over LEAF parameters that require grad: the packed attention matrices and the other derived
    weights are rebuilt (differentiably, on the device) by `refresh()` once per training step."""

    def __init__(self, params: Dict[str, Tensor], cfg: dict, sizes: dict, device="cuda", precision: int = 0):
        assert precision in (0, 1)
        self.cfg, self.sz, self.dev, self.precision = cfg, sizes, torch.device(device), precision
        self.d, self.W = cfg["hidden_dim"], cfg["temp_window_size"]
        self.kv_half = False  # fp32 rows everywhere; precision 1 only switches the GEMMs to tf32 tcgen05
        self.fp32_tc = False  # precision 0 of the training path is the FFMA kernel (forward and backward)
        self.P = params
        self.detach_tl_feature = cfg["tl_state_predictor"]["detach_tl_feature"]
        self.freq_rpe = ops.pe_freq_xy(self.d, cfg["pose_rpe"]["theta_xy"], self.dev)
        self.freq_ag = ops.pe_freq_xy(self.d // 2, cfg["ag_encoder"]["pose_emb"]["theta_xy"], self.dev)
        self.refresh()

    def refresh(self) -> None:
        P = self.P
        self.fa = {k[: -len(".in_proj_weight")]: fuse_attention(P, k[: -len(".in_proj_weight")], self.d, differentiable=True)
                   for k in P if k.endswith(".in_proj_weight")}
        if "action_head.mlp_mean.0.fc_layers.0.weight" in P:
            self.act_w0 = torch.cat([P[f"action_head.mlp_mean.{t}.fc_layers.0.weight"] for t in range(3)], 0)
            self.act_b0 = torch.cat([P[f"action_head.mlp_mean.{t}.fc_layers.0.bias"] for t in range(3)], 0)
            self.act_w4 = torch.block_diag(*[P[f"action_head.mlp_mean.{t}.fc_layers.4.weight"] for t in range(3)])
            self.act_b4 = torch.cat([P[f"action_head.mlp_mean.{t}.fc_layers.4.bias"] for t in range(3)], 0)
        self.drop_caches()

    def drop_caches(self) -> None:
        for c in ("_pn_split", "_w_il", "_w_half", "_chain", "_ag_front"):
            if hasattr(self, c):
                delattr(self, c)

    # functional head chain (no persistent cat buffers: every step's activations stay alive for the backward)
    def heads_train(self, x: Tensor, pose: Tensor, navi: dict, navi_inv: Tensor, lat_feat: Tensor, lat_inv: Tensor
                    ) -> Tensor:
        """navi_encoder (per-step half) -> add_navi -> add_latent -> action-head branches (traffic_bots.py:191-217,
        navigation.py:73-79, add_navi_latent.py:46-64, action_head.py:78-82). x [M, d] -> act_branch [M, 6]."""
        d = self.d
        pe = ops.pose_emb(navi["pose"], self.freq_rpe, d, frame=pose, frame_div=1)
        nf = self.lin(pe, "navi_encoder.mlp_pe.fc_layers.0", res=navi["feat"])
        a = self.mlp(nf, "add_navi.mlp_in", (0, 3, 6), True, mask_post=navi_inv)
        x2 = self.mlp(torch.cat([x, a], 1), "add_navi.mlp", (0, 3, 6), True, mask_pre=navi_inv, res=x)
        x3 = self.mlp(torch.cat([x2, lat_feat], 1), "add_latent.mlp", (0, 3, 6), True, mask_pre=lat_inv, res=x2)
        h0 = ops.linear(x3, self.act_w0, self.act_b0, relu=True, precision=self.precision)
        h1 = torch.cat([self.lin(h0[:, t * d:(t + 1) * d], f"action_head.mlp_mean.{t}.fc_layers.2", relu=True)
                        for t in range(3)], 1)
        return ops.linear(h1, self.act_w4, self.act_b4, precision=self.precision)


def _remap(params: Dict[str, Tensor], mapping: Dict[str, str]) -> Dict[str, Tensor]:
    """View of `params` under other names (same leaf tensors): e.g. latent_encoder.ag_encoder_post.* -> ag_encoder.*"""
    out = {}
    for k, v in params.items():
        for src, dst in mapping.items():
            if k.startswith(src):
                out[dst + k[len(src):]] = v
    return out


def latent_window(cfg: dict) -> int:
    """latent_encoder.py:33-37: number of (down-sampled) steps the posterior / prior encoders see."""
    rate = cfg["latent_encoder"]["temporal_down_sample_rate"]
    n = cfg["time_step_gt"] + 1
    return n // rate + 1 if rate > 1 else n


class TrainStep:
    """loss + gradients of one training step. `params`: the reference's state_dict names (hot-path modules, optionally
    `latent_encoder.{tl,ag}_encoder_post.*` + `latent_encoder.latent_dist_post.*` and `navi_predictor.*`)."""

    def __init__(self, P: Dict[str, Tensor], cfg: Optional[dict] = None, device="cuda", precision: int = 0,
                 train_cfg: Optional[dict] = None, dynamics_cfg: Optional[dict] = None, share_leaves: bool = False):
        """`share_leaves`: use the given tensors themselves as the leaves (an nn.Module's CUDA fp32 parameters: their
        `.grad` is then what a torch optimiser over the module updates) instead of private device copies."""
        L.load()
        self.cfg = cfg or C.default_model_cfg()
        self.sz = C.derived_sizes(self.cfg)
        self.dev = torch.device(device)
        self.tc = dict(TRAIN_CFG, **(train_cfg or {}))
        self.T = self.tc["time_step_end"]
        if share_leaves:
            for k, v in P.items():
                if not (v.is_cuda and v.dtype == torch.float32 and v.is_contiguous() and v.is_leaf):
                    raise ValueError(f"share_leaves: {k} must be a contiguous fp32 CUDA leaf tensor")
                v.requires_grad_(True)
            self.params = dict(P)
        else:
            self.params = {k: v.detach().to(self.dev, torch.float32).contiguous().requires_grad_(True)
                           for k, v in P.items()}
        main = {k: v for k, v in self.params.items() if not k.startswith("latent_encoder.")}
        self.eng = RolloutEngine({k: v.detach() for k, v in main.items() if not k.startswith("navi_predictor.")},
                                 self.cfg, device, precision=0, n_rollout=1, step_end=self.T, use_graph=False,
                                 dynamics_cfg=dynamics_cfg)
        self.eng.tf_steps = (self.tc["step_spawn_agent"], self.tc["step_warm_start"])
        self.model = TrainModel(main, self.cfg, self.sz, device, precision)
        self.eng.model = self.model
        self.post = None
        if any(k.startswith("latent_encoder.ag_encoder_post.") for k in self.params):
            cfg_l = copy.deepcopy(self.cfg)
            cfg_l["temp_window_size"] = latent_window(self.cfg)
            self.post = TrainModel(_remap(self.params, {"latent_encoder.tl_encoder_post.": "tl_encoder.",
                                                        "latent_encoder.ag_encoder_post.": "ag_encoder.",
                                                        "latent_encoder.latent_dist_post.": "latent_dist."}),
                                   cfg_l, self.sz, device, precision)
        self.has_navi = any(k.startswith("navi_predictor.") for k in self.params)
        self.bucket = None
        self.compact_agents = True

    def zero_grad(self) -> None:
        if self.bucket is not None:
            self.bucket.zero()
        else:
            for p in self.params.values():
                p.grad = None

    def data_parallel(self) -> None:
        """Keep all gradients in one flat bucket and average them over the ranks at the end of every step
        (parallel.GradBucket: one NCCL all-reduce; the reference trains with Lightning DDP)."""
        from .parallel import GradBucket
        self.bucket = GradBucket(self.params)

    # ------------------------------------------------------------------------------------------ latent posterior
    def _posterior(self, g, mp: dict) -> Tensor:
        """LatentEncoder.forward(posterior=True) (latent_encoder.py:56-122) + DistEncoder diag_gaus (:222-233):
        the TL and agent encoders with their own weights over the ground-truth future, every 5th step."""
        m, rate = self.post, self.cfg["latent_encoder"]["temporal_down_sample_rate"]
        W = m.W
        v, po, mo = (g(k)[:, :, ::rate].contiguous() for k in ("gt/ag_valid", "gt/ag_pose", "gt/ag_motion"))
        tls = g("gt/tl_state")[:, :, ::rate].contiguous()
        n_sc, A, n_step = v.shape
        assert n_step == W, (n_step, W)
        mp_det = dict(mp, mp_token_feature=mp["mp_token_feature"].detach())
        tl = m.tl_pre_compute(g("gt/tl_valid") if "gt/tl_valid" in self._batch else g("sc/tl_valid"), g("sc/tl_attr"),
                              g("sc/tl_pose"), mp_det)
        d_step = torch.full((1,), W, dtype=torch.int32, device=self.dev)  # full window: slot = window position
        tl_feat, _ = m.tl_forward(tls.to(torch.uint8), d_step, tl, with_logits=False)
        st = dict(B=n_sc, A=A, hist_valid=v.to(torch.uint8), hist_pose=po, hist_motion=mo,
                  ag_attr=g("sc/ag_attr").contiguous(), d_step=d_step)
        kv_mp = m.ag_static(mp)
        x = m.ag_forward(st, mp, kv_mp, tl, tl_feat, 1)
        valid = v.any(-1)
        return m.mlp(x, "latent_dist.mlp_mean", (0, 2, 4), False, mask_post=~valid.reshape(-1)), valid   # :232

    # ------------------------------------------------------------------------------------------ one training step
    def step(self, batch: Dict[str, Tensor], n_steps: Optional[int] = None, backward: bool = True) -> Dict[str, Tensor]:
        """batch: "sc/*", "gt/*", "ref/*", "map/*" as SceneCentricPreProcessing hands them to training_step
        (scene_centric.py:39-147) plus the step's random draws: "tf/forcing_agent" [n_sc, n_ag] bool
        (teacher_forcing.py:87-92), "ag_latent_eps" [n_sc, n_ag, latent_dim] (rsample noise), "rollout_prior" bool.
        Returns the loss terms (training.py:162-189); gradients are left in `self.params[k].grad`."""
        m, eng, tc, dev = self.model, self.eng, self.tc, self.dev
        T = n_steps or self.T
        assert T <= self.T
        batch, perm, A_full = self._compact(batch)
        self._batch = batch
        g = lambda k: batch[k].to(dev)  # noqa: E731
        self._marks = [("start", self._event())]
        m.refresh()
        if self.post is not None:
            self.post.refresh()
        d = m.d
        # ! map, traffic lights (waymo_motion.py:317-324); TL tokens see detached map features (traffic_light.py:113-115)
        mp = m.map_encoder(g("sc/mp_valid"), g("sc/mp_attr"), g("sc/mp_pose"))
        mp_det = dict(mp, mp_token_feature=mp["mp_token_feature"].detach())
        tl = m.tl_pre_compute(g("sc/tl_valid"), g("sc/tl_attr"), g("sc/tl_pose"), mp_det)
        kv_mp = m.ag_static(mp)
        n_sc, A, n_gt = batch["gt/ag_valid"].shape
        n_tl = batch["sc/tl_valid"].shape[1]
        # ! latent (:326-350): posterior N(mu, exp(log_std)) vs the unit-Gaussian prior; rollout on a reparameterised sample
        out = {}
        eps = g("ag_latent_eps")
        lat_valid = post_valid = g("gt/ag_valid").any(-1)
        if self.post is not None:
            mu, lat_valid = self._posterior(g, mp)
            mu, post_valid = mu.view(n_sc, A, -1), lat_valid
            log_std = self.params["latent_encoder.latent_dist_post.log_std"]
            if bool(batch.get("rollout_prior", False)):
                latent, lat_valid = eps, g("sc/ag_valid").any(-1)                                      # prior sample
            else:
                latent = mu + log_std.exp() * eps
        else:
            mu, latent = None, eps
        # ! destination predictor (:352-359), inputs detached (navigation.py:189-192)
        navi_logits = None
        if self.has_navi:
            navi_logits = m.navi_predictor(g("sc/ag_valid"), g("sc/ag_attr"), g("sc/ag_motion"), g("sc/ag_pose"), mp_det,
                                           g("ref/ag_type"), g("ref/mp_type"))
        self._mark("scene_encoders")
        # ! rollout state (Dynamics.init / TeacherForcing.init, dynamics.py:29-64, teacher_forcing.py:51-92)
        n_mp, n_node = batch["sc/mp_valid"].shape[1:]
        rb = {"sc/ag_valid": batch["gt/ag_valid"], "sc/ag_pose": batch["gt/ag_pose"], "sc/ag_motion": batch["gt/ag_motion"],
              "sc/tl_state": batch["gt/tl_state"], "ag_latent": latent.detach()[:, None],
              "ag_latent_valid": lat_valid, "agent/dest": batch["gt/ag_navi"],
              "ag_navi_valid": batch["gt/ag_valid"].any(-1)}
        rb = {**batch, **rb}
        eng.T = self.T
        shape = (n_sc, A, n_tl, n_gt, n_mp, n_node)
        if eng._shape != shape:
            eng._st, eng._shape = eng._alloc(*shape), shape
        st = eng._st
        eng._load_state(st, rb)
        tf = teacher_forcing_mask(g("gt/ag_valid"), *eng.tf_steps)
        if "tf/forcing_agent" in batch:
            tf = tf | (g("tf/forcing_agent")[:, :, None] & g("gt/ag_valid"))
        st["tf_mask"].copy_(tf)
        self._reset(st)
        pose0, motion0 = st["pose"].clone(), st["motion"].clone()
        navi = m.navi_static(mp_det, st["dest_idx"], 1)                                                 # navigation.py:65-71
        lat_inv = (~lat_valid).reshape(-1).contiguous()
        lat_feat = m.mlp(latent.reshape(n_sc * A, -1), "add_latent.mlp_in", (0, 3, 6), True, mask_post=lat_inv)
        if not T < n_gt:
            raise NotImplementedError("the training rollout needs ground-truth traffic-light states for every step "
                                      "(teacher_forcing.py:65): time_step_end < number of gt steps")
        # ---- pass 1 (no grad): advance the closed loop and record the state every policy step sees. The policy inputs
        # are detached (waymo_motion.py:158-161), so this IS the rollout; nothing of it is kept for the backward.
        W = m.W
        tidx, tmask = self._ring_index(T, W)
        tlb, hist_tl_all = self._tl_batch(st, tl, T, tidx, tmask)
        d_rows = torch.arange(1, T + 1, dtype=torch.int32, device=dev).repeat(n_sc).contiguous()        # step of row (sc, s)
        by_step = lambda t, n: t.view(n_sc, T, n, -1).permute(1, 0, 2, 3).contiguous().view(T, n_sc * n, -1)  # noqa: E731
        S_valid = torch.empty(T, n_sc, A, dtype=torch.uint8, device=dev)
        S_pose, S_motion = torch.empty(T, n_sc, A, 3, device=dev), torch.empty(T, n_sc, A, 3, device=dev)
        N_inv = torch.empty(T, n_sc, A, dtype=torch.bool, device=dev)
        lib, dy, order = L.load(), eng.dyn, ("veh", "ped", "cyc")
        with torch.no_grad():
            tl_feat_ng, logits_ng = m.tl_forward(hist_tl_all, d_rows, tlb)                              # all steps at once
            kv_tl_ng = [by_step(t, n_tl) for t in m.ag_tl_tables(tl_feat_ng)]
            tl_feat_ng, logits_ng = by_step(tl_feat_ng, n_tl), by_step(logits_ng, n_tl)
            for s in range(1, T + 1):                                                                   # :233
                S_valid[s - 1], S_pose[s - 1], S_motion[s - 1] = st["valid"], st["pose"], st["motion"]
                N_inv[s - 1] = st["navi_invalid"]
                x = m.ag_forward(st, mp, kv_mp, tl, tl_feat_ng[s - 1], 1, kv_tl=[t[s - 1] for t in kv_tl_ng])
                act = m.heads_train(x, st["pose"].reshape(-1, 3), navi, st["navi_invalid"].reshape(-1), lat_feat, lat_inv)
                L.check(lib.tb_tl_step_ex(L.ptr(logits_ng[s - 1]), L.ptr(ops._u8(tl["tl_token_invalid"])),
                                          L.ptr(st["gt_tl"]), st["n_gt"], L.ptr(st["d_step"]), st["Bt"], st["n_tl"], W,
                                          self.T, L.ptr(st["hist_tl"]), L.ptr(st["tl_out"]), None, L.stream()),
                        "tb_tl_step_ex")
                L.check(lib.tb_dyn_step_ex(
                    L.ptr(act), L.ptr(st["ag_type"]), ops.host_f3([dy[k]["max_acc"] for k in order]),
                    ops.host_f3([dy[k]["max_yaw_rate"] for k in order]), dy["dt"], L.ptr(st["valid"]),
                    L.ptr(st["disabled"]), L.ptr(ops._u8(st["navi_invalid"])), L.ptr(st["dest_reached"]), L.ptr(st["pose"]),
                    L.ptr(st["motion"]), L.ptr(st["gt_valid"]), L.ptr(st["gt_pose"]), L.ptr(st["gt_motion"]),
                    L.ptr(st["tf_mask"]), st["n_gt"], 1, L.ptr(st["boundary"]), L.ptr(st["dest_idx"]), L.ptr(st["mp_pos"]),
                    L.ptr(st["mp_dirn"]), L.ptr(st["mp_node_invalid"]), L.ptr(st["mp_kind"]), st["n_mp"], st["n_node"],
                    eng.thresh_lane, eng.thresh_edge, eng.cos_rot, L.ptr(st["d_step"]), st["B"], st["A"], W, self.T,
                    L.ptr(st["hist_valid"]), L.ptr(st["hist_pose"]), L.ptr(st["hist_motion"]), L.ptr(st["pred_valid"]),
                    L.ptr(st["pred_pose"]), L.ptr(st["pred_motion"]), None, None, L.stream()), "tb_dyn_step_ex")
                L.check(lib.tb_step_advance(L.ptr(st["d_step"]), L.stream()), "tb_step_advance")
                ops._count(3)
            del tl_feat_ng, logits_ng, kv_tl_ng
        m.drop_caches()  # derived weights cached during the no-grad pass carry no graph
        self._mark("rollout")
        # ---- pass 2 (differentiable): the policy of ALL T steps as one batch of n_sc x T rollout-scenes, row (sc, s).
        # Every step's input is the recorded state, so forward and backward are ~300 large launches in total instead of
        # ~270 small ones per step; K|V tables of the map are indexed with div = T, TL tables are per (scene, step).
        Bq = n_sc * T
        ring = lambda S: S[tidx]  # noqa: E731  [T, W, n_sc, A, ...]: ring slot k of step s holds time tidx[s-1, k]
        hv = (ring(S_valid).permute(2, 0, 3, 1) * tmask[None, :, None, :].to(torch.uint8)).contiguous().view(Bq, A, W)
        hp = ring(S_pose).permute(2, 0, 3, 1, 4).contiguous().view(Bq, A, W, 3)
        hm = ring(S_motion).permute(2, 0, 3, 1, 4).contiguous().view(Bq, A, W, 3)
        over_t = lambda t: t.view(n_sc, 1, A, -1).expand(-1, T, -1, -1).reshape(Bq * A, -1)  # noqa: E731
        stb = dict(B=Bq, A=A, hist_valid=hv, hist_pose=hp, hist_motion=hm, d_step=d_rows,
                   ag_attr=st["ag_attr"].view(n_sc, 1, A, 6).expand(-1, T, -1, -1).contiguous().view(Bq, A, 6))
        tl_feat, logits = m.tl_forward(hist_tl_all, d_rows, tlb)
        x = m.ag_forward(stb, mp, kv_mp, tl, tl_feat, T, kv_tl=m.ag_tl_tables(tl_feat), tl_pose_div=T)
        navi_b = dict(feat=over_t(navi["feat"]), pose=over_t(navi["pose"]).contiguous())
        act = m.heads_train(x, S_pose.permute(1, 0, 2, 3).reshape(-1, 3), navi_b, N_inv.permute(1, 0, 2).reshape(-1),
                            over_t(lat_feat), over_t(lat_inv.view(-1, 1)).reshape(-1))
        acts = act.view(n_sc, T, A, 6).permute(1, 0, 2, 3).reshape(T, n_sc * A, 6)
        logit_l = by_step(logits, n_tl)
        # ! losses (training.py:76-189)
        pred_valid = st["pred_valid"][:, :, :T].contiguous()
        rec = dict(B=n_sc, A=A, ag_type=st["ag_type"], pred_valid=pred_valid, pose0=pose0, motion0=motion0,
                   gt_valid=st["gt_valid"], gt_pose=st["gt_pose"], gt_motion=st["gt_motion"], tf_mask=st["tf_mask"],
                   n_gt=n_gt, sc_div=1)
        t0 = tc["step_training_start"]
        il, _ = AG.il_loss(acts, rec, dy, (tc["w_pos"], tc["w_rot"], tc["w_spd"]), t0)
        loss = torch.zeros((), device=dev)
        if float(il.detach()[1]) > 0:                                                                            # :170-181
            out["diffbar_reward"] = -tc["w_diffbar_reward"] * il[0] / il[1]
            loss = loss - out["diffbar_reward"]
        nll = AG.tl_nll(logit_l, tl["tl_token_invalid"].reshape(-1), st["gt_tl"], n_gt)
        if float(nll.detach()[1]) > 0:                                                                           # :186-188
            out["tl_state_loss"] = tc["w_tl_state"] * nll[0] / nll[1]
            loss = loss + out["tl_state_loss"]
        loss_any = pred_valid[:, :, t0:].bool().any(-1)                                                 # loss_valid.any(-1)
        if mu is not None:                                                                              # :108-121, loss.py:40-76
            kl_valid = post_valid & loss_any                                                            # kl_for_unseen_agent
            kl = (0.5 * ((2 * log_std).exp() + mu * mu - 1.0) - log_std).sum(-1)                        # KL(post || N(0, I))
            free = tc["kl_free_nats"]
            err = torch.clamp(kl.detach(), min=free) + tc["kl_balance_scale"] * torch.clamp(kl, min=free)
            if int(kl_valid.sum()) > 0:
                out["vae_kl"] = tc["w_vae_kl"] * err.masked_fill(~kl_valid, 0.0).sum() / kl_valid.sum()
                loss = loss + out["vae_kl"]
        if navi_logits is not None:                                                                     # :146-153
            navi_valid = g("sc/ag_valid").any(-1) & loss_any
            n_mp_ = navi_logits.shape[-1]
            nn = AG.softmax_nll(navi_logits.reshape(-1, n_mp_), g("gt/ag_navi").reshape(-1), navi_valid.reshape(-1))
            if float(nn.detach()[1]) > 0:
                out["navi_loss"] = tc["w_navi"] * nn[0] / nn[1]
                loss = loss + out["navi_loss"]
        out["loss"] = loss
        self._mark("policy_forward_losses")
        if backward:
            loss.backward()
            self._mark("backward")
            if self.bucket is not None:
                self.bucket.all_reduce()
                self._mark("grad_all_reduce")
        out["pred_pose"], out["pred_valid"] = st["pred_pose"][:, :, :T], pred_valid.bool()
        if perm is not None:  # back to the caller's agent order (dropped slots: zeros / False)
            sc_i = torch.arange(n_sc, device=dev)[:, None]
            for k in ("pred_pose", "pred_valid"):
                full = out[k].new_zeros((n_sc, A_full) + tuple(out[k].shape[2:]))
                full[sc_i, perm.to(dev)] = out[k]
                out[k] = full
        return out

    def _event(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def _mark(self, name: str) -> None:
        self._marks.append((name, self._event()))

    def timings_ms(self) -> Dict[str, float]:
        """Device time of the phases of the last step() (CUDA events on the current stream; synchronises)."""
        torch.cuda.synchronize()
        return {n: self._marks[i][1].elapsed_time(e) for i, (n, e) in enumerate(self._marks[1:])}

    def _ring_index(self, T: int, W: int):
        """History ring of policy step s (1-based): slot k holds time t = s-1 - ((s-1-k) mod W) when t >= 0
        (traffic_bots.py:123-143 keeps the last W states; the kernels read slot = time % W). Returns the time index
        [T, W] (clamped to 0) and its validity mask."""
        s1 = torch.arange(T, device=self.dev)[:, None]              # s - 1
        k = torch.arange(W, device=self.dev)[None, :]
        t = s1 - torch.remainder(s1 - k, W)
        return t.clamp(min=0), t >= 0

    def _tl_batch(self, st: dict, tl: dict, T: int, tidx: Tensor, tmask: Tensor):
        """Static TL dict + TL history rings for the batch of n_sc x T rollout-scenes (row (sc, s)). With ground-truth
        light states for every step (always in training, teacher_forcing.py:65,159-160) the TL history of every step
        is known before the rollout and the TL tokens never depend on the agents."""
        m = self.model
        W, d = m.W, m.d
        n_sc, n_tl = st["Bt"], st["n_tl"]
        hist = st["gt_tl"][:, :, tidx]                                                                  # [n_sc,n_tl,T,W,5]
        hist = (hist.permute(0, 2, 1, 3, 4) * tmask[None, :, None, :, None].to(torch.uint8)).contiguous()
        rep = lambda t: t.repeat_interleave(T, 0).contiguous()  # noqa: E731
        knn = lambda k: dict(idx=rep(k["idx"]), inv=rep(k["inv"]), rel=rep(k["rel"]))  # noqa: E731
        c0 = knn(tl["cross"][0])
        tlb = dict(n_sc=n_sc * T, n_tl=n_tl, tl_token_invalid=rep(tl["tl_token_invalid"]),
                   tl_attr_rows=tl["tl_token_attr"].view(n_sc, 1, n_tl, 1, d).expand(-1, T, -1, W, -1).reshape(-1, d),
                   knn_self=knn(tl["knn_self"]),
                   cross=[dict(c0, kv0=c["kv0"], T0=c["T0"], div0=T, K0=c["K0"]) for c in tl["cross"]])
        return tlb, hist.view(n_sc * T, n_tl, W, 5)

    _AGENT_KEYS = ("gt/ag_valid", "gt/ag_pose", "gt/ag_motion", "gt/ag_navi", "sc/ag_valid", "sc/ag_pose", "sc/ag_motion",
                   "sc/ag_attr", "ref/ag_type", "ref/ag_size", "ref/ag_role", "tf/forcing_agent", "ag_latent_eps",
                   "agent/dest", "ag_navi_valid", "ag_latent_valid")

    def _compact(self, batch: Dict[str, Tensor]):
        """Agent compaction as in RolloutEngine._compact: slots without a valid ground-truth step never become valid and
        contribute neither loss nor gradient; they are dropped for the whole step (valid-first reorder per scene, the
        largest valid count of the batch rounded up to 4). Returns (batch to use, perm [n_sc, A_eff] or None, A)."""
        gt_valid = batch["gt/ag_valid"]
        n_sc, A, _ = gt_valid.shape
        if not self.compact_agents:
            return batch, None, A
        ever = gt_valid.any(-1)
        a_eff = max(int(ever.sum(1).max()), self.sz["k_ag2ag"] + 1)
        a_eff = min(A, (a_eff + 3) // 4 * 4)
        if a_eff >= A:
            return batch, None, A
        perm = torch.sort((~ever).to(torch.uint8), dim=1, stable=True)[1][:, :a_eff]
        out = dict(batch)
        for k in self._AGENT_KEYS:
            if k in batch and torch.is_tensor(batch[k]) and batch[k].dim() >= 2 and batch[k].shape[:2] == (n_sc, A):
                t = batch[k]
                idx = perm.to(t.device).view((n_sc, a_eff) + (1,) * (t.dim() - 2)).expand((n_sc, a_eff) + tuple(t.shape[2:]))
                out[k] = torch.gather(t, 1, idx)
        return out, perm, A

    def _reset(self, st: dict) -> None:
        """time 0 of the rollout (waymo_motion.py:219-227): RolloutEngine._reset without the TL prologue of the
        software-pipelined inference loop."""
        for k in ("disabled", "dest_reached", "hist_valid", "hist_pose", "hist_motion", "hist_tl", "pred_valid",
                  "pred_pose", "pred_motion", "tl_out"):
            st[k].zero_()
        st["valid"].copy_(st["gt_valid"][:, :, 0])
        st["pose"].copy_(st["gt_pose"][:, :, 0])
        st["motion"].copy_(st["gt_motion"][:, :, 0])
        st["navi_invalid"].copy_(~st["init_navi_valid"])
        st["hist_valid"][:, :, 0] = st["valid"]
        st["hist_pose"][:, :, 0] = st["pose"]
        st["hist_motion"][:, :, 0] = st["motion"]
        st["hist_tl"][:, :, 0] = st["gt_tl"][:, :, 0]
        st["knn_state"].fill_(float("inf"))
        st["knn_state_tl"].fill_(float("inf"))
        st["d_step"].fill_(1)
