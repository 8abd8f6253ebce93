"""Rollout-level and dynamics-level drop-ins: the reference's own call surface for

  utils.dynamics.MultiPathPP / Dynamics                     src/utils/dynamics.py:11-274
  utils.teacher_forcing.TeacherForcing (configuration)      src/utils/teacher_forcing.py:8-82
  utils.traffic_rule_checker.TrafficRuleChecker (inputs)    src/utils/traffic_rule_checker.py:10-84
  utils.buffer.RolloutBuffer                                src/utils/buffer.py:7-146
  WaymoMotion.rollout / joint_future_pred                   src/pl_modules/waymo_motion.py:206-311, 439-524

A maintainer swaps `self.rollout(...)` of the Lightning module for `WaymoMotionRollout.rollout(...)` (same arguments,
same `RolloutBuffer` fields back); the loop itself runs inside `RolloutEngine` (CUDA graph, device-resident state)
instead of 90 x Python. `Dynamics` is the stand-alone class for callers that keep the reference's Python loop around the
`TrafficBots` drop-in (traffic_bots.py): its agent update is the `tb_dyn_update` kernel, the overrides are mask plumbing.

Only the inference configuration of the reference is implemented (teacher forcing = warm start + spawn, deterministic
actions, navi_mode "dest", no player policy, `pred_navi_after_reached: False`, sim_agent.yaml:23-25,262-264);
anything else raises NotImplementedError — there is no PyTorch fallback of the loop.
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor
from torch.distributions import Categorical, Independent

from . import config as C
from . import lib as L
from . import ops
from .engine import RolloutEngine, teacher_forcing_mask


# ---------------------------------------------------------------------------------------------------- dynamics
class MultiPathPP:
    """utils/dynamics.py:226-274: unicycle model with bounded acceleration / yaw rate."""

    def __init__(self, dt: float, max_acc: float = 4, max_yaw_rate: float = 1) -> None:
        self.dt, self._max_acc, self._max_yaw_rate = dt, max_acc, max_yaw_rate

    def _one(self, action_unbounded, pose, motion):
        n = action_unbounded.shape[0] * action_unbounded.shape[1]
        dev = action_unbounded.device
        ones = torch.ones(n, dtype=torch.uint8, device=dev)
        ty = torch.zeros(n, 3, dtype=torch.uint8, device=dev)
        ty[:, 0] = 1
        out_p, out_m, out_a = torch.empty_like(pose), torch.empty_like(motion), torch.empty_like(action_unbounded)
        L.check(L.load().tb_dyn_update(L.ptr(action_unbounded), L.ptr(ty), L.ptr(ones), None, None,
                                       ops.host_f3([self._max_acc] * 3), ops.host_f3([self._max_yaw_rate] * 3), self.dt, n,
                                       L.ptr(pose), L.ptr(motion), L.ptr(out_p), L.ptr(out_m), L.ptr(out_a), L.stream()),
                "tb_dyn_update")
        ops._count()
        return out_p, out_m, out_a

    @torch.no_grad()
    def process_action(self, action: Tensor) -> Tensor:
        """[n_sc, n_ag, 2] unbounded -> (acc m/s^2, yaw rate rad/s), :237-246."""
        a = action.float().contiguous()
        z = torch.zeros(*a.shape[:-1], 3, device=a.device)
        return self._one(a, z, z)[2]

    @torch.no_grad()
    def update(self, pose: Tensor, motion: Tensor, action: Tensor) -> Tuple[Tensor, Tensor]:
        """:248-274 with a PHYSICAL action: atanh is not needed — the kernel is fed through the player-override slot."""
        p, m, a = pose.float().contiguous(), motion.float().contiguous(), action.float().contiguous()
        n = p.shape[0] * p.shape[1]
        ones = torch.ones(n, dtype=torch.uint8, device=p.device)
        ty = torch.zeros(n, 3, dtype=torch.uint8, device=p.device)
        ty[:, 0] = 1
        out_p, out_m, out_a = torch.empty_like(p), torch.empty_like(m), torch.empty_like(a)
        L.check(L.load().tb_dyn_update(L.ptr(a), L.ptr(ty), L.ptr(ones), L.ptr(ones), L.ptr(a),
                                       ops.host_f3([self._max_acc] * 3), ops.host_f3([self._max_yaw_rate] * 3), self.dt, n,
                                       L.ptr(p), L.ptr(m), L.ptr(out_p), L.ptr(out_m), L.ptr(out_a), L.stream()),
                "tb_dyn_update")
        ops._count()
        return out_p, out_m


def _instantiate(cfg, dt: float) -> MultiPathPP:
    """hydra.utils.instantiate(cfg, dt=dt) for the one dynamics class the rollout configuration uses."""
    if isinstance(cfg, MultiPathPP):
        return cfg
    cfg = dict(cfg)
    target = cfg.pop("_target_", "utils.dynamics.MultiPathPP")
    if not str(target).endswith("MultiPathPP"):
        raise NotImplementedError(f"dynamics {target!r}: only MultiPathPP (sim_agent.yaml:156-167) is implemented")
    return MultiPathPP(dt=dt, **cfg)


class Dynamics:
    """utils/dynamics.py:11-222 — same constructor, attributes (`ag_valid`, `ag_pose`, `ag_motion`, `tl_state`,
    `ag_navi`, `ag_navi_valid`, `ag_disabled`, `mask_navi_reached`, `ag_navi_updated`, ...) and methods."""

    def __init__(self, veh, ped, cyc, navi_mode: str, use_veh_dynamics_for_all: bool = False) -> None:
        self.dt = 0.1
        self.action_dim = 2
        self.navi_mode = navi_mode
        self.use_veh_dynamics_for_all = use_veh_dynamics_for_all
        if use_veh_dynamics_for_all:
            self.ag_dynamics = _instantiate(veh, self.dt)
        else:
            self.ag_dynamics = (_instantiate(veh, self.dt), _instantiate(ped, self.dt), _instantiate(cyc, self.dt))

    @classmethod
    def default(cls, navi_mode: str = "dest") -> "Dynamics":
        """The dynamics of configs/model/sim_agent.yaml:156-167."""
        d = C.DYNAMICS_CFG
        return cls(veh=d["veh"], ped=d["ped"], cyc=d["cyc"], navi_mode=navi_mode)

    def limits(self) -> Tuple[List[float], List[float]]:
        dyn = (self.ag_dynamics,) * 3 if self.use_veh_dynamics_for_all else self.ag_dynamics
        return [d._max_acc for d in dyn], [d._max_yaw_rate for d in dyn]

    def init(self, tl_state: Tensor, gt_valid: Tensor, gt_pose: Tensor, gt_motion: Tensor, ag_type: Tensor,
             ag_attr: Tensor, ag_latent: Optional[Tensor], ag_latent_valid: Optional[Tensor], ag_navi: Optional[Tensor],
             ag_navi_valid: Tensor, **kwargs) -> None:
        """:29-64."""
        self.ag_type, self.ag_attr = ag_type, ag_attr
        self.ag_latent, self.ag_latent_valid = ag_latent, ag_latent_valid
        self.ag_valid = gt_valid[:, :, 0]
        self.ag_disabled = torch.zeros_like(self.ag_valid)
        self.ag_pose = gt_pose[:, :, 0]
        self.ag_motion = gt_motion[:, :, 0]
        self.tl_state = tl_state[:, :, 0]
        self.ag_navi = ag_navi
        self.ag_navi_valid = ag_navi_valid
        self.mask_navi_reached = torch.zeros_like(self.ag_navi_valid)
        self.ag_navi_updated = True

    @torch.no_grad()
    def update_ag(self, action_dist: Independent, deterministic: bool = True,
                  player_override: Optional[Dict[str, Tensor]] = None) -> Tuple[Tensor, Tensor]:
        """:66-120. Returns (action [n_sc,n_ag,2] physical, action_log_prob [n_sc,n_ag])."""
        if torch.is_grad_enabled() and action_dist.mean.requires_grad:
            raise NotImplementedError("the B200 Dynamics drop-in is the inference path (no gradient through update_ag)")
        ag_invalid = ~self.ag_valid
        act_u = (action_dist.mean if deterministic else action_dist.rsample()).float().contiguous()
        log_prob = action_dist.log_prob(act_u).masked_fill(ag_invalid, 0)
        B, A, _ = act_u.shape
        max_acc, max_yaw = self.limits()
        ty = (self.ag_type if not self.use_veh_dynamics_for_all
              else torch.ones_like(self.ag_type) & torch.tensor([True, False, False], device=act_u.device))
        pose, motion = self.ag_pose.float().contiguous(), self.ag_motion.float().contiguous()
        out_p, out_m, out_a = torch.empty_like(pose), torch.empty_like(motion), torch.empty_like(act_u)
        pv = pa = None
        if player_override is not None:
            pv = player_override["valid"].to(torch.uint8).contiguous()
            pa = player_override["action"].float().contiguous()
        ty8, valid8 = ty.to(torch.uint8).contiguous(), self.ag_valid.to(torch.uint8).contiguous()  # keep alive
        L.check(L.load().tb_dyn_update(L.ptr(act_u), L.ptr(ty8), L.ptr(valid8), L.ptr(pv), L.ptr(pa),
                                       ops.host_f3(max_acc), ops.host_f3(max_yaw), self.dt, B * A, L.ptr(pose), L.ptr(motion),
                                       L.ptr(out_p), L.ptr(out_m), L.ptr(out_a), L.stream()), "tb_dyn_update")
        ops._count()
        self.ag_pose, self.ag_motion = out_p, out_m
        return out_a, log_prob

    def override_ag(self, ag_override: Dict[str, Tensor]) -> None:
        """:122-141 (teacher forcing / spawn)."""
        valid = ag_override["valid"] & (~self.ag_disabled)
        self.ag_valid = self.ag_valid | valid
        self.ag_pose = torch.where(valid.unsqueeze(-1), ag_override["pose"], self.ag_pose)
        self.ag_motion = torch.where(valid.unsqueeze(-1), ag_override["motion"], self.ag_motion)

    @torch.no_grad()
    def override_tl(self, tl_state_dist: Categorical, tl_override: Dict[str, Tensor]) -> None:
        """:143-163."""
        new = torch.nn.functional.one_hot(tl_state_dist.probs.argmax(-1), self.tl_state.shape[-1]).to(self.tl_state.dtype)
        self.tl_state = torch.where(tl_override["valid"].unsqueeze(-1), tl_override["state"], new)

    @torch.no_grad()
    def disable_ag(self, traffic_rule_violation: Dict[str, Tensor], gt_valid: Optional[Tensor] = None) -> None:
        """:165-181."""
        mask_disable = traffic_rule_violation["outside_map_this_step"]
        if gt_valid is not None:
            mask_disable = mask_disable & (~gt_valid)
        self.ag_disabled = self.ag_disabled | mask_disable
        self.ag_valid = self.ag_valid & (~mask_disable)

    @torch.no_grad()
    def disable_navi(self, traffic_rule_violation: Dict[str, Tensor]) -> None:
        """:183-204."""
        if self.navi_mode in ("dest", "goal"):
            self.mask_navi_reached = traffic_rule_violation[f"{self.navi_mode}_reached_this_step"]
            self.ag_navi_valid = self.ag_navi_valid & (~self.mask_navi_reached)

    @torch.no_grad()
    def override_navi(self, navi: Tensor) -> None:
        """:206-222."""
        valid = self.mask_navi_reached
        if self.navi_mode in ("cmd", "goal"):
            valid = valid.unsqueeze(-1)
        self.ag_navi = torch.where(valid, navi, self.ag_navi)
        self.ag_navi_valid = self.ag_navi_valid | self.mask_navi_reached
        self.ag_navi_updated = True


# ---------------------------------------------------------------------------------------------------- loop inputs
class TeacherForcing:
    """utils/teacher_forcing.py:8-49 — the configuration object `rollout` receives. The engine evaluates the
    inference-time schedule (spawn up to `step_spawn_agent`, warm start up to `step_warm_start`, ground-truth traffic
    lights while available, :51-82,126-160) on the device; the training-time schedules are not implemented."""

    def __init__(self, step_spawn_agent: int = 10, step_warm_start: int = 10, step_horizon: int = 0,
                 step_horizon_decrease_per_epoch: int = 0, prob_forcing_agent: float = 0,
                 prob_forcing_agent_decrease_per_epoch: float = 0, prob_scheduled_sampling: float = 0,
                 prob_scheduled_sampling_decrease_per_epoch: float = 0, gt_sdc: bool = False, threshold_xy: float = -1.0,
                 threshold_yaw: float = -1.0, threshold_spd: float = -1.0) -> None:
        self.step_spawn_agent, self.step_warm_start = step_spawn_agent, step_warm_start
        self.step_horizon, self.step_horizon_decrease_per_epoch = step_horizon, step_horizon_decrease_per_epoch
        self.prob_forcing_agent = prob_forcing_agent
        self.prob_forcing_agent_decrease_per_epoch = prob_forcing_agent_decrease_per_epoch
        self.prob_scheduled_sampling = prob_scheduled_sampling
        self.prob_scheduled_sampling_decrease_per_epoch = prob_scheduled_sampling_decrease_per_epoch
        self.gt_sdc = gt_sdc
        self.threshold_xy, self.threshold_yaw, self.threshold_spd = threshold_xy, threshold_yaw, threshold_spd


def _check_teacher_forcing(tf) -> Tuple[int, int]:
    """Works on this module's TeacherForcing and on the reference's own object (same attribute names)."""
    off = (tf.step_horizon <= 0 and tf.prob_forcing_agent <= 0 and tf.prob_scheduled_sampling <= 0 and not tf.gt_sdc
           and tf.threshold_xy < 0 and tf.threshold_yaw < 0 and tf.threshold_spd < 0)
    if not off:
        raise NotImplementedError("only the inference teacher forcing (step_spawn_agent / step_warm_start; "
                                  "sim_agent.yaml:262-264) runs inside the rollout engine")
    return int(tf.step_spawn_agent), int(tf.step_warm_start)


class TrafficRuleChecker:
    """utils/traffic_rule_checker.py:10-84 — constructor-compatible holder of the checker's inputs. Its `check` is not
    a Python method here: outside-map / destination-reached feed back inside `tb_dyn_step`, the five logging checks
    are `tb_rule_check`, both launched by the engine every step."""

    def __init__(self, mp_boundary: Tensor, mp_valid: Tensor, mp_type: Tensor, mp_pos: Tensor, mp_dir: Tensor,
                 ag_type: Tensor, ag_size: Tensor, ag_goal: Optional[Tensor], ag_dest: Optional[Tensor], tl_valid: Tensor,
                 tl_pose: Tensor, disable_check: bool, collision_size_scale: float = 1.1) -> None:
        if ag_goal is not None or ag_dest is None:
            raise NotImplementedError('navi_mode "dest" only (sim_agent.yaml:5): pass ag_dest, not ag_goal')
        if collision_size_scale != 1.1:
            raise NotImplementedError("collision_size_scale is fixed at the reference default 1.1")
        self.mp_boundary, self.mp_valid, self.mp_type = mp_boundary, mp_valid, mp_type
        self.mp_pos, self.mp_dir = mp_pos[..., :2], mp_dir[..., :2]
        self.ag_type, self.ag_size_raw, self.ag_dest = ag_type, ag_size, ag_dest
        self.ag_size = ag_size[..., :2] * collision_size_scale
        self.tl_valid, self.tl_pose = tl_valid, tl_pose
        self.disable_check = disable_check


# ---------------------------------------------------------------------------------------------------- buffer
class RolloutBuffer:
    """utils/buffer.py:7-146 — same attributes after `finish()`; filled from the engine's device buffers in one go
    instead of 90 `add` calls (so `add` / `finish` are not part of this class)."""

    def __init__(self, step_end: int, step_current: int) -> None:
        self.step_start = 1
        self.step_end = step_end
        self.step_future_start = step_current
        self.pred_valid = self.pred_pose = self.pred_motion = self.action_log_prob = None
        self.navi_log_prob: List[Tensor] = []
        self.navi_log_prob_valid: List[Tensor] = []
        self.tl_state_nll = self.tl_state_nll_invalid = None
        self.diffbar_reward: Dict[str, Tensor] = {}  # training-only (rewards.py); empty at inference
        self.mask_teacher_forcing = None
        self.violation: Dict[str, Tensor] = {}
        self.vis_dict: Dict[str, Tensor] = {}
        self.log_prob = None

    def add_navi_log_prob(self, ag_navi_log_prob: Tensor, mask_navi_reached: Tensor) -> None:
        self.navi_log_prob.append(ag_navi_log_prob)
        self.navi_log_prob_valid.append(mask_navi_reached)

    def _finish(self) -> None:
        self.navi_log_prob = torch.stack(self.navi_log_prob, dim=2)
        self.navi_log_prob_valid = torch.stack(self.navi_log_prob_valid, dim=2)

    def compute_log_prob(self, latent_log_prob: Optional[Tensor]) -> None:
        """:103-110."""
        self.log_prob = (self.navi_log_prob * self.navi_log_prob_valid).sum(-1)
        self.log_prob = self.log_prob / self.navi_log_prob_valid.sum(-1)
        self.log_prob = self.log_prob.masked_fill(~self.navi_log_prob_valid.any(-1), 0)
        if latent_log_prob is not None:
            self.log_prob = self.log_prob + latent_log_prob.view(self.log_prob.shape)

    def flatten_joint_future(self, n_joint_future: int) -> None:
        """:112-146."""
        B, n_ag, n_step = self.pred_valid.shape
        n_sc = B // n_joint_future
        v = lambda t: t.view(n_sc, n_joint_future, *t.shape[1:])  # noqa: E731
        self.pred_valid, self.pred_pose, self.pred_motion = v(self.pred_valid), v(self.pred_pose), v(self.pred_motion)
        self.navi_log_prob, self.navi_log_prob_valid = v(self.navi_log_prob), v(self.navi_log_prob_valid)
        self.tl_state_nll, self.tl_state_nll_invalid = v(self.tl_state_nll), v(self.tl_state_nll_invalid)
        self.violation = {k: v(t) for k, t in self.violation.items()}
        self.diffbar_reward = {k: v(t) for k, t in self.diffbar_reward.items()}
        self.action_log_prob = v(self.action_log_prob)
        self.vis_dict = {k: v(t) for k, t in self.vis_dict.items()}
        self.mask_teacher_forcing = v(self.mask_teacher_forcing)


# ---------------------------------------------------------------------------------------------------- rollout
_VIO = ("outside_map", "collided", "collided_wosac", "run_road_edge", "run_red_light", "passive", "goal_reached",
        "dest_reached")


class WaymoMotionRollout:
    """`WaymoMotion.rollout` / `joint_future_pred` (waymo_motion.py:206-311, 439-524) on the rollout engine.

    model: the `TrafficBots` drop-in (traffic_bots.py; a reference checkpoint loads into it unchanged) — its
      `mp_encoder` / `tl_encoder.pre_compute` produce the token dicts `rollout` consumes.
    dynamics: a `Dynamics` (this module or the reference's: only `ag_dynamics[i]._max_acc/_max_yaw_rate`, `dt`,
      `navi_mode` are read)."""

    def __init__(self, model, dynamics=None, time_step_current: int = 10, time_step_end: int = 90, n_joint_future: int = 1,
                 current_epoch: int = 0) -> None:
        self.model = model
        self.dynamics = dynamics if dynamics is not None else Dynamics.default()
        if self.dynamics.navi_mode != "dest":
            raise NotImplementedError('navi_mode "dest" only')
        self.time_step_current, self.time_step_end = time_step_current, time_step_end
        self.n_joint_future, self.current_epoch = n_joint_future, current_epoch
        self.training = False
        self._engines: Dict[tuple, RolloutEngine] = {}
        self._train_step = None

    def _engine(self, R: int, step_end: int, rule_checks: bool) -> RolloutEngine:
        hp = self.model._runner()  # (re)builds the fused weights when a parameter changed
        key = (R, step_end, rule_checks, id(hp))
        if key not in self._engines:
            acc, yaw = self.dynamics.limits() if hasattr(self.dynamics, "limits") else (
                [d._max_acc for d in self.dynamics.ag_dynamics], [d._max_yaw_rate for d in self.dynamics.ag_dynamics])
            dyn = dict(veh=dict(max_acc=acc[0], max_yaw_rate=yaw[0]), ped=dict(max_acc=acc[1], max_yaw_rate=yaw[1]),
                       cyc=dict(max_acc=acc[2], max_yaw_rate=yaw[2]), dt=self.dynamics.dt)
            self._engines = {k: e for k, e in self._engines.items() if k[3] == id(hp)}  # drop engines of old weights
            self._engines[key] = RolloutEngine(hp.P, self.model.cfg, hp.dev, precision=hp.precision, n_rollout=R,
                                               step_end=step_end, rule_checks=rule_checks, record_feedback=True,
                                               dynamics_cfg=dyn)
        return self._engines[key]

    def training_step(self, batch: Dict[str, Tensor], batch_idx: int = 0, n_steps: Optional[int] = None,
                      train_cfg: Optional[dict] = None) -> Dict[str, Tensor]:
        """Body of `WaymoMotion.training_step` (waymo_motion.py:313-385) on the CUDA training path (training.TrainStep;
        the model must have been built with `training_modules=True`). The module's own parameters are the leaves: after
        the call their `.grad` holds the gradients (a torch optimiser over `model.parameters()` steps on them; call
        `model.zero_grad()` between steps). `batch`: the pre-processed dict plus the step's random draws
        ("tf/forcing_agent", "ag_latent_eps", optional "rollout_prior"). Returns the loss terms (training.py:162-189)."""
        from .training import TrainStep
        if self._train_step is None:
            leaves = {k: p for k, p in self.model.named_parameters()}
            acc, yaw = self.dynamics.limits() if hasattr(self.dynamics, "limits") else (
                [d._max_acc for d in self.dynamics.ag_dynamics], [d._max_yaw_rate for d in self.dynamics.ag_dynamics])
            dyn = dict(veh=dict(max_acc=acc[0], max_yaw_rate=yaw[0]), ped=dict(max_acc=acc[1], max_yaw_rate=yaw[1]),
                       cyc=dict(max_acc=acc[2], max_yaw_rate=yaw[2]), dt=self.dynamics.dt)
            dev = next(iter(leaves.values())).device
            self._train_step = TrainStep(leaves, self.model.cfg, dev, precision=min(self.model.precision, 1),
                                         train_cfg=dict(train_cfg or {}, time_step_end=self.time_step_end),
                                         dynamics_cfg=dyn, share_leaves=True)
        return self._train_step.step(batch, n_steps=n_steps)

    @torch.no_grad()
    def rollout(self, ag_tokens: Dict[str, Tensor], mp_tokens: Dict[str, Tensor], tl_tokens: Dict[str, Tensor],
                tl_state_gt: Tensor, teacher_forcing, rule_checker, step_end: int, deterministic_action: bool,
                player_policy=None, n_joint_future: Optional[int] = None) -> RolloutBuffer:
        """Same arguments as the reference. Every tensor arrives repeated `n_joint_future` times along dim 0
        (waymo_motion.py:458-462); the scene-level ones (map / traffic-light tokens, ground truth, map tables of the
        rule checker) are de-duplicated by stride so that the 32 rollouts of a scene share them on the device."""
        if not deterministic_action or player_policy is not None or self.training:
            raise NotImplementedError("inference rollout: deterministic_action=True, no player policy, eval mode")
        R = self.n_joint_future if n_joint_future is None else n_joint_future
        B, A, n_gt = ag_tokens["gt_valid"].shape
        if B % R:
            raise ValueError(f"batch of {B} rollout-scenes is not a multiple of n_joint_future={R}")
        n_sc = B // R
        step_spawn, step_warm = _check_teacher_forcing(teacher_forcing)
        sc = lambda t: t[::R]  # noqa: E731  scene-level view of a repeat_interleave'd tensor
        mp_type = rule_checker.mp_type
        batch = {
            "sc/ag_valid": sc(ag_tokens["gt_valid"]), "sc/ag_pose": sc(ag_tokens["gt_pose"]),
            "sc/ag_motion": sc(ag_tokens["gt_motion"]), "sc/ag_attr": sc(ag_tokens["ag_attr"]),
            "ref/ag_type": sc(ag_tokens["ag_type"]), "ref/ag_size": sc(ag_tokens["ag_size"]),
            "sc/tl_state": sc(tl_state_gt), "sc/tl_valid": sc(~tl_tokens["tl_token_invalid"]),
            "sc/mp_valid": sc(rule_checker.mp_valid),  # shape carrier only: the map tokens are given
            "map/boundary": sc(rule_checker.mp_boundary), "map/valid": sc(rule_checker.mp_valid), "map/type": sc(mp_type),
            "map/pos": sc(rule_checker.mp_pos), "map/dir": sc(rule_checker.mp_dir),
            "ag_latent": ag_tokens["ag_latent"].view(n_sc, R, A, -1), "ag_latent_valid": ag_tokens["ag_latent_valid"],
            "agent/dest": ag_tokens["ag_navi"].view(n_sc, R, A), "ag_navi_valid": ag_tokens["ag_navi_valid"],
        }
        eng = self._engine(R, step_end, not rule_checker.disable_check)
        eng.tf_steps = (step_spawn, step_warm)
        static = eng.static_from_tokens({k: sc(v) for k, v in mp_tokens.items()},
                                        {k: sc(v) for k, v in tl_tokens.items() if v is not None})
        eng.prepare(batch, static=static)
        res = eng.run(step_end)
        if eng.model.kv_half:
            eng.check_fp16_range()
        return self._buffer(eng, res, ag_tokens, tl_tokens, step_end, n_gt, R, step_spawn, step_warm)

    def _buffer(self, eng, res, ag_tokens, tl_tokens, step_end, n_gt, R, step_spawn, step_warm) -> RolloutBuffer:
        buf = RolloutBuffer(step_end, self.time_step_current)
        buf.add_navi_log_prob(ag_tokens["ag_navi_log_prob"], ag_tokens["ag_navi_valid"])
        buf._finish()
        buf.pred_valid, buf.pred_pose, buf.pred_motion = res["pred_valid"], res["pred_pose"], res["pred_motion"]
        B, A, T = buf.pred_valid.shape
        dev = buf.pred_valid.device
        # action_log_prob: the Normal's log-density at its own mean (dynamics.py:90), 0 for invalid agents
        log_std = torch.stack([eng.model.P[f"action_head.log_std.{t}"] for t in range(3)], 0)        # action_head.py:48-50
        lp_type = -(log_std.sum(-1)) - math.log(2 * math.pi)                                          # [3]
        lp = (ag_tokens["ag_type"].float() @ lp_type)[:, :, None].expand(-1, -1, T)
        buf.action_log_prob = lp.masked_fill(~buf.pred_valid, 0.0)
        # violations: *_this_step from the kernels, cumulative flags by a running OR over the steps (:361-451)
        zeros = torch.zeros(B, A, T, dtype=torch.bool, device=dev)
        for k in _VIO:
            this = res.get(k, zeros)
            buf.violation[k] = torch.cummax(this.to(torch.uint8), dim=2)[0].bool()
            buf.violation[f"{k}_this_step"] = this
        # teacher forcing mask of every step (ag_override["valid"], teacher_forcing.py:126-147)
        tf = teacher_forcing_mask(ag_tokens["gt_valid"], step_spawn, step_warm)
        mtf = torch.zeros(B, A, T, dtype=torch.bool, device=dev)
        n = min(T, n_gt - 1)
        mtf[:, :, :n] = tf[:, :, 1:n + 1]
        buf.mask_teacher_forcing = mtf
        st = eng._st
        buf.tl_state_nll = st["tl_nll"].repeat_interleave(R, 0)
        inv = tl_tokens["tl_token_invalid"][:, :, None].expand(-1, -1, T).clone()
        inv[:, :, max(n_gt - 1, 0):] = True                                                           # waymo_motion.py:270-272
        buf.tl_state_nll_invalid = inv
        buf.vis_dict = {"tl_state": res["tl_state"]}
        return buf

    @torch.no_grad()
    def joint_future_pred(self, batch: Dict[str, Tensor], mp_tokens: Dict[str, Tensor], tl_tokens: Dict[str, Tensor],
                          ag_latent: Tensor, ag_latent_valid: Tensor, ag_navi: Tensor, ag_navi_valid: Tensor,
                          teacher_forcing, n_joint_future: int, ag_navi_log_prob: Optional[Tensor] = None,
                          ag_latent_log_prob: Optional[Tensor] = None) -> RolloutBuffer:
        """waymo_motion.py:439-524 with the latent / destination SAMPLES passed in (the reference samples them from
        `MyDist` objects right here, :467-495; `RolloutEngine.predict_destinations` is the CUDA destination sampler):
        ag_latent [n_sc, n_joint_future, n_ag, latent_dim], ag_navi [n_sc, n_joint_future, n_ag] or [n_sc, n_ag]."""
        R = n_joint_future
        rep = lambda t: t.repeat_interleave(R, 0)  # noqa: E731
        n_sc, A = batch["sc/ag_valid"].shape[:2]
        ag_tokens = {"ag_type": rep(batch["ref/ag_type"]), "ag_size": rep(batch["ref/ag_size"]),
                     "ag_attr": rep(batch["sc/ag_attr"]), "gt_valid": rep(batch["sc/ag_valid"]),
                     "gt_pose": rep(batch["sc/ag_pose"]), "gt_motion": rep(batch["sc/ag_motion"])}
        ag_tokens["ag_latent"] = ag_latent[:, :R].reshape(n_sc * R, A, -1)
        ag_tokens["ag_latent_valid"] = rep(ag_latent_valid)
        if ag_navi.dim() == 2:
            ag_navi = ag_navi[:, None].expand(-1, R, -1)
        ag_tokens["ag_navi"] = ag_navi.reshape(n_sc * R, A)
        ag_tokens["ag_navi_valid"] = rep(ag_navi_valid)
        ag_tokens["ag_navi_log_prob"] = (torch.zeros(n_sc * R, A, device=ag_navi.device) if ag_navi_log_prob is None
                                         else ag_navi_log_prob.reshape(n_sc * R, A))
        mp_tokens = {k: rep(v) for k, v in mp_tokens.items()}
        tl_tokens = {k: (rep(v) if v is not None else None) for k, v in tl_tokens.items()}
        rule_checker = TrafficRuleChecker(
            mp_boundary=rep(batch["map/boundary"]), mp_valid=rep(batch["map/valid"]), mp_type=rep(batch["map/type"]),
            mp_pos=rep(batch["map/pos"]), mp_dir=rep(batch["map/dir"]), ag_type=ag_tokens["ag_type"],
            ag_size=ag_tokens["ag_size"], ag_goal=None, ag_dest=ag_tokens["ag_navi"], tl_valid=tl_tokens["tl_token_valid"],
            tl_pose=tl_tokens["tl_token_pose"], disable_check=self.training)
        buf = self.rollout(ag_tokens=ag_tokens, mp_tokens=mp_tokens, tl_tokens=tl_tokens,
                           tl_state_gt=rep(batch["sc/tl_state"]), teacher_forcing=teacher_forcing, rule_checker=rule_checker,
                           step_end=self.time_step_end, deterministic_action=True, n_joint_future=R)
        buf.flatten_joint_future(R)
        buf.compute_log_prob(ag_latent_log_prob)
        return buf
