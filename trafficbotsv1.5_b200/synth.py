"""Seeded synthetic WOMD-shaped scenes (there is no dataset in the build/GPU containers).

The dict layout is what the reference's `SceneCentricPreProcessing.forward` hands to the model at
test time (`src/data_modules/scene_centric.py:39-147`; tensor sizes `data_h5_womd.py:95-140`), i.e.
the "sc/*", "ref/*" and "map/*" keys consumed by `WaymoMotion.test_step` / `joint_future_pred`
(`src/pl_modules/waymo_motion.py:843-876, 439-524`). Recipe: SURVEY.md §8(d).
"""
import math
from typing import Dict

import torch


def make_scene_batch(n_sc: int, n_ag: int = 128, n_mp: int = 1024, n_tl: int = 40, n_node: int = 20,
                     n_hist: int = 11, seed: int = 1000, boundary: float = 400.0,
                     n_rollout: int = 32, latent_dim: int = 16, scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """All tensors fp32/bool/int64 on CPU. Scene s is drawn from seed `seed + s`. `scale` shrinks/expands the
    spatial extent of map and agents (dense scenes exercise the collision / road-edge checks)."""
    out = {}
    scenes = [_one_scene(n_ag, n_mp, n_tl, n_node, n_hist, seed + s, boundary, n_rollout, latent_dim, scale)
              for s in range(n_sc)]
    for k in scenes[0]:
        out[k] = torch.stack([sc[k] for sc in scenes], 0)
    return out


def _one_scene(n_ag, n_mp, n_tl, n_node, n_hist, seed, boundary, n_rollout, latent_dim, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    U = lambda *s, lo=0.0, hi=1.0: torch.rand(*s, generator=g) * (hi - lo) + lo  # noqa: E731
    d = {}
    # ---- map polylines: straight, 1 m node spacing, node 0 is the token pose (map_encoder.py:65)
    p0 = U(n_mp, 2, lo=-150.0 * scale, hi=150.0 * scale)
    hd = U(n_mp, lo=-math.pi, hi=math.pi)
    dirv = torch.stack([hd.cos(), hd.sin()], -1)  # [n_mp, 2]
    t = torch.arange(n_node, dtype=torch.float32)
    pos = p0[:, None, :] + t[None, :, None] * dirv[:, None, :]  # [n_mp, n_node, 2]
    mp_valid = torch.ones(n_mp, n_node, dtype=torch.bool)
    tail = (U(n_mp) < 0.2)
    n_keep = torch.randint(1, n_node, (n_mp,), generator=g)
    mp_valid &= ~(tail[:, None] & (t[None, :] >= n_keep[:, None]))
    mp_valid &= ~(U(n_mp) < 0.05)[:, None]
    mp_type = torch.nn.functional.one_hot(torch.randint(0, 11, (n_mp,), generator=g), 11).bool()
    d["map/valid"] = mp_valid
    d["map/type"] = mp_type
    d["map/pos"] = torch.cat([pos, torch.zeros(n_mp, n_node, 1)], -1)
    d["map/dir"] = torch.cat([dirv[:, None, :].expand(-1, n_node, -1), torch.zeros(n_mp, n_node, 1)], -1).contiguous()
    d["map/boundary"] = torch.tensor([-boundary, boundary, -boundary, boundary])
    d["sc/mp_valid"] = mp_valid.clone()
    d["sc/mp_attr"] = mp_type.float()
    d["sc/mp_pose"] = torch.cat([pos, hd[:, None, None].expand(-1, n_node, 1)], -1).contiguous()
    d["ref/mp_type"] = mp_type
    # ---- traffic lights on lanes (tl_mode == "lane"; scene_centric.py:97-100)
    tl_idx = torch.randperm(n_mp, generator=g)[:n_tl]
    d["sc/tl_attr"] = tl_idx
    d["sc/tl_pose"] = d["sc/mp_pose"][tl_idx, 0].clone()
    d["sc/tl_valid"] = U(n_tl) < 0.9
    st0 = torch.randint(0, 5, (n_tl,), generator=g)
    st1 = torch.randint(0, 5, (n_tl,), generator=g)
    t_sw = torch.randint(0, n_hist + 1, (n_tl,), generator=g)
    st = torch.where(torch.arange(n_hist)[None, :] < t_sw[:, None], st0[:, None], st1[:, None])
    d["sc/tl_state"] = torch.nn.functional.one_hot(st, 5).bool()  # [n_tl, n_hist, 5]
    # ---- agents: constant-velocity history, dt = 0.1 s
    a0 = U(n_ag, 2, lo=-100.0 * scale, hi=100.0 * scale)
    yaw = U(n_ag, lo=-math.pi, hi=math.pi)
    spd = U(n_ag, lo=0.0, hi=10.0)
    th = torch.arange(n_hist, dtype=torch.float32) * 0.1
    vel = torch.stack([yaw.cos(), yaw.sin()], -1) * spd[:, None]
    apos = a0[:, None, :] + th[None, :, None] * vel[:, None, :]
    d["sc/ag_pose"] = torch.cat([apos, yaw[:, None, None].expand(-1, n_hist, 1)], -1).contiguous()
    d["sc/ag_motion"] = torch.stack([spd[:, None].expand(-1, n_hist), torch.zeros(n_ag, n_hist),
                                     torch.zeros(n_ag, n_hist)], -1).contiguous()
    ag_valid = torch.ones(n_ag, n_hist, dtype=torch.bool)
    ag_valid &= ~(U(n_ag) < 0.1)[:, None]
    late = U(n_ag) < 0.1
    t_in = torch.randint(1, n_hist, (n_ag,), generator=g)
    ag_valid &= ~(late[:, None] & (torch.arange(n_hist)[None, :] < t_in[:, None]))
    d["sc/ag_valid"] = ag_valid
    ag_type = torch.nn.functional.one_hot(torch.multinomial(torch.tensor([0.7, 0.2, 0.1]), n_ag, True, generator=g),
                                          3).bool()
    size = U(n_ag, 3, lo=1.0, hi=5.0)
    d["ref/ag_type"] = ag_type
    d["ref/ag_size"] = size
    d["sc/ag_attr"] = torch.cat([size, ag_type.float()], -1)
    # ---- fixed navigation + latent samples (north star: "fixed latent samples")
    d["agent/dest"] = torch.randint(0, n_mp, (n_ag,), generator=g)
    d["ag_navi_valid"] = ag_valid.any(-1)
    gl = torch.Generator().manual_seed(seed + 1000)
    d["ag_latent"] = torch.randn(n_rollout, n_ag, latent_dim, generator=gl)
    d["ag_latent_valid"] = ag_valid.any(-1)
    return d


def make_train_batch(n_sc: int, n_ag: int = 128, n_mp: int = 1024, n_tl: int = 40, n_step: int = 91, n_hist: int = 11,
                     seed: int = 3000, p_forcing_agent: float = 0.3, **kw) -> Dict[str, torch.Tensor]:
    """The dict `SceneCentricPreProcessing` hands to `training_step` (scene_centric.py:39-147): "gt/*" = the whole
    n_step-long track (here: speed ramps and constant yaw rates, so position, heading and speed errors are all
    exercised), "sc/*" = its first n_hist steps, plus the step's random draws as inputs — "tf/forcing_agent"
    (teacher_forcing.py:87-92), "ag_latent_eps" (rsample noise of the latent), "gt/ag_navi" a destination that passes
    the type masks of the destination classifier (navigation.py:265-278), "ref/ag_role"."""
    b = make_scene_batch(n_sc=n_sc, n_ag=n_ag, n_mp=n_mp, n_tl=n_tl, n_hist=n_step, seed=seed, n_rollout=1, **kw)
    g = torch.Generator().manual_seed(seed + 31)
    t = torch.arange(n_step, dtype=torch.float32) * 0.1
    spd0 = b["sc/ag_motion"][:, :, 0, 0]
    acc = (torch.rand(n_sc, n_ag, generator=g) * 2 - 1) * 1.0
    yr = (torch.rand(n_sc, n_ag, generator=g) * 2 - 1) * 0.2
    spd = (spd0[..., None] + acc[..., None] * t).clamp(min=0.0)
    yaw = b["sc/ag_pose"][:, :, 0, 2][..., None] + yr[..., None] * t
    vel = torch.stack([yaw.cos(), yaw.sin()], -1) * spd[..., None]
    xy = b["sc/ag_pose"][:, :, :1, :2] + torch.cumsum(vel * 0.1, 2) - vel[:, :, :1] * 0.1
    b["gt/ag_pose"] = torch.cat([xy, yaw[..., None]], -1).contiguous()
    b["gt/ag_motion"] = torch.stack([spd, acc[..., None].expand(-1, -1, n_step), yr[..., None].expand(-1, -1, n_step)],
                                    -1).contiguous()
    gv = b["sc/ag_valid"].clone()
    drop = torch.rand(n_sc, n_ag, generator=g) < 0.15   # tracks that end early
    t_end = torch.randint(n_hist + 5, n_step, (n_sc, n_ag), generator=g)
    gv &= ~(drop[..., None] & (torch.arange(n_step)[None, None] >= t_end[..., None]))
    b["gt/ag_valid"] = gv
    b["gt/tl_state"] = b["sc/tl_state"]
    for k in ("ag_valid", "ag_pose", "ag_motion"):
        b[f"sc/{k}"] = b[f"gt/{k}"][:, :, :n_hist].contiguous()
    b["sc/tl_state"] = b["gt/tl_state"][:, :, :n_hist].contiguous()
    # a destination that survives the type masks: FREEWAY..ROAD_EDGE_BOUNDARY minus the per-type exclusions
    mt, at = b["ref/mp_type"], b["ref/ag_type"]
    ok = (mt[:, :, :5].any(-1) & b["sc/mp_valid"][:, :, 0])[:, None] & ~(
        (at[:, :, [0]] & mt[:, :, 3][:, None]) | (at[:, :, [1]] & mt[:, :, :4].any(-1)[:, None])
        | (at[:, :, [2]] & mt[:, :, :3].any(-1)[:, None]))
    w = ok.float() + 1e-9
    b["gt/ag_navi"] = torch.multinomial(w.view(n_sc * n_ag, -1), 1, generator=g).view(n_sc, n_ag)
    b["agent/dest"] = b["gt/ag_navi"]
    b["tf/forcing_agent"] = torch.rand(n_sc, n_ag, generator=g) < p_forcing_agent
    b["ag_latent_eps"] = torch.randn(n_sc, n_ag, b["ag_latent"].shape[-1], generator=g)
    b["ref/ag_role"] = torch.rand(n_sc, n_ag, 3, generator=g) < 0.2
    b["ag_navi_valid"] = gv.any(-1)
    b["ag_latent_valid"] = gv.any(-1)
    return b


def make_rule_scene_batch(n_sc: int, n_ag: int = 48, n_mp: int = 96, n_tl: int = 40, seed: int = 5000,
                          boundary: float = 80.0, scale: float = 0.4, **kw) -> Dict[str, torch.Tensor]:
    """A `make_scene_batch` scene edited so that the two rare TrafficRuleChecker events fire often
    (traffic_rule_checker.py:176-274): the first n_tl agents are long, fast vehicles with a RED traffic light 0.5-3 m
    ahead of their start pose (they drive over it during the ground-truth warm start -> `run_red_light`), the next
    n_ag/4 agents are slow vehicles standing on distinct lane-centre polylines with nothing ahead (-> `passive` once
    their counter passes 20 steps)."""
    b = make_scene_batch(n_sc=n_sc, n_ag=n_ag, n_mp=n_mp, n_tl=n_tl, seed=seed, boundary=boundary, scale=scale, **kw)
    n_hist = b["sc/ag_valid"].shape[-1]
    th = torch.arange(n_hist, dtype=torch.float32) * 0.1
    veh = torch.tensor([True, False, False])
    n_red, n_slow = min(n_tl, n_ag // 2), n_ag // 4
    for s in range(n_sc):
        g = torch.Generator().manual_seed(seed + 77 * (s + 1))
        U = lambda n, lo, hi: torch.rand(n, generator=g) * (hi - lo) + lo  # noqa: E731

        def put(i, xy0, yaw, spd, length):
            hd = torch.stack([torch.cos(yaw), torch.sin(yaw)])
            b["sc/ag_pose"][s, i, :, :2] = xy0[None] + th[:, None] * spd * hd[None]
            b["sc/ag_pose"][s, i, :, 2] = yaw
            b["sc/ag_motion"][s, i] = 0.0
            b["sc/ag_motion"][s, i, :, 0] = spd
            b["sc/ag_valid"][s, i] = True
            b["ref/ag_type"][s, i] = veh
            b["ref/ag_size"][s, i, 0] = length
            b["ref/ag_size"][s, i, 1] = 2.0

        spd, length, dist = U(n_red, 5.0, 10.0), U(n_red, 3.0, 5.0), U(n_red, 0.5, 3.0)
        for j in range(n_red):  # a red light ahead of agent j
            xy0, yaw = b["sc/ag_pose"][s, j, 0, :2].clone(), b["sc/ag_pose"][s, j, 0, 2].clone()
            put(j, xy0, yaw, spd[j], length[j])
            hd = torch.stack([torch.cos(yaw), torch.sin(yaw)])
            b["sc/tl_pose"][s, j, :2] = xy0 + dist[j] * hd
            b["sc/tl_pose"][s, j, 2] = yaw
            b["sc/tl_valid"][s, j] = True
            b["sc/tl_state"][s, j] = False
            b["sc/tl_state"][s, j, :, 1] = True  # LANE_STATE_STOP during the whole history
        lanes = torch.nonzero(b["map/type"][s, :, :3].any(-1) & b["map/valid"][s].all(-1)).flatten()
        lanes = lanes[torch.randperm(len(lanes), generator=g)][:n_slow]
        slow = U(len(lanes), 0.0, 1.0)
        for k, pl in enumerate(lanes.tolist()):  # a slow vehicle on lane polyline pl, node 5
            put(n_red + k, b["sc/mp_pose"][s, pl, 5, :2].clone(), b["sc/mp_pose"][s, pl, 0, 2].clone(), slow[k],
                torch.tensor(4.0))
    b["sc/ag_attr"] = torch.cat([b["ref/ag_size"], b["ref/ag_type"].float()], -1)
    b["ag_navi_valid"] = b["sc/ag_valid"].any(-1)
    b["ag_latent_valid"] = b["sc/ag_valid"].any(-1)
    return b


def make_wosac_post_inputs(seed: int, n_sc: int, K: int, A: int, T: int):
    """Seeded inputs of the WOSAC post-processing fixture (wosac_post_processing.py:31-75; shared by the golden
    generator and the tests): scene-centric joint futures, violation
    flags whose role-weighted counts are distinct per future (so the reference's top-k has no ties), roles, and the
    scene -> global transform."""
    g = torch.Generator().manual_seed(seed)
    pose = torch.cat([(torch.rand(n_sc * K, A, T, 2, generator=g) * 2 - 1) * 150,
                      (torch.rand(n_sc * K, A, T, 1, generator=g) * 2 - 1) * 3.1], -1)
    role = torch.rand(n_sc, A, 3, generator=g) < 0.5
    role[:, :, 0] |= ~role.any(-1)  # every agent has a role here: counts 0..K-1 need K-1 counted agents
    col = torch.zeros(n_sc, K, A, T, dtype=torch.bool)
    for s in range(n_sc):
        perm = torch.randperm(K, generator=g)
        for k in range(K):  # future k: perm[k] agents collide at some future step
            ags = torch.randperm(A, generator=g)[: int(perm[k])]
            col[s, k, ags, torch.randint(4, T, (len(ags),), generator=g)] = True
    col[:, :, :, :4] |= torch.rand(n_sc, K, A, 4, generator=g) < 0.3  # history steps must be ignored
    edge = torch.rand(n_sc, K, A, T, generator=g) < 0.02
    center = (torch.rand(n_sc, 2, generator=g) * 2 - 1) * 5000
    yaw = (torch.rand(n_sc, generator=g) * 2 - 1) * 3.1
    return dict(pose=pose, role=role, collided=col, run_road_edge=edge, center=center, yaw=yaw)


def make_womd_post_inputs(seed: int, n_sc: int, K: int, A: int, T: int):
    """Seeded inputs of the WOMD post-processing fixture (womd_post_processing.py:36-72): joint futures that form a few
    clusters per agent (so the 2 m ADE test of `mpa_nms` has both outcomes), distinct log-prob scores, agent types."""
    g = torch.Generator().manual_seed(seed)
    base = (torch.rand(n_sc, 4, A, 1, 2, generator=g) * 2 - 1) * 30           # 4 cluster end points per agent
    which = torch.randint(0, 4, (n_sc, K, A), generator=g)
    end = torch.gather(base.expand(-1, -1, -1, 1, -1), 1, which[..., None, None].expand(-1, -1, -1, 1, 2))
    t = torch.linspace(0, 1, T).view(1, 1, 1, T, 1)
    xy = end * t + torch.randn(n_sc, K, A, T, 2, generator=g) * 0.4
    yaw = (torch.rand(n_sc, K, A, T, 1, generator=g) * 2 - 1) * 3.1
    trajs = torch.cat([xy, yaw], -1).contiguous()
    scores = torch.randn(n_sc, K, A, generator=g) * 1.5                        # log-probs, distinct with probability 1
    ag_type = torch.zeros(n_sc, A, 3, dtype=torch.bool)
    ag_type[torch.arange(n_sc)[:, None], torch.arange(A)[None], torch.randint(0, 3, (n_sc, A), generator=g)] = True
    return dict(trajs=trajs, scores=scores, ag_type=ag_type)
