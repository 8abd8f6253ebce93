"""Neighbour-list statistics of the pair attention kernel at config 3 (how many 8-row groups a token pair runs, against
what its valid neighbours need): python profiles/nbr_stats.py > gpurun_out/nbr_stats.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tbpkg  # noqa: E402,F401
from trafficbotsv1_5_b200 import config, params, synth  # noqa: E402
from trafficbotsv1_5_b200.engine import RolloutEngine  # noqa: E402

cfg = config.default_model_cfg()
eng = RolloutEngine(params.init_params(cfg, 0), cfg, "cuda", precision=1, n_rollout=32, step_end=90, use_graph=False)
eng.prepare(synth.make_scene_batch(n_sc=16, seed=1000))
st = eng._st
eng._reset(st)
for s_ in range(1, 61):
    aux = {} if s_ in (12, 30, 60) else None
    eng._step(st, eng._static, eng._navi, aux)
    if aux is None:
        continue
    torch.cuda.synchronize()
    for name, inv in (("self (agent->agent)", aux["knn_self"]["inv"]), ("cross (agent->map+TL)", aux["cinv"])):
        n = (~inv.bool()).sum(-1).reshape(-1).float()  # valid neighbours per token
        K = inv.shape[-1]
        a, b = n[0::2], n[1::2]
        now = torch.ceil(torch.maximum(a, b) / 8)
        flex = torch.ceil((a + b) / 16)
        srt = n.sort().values
        now_sorted = torch.ceil(torch.maximum(srt[0::2], srt[1::2]) / 8)
        print(f"step {s_:2d} {name:22s} K={K:3d}: valid/token mean {n.mean():5.1f} (zero: {100 * (n == 0).float().mean():4.1f} %), "
              f"groups/pair now {now.mean():5.2f}, rows free to mix {flex.mean():5.2f}, pairs sorted by count "
              f"{now_sorted.mean():5.2f}, lower bound {(a + b).mean() / 16:5.2f}")
