"""Profiling driver: one eager (non-graph) policy iteration of the BASELINE config-3 workload (16 scenes x 32
rollouts) bracketed by cudaProfilerStart/Stop, after 12 unprofiled iterations (full 11-step histories).

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python profiles/profile_step.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tbpkg  # noqa: E402,F401
from trafficbotsv1_5_b200 import config, params, synth  # noqa: E402
from trafficbotsv1_5_b200.engine import RolloutEngine  # noqa: E402

n_sc = int(os.environ.get("TB_SCENES", "16"))
prec = int(os.environ.get("TB_PRECISION", "0"))
cfg = config.default_model_cfg()
eng = RolloutEngine(params.init_params(cfg, 0), cfg, "cuda", precision=prec, n_rollout=32, step_end=90, use_graph=False,
                    rule_checks=bool(int(os.environ.get("TB_RULE_CHECKS", "0"))))
eng.prepare(synth.make_scene_batch(n_sc=n_sc, seed=1000))
st = eng._st
eng._reset(st)
for _ in range(12):
    eng._step(st, eng._static, eng._navi)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(int(os.environ.get("TB_PROFILE_ITERS", "1"))):
    eng._step(st, eng._static, eng._navi)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled iterations done")
