"""Per-kernel device time of one training step (BASELINE config 4), CUPTI via torch.profiler (not a bench number):
  gpurun -- python profiles/train_profile.py 64 > gpurun_out/train_profile.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tbpkg  # noqa: E402,F401
from trafficbotsv1_5_b200 import config, params, synth  # noqa: E402
from trafficbotsv1_5_b200.training import TrainStep  # noqa: E402

n_sc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = config.default_model_cfg()
P = params.init_params(cfg, 0, with_navi_predictor=True, with_latent_post=True)
batch = synth.make_train_batch(n_sc, seed=3000)
ts = TrainStep(P, cfg, "cuda:0", precision=prec)
ts.step(batch)
ts.zero_grad()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    ts.step(batch)
    torch.cuda.synchronize()
print("phases_ms", ts.timings_ms())
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.device_time_total for e in ev)
print(f"total device time {tot / 1e3:.1f} ms over {sum(e.count for e in ev)} kernels")
for e in sorted(ev, key=lambda e: -e.device_time_total)[:40]:
    print(f"{e.device_time_total / 1e3:9.2f} ms {100 * e.device_time_total / tot:5.1f}%  n={e.count:6d}  {e.key[:110]}")
