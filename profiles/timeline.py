"""Kernel timeline of graph-replayed policy iterations (torch.profiler / CUPTI): per-stream busy time, gaps on the main
stream, overlap of the side streams. Not a timing source (profiler overhead) - used to see WHERE the iteration waits.
  python profiles/timeline.py > gpurun_out/timeline.txt"""
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tbpkg  # noqa: E402,F401
from trafficbotsv1_5_b200 import config, params, synth  # noqa: E402
from trafficbotsv1_5_b200.engine import RolloutEngine  # noqa: E402

cfg = config.default_model_cfg()
eng = RolloutEngine(params.init_params(cfg, 0), cfg, "cuda", precision=1, n_rollout=32, step_end=90, use_graph=True)
eng.prepare(synth.make_scene_batch(n_sc=16, seed=1000))
eng.run()
torch.cuda.synchronize()
eng._reset(eng._st)
for s_ in range(1, 15):
    eng._graph[(s_ + 1) % 2].replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for s_ in range(15, 18):
        eng._graph[(s_ + 1) % 2].replay()
    torch.cuda.synchronize()
import json  # noqa: E402
import tempfile  # noqa: E402

path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
tr = json.load(open(path))
ev = [e for e in tr["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
print(f"{len(ev)} kernels in {t1 - t0:.0f} us over 3 iterations ({(t1 - t0) / 3:.0f} us per iteration under the profiler)")
streams = defaultdict(list)
for e in ev:
    streams[e["args"]["stream"]].append(e)
main_id = max(streams, key=lambda k: sum(e["dur"] for e in streams[k]))
for sid, lst in sorted(streams.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    busy = sum(e["dur"] for e in lst)
    print(f"stream {sid}{' (main)' if sid == main_id else ''}: {len(lst)} kernels, busy {busy / 3:.0f} us / iteration "
          f"({100 * busy / (t1 - t0):.1f} % of the window)")
main = streams[main_id]
def short(n):
    n = n.replace("void ", "").replace("(anonymous namespace)::", "")
    return n.split("(")[0][:34]


gaps = [(b["ts"] - a["ts"] - a["dur"], short(a["name"]), short(b["name"])) for a, b in zip(main, main[1:])]
small = [g[0] for g in gaps if g[0] < 20]
print(f"main-stream gaps: total {sum(g[0] for g in gaps) / 3:.0f} us / iteration; {len(small)} gaps < 20 us average "
      f"{sum(small) / len(small):.2f} us")
for g in sorted(gaps, reverse=True)[:10]:
    print(f"  gap {g[0]:7.1f} us  after {g[1]:34s} before {g[2]}")
by = defaultdict(float)
for e in ev:
    by[short(e["name"])] += e["dur"]
for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:12]:
    print(f"  {v / 3:8.1f} us/iter  {k}")

# one iteration as a compact timeline: offset [stream] kernel duration
it = [e for e in ev if e["ts"] >= main[len(main) // 3]["ts"] and e["ts"] < main[2 * len(main) // 3]["ts"]]
base = it[0]["ts"]
ids = {sid: i for i, sid in enumerate(sorted(streams, key=lambda k: -sum(e["dur"] for e in streams[k])))}
print("timeline of one iteration: offset_us [stream] kernel dur_us")
for e in it:
    print(f"{e['ts'] - base:8.1f} [{ids[e['args']['stream']]}] {short(e['name']):34s} {e['dur']:7.1f}")
