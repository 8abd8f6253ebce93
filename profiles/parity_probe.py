"""Closed-loop parity probe: engine (precision 0 and 1) vs the CPU oracle over all 90 policy iterations at the
BASELINE config-3 scene shape (128 agents, 1024 polylines x 20, 40 TL), several seeds. Prints the per-10-step maximum
position / yaw error and every mask mismatch (pred_valid, tl_state, final_navi_valid) with the step it first appears
at. Diagnostic companion of tests/test_rollout_gpu.py::test_rollout_config3_shape_90_steps_vs_oracle.

    python profiles/parity_probe.py [--seeds 31 32 33] [--rollouts 4] [--agents 128] [--polylines 1024]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tbpkg  # noqa: E402,F401
from oracle import tb_oracle as O  # noqa: E402
from trafficbotsv1_5_b200 import config, params, synth  # noqa: E402
from trafficbotsv1_5_b200.engine import RolloutEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs="+", default=[31, 32, 33])
    ap.add_argument("--rollouts", type=int, default=4)
    ap.add_argument("--agents", type=int, default=128)
    ap.add_argument("--polylines", type=int, default=1024)
    ap.add_argument("--tl", type=int, default=40)
    ap.add_argument("--steps", type=int, default=90)
    ap.add_argument("--boundary", type=float, default=400.0)
    ap.add_argument("--variants", action="store_true", help="also probe precision-1 variants (error attribution)")
    a = ap.parse_args()
    cfg = config.default_model_cfg()
    sz = config.derived_sizes(cfg)
    P = params.init_params(cfg, 0)
    R, T = a.rollouts, a.steps
    for seed in a.seeds:
        batch = synth.make_scene_batch(n_sc=1, n_ag=a.agents, n_mp=a.polylines, n_tl=a.tl, seed=seed, n_rollout=R,
                                       boundary=a.boundary)
        t0 = time.time()
        ref = O.rollout(P, cfg, sz, config.DYNAMICS_CFG, config.ROLLOUT_CFG, batch, R, T)
        print(f"seed {seed}: oracle {time.time() - t0:.1f} s; final disabled {int((~ref['final_valid']).sum())}, "
              f"dest reached {int((~ref['final_navi_valid']).sum())}", flush=True)
        runs = [(0, ""), (1, "")]
        if a.variants:
            runs += [(1, "kv_half=False (tf32 projections, fp32 intermediates, SIMT attention)"),
                     (1, "heads fp32"), (1, "heads fp32 + kv_half=False")]
        for prec, variant in runs:
            eng = RolloutEngine(P, cfg, "cuda", precision=prec, n_rollout=R, step_end=T)
            if "kv_half=False" in variant:
                eng.model.kv_half = False
            if "heads fp32" in variant:
                eng.model._probe_heads_fp32 = True
            if variant:
                print("  variant:", variant)
            res = {k: v.cpu() for k, v in eng.rollout(batch).items() if torch.is_tensor(v)}
            err = (res["pred_pose"] - ref["pred_pose"]).abs()
            both = res["pred_valid"] & ref["pred_valid"]
            err = err * both[..., None]
            xy = err[..., :2].amax(dim=(0, 1, 3))
            yaw = err[..., 2].amax(dim=(0, 1))
            print(f"  precision {prec}: xy  ", " ".join(f"{float(v):.1e}" for v in xy[9::10]))
            print(f"               yaw ", " ".join(f"{float(v):.1e}" for v in yaw[9::10]))
            for k in ("pred_valid", "tl_state", "final_valid", "final_navi_valid"):
                diff = res[k] != ref[k]
                n = int(diff.sum())
                msg = f"  precision {prec}: {k} mismatches {n}"
                if n and diff.dim() >= 3:
                    steps = diff.flatten(3).any(-1) if diff.dim() > 3 else diff
                    first = int(steps.any(0).any(0).float().argmax())
                    msg += f" (first at step index {first})"
                print(msg, flush=True)
            del eng
            torch.cuda.empty_cache()


if __name__ == "__main__":
    with torch.no_grad():
        main()
