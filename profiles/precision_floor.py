"""Precision floor of the closed loop (CPU, oracle only): the reference-order oracle with ONLY its nn.Linear weights
rounded to fp16 (10-bit mantissa, = tf32) and, separately, only the activations entering its nn.Linear calls rounded,
over 90 policy iterations on the config-1 scene. Weight rounding is a fixed perturbation of the policy and integrates
to a ~t^2 position drift; activation rounding changes from step to step and averages out. This is the noise floor the
16-bit mode's stated tolerance (tests/test_rollout_gpu.py) is derived from. Output: profiles/r2/precision_floor.txt
(max |xy - fp32 oracle| at policy iterations 10, 20, ..., 90, metres)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tbpkg  # noqa: E401,F401
import torch.nn.functional as F
from oracle import tb_oracle as O
from trafficbotsv1_5_b200 import config, params, synth
torch.set_num_threads(8)
cfg=config.default_model_cfg(); sz=config.derived_sizes(cfg); P=params.init_params(cfg,0)
shape=dict(n_sc=1,n_ag=64,n_mp=256,n_tl=40,seed=31,boundary=120.0)
batch=synth.make_scene_batch(**shape)
R,T=2,90
orig=F.linear
def h16(x): return x.half().float()
idn=lambda x:x
def mk(f, fw):
    def lin(x,w,b=None):
        return orig(f(x.contiguous()), fw(w.contiguous()), b)
    return lin
with torch.no_grad():
    ref=O.rollout(P,cfg,sz,config.DYNAMICS_CFG,config.ROLLOUT_CFG,batch,R,T)
    for name,f,fw in (("fp16 weights only",idn,h16),("fp16 activations only",h16,idn)):
        O.F.linear=mk(f,fw)
        try:
            r=O.rollout(P,cfg,sz,config.DYNAMICS_CFG,config.ROLLOUT_CFG,batch,R,T)
        finally:
            O.F.linear=orig
        err=(r["pred_pose"]-ref["pred_pose"]).abs()
        xy=err[...,:2].amax(dim=(0,1,3)); 
        print(name, " ".join(f"{float(v):.1e}" for v in xy[9::10]), flush=True)
