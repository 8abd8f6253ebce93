"""Module-level drop-in for `models.traffic_bots.TrafficBots` (src/models/traffic_bots.py:17-221), HPTR / eval path.

Same `state_dict` keys (hot-path subset), same call surface as used by
`WaymoMotion.forward/rollout/test_step` (src/pl_modules/waymo_motion.py:163-177, 227, 847-851):

    model.mp_encoder(mp_valid, mp_attr, mp_pose, mp_type)            -> mp_tokens   (dict of [n_sc, ...] tensors)
    model.tl_encoder.pre_compute(tl_valid, tl_attr, tl_pose, **mp_tokens) -> tl_tokens
    model.init()
    model(ag_valid, ag_pose, ag_motion, ag_attr, ag_type, ag_latent, ag_latent_valid, ag_navi, ag_navi_valid,
          ag_navi_updated, tl_state, tl_tokens, mp_tokens)           -> (Independent(Normal), Categorical)

so the reference's own Python rollout loop (Dynamics, TeacherForcing, TrafficRuleChecker, RolloutBuffer) can drive the
CUDA policy step unchanged, including its `repeat_interleave(n_joint_future, 0)` of every token tensor (:458-462): all
dict values are tensors whose leading dimension is the scene. The dicts carry extra `b200_*` entries (per-layer K/V
tables of the map tokens) next to the reference's keys. For full speed use `engine.RolloutEngine` instead, which keeps
the whole loop on the device.
"""
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn
from torch.distributions import Categorical, Independent, Normal

from . import config as C
from . import lib as L
from . import ops
from . import params
from .model import HotPathModel


class _Tree(nn.Module):
    """Generic container reproducing the reference's module tree from dotted parameter names."""

    def __init__(self):
        super().__init__()
        self._fn = None

    def forward(self, *a, **k):
        if self._fn is None:
            raise NotImplementedError("this sub-module is a parameter container in the B200 drop-in")
        return self._fn(*a, **k)

    def add(self, dotted: str, tensor: Tensor, buffer: bool = False) -> None:
        head, _, rest = dotted.partition(".")
        if not rest:
            if buffer:
                self.register_buffer(head, tensor)
            else:
                self.register_parameter(head, nn.Parameter(tensor))
            return
        if head not in self._modules:
            self.add_module(head, _Tree())
        self._modules[head].add(rest, tensor, buffer)


class TrafficBots(_Tree):
    def __init__(self, cfg: Optional[dict] = None, precision: int = 0, seed: int = 0, training_modules: bool = False,
                 **overrides) -> None:
        """`training_modules`: also register `navi_predictor.*` and the posterior half of `latent_encoder.*` (the
        modules only the training step uses; a full reference checkpoint then loads with strict=False leaving only the
        unused prior encoders unmatched)."""
        super().__init__()
        self.cfg = cfg or C.default_model_cfg()
        self.cfg.update(overrides)
        self.sz = C.derived_sizes(self.cfg)
        self.precision = precision
        self.temp_window_size = self.cfg["temp_window_size"]
        for k, v in params.init_params(self.cfg, seed, with_navi_predictor=training_modules,
                                       with_latent_post=training_modules).items():
            self.add(k, v)
        d, W, Ln = self.cfg["hidden_dim"], self.temp_window_size, self.cfg["n_mp_pl_node"]
        # persistent buffers of the reference (App. B); one shared PoseEmb is registered under several names
        fx = lambda pe: (1.0 / (1e3 ** (torch.arange(0, pe // 4, 2)[: pe // 8].float() / (pe // 4)))).repeat_interleave(2)  # noqa
        fy = lambda pe: (torch.arange(0, pe // 4) + 1.0).repeat_interleave(2, 0)  # noqa: E731
        for p in ("pose_rpe", "mp_encoder.pose_rpe", "tl_encoder.pose_rpe", "ag_encoder.pose_rpe", "navi_encoder.pose_emb"):
            self.add(f"{p}.pe_xy.freqs", fx(d), buffer=True)
            self.add(f"{p}.pe_yaw.freqs", fy(d), buffer=True)
        self.add("ag_encoder.pose_emb.pe_xy.freqs", fx(d // 2), buffer=True)
        self.add("ag_encoder.pose_emb.pe_yaw.freqs", fy(d // 2), buffer=True)
        self.add("mp_encoder.pl_node_ohe", torch.eye(Ln)[None, None], buffer=True)
        self.add("tl_encoder.hist_ohe", torch.eye(W), buffer=True)
        self.add("ag_encoder.hist_ohe", torch.eye(W), buffer=True)
        self.mp_encoder._fn = self._mp_encoder
        self.tl_encoder.pre_compute = self._tl_pre_compute
        self._hp: Optional[HotPathModel] = None
        self._hp_ver = None
        self.init()

    # ------------------------------------------------------------------------------------------ weights -> kernels
    def _runner(self) -> HotPathModel:
        sd = {k: v for k, v in self.state_dict().items() if not k.endswith(("freqs", "_ohe"))
              and not k.startswith("latent_encoder.")}
        ver = tuple((k, v._version, v.data_ptr()) for k, v in sd.items())
        if self._hp_ver != ver:
            dev = next(iter(sd.values())).device
            if dev.type != "cuda":
                raise RuntimeError("TrafficBots (B200) runs on CUDA only — there is no CPU fallback; call .cuda()")
            self._hp = HotPathModel(sd, self.cfg, self.sz, dev, self.precision)
            self._hp_ver = ver
        return self._hp

    # ------------------------------------------------------------------------------------------ scene encoders
    @torch.no_grad()
    def _mp_encoder(self, mp_valid: Tensor, mp_attr: Tensor, mp_pose: Tensor, mp_type: Tensor) -> Dict[str, Tensor]:
        """MapEncoder.forward (map_encoder.py:50-113) + per-layer map K/V tables of the agent decoder."""
        m = self._runner()
        mp = m.map_encoder(mp_valid, mp_attr.float(), mp_pose)
        n_sc, n_mp, d = mp["mp_token_feature"].shape
        out = dict(mp_token_invalid=mp["mp_token_invalid"], mp_token_feature=mp["mp_token_feature"],
                   mp_token_pose=mp["mp_token_pose"], mp_token_type=mp_type)
        for i, kv in enumerate(m.ag_static(mp)):
            out[f"b200_kv_ag_{i}"] = kv.view(n_sc, n_mp, 2 * d)
        return out

    @torch.no_grad()
    def _tl_pre_compute(self, tl_valid: Tensor, tl_attr: Tensor, tl_pose: Tensor, mp_token_invalid: Tensor,
                        mp_token_feature: Tensor, mp_token_pose: Tensor, **kwargs) -> Dict[str, Tensor]:
        """TrafficLightEncoder.pre_compute (traffic_light.py:76-154)."""
        m = self._runner()
        mp = dict(mp_token_invalid=mp_token_invalid, mp_token_feature=mp_token_feature, mp_token_pose=mp_token_pose)
        tl = m.tl_pre_compute(tl_valid, tl_attr, tl_pose, mp)
        n_sc, n_tl = tl_valid.shape
        n_mp, d = mp_token_pose.shape[1], m.d
        ks, cr = tl["knn_self"], tl["cross"][0]
        out = dict(tl_token_valid=tl_valid, tl_token_invalid=tl["tl_token_invalid"], tl_token_pose=tl["tl_token_pose"],
                   tl_token_attr=tl["tl_token_attr"].view(n_sc, n_tl, d),
                   knn_idx_tl2tl=ks["idx"].long(), knn_invalid_tl2tl=ks["inv"], rpe_tl2tl=ks["rel"],
                   knn_idx_tl2mp=cr["idx"].long(), knn_invalid_tl2mp=cr["inv"], rpe_tl2mp=cr["rel"])
        for i, c in enumerate(tl["cross"]):
            out[f"b200_kv_tl_{i}"] = c["kv0"].view(n_sc, n_mp, 2 * d)
        return out

    # ------------------------------------------------------------------------------------------ stateful policy step
    def init(self) -> None:
        """traffic_bots.py:145-149."""
        self._t = 0
        self._st = None
        self._rt = None
        self.navi_feature = None

    def _runtime(self, tl_tokens, mp_tokens, B, A, dev):
        m, W, d = self._hp, self.temp_window_size, self._hp.d
        n_mp, n_tl = mp_tokens["mp_token_pose"].shape[1], tl_tokens["tl_token_pose"].shape[1]
        mp = dict(mp_token_invalid=mp_tokens["mp_token_invalid"].contiguous(),
                  mp_token_feature=mp_tokens["mp_token_feature"].contiguous(),
                  mp_token_pose=mp_tokens["mp_token_pose"].contiguous())
        kv_mp = [mp_tokens[f"b200_kv_ag_{i}"].reshape(B * n_mp, 2 * d) for i in range(self.cfg["ag_encoder"]["n_layer_tf"])]
        attr = tl_tokens["tl_token_attr"].reshape(B * n_tl, d)
        knn = lambda i, v, r: dict(idx=tl_tokens[i].to(torch.int32).contiguous(), inv=tl_tokens[v].contiguous(),  # noqa
                                   rel=tl_tokens[r].contiguous())
        c0 = knn("knn_idx_tl2mp", "knn_invalid_tl2mp", "rpe_tl2mp")
        tl = dict(n_sc=B, n_tl=n_tl, tl_token_invalid=tl_tokens["tl_token_invalid"].contiguous(),
                  tl_token_pose=tl_tokens["tl_token_pose"].contiguous(),
                  tl_attr_rows=attr.view(B * n_tl, 1, d).expand(-1, W, -1).reshape(-1, d).contiguous(),
                  knn_self=knn("knn_idx_tl2tl", "knn_invalid_tl2tl", "rpe_tl2tl"),
                  cross=[dict(c0, kv0=tl_tokens[f"b200_kv_tl_{i}"].reshape(B * n_mp, 2 * d), T0=n_mp, div0=1,
                              K0=self.sz["k_tl2mp"]) for i in range(self.cfg["tl_encoder"]["n_layer_tf"])])
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)  # noqa: E731
        st = dict(B=B, A=A, d_step=z(1, dt=torch.int32), hist_valid=z(B, A, W, dt=torch.uint8), hist_pose=z(B, A, W, 3),
                  hist_motion=z(B, A, W, 3), hist_tl=z(B, n_tl, W, 5, dt=torch.uint8), x_cat=z(B * A, 2 * d))
        return dict(mp=mp, kv_mp=kv_mp, tl=tl), st

    @torch.no_grad()
    def forward(self, ag_valid: Tensor, ag_pose: Tensor, ag_motion: Tensor, ag_attr: Tensor, ag_type: Tensor,
                ag_latent: Optional[Tensor], ag_latent_valid: Optional[Tensor], ag_navi: Optional[Tensor],
                ag_navi_valid: Tensor, ag_navi_updated: bool, tl_state: Tensor, tl_tokens: Dict[str, Tensor],
                mp_tokens: Dict[str, Tensor]) -> Tuple[Independent, Categorical]:
        """TrafficBots.forward (traffic_bots.py:151-221), eval mode, navi_mode "dest", latent_dim > 0."""
        if self.training:
            raise NotImplementedError("the B200 drop-in implements the inference rollout path: call .eval()")
        m = self._runner()
        B, A = ag_valid.shape
        dev, W, d = ag_pose.device, self.temp_window_size, m.d
        if self._rt is None:
            self._rt, self._st = self._runtime(tl_tokens, mp_tokens, B, A, dev)
            self._navi = None
        rt, st = self._rt, self._st
        # _append_hist (:123-143): ring slot = call index % W; the kernels read the window through d_step
        slot = self._t % W
        st["hist_valid"][:, :, slot] = ag_valid
        st["hist_pose"][:, :, slot] = ag_pose
        st["hist_motion"][:, :, slot] = ag_motion
        st["hist_tl"][:, :, slot] = tl_state
        st["d_step"].fill_(self._t + 1)
        st.update(ag_attr=ag_attr.float().contiguous(), pose=ag_pose.contiguous(),
                  navi_invalid=(~ag_navi_valid).contiguous())
        if self._navi is None or ag_navi_updated:                                                   # :191-194
            self._navi = m.navi_static(rt["mp"], ag_navi.to(torch.int32).contiguous(), 1)
            st.update(latent=ag_latent.reshape(B * A, -1).float().contiguous(),
                      latent_invalid=(~ag_latent_valid).contiguous())
            m.latent_static(self._navi, st)
        tl_feat, logits = m.tl_forward(st["hist_tl"], st["d_step"], rt["tl"])                        # :197
        m.ag_forward(st, rt["mp"], rt["kv_mp"], rt["tl"], tl_feat, 1, out=st["x_cat"][:, :d])        # :200
        act = m.heads(st["x_cat"], st, self._navi)                                                   # :213-217
        mean = torch.empty(B * A, 2, device=dev)
        ty = ag_type.to(torch.uint8).contiguous()
        L.check(L.load().tb_action_mean(L.ptr(act), L.ptr(ty), L.ptr(ag_valid.to(torch.uint8).contiguous()), B * A,
                                        L.ptr(mean), L.stream()), "tb_action_mean")
        ops._count()
        log_std = torch.stack([m.P[f"action_head.log_std.{t}"] for t in range(3)], 0)                # action_head.py:91-95
        log_std = (ag_type & ag_valid.unsqueeze(-1)).float() @ log_std
        n_tl = rt["tl"]["n_tl"]
        lg = logits.view(B, n_tl, -1).masked_fill(rt["tl"]["tl_token_invalid"].unsqueeze(-1), 0.0).clamp(-3, 3)  # :284-286
        self._t += 1
        return Independent(Normal(mean.view(B, A, 2), log_std.exp()), 1), Categorical(logits=lg)
