"""Parameter inventory of the hot-path modules, keyed by the reference's `state_dict` names
(dumped from `models.traffic_bots.TrafficBots`, src/models/traffic_bots.py:17-121).

`param_shapes` is the boundary contract (a reference checkpoint's tensors of these names load
unchanged); `init_params` draws a seeded random-init set for benchmarks and parity tests — there is no
network for checkpoints. `latent_encoder.*` / `navi_predictor.*` are off the rollout path and omitted.
"""
from collections import OrderedDict
from typing import Dict, Tuple

import torch


def _tf_layer(shapes, p, d, d_rpe, dec):
    def attn(q):
        shapes[f"{q}.in_proj_weight"] = (3 * d, d)
        shapes[f"{q}.out_proj_weight"] = (d, d)
        shapes[f"{q}.in_proj_bias"] = (3 * d,)
        shapes[f"{q}.out_proj_bias"] = (d,)
        shapes[f"{q}.linear_rpe.weight"] = (2 * d, d_rpe)
        shapes[f"{q}.linear_rpe.bias"] = (2 * d,)

    def ln(q):
        shapes[f"{q}.weight"] = (d,)
        shapes[f"{q}.bias"] = (d,)

    ln(f"{p}.norm1"); ln(f"{p}.norm_tgt")
    if dec:
        attn(f"{p}.attn_src"); ln(f"{p}.norm_src")
    attn(f"{p}.attn")
    shapes[f"{p}.linear1.weight"] = (4 * d, d); shapes[f"{p}.linear1.bias"] = (4 * d,)
    shapes[f"{p}.linear2.weight"] = (d, 4 * d); shapes[f"{p}.linear2.bias"] = (d,)
    ln(f"{p}.norm2")


def _mlp(shapes, p, dims, idxs):
    for i, (a, b) in zip(idxs, zip(dims[:-1], dims[1:])):
        shapes[f"{p}.fc_layers.{i}.weight"] = (b, a)
        shapes[f"{p}.fc_layers.{i}.bias"] = (b,)


def param_shapes(cfg: dict) -> "OrderedDict[str, Tuple[int, ...]]":
    d = cfg["hidden_dim"]
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    L = cfg["n_mp_pl_node"]
    W = cfg["temp_window_size"]
    # map encoder (map_encoder.py:29-48): input MLP [attr+L -> d-7]*3 (cat 7 polyline feats), PointNet, 8 enc layers
    _mlp(s, "mp_encoder.input_encoder.mlp", [cfg["mp_attr_dim"] + L] + [d - 7] * 3, (0, 2, 4))
    for i in range(3):
        _mlp(s, f"mp_encoder.pl_encoder.mlp_layers.{i}", [d, d // 2], (0,))
    for i in range(cfg["mp_encoder"]["n_layer_tf"]):
        _tf_layer(s, f"mp_encoder.tf_mp2mp.layers.{i}", d, d, False)
    # traffic-light encoder (traffic_light.py:40-74): input MLP [5+W -> d]*3 (add lane feature)
    _mlp(s, "tl_encoder.input_encoder.mlp", [cfg["tl_state_dim"] + W] + [d] * 3, (0, 2, 4))
    for i in range(3):
        _mlp(s, f"tl_encoder.temp_encoder.mlp_layers.{i}", [d, d // 2], (0,))
    for i in range(cfg["tl_encoder"]["n_layer_tf"]):
        _tf_layer(s, f"tl_encoder.tf_tl2tlmp.layers.{i}", d, d, True)
    _mlp(s, "tl_state_predictor.mlp", [d, d, d, cfg["tl_state_dim"]], (0, 2, 4))
    # agent encoder (agent_encoder.py:39-70): input MLP [6+3+W -> d/2]*3 (cat d/2 pose emb)
    _mlp(s, "ag_encoder.input_encoder.mlp", [cfg["ag_attr_dim"] + cfg["ag_motion_dim"] + W] + [d // 2] * 3, (0, 2, 4))
    for i in range(3):
        _mlp(s, f"ag_encoder.temp_encoder.mlp_layers.{i}", [d, d // 2], (0,))
    for i in range(cfg["ag_encoder"]["n_layer_tf"]):
        _tf_layer(s, f"ag_encoder.tf_ag2agmptl.layers.{i}", d, d, True)
    # heads (navigation.py:37-40, add_navi_latent.py:27-31, action_head.py:25-50)
    _mlp(s, "navi_encoder.mlp_mp", [d, d], (0,))
    _mlp(s, "navi_encoder.mlp_pe", [d, d], (0,))
    _mlp(s, "add_navi.mlp_in", [d, d, d, d], (0, 3, 6))
    _mlp(s, "add_navi.mlp", [2 * d, d, d, d], (0, 3, 6))
    _mlp(s, "add_latent.mlp_in", [cfg["latent_encoder"]["latent_dim"], d, d, d], (0, 3, 6))
    _mlp(s, "add_latent.mlp", [2 * d, d, d, d], (0, 3, 6))
    for t in range(3):
        _mlp(s, f"action_head.mlp_mean.{t}", [d, d, d, cfg["action_dim"]], (0, 2, 4))
        s[f"action_head.log_std.{t}"] = (cfg["action_dim"],)
    return s


def navi_predictor_shapes(cfg: dict) -> "OrderedDict[str, Tuple[int, ...]]":
    """`navi_predictor.*` (NaviPredictor in "dest" mode, navigation.py:100-160): the once-per-scene destination
    classifier that runs immediately before the rollout loop."""
    d, W = cfg["hidden_dim"], cfg["temp_window_size"]
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    for i in range(3):
        _mlp(s, f"navi_predictor.temp_encoder.mlp_layers.{i}", [d, d // 2], (0,))
    _mlp(s, "navi_predictor.input_encoder.mlp", [cfg["ag_attr_dim"] + cfg["ag_motion_dim"] + W] + [d // 2] * 3, (0, 2, 4))
    # MLP [2d + d_rpe, d, d, 1] with LayerNorm after the first two Linears (mlp.py:47-51: Linear, LN, ReLU)
    _mlp(s, "navi_predictor.mlp", [3 * d, d, d, 1], (0, 3, 6))
    for i in (1, 4):
        s[f"navi_predictor.mlp.fc_layers.{i}.weight"] = (d,)
        s[f"navi_predictor.mlp.fc_layers.{i}.bias"] = (d,)
    return s


def latent_post_shapes(cfg: dict) -> "OrderedDict[str, Tuple[int, ...]]":
    """`latent_encoder.{tl,ag}_encoder_post.*` + `latent_encoder.latent_dist_post.*` (LatentEncoder with
    share_post_prior_encoders False and a std_gaus prior, latent_encoder.py:27-54,124-190): the posterior's own
    traffic-light and agent encoders over the down-sampled ground truth (window (time_step_gt + 1) // rate + 1 = 19
    steps) and the diag-Gaussian head. Used by the training step only."""
    d = cfg["hidden_dim"]
    rate = cfg["latent_encoder"]["temporal_down_sample_rate"]
    n = cfg["time_step_gt"] + 1
    W = n // rate + 1 if rate > 1 else n
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    tp, ap = "latent_encoder.tl_encoder_post", "latent_encoder.ag_encoder_post"
    _mlp(s, f"{tp}.input_encoder.mlp", [cfg["tl_state_dim"] + W] + [d] * 3, (0, 2, 4))
    for i in range(3):
        _mlp(s, f"{tp}.temp_encoder.mlp_layers.{i}", [d, d // 2], (0,))
    for i in range(cfg["tl_encoder"]["n_layer_tf"]):
        _tf_layer(s, f"{tp}.tf_tl2tlmp.layers.{i}", d, d, True)
    _mlp(s, f"{ap}.input_encoder.mlp", [cfg["ag_attr_dim"] + cfg["ag_motion_dim"] + W] + [d // 2] * 3, (0, 2, 4))
    for i in range(3):
        _mlp(s, f"{ap}.temp_encoder.mlp_layers.{i}", [d, d // 2], (0,))
    for i in range(cfg["ag_encoder"]["n_layer_tf"]):
        _tf_layer(s, f"{ap}.tf_ag2agmptl.layers.{i}", d, d, True)
    _mlp(s, "latent_encoder.latent_dist_post.mlp_mean", [d, d, d, cfg["latent_encoder"]["latent_dim"]], (0, 2, 4))
    s["latent_encoder.latent_dist_post.log_std"] = (cfg["latent_encoder"]["latent_dim"],)
    return s


def init_params(cfg: dict, seed: int = 0, bias_scale: float = 1.0, with_navi_predictor: bool = False,
                with_latent_post: bool = False) -> Dict[str, torch.Tensor]:
    """Seeded random init (CPU generator, fp32): Linear-style U(+-1/sqrt(fan_in)) for matrices and
    biases (so attention biases are exercised, unlike the reference's zero init, attention_rpe.py:50-56),
    LayerNorm gamma 1+0.1N / beta 0.1N, log_std -2 (action_head.py:48-50). `navi_predictor.*` tensors are drawn
    from a separate stream so that the hot-path weights of a seed do not depend on the flag."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    shapes = param_shapes(cfg)
    if with_navi_predictor:
        P.update(rand_like_state_dict({k.replace("fc_layers.1.", "fc_layers.1.norm.").replace("fc_layers.4.", "fc_layers.4.norm."): v
                                       for k, v in navi_predictor_shapes(cfg).items()}, seed + 7919))
        P = {k.replace(".norm.", "."): v for k, v in P.items()}
    if with_latent_post:
        lp = rand_like_state_dict(latent_post_shapes(cfg), seed + 104729)
        lp["latent_encoder.latent_dist_post.log_std"] = torch.full_like(lp["latent_encoder.latent_dist_post.log_std"],
                                                                        float(cfg["latent_encoder"]["latent_post"]["log_std"]))
        P.update(lp)
    for k, shp in shapes.items():
        if "log_std" in k:
            P[k] = torch.full(shp, -2.0)
        elif ".norm" in k:
            n = torch.randn(shp, generator=g) * 0.1
            P[k] = (1.0 + n) if k.endswith("weight") else n
        elif len(shp) == 2:
            b = 1.0 / shp[1] ** 0.5
            P[k] = (torch.rand(shp, generator=g) * 2 - 1) * b
        else:
            wk = k.replace("in_proj_bias", "in_proj_weight").replace("out_proj_bias", "out_proj_weight")
            wk = wk[:-4] + "weight" if wk.endswith("bias") else wk
            b = bias_scale / shapes[wk][1] ** 0.5
            P[k] = (torch.rand(shp, generator=g) * 2 - 1) * b
    return P


def rand_like_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int) -> Dict[str, torch.Tensor]:
    """Seeded tensors for an arbitrary {name: shape or tensor} dict (sorted-name order): matrices
    U(+-1/sqrt(fan_in)), LayerNorm-like vectors 1+0.1N (weight) / 0.1N (other vectors). Used so golden
    fixtures can store a seed instead of weights."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k].shape) if hasattr(shapes[k], "shape") else tuple(shapes[k])
        if len(shp) >= 2:
            out[k] = (torch.rand(shp, generator=g) * 2 - 1) / shp[-1] ** 0.5
        elif ".norm" in k and k.endswith("weight"):
            out[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            out[k] = 0.1 * torch.randn(shp, generator=g)
    return out
