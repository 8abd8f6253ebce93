"""Default hyper-parameters of the hot path, as plain nested dicts.

Values mirror the reference's `configs/model/sim_agent.yaml:26-167` (model, dynamics,
teacher_forcing_joint_future_pred) with the OmegaConf interpolations resolved by hand, plus the
data-derived kwargs that `SceneCentricPreProcessing.model_kwargs` supplies
(`src/data_modules/scene_centric.py:28-37`). No Hydra / OmegaConf is needed at run time.
"""
import copy


def default_model_cfg(hidden_dim: int = 128) -> dict:
    tf_cfg = dict(d_model=hidden_dim, n_head=4, k_feedforward=4, dropout_p=0.1, bias=True,
                  activation="relu", out_layernorm=False, apply_q_rpe=False)
    pose_rpe = dict(mode="pe_xy_yaw", theta_xy=1e3, theta_cs=1e1)
    inp = lambda mode: dict(mode=mode, n_layer=3, mlp_dropout_p=0, mlp_use_layernorm=False)  # noqa: E731
    latent = dict(dist_type="diag_gaus", n_cat=8, log_std=0.0, mlp_use_layernorm=False, n_layer=3, branch_type=False)
    cfg = dict(
        hidden_dim=hidden_dim,
        pairwise_relative=True,
        temp_window_size=11,
        n_tgt_knn=32,
        dist_limit=500,
        tf_cfg=tf_cfg,
        pose_rpe=pose_rpe,
        mp_encoder=dict(
            n_layer_tf=8,
            pose_emb=dict(mode="mpa_pl", theta_xy=1e3, theta_cs=1e1),
            input_encoder=inp("cat"),
            pl_encoder=dict(pooling_mode="max_valid", n_layer=3, mlp_dropout_p=0.1, mlp_use_layernorm=False,
                            use_pointnet=True),
        ),
        tl_encoder=dict(
            temp_stack_input=False, tl_lane_detach_mp_feature=True, n_layer_tf=4,
            k_tgt_knn_tl2tl=0.75, k_tgt_knn_tl2mp=0.75, k_dist_limit=0.5,
            pose_emb=dict(mode="pe_xy_yaw", theta_xy=1e3, theta_cs=1e1),
            input_encoder=inp("add"),
        ),
        tl_state_predictor=dict(detach_tl_feature=True, n_layer=3, rnn_dropout_p=0.1),
        ag_encoder=dict(
            n_layer_tf=4, k_tgt_knn_ag2mp=2.0, k_tgt_knn_ag2tl=0.8, k_tgt_knn_ag2ag=0.8, k_dist_limit=1.0,
            rnn_latent_temp_pool_mode="max_valid",
            pose_emb=dict(mode="pe_xy_yaw", theta_xy=1e3, theta_cs=1e1),
            input_encoder=inp("cat"),
        ),
        latent_encoder=dict(
            latent_dim=16, temporal_down_sample_rate=5, share_post_prior_encoders=False,
            latent_post=dict(latent), latent_prior=dict(latent, dist_type="std_gaus"),
        ),
        navi_encoder=dict(dest_detach_mp_feature=True),
        navi_predictor=dict(detach_input=True, rnn_res_add=True, n_layer_tf=3, n_layer_mlp=3, mlp_use_layernorm=True,
                            k_tgt_knn=1.0, k_dist_limit=1000, goal_log_std=2.0),
        add_navi_latent=dict(mode="cat", res_add=True, n_layer=3, mlp_use_layernorm=False, mlp_dropout_p=0.1),
        action_head=dict(log_std=-2, n_layer=3, branch_type=True, mlp_use_layernorm=False),
        # data-derived (scene_centric.py:28-37) and module-level (sim_agent.yaml:5-7, dynamics.py:16)
        mp_attr_dim=11, tl_state_dim=5, ag_attr_dim=6, ag_motion_dim=3, navi_mode="dest", navi_dim=None,
        n_mp_pl_node=20, tl_mode="lane", time_step_gt=90, action_dim=2,
    )
    return copy.deepcopy(cfg)


# sim_agent.yaml:156-167; tuple order (veh, ped, cyc) follows dynamics.py:23-27
DYNAMICS_CFG = dict(
    veh=dict(max_acc=5.0, max_yaw_rate=1.5),
    ped=dict(max_acc=7.0, max_yaw_rate=7.0),
    cyc=dict(max_acc=6.0, max_yaw_rate=3.0),
    dt=0.1,
)

# sim_agent.yaml:262-264 (teacher_forcing_joint_future_pred) and :5-11
ROLLOUT_CFG = dict(step_spawn_agent=10, step_warm_start=10, time_step_current=10, time_step_end=90,
                   n_joint_future_wosac=32)


def derived_sizes(cfg: dict) -> dict:
    """KNN sizes / distance limits the encoders derive from the config
    (map_encoder.py:33-34, traffic_light.py:62-64, agent_encoder.py:41-44)."""
    n, dl = cfg["n_tgt_knn"], cfg["dist_limit"]
    tl, ag = cfg["tl_encoder"], cfg["ag_encoder"]
    return dict(
        k_mp2mp=n, dl_mp=float(dl),
        k_tl2tl=int(n * tl["k_tgt_knn_tl2tl"]), k_tl2mp=int(n * tl["k_tgt_knn_tl2mp"]),
        dl_tl=float(dl * tl["k_dist_limit"]),
        k_ag2ag=int(n * ag["k_tgt_knn_ag2ag"]), k_ag2mp=int(n * ag["k_tgt_knn_ag2mp"]),
        k_ag2tl=int(n * ag["k_tgt_knn_ag2tl"]), dl_ag=float(dl * ag["k_dist_limit"]),
    )
