"""One tiny learner iteration on cuda:0 (compute_returns + PPO.update on a 128x128 RGB-D rollout)."""
from __future__ import annotations

import torch


def run() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("hb200 smoke: no CUDA device (the hot path has no CPU fallback)")
    import habitat_lab_b200 as hb
    from habitat_lab_b200.synthetic import fill_rollout_, pointnav_spaces

    hb.load()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    obs_space, act_space = pointnav_spaces(128, 128)
    policy = hb.PointNavResNetPolicy(obs_space, act_space, hidden_size=512, num_recurrent_layers=2, rnn_type="LSTM",
                                     resnet_baseplanes=32, backbone="resnet18", normalize_visual_inputs=True).to(dev)
    ppo = hb.PPO(policy, clip_param=0.2, ppo_epoch=1, num_mini_batch=2, value_loss_coef=0.5, entropy_coef=0.01,
                 lr=2.5e-4, eps=1e-5, max_grad_norm=0.2, use_clipped_value_loss=True, use_normalized_advantage=False)
    st = hb.RolloutStorage(8, 4, obs_space, act_space, policy)
    st.to(dev)
    next_value = fill_rollout_(st, seed=1, p_done=0.05)
    st.compute_returns(next_value, True, 0.99, 0.95)
    metrics = ppo.update(st)
    torch.cuda.synchronize()
    for k in ("value_loss", "action_loss", "dist_entropy", "grad_norm"):
        v = metrics[k]
        if not (v == v) or abs(v) > 1e6:
            raise RuntimeError(f"hb200 smoke: metric {k} = {v}")
    print("hb200 smoke ok:", {k: round(v, 5) for k, v in metrics.items()}, "launches", hb.load().hb200_launch_count())
