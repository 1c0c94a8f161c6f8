"""Synthetic PointNav RGB-D rollouts of the reference's shapes/dtypes (habitat-sim is stubbed out of
the learner loop): rgb u8 HxWx3 (habitat_simulator.py:116-127), depth f32 HxWx1 in [0,1] (:157-162),
pointgoal_with_gps_compass f32[2]."""
from __future__ import annotations

import math

import numpy as np
import torch

from .common import spaces


def pointnav_spaces(H=256, W=256, n_actions=4):
    obs = spaces.Dict({
        "rgb": spaces.Box(0, 255, (H, W, 3), np.uint8),
        "depth": spaces.Box(0.0, 1.0, (H, W, 1), np.float32),
        "pointgoal_with_gps_compass": spaces.Box(np.finfo(np.float32).min, np.finfo(np.float32).max, (2,), np.float32),
    })
    return obs, spaces.Discrete(n_actions)


def fill_rollout_(storage, seed: int, p_done: float = 1.0 / 250.0, device=None, chunk_steps: int = 8):
    """Fill every buffer of a RolloutStorage in place with synthetic data generated ON THE STORAGE'S
    DEVICE chunk by chunk (the full rgb+depth buffer of config #2 is 3.8 GB)."""
    b = storage.buffers
    dev = b["rewards"].device
    g = torch.Generator(device=dev).manual_seed(seed)
    T1, N = b["rewards"].shape[:2]
    obs = b["observations"]
    for t0 in range(0, T1, chunk_steps):
        t1 = min(T1, t0 + chunk_steps)
        if "rgb" in obs:
            obs["rgb"][t0:t1] = torch.randint(0, 256, obs["rgb"][t0:t1].shape, generator=g, device=dev,
                                              dtype=torch.uint8)
        if "depth" in obs:
            obs["depth"][t0:t1] = torch.rand(obs["depth"][t0:t1].shape, generator=g, device=dev)
    goal = torch.rand(T1, N, 2, generator=g, device=dev)
    goal[..., 0] *= 10.0
    goal[..., 1] = goal[..., 1] * 2 * math.pi - math.pi
    obs["pointgoal_with_gps_compass"].copy_(goal)
    A = 4
    b["masks"].copy_(torch.rand(T1, N, 1, generator=g, device=dev) > p_done)
    b["rewards"].copy_(torch.randn(T1, N, 1, generator=g, device=dev) * 0.1 + 2.5 * (~b["masks"]).float())
    b["value_preds"].copy_(torch.randn(T1, N, 1, generator=g, device=dev) * 0.5)
    b["action_log_probs"].copy_(-math.log(A) + 0.05 * torch.randn(T1, N, 1, generator=g, device=dev))
    b["actions"].copy_(torch.randint(0, A, (T1, N, 1), generator=g, device=dev))
    b["prev_actions"].copy_(torch.randint(0, A, (T1, N, 1), generator=g, device=dev))
    b["recurrent_hidden_states"].copy_(torch.randn(b["recurrent_hidden_states"].shape, generator=g, device=dev) * 0.1)
    storage.current_rollout_step_idxs = [storage.num_steps for _ in storage.current_rollout_step_idxs]
    return torch.randn(N, 1, generator=g, device=dev) * 0.5  # next_value
