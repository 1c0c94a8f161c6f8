"""Deterministic, reference-independent recipes for weights and synthetic rollouts, shared by
tests/golden/make_golden.py (which feeds them to the REAL reference classes in the build
container) and by the tests / bench (which feed them to the oracle and to the CUDA path).
Only torch's CPU generator is used, so every machine with the same torch build regenerates
bit-identical tensors and the fixtures only need to store the reference's OUTPUTS."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch


def recipe_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int) -> Dict[str, torch.Tensor]:
    out = {}
    for i, k in enumerate(sorted(shapes)):
        shape = tuple(shapes[k])
        g = torch.Generator().manual_seed(seed * 1000003 + i)
        if k.endswith("running_mean_and_var._count"):
            v = torch.tensor(5.0)
        elif k.endswith("running_mean_and_var._mean"):
            v = torch.rand(shape, generator=g) * 0.5 + 0.2
        elif k.endswith("running_mean_and_var._var"):
            v = torch.rand(shape, generator=g) * 0.1 + 0.03
        elif len(shape) == 4:  # conv OIHW
            fan_in = shape[1] * shape[2] * shape[3]
            v = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif "rnn.weight" in k:
            v = torch.randn(shape, generator=g) / math.sqrt(shape[1])
        elif "rnn.bias" in k:
            v = torch.randn(shape, generator=g) * 0.1
        elif k.startswith("action_distribution") and k.endswith("weight"):
            v = torch.randn(shape, generator=g) * 0.05
        elif k.endswith("embedding.weight"):
            v = torch.randn(shape, generator=g)
        elif len(shape) == 2:  # linear
            v = torch.randn(shape, generator=g) / math.sqrt(shape[1])
        elif k.endswith(".weight"):  # GroupNorm gamma
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:  # biases / GroupNorm beta
            v = 0.1 * torch.randn(shape, generator=g)
        out[k] = v.float()
    return out


def synthetic_rollout(T: int, N: int, H: int, W: int, n_actions: int, hidden_layers: int, hidden: int,
                      seed: int, p_done: float = 1 / 25, rgb: bool = True, depth: bool = True):
    """Buffers [T+1, N, ...] with the reference RolloutStorage's dtypes
    (habitat-baselines/habitat_baselines/common/rollout_storage.py:38-86)."""
    g = torch.Generator().manual_seed(seed)
    b = {}
    obs = {}
    if rgb:
        obs["rgb"] = torch.randint(0, 256, (T + 1, N, H, W, 3), generator=g, dtype=torch.uint8)
    if depth:
        obs["depth"] = torch.rand(T + 1, N, H, W, 1, generator=g)
    goal = torch.rand(T + 1, N, 2, generator=g)
    goal[..., 0] *= 10.0
    goal[..., 1] = goal[..., 1] * 2 * math.pi - math.pi
    obs["pointgoal_with_gps_compass"] = goal
    b["observations"] = obs
    b["recurrent_hidden_states"] = torch.randn(T + 1, N, hidden_layers, hidden, generator=g) * 0.5
    b["masks"] = torch.rand(T + 1, N, 1, generator=g) > p_done
    b["rewards"] = torch.randn(T + 1, N, 1, generator=g) * 0.1 + 2.5 * (~b["masks"]).float()
    b["value_preds"] = torch.randn(T + 1, N, 1, generator=g) * 0.5
    b["returns"] = torch.zeros(T + 1, N, 1)
    b["action_log_probs"] = -math.log(n_actions) + 0.2 * torch.randn(T + 1, N, 1, generator=g)
    b["actions"] = torch.randint(0, n_actions, (T + 1, N, 1), generator=g)
    b["prev_actions"] = torch.randint(0, n_actions, (T + 1, N, 1), generator=g)
    next_value = torch.randn(N, 1, generator=g) * 0.5
    return b, next_value


def objectnav_rollout(T: int, N: int, H: int, W: int, n_actions: int, hidden_layers: int, hidden: int, seed: int,
                      n_categories: int = 21, imagegoal: bool = False):
    """Config #3 / #4 sensor sets on top of synthetic_rollout: rgb (+ depth) + `semantic` (int32 class ids) +
    objectgoal / compass / gps (ObjectNav), or rgb + imagegoal + compass / gps (ImageNav).  Same generator discipline."""
    b, next_value = synthetic_rollout(T, N, H, W, n_actions, hidden_layers, hidden, seed, rgb=True, depth=not imagegoal)
    g = torch.Generator().manual_seed(seed + 7919)
    obs = b["observations"]
    del obs["pointgoal_with_gps_compass"]
    if imagegoal:
        obs["imagegoal"] = torch.randint(0, 256, (T + 1, N, H, W, 3), generator=g, dtype=torch.uint8)
    else:
        obs["semantic"] = torch.randint(0, 40, (T + 1, N, H, W, 1), generator=g, dtype=torch.int32)
        obs["objectgoal"] = torch.randint(0, n_categories, (T + 1, N, 1), generator=g)
    obs["compass"] = torch.rand(T + 1, N, 1, generator=g) * 2 * math.pi - math.pi
    obs["gps"] = torch.randn(T + 1, N, 2, generator=g) * 3.0
    return b, next_value
