"""Shared test helpers: rebuild recipe inputs and the reference's minibatch layout."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from recipe import recipe_state_dict, synthetic_rollout  # noqa: E402,F401

GOLDEN = os.path.join(HERE, "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def minibatch_env_inds(seed, n_envs, num_mini_batch):
    """data_generator's env split: torch.randperm(N).chunk(num_mini_batch) on the CPU RNG
    (habitat-baselines/habitat_baselines/common/rollout_storage.py:236)."""
    torch.manual_seed(seed)
    return list(torch.randperm(n_envs).chunk(num_mini_batch))


def gather_minibatch(bufs, advantages, inds, T):
    """buffers[(slice(0,T), inds)] flattened (t, n) -> [T*n, ...]; hidden state from row 0
    (rollout_storage.py:237-246)."""
    def sel(v):
        return v[0:T, inds].flatten(0, 1)

    batch = {k: sel(v) for k, v in bufs.items() if k not in ("observations", "recurrent_hidden_states")}
    batch["observations"] = {k: sel(v) for k, v in bufs["observations"].items()}
    batch["recurrent_hidden_states"] = bufs["recurrent_hidden_states"][0, inds]
    batch["advantages"] = sel(advantages)
    return batch


POLICY_CFG = dict(visual_keys=["rgb", "depth"], ngroups=16, rnn_type="LSTM", num_layers=2)
