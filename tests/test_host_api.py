"""CPU: host-side API surface that mirrors the reference (no kernels): TensorDict semantics
(test/test_tensor_dict.py), registry overwrite, RolloutStorage bookkeeping, ddp_utils env parsing,
FusedAdam's torch.optim.Adam-compatible state_dict, policy state_dict layout."""
import os

import numpy as np
import pytest
import torch

import habitat_lab_b200 as hb
from habitat_lab_b200.synthetic import pointnav_spaces
from helpers import load_golden


def test_tensor_dict_indexing_and_set():
    td = hb.TensorDict.from_tree({"a": torch.arange(12.).view(3, 4), "n": {"b": torch.zeros(3, 2)}})
    sub = td[1:3]
    assert sub["a"].shape == (2, 4) and sub["n"]["b"].shape == (2, 2)
    td[0] = {"a": torch.ones(4), "n": {"b": torch.ones(2)}}
    assert td["a"][0].tolist() == [1, 1, 1, 1] and td["n"]["b"][0].tolist() == [1, 1]
    with pytest.raises(KeyError):
        td.set(0, {"a": torch.ones(4)}, strict=True)
    td.set(0, {"a": torch.zeros(4)}, strict=False)
    assert td["a"][0].sum() == 0
    doubled = td.map(lambda v: v * 2)
    assert torch.equal(doubled["a"], td["a"] * 2)
    assert isinstance(td.to_tree()["n"], dict)


def test_registry_names_match_the_reference_yaml_keys():
    import habitat_lab_b200.rl.ppo_trainer  # noqa: F401  (registers the trainer)

    reg = hb.baseline_registry
    assert reg.get_policy("PointNavResNetPolicy") is hb.PointNavResNetPolicy
    assert reg.get_policy("PointNavBaselinePolicy") is hb.PointNavBaselinePolicy
    assert reg.get_updater("PPO") is hb.PPO and reg.get_updater("DDPPO") is hb.DDPPO
    assert reg.get_storage("RolloutStorage") is hb.RolloutStorage
    assert reg.get_trainer("ddppo") is hb.PPOTrainer and reg.get_trainer("ppo") is hb.PPOTrainer


def test_policy_state_dict_layout_matches_reference():
    G = load_golden("full256")
    obs, act = pointnav_spaces(256, 256)
    pol = hb.PointNavResNetPolicy(obs, act, hidden_size=512, num_recurrent_layers=2, rnn_type="LSTM",
                                  normalize_visual_inputs=True)
    assert {k: tuple(v.shape) for k, v in pol.state_dict().items()} == {k: tuple(v) for k, v in G["shapes"].items()}
    assert sum(p.numel() for p in pol.parameters()) == 8481125
    assert pol.num_recurrent_layers == 4 and pol.recurrent_hidden_size == 512
    assert pol.net.visual_encoder.output_shape == (128, 4, 4)


def test_rollout_storage_bookkeeping_on_cpu():
    obs, act = pointnav_spaces(64, 64)
    pol = hb.PointNavResNetPolicy(obs, act, hidden_size=512, num_recurrent_layers=2, rnn_type="LSTM",
                                  normalize_visual_inputs=True)
    st = hb.RolloutStorage(4, 3, obs, act, pol)
    assert st.buffers["observations"]["rgb"].shape == (5, 3, 64, 64, 3) and st.buffers["observations"]["rgb"].dtype == torch.uint8
    assert st.buffers["recurrent_hidden_states"].shape == (5, 3, 4, 512) and st.buffers["actions"].dtype == torch.int64
    st.insert(next_observations={k: torch.ones_like(v[0]) for k, v in st.buffers["observations"].items()},
              actions=torch.full((3, 1), 2), rewards=torch.ones(3, 1), next_masks=torch.ones(3, 1, dtype=torch.bool))
    st.advance_rollout()
    assert st.current_rollout_step_idx == 1
    assert st.buffers["prev_actions"][1].flatten().tolist() == [2, 2, 2] and st.buffers["actions"][0].flatten().tolist() == [2, 2, 2]
    assert st.buffers["observations"]["depth"][1].min() == 1
    st.after_update()
    assert st.current_rollout_step_idx == 0 and st.buffers["observations"]["depth"][0].min() == 1
    with pytest.raises(hb.Hb200Error):
        st.compute_returns(torch.zeros(3, 1), True, 0.99, 0.95)  # CPU buffers: the hot path refuses, no fallback
    with pytest.raises(AssertionError):
        next(iter(hb.RolloutStorage(4, 1, obs, act, pol).data_generator(None, 2)))  # num_envs < num_mini_batch


def test_fused_adam_state_dict_is_adam_compatible():
    params = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5))]
    opt = hb.FusedAdam(params, lr=2.5e-4, eps=1e-5)
    ref = torch.optim.Adam([torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5))], lr=2.5e-4, eps=1e-5)
    for p in ref.param_groups[0]["params"]:
        p.grad = torch.ones_like(p)
    ref.step()
    opt.load_state_dict(ref.state_dict())          # Adam -> FusedAdam
    sd = opt.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and sd["param_groups"][0]["lr"] == 2.5e-4
    ref.load_state_dict(sd)                        # FusedAdam -> Adam
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda _: 0.5)
    assert opt.param_groups[0]["lr"] == pytest.approx(1.25e-4)
    del sched


def test_ddp_utils_env_contract(monkeypatch):
    from habitat_lab_b200.rl import ddp_utils

    for k in ("LOCAL_RANK", "RANK", "WORLD_SIZE", "SLURM_JOBID"):
        monkeypatch.delenv(k, raising=False)
    assert ddp_utils.get_distrib_size() == (0, 0, 1)
    monkeypatch.setenv("LOCAL_RANK", "3"); monkeypatch.setenv("RANK", "11"); monkeypatch.setenv("WORLD_SIZE", "16")
    assert ddp_utils.get_distrib_size() == (3, 11, 16)
    monkeypatch.delenv("LOCAL_RANK")
    monkeypatch.setenv("SLURM_JOBID", "7"); monkeypatch.setenv("SLURM_LOCALID", "1")
    monkeypatch.setenv("SLURM_PROCID", "5"); monkeypatch.setenv("SLURM_NTASKS", "8")
    assert ddp_utils.get_distrib_size() == (1, 5, 8)
    assert ddp_utils.rank0_only() is True
    assert isinstance(ddp_utils.find_free_port(), int)


def test_trainer_percent_done_and_config_defaults():
    from habitat_lab_b200.rl.ppo_trainer import PPOConfig, PPOTrainer, make_config

    d = PPOConfig()
    assert (d.clip_param, d.ppo_epoch, d.num_mini_batch, d.max_grad_norm, d.use_clipped_value_loss,
            d.use_normalized_advantage) == (0.2, 4, 2, 0.5, True, False)   # default_structured_configs.py:288-315
    tr = PPOTrainer(make_config(num_updates=10))
    tr.num_updates_done = 4
    assert tr.percent_done() == 0.4 and not tr.is_done()
    assert tr.should_end_early(100) is False  # not distributed


def test_rnn_state_encoder_factory_surface():
    from habitat_lab_b200.rl.models.rnn_state_encoder import build_rnn_state_encoder

    gru = build_rnn_state_encoder(514, 512, rnn_type="GRU", num_layers=1)
    assert gru.num_recurrent_layers == 1 and "rnn.weight_hh_l0" in gru.state_dict()
    lstm = build_rnn_state_encoder(576, 512, rnn_type="lstm", num_layers=2)
    assert lstm.num_recurrent_layers == 4
    assert float(lstm.rnn.bias_ih_l0.detach().abs().sum()) == 0.0  # zero biases, orthogonal weights (:288-293)
    w = lstm.rnn.weight_hh_l1.detach()          # [2048, 512]: orthonormal columns
    torch.testing.assert_close(w.t() @ w, torch.eye(512), rtol=1e-4, atol=1e-4)
    with pytest.raises(RuntimeError):
        build_rnn_state_encoder(8, 8, rnn_type="transformer")
    with pytest.raises(hb.Hb200Error):   # CPU tensors: the product path has no CPU fallback
        lstm(torch.zeros(4, 576), torch.zeros(2, 4, 512), torch.ones(4, 1, dtype=torch.bool))


def test_batch_obs_stacks_per_env_observations():
    from habitat_lab_b200.utils.common import ObservationBatchingCache, batch_obs

    rng = np.random.default_rng(0)
    obs = [{"rgb": rng.integers(0, 255, (8, 8, 3), dtype=np.uint8), "depth": rng.random((8, 8, 1)),
            "pointgoal_with_gps_compass": np.array([1.0 + i, -0.5], dtype=np.float32),
            "nested": {"gps": np.array([i, i], dtype=np.float32)}} for i in range(3)]
    cache = ObservationBatchingCache()
    b = batch_obs(obs, device="cpu", cache=cache)
    assert b["rgb"].shape == (3, 8, 8, 3) and b["rgb"].dtype == torch.uint8
    assert b["depth"].dtype == torch.float32            # float64 sensors are narrowed, uint8 kept (common.py:262-330)
    assert torch.equal(b["rgb"][1], torch.from_numpy(obs[1]["rgb"]))
    assert torch.equal(b["nested"]["gps"][2], torch.tensor([2.0, 2.0]))
    b2 = batch_obs(obs, device="cpu", cache=cache)      # staging buffers are reused, results are independent copies
    assert torch.equal(b2["pointgoal_with_gps_compass"], b["pointgoal_with_gps_compass"])
    assert len(cache._pool) == 4


@pytest.mark.parametrize("golden,backbone,cin", [("r50_objectnav", "resnet50", 5), ("rx50_imagenav", "resneXt50", 3)])
def test_backbone_holders_match_reference_checkpoint_layout(golden, backbone, cin):
    """Bottleneck / ResNeXt holders (no kernels yet: SURVEY 8f-2) expose exactly the reference's backbone state_dict --
    including its quirk that only the first block of a ResNeXt stage is grouped."""
    from habitat_lab_b200.rl.backbones import make_backbone

    G = load_golden(golden)
    pre = "net.visual_encoder.backbone."
    ref = {k[len(pre):]: tuple(v) for k, v in G["shapes"].items() if k.startswith(pre)}
    mine = {k: tuple(v.shape) for k, v in make_backbone(backbone, cin, 32, 16).state_dict().items()}
    assert mine == ref
    with pytest.raises(ValueError):
        make_backbone("resnet1000", 3, 32, 16)
    se = make_backbone("se_resneXt50", 3, 32, 16).state_dict()
    assert tuple(se["layer1.0.se.excite.0.weight"].shape) == (8, 128)   # planes*expansion = 128, r = 16
