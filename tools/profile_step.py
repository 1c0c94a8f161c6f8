"""One config-#2 minibatch pass (forward + backward + clip/Adam) for ncu / timing:
warm-up pass, then the profiled pass between cudaProfilerStart/Stop
(run ncu with --profile-from-start off).  Optional argv: frames-per-minibatch envs (default 32 -> 4096 frames)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import habitat_lab_b200 as hb  # noqa: E402
from habitat_lab_b200.synthetic import fill_rollout_, pointnav_spaces  # noqa: E402

n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
hb.load()
torch.manual_seed(100)
obs_space, act_space = pointnav_spaces(256, 256)
policy = hb.PointNavResNetPolicy(obs_space, act_space, hidden_size=512, num_recurrent_layers=2, rnn_type="LSTM",
                                 normalize_visual_inputs=True).to(dev)
ppo = hb.PPO(policy, clip_param=0.2, ppo_epoch=1, num_mini_batch=2, value_loss_coef=0.5, entropy_coef=0.01, lr=2.5e-4,
             eps=1e-5, max_grad_norm=0.2, use_clipped_value_loss=True, use_normalized_advantage=False)
policy.train()
st = hb.RolloutStorage(T, n_envs, obs_space, act_space, policy)
st.to(dev)
nv = fill_rollout_(st, seed=100)
st.compute_returns(nv, True, 0.99, 0.95)
adv = ppo.get_advantages(st)
gen = st.data_generator(adv, 2)
b0 = next(gen)
b1 = next(gen)
import collections  # noqa: E402

lm = collections.defaultdict(list)
ppo._update_from_batch(b0, 0, st, lm)  # warm-up (allocations, attribute setup)
torch.cuda.synchronize()
torch.cuda.profiler.start()
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ppo._update_from_batch(b1, 0, st, lm)
e1.record()
t_enq = (time.perf_counter() - t0) * 1e3
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"minibatch of {T * n_envs // 2} frames: {e0.elapsed_time(e1):.2f} ms device, {(time.perf_counter() - t0) * 1e3:.2f} ms wall, "
      f"{t_enq:.2f} ms host enqueue")
