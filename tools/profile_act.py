"""One eager act() step at rollout batch 64 (config #2 shapes) between cudaProfilerStart/Stop, for an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off ...`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import habitat_lab_b200 as hb  # noqa: E402
from habitat_lab_b200.synthetic import fill_rollout_, pointnav_spaces  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
hb.load()
torch.manual_seed(100)
obs_space, act_space = pointnav_spaces(256, 256)
policy = hb.PointNavResNetPolicy(obs_space, act_space, hidden_size=512, num_recurrent_layers=2, rnn_type="LSTM",
                                 normalize_visual_inputs=True).to(dev)
policy.eval()
st = hb.RolloutStorage(4, N, obs_space, act_space, policy)
st.to(dev)
fill_rollout_(st, seed=100)
b = st.buffers
step = lambda t: ({k: v[t] for k, v in b["observations"].items()}, b["recurrent_hidden_states"][t], b["prev_actions"][t],  # noqa: E731
                  b["masks"][t])
for t in (0, 1):
    policy.act(*step(t))
torch.cuda.synchronize()
torch.cuda.profiler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
policy.act(*step(2))
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"act() at batch {N}: {e0.elapsed_time(e1):.3f} ms device")
# CUDA-graph replay of the same step (what the trainer's rollout loop runs)
ga = hb.GraphedActor(policy, *step(0))
for _ in range(5):
    ga(*step(1))
torch.cuda.synchronize()
e0.record()
for _ in range(50):
    ga.graph.replay()
e1.record()
torch.cuda.synchronize()
print(f"graph replay at batch {N}: {e0.elapsed_time(e1) / 50:.3f} ms device per step")
