"""GPU box diagnostic: per-tensor gradient agreement (norm ratio, cosine) between the native path and
the CPU oracle on a golden case; writes gpurun_out/grad_diag_<case>.txt."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import POLICY_CFG, gather_minibatch, load_golden, minibatch_env_inds, recipe_state_dict, synthetic_rollout  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402
import habitat_lab_b200 as hb  # noqa: E402
from test_gpu_policy import _make  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "small128"
G = load_golden(name)
pol, st, next_value, c = _make(hb, G)
pol.train()
st.buffers["value_preds"].copy_(G["value_preds_after"])
st.buffers["returns"].copy_(G["returns"])
torch.manual_seed(G["mb_env_inds_seed"])
batch = next(iter(st.data_generator(G["advantages"].cuda(), c["mb"])))
pol.loss_and_backward(batch, 0.2, 0.5, 0.01, True)
torch.cuda.synchronize()
# oracle
bufs, _ = synthetic_rollout(c["T"], c["N"], c["H"], c["W"], 4, 2 * c["layers"], 512, c["seed"])
bufs["value_preds"] = G["value_preds_after"].clone()
bufs["returns"] = G["returns"].clone()
inds = minibatch_env_inds(G["mb_env_inds_seed"], c["N"], c["mb"])[0]
b = gather_minibatch(bufs, G["advantages"], inds, c["T"])
sd = recipe_state_dict(G["shapes"], c["seed"])
sdr = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running_mean" not in k else v) for k, v in sd.items()}
value, lp, ent, hid, ns, _ = O.evaluate_actions(b["observations"], b["recurrent_hidden_states"], b["prev_actions"], b["masks"], b["actions"], sdr, POLICY_CFG, True)
res = O.ppo_loss(value, lp, ent, b, 0.2, 0.5, 0.01, True)
res["total_loss"].backward()
lines = []
for k, p in pol.named_parameters():
    g, r = p.grad.detach().cpu().flatten().double(), sdr[k].grad.flatten().double()
    cos = (g @ r / (g.norm() * r.norm() + 1e-30)).item()
    lines.append(f"{k:70s} n={p.numel():8d} |g|={g.norm().item():.5e} |ref|={r.norm().item():.5e} ratio={g.norm().item() / (r.norm().item() + 1e-30):.4f} cos={cos:.5f}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", f"grad_diag_{name}.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
