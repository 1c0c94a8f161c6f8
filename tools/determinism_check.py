"""run the same 4096-frame minibatch twice and report which parameter gradients differ run to run"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import habitat_lab_b200 as hb
from habitat_lab_b200.synthetic import fill_rollout_, pointnav_spaces
DEV = torch.device("cuda:0")
T, N = 128, int(os.environ.get("DC_N", "32"))
torch.manual_seed(3)
obs_space, act_space = pointnav_spaces(256, 256)
pol = hb.PointNavResNetPolicy(obs_space, act_space, hidden_size=512, num_recurrent_layers=2, rnn_type="LSTM",
                              normalize_visual_inputs=True).to(DEV)
pol.eval()
ppo = hb.PPO(pol, clip_param=0.2, ppo_epoch=1, num_mini_batch=1, value_loss_coef=0.5, entropy_coef=0.01, lr=2.5e-4,
             eps=1e-5, max_grad_norm=0.2, use_clipped_value_loss=True, use_normalized_advantage=False)
st = hb.RolloutStorage(T, N, obs_space, act_space, pol)
st.to(DEV)
nv = fill_rollout_(st, seed=9)
st.compute_returns(nv, True, 0.99, 0.95)
adv = ppo.get_advantages(st)
gs = []
for rep in range(3):
    torch.manual_seed(77)
    batch = next(iter(st.data_generator(adv, 1)))
    m = pol.loss_and_backward(batch, 0.2, 0.5, 0.01, True)
    torch.cuda.synchronize()
    gs.append(pol._flat["grads"].double().clone())
    print("rep", rep, [round(float(x), 6) for x in m[:3]])
for a, b in ((0, 1), (1, 2)):
    rel = (gs[a] - gs[b]).norm().item() / gs[b].norm().item()
    worst = []
    for (name, p_), off in zip(pol.named_parameters(), pol._flat["offsets"]):
        x, y = gs[a][off: off + p_.numel()], gs[b][off: off + p_.numel()]
        worst.append(((x - y).norm().item() / (y.norm().item() + 1e-30), name))
    worst.sort(reverse=True)
    print(f"runs {a},{b}: rel diff {rel:.3e}; worst:", [(f"{w:.2e}", n) for w, n in worst[:6]], "best:", [(f"{w:.2e}", n) for w, n in worst[-2:]])

# ---- forward intermediates run to run
eng = pol._engine_()
snaps = []
for rep in range(2):
    torch.manual_seed(77)
    batch = next(iter(st.data_generator(adv, 1)))
    m = pol.loss_and_backward(batch, 0.2, 0.5, 0.01, True)
    torch.cuda.synchronize()
    ws = eng._ws[(T * N, True)]
    snap = {k: v.detach().float().clone() for k, v in ws.items() if k.startswith(("y", "st", "a", "o", "x0", "x1", "feat"))}
    snap["values"] = pol._last["values"].float().clone()
    snaps.append(snap)
for k in sorted(snaps[0], key=lambda s: (s.rstrip("0123456789"), int("0" + "".join(c for c in s if c.isdigit())))):
    a, b = snaps[0][k], snaps[1][k]
    d = (a - b).abs()
    nz = int((d > 0).sum())
    print(f"{k:8s} numel {a.numel():11d}  differing {nz:9d}  max|d| {d.max().item():.3e}  rel-norm {(d.norm() / (b.norm() + 1e-30)).item():.3e}")
