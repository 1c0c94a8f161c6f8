"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's CPU learner iteration restated on top of
oracle/torch_oracle.py: compute_returns -> get_advantages -> ppo_epoch x num_mini_batch x
(evaluate_actions, loss, autograd backward, clip_grad_norm_, Adam)  exactly as
habitat-baselines/habitat_baselines/rl/ppo/ppo.py:301-332 + rollout_storage.py:207-257 run it on
`device="cpu"`.  Used by bench.py (`cpu_baseline`, `--impl reference`) and by tests; never by the
product package."""
from __future__ import annotations

import time
from typing import Dict

import torch

from . import torch_oracle as O


class CpuLearner:
    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: dict, lr=2.5e-4, eps=1e-5, max_grad_norm=0.2,
                 clip_param=0.2, value_loss_coef=0.5, entropy_coef=0.01, ppo_epoch=2, num_mini_batch=2,
                 use_clipped_value_loss=True, use_normalized_advantage=False):
        self.sd = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running_mean" not in k
                       else v.clone()) for k, v in state_dict.items()}
        self.params = [v for v in self.sd.values() if v.requires_grad]
        self.opt = torch.optim.Adam(self.params, lr=lr, eps=eps)  # what ppo.py:112-137 constructs
        self.cfg = cfg
        self.h = dict(max_grad_norm=max_grad_norm, clip=clip_param, c_v=value_loss_coef, c_e=entropy_coef,
                      epochs=ppo_epoch, mb=num_mini_batch, clip_v=use_clipped_value_loss, norm=use_normalized_advantage)

    def update(self, bufs, next_value, T, gamma=0.99, tau=0.95) -> Dict[str, float]:
        h = self.h
        v = bufs["value_preds"]
        returns = O.compute_returns(bufs["rewards"], v, bufs["masks"], next_value, T, True, gamma, tau)
        bufs = dict(bufs, returns=returns)
        adv = O.get_advantages(returns, v, h["norm"])
        N = v.shape[1]
        out = []
        for _ in range(h["epochs"]):
            for inds in torch.randperm(N).chunk(h["mb"]):
                sel = lambda t: t[0:T, inds].flatten(0, 1)  # noqa: E731
                batch = {k: sel(bufs[k]) for k in ("value_preds", "returns", "action_log_probs", "actions",
                                                    "prev_actions", "masks")}
                batch["advantages"] = sel(adv)
                obs = {k: sel(t) for k, t in bufs["observations"].items()}
                hidden = bufs["recurrent_hidden_states"][0, inds]
                self.opt.zero_grad(set_to_none=True)
                value, lp, ent, _, new_stats, _ = O.evaluate_actions(obs, hidden, batch["prev_actions"],
                                                                     batch["masks"], batch["actions"], self.sd,
                                                                     self.cfg, training=True)
                if new_stats is not None:
                    p = "net.visual_encoder.running_mean_and_var."
                    self.sd[p + "_mean"], self.sd[p + "_var"], self.sd[p + "_count"] = (t.detach() for t in new_stats)
                res = O.ppo_loss(value, lp, ent, batch, h["clip"], h["c_v"], h["c_e"], h["clip_v"])
                res["total_loss"].backward()
                gn = torch.nn.utils.clip_grad_norm_(self.params, h["max_grad_norm"])
                self.opt.step()
                out.append({**{k: float(x) for k, x in res.items()}, "grad_norm": float(gn)})
        return {k: sum(o[k] for o in out) / len(out) for k in out[0]}


def time_cpu_learner(make_inputs, state_dict, cfg, T, N, updates: int, threads: int, **kw):
    """frames/s of the CPU learner on a bounded sample: `updates` iterations after one warm-up."""
    torch.set_num_threads(threads)
    learner = CpuLearner(state_dict, cfg, **kw)
    bufs, next_value = make_inputs()
    learner.update({k: (dict(v) if isinstance(v, dict) else v.clone()) for k, v in bufs.items()}, next_value, T)
    t0 = time.perf_counter()
    for _ in range(updates):
        learner.update({k: (dict(v) if isinstance(v, dict) else v.clone()) for k, v in bufs.items()}, next_value, T)
    dt = time.perf_counter() - t0
    return updates * T * N / dt, dt
