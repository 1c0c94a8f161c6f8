"""CPU experiment (no GPU needed): how much per-tensor gradient error does the hb200 precision scheme cause BY ITSELF?

Runs the oracle's config-#2 minibatch twice on the same recipe inputs -- once in plain fp32 and once with the
storage roundings of the CUDA path emulated (conv weights -> bf16; every stored activation y / a / o / pooled x1 /
input x0 -> bf16; every stored gradient at the same points -> bf16) -- and prints the per-tensor cosine / norm ratio.
Switches isolate the contribution of each rounding site:

    python tools/precision_emulation.py full256 [w] [act] [grad]
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import POLICY_CFG, gather_minibatch, load_golden, minibatch_env_inds, recipe_state_dict, synthetic_rollout  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402

MODE = {"w": False, "act": False, "grad": False, "tf32": False, "fp16": False}


def _round(x, is_grad=False):
    if MODE["fp16"] and not is_grad:   # fp16 forward operands (11-bit significand = TF32's), gradients stay bf16
        return x.half().float()
    if MODE["tf32"]:   # 10-bit mantissa, round to nearest even (what cuDNN/cuBLAS TF32 does to fp32 operands)
        i = x.contiguous().view(torch.int32)
        i = (i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF
        return i.view(torch.float32)
    return x.bfloat16().float()


class _Q(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return _round(x) if fwd else x

    @staticmethod
    def backward(ctx, g):
        return (_round(g, True) if ctx.bwd else g), None, None


def qa(x):   # a stored activation (and the gradient stored at the same point)
    return _Q.apply(x, MODE["act"], MODE["grad"])


def qw(w):   # bf16 weight image; the weight gradient itself is fp32
    return _Q.apply(w, MODE["w"], False)


def resnet18_q(x, sd, prefix, ngroups):
    p = prefix
    x = qa(x)
    x = qa(F.conv2d(x, qw(sd[p + "conv1.0.weight"]), stride=2, padding=3))
    x = F.relu(O._gn(x, sd, p + "conv1.1", ngroups))
    x = qa(F.max_pool2d(x, kernel_size=3, stride=2, padding=1))
    for layer in (1, 2, 3, 4):
        for blk in (0, 1):
            q = f"{p}layer{layer}.{blk}."
            stride = 2 if (layer > 1 and blk == 0) else 1
            out = qa(F.conv2d(x, qw(sd[q + "convs.0.weight"]), stride=stride, padding=1))
            out = qa(F.relu(O._gn(out, sd, q + "convs.1", ngroups)))
            out = qa(F.conv2d(out, qw(sd[q + "convs.3.weight"]), stride=1, padding=1))
            out = O._gn(out, sd, q + "convs.4", ngroups)
            if (q + "downsample.0.weight") in sd:
                res = qa(F.conv2d(x, qw(sd[q + "downsample.0.weight"]), stride=stride))
                res = O._gn(res, sd, q + "downsample.1", ngroups)
            else:
                res = x
            x = qa(F.relu(out + res))
    return x


def run(name, emulate):
    G = load_golden(name)
    c = G["case"]
    bufs, _ = synthetic_rollout(c["T"], c["N"], c["H"], c["W"], 4, 2 * c["layers"], 512, c["seed"],
                                p_done=c.get("p_done", 1 / 25))
    bufs["value_preds"], bufs["returns"] = G["value_preds_after"].clone(), G["returns"].clone()
    inds = minibatch_env_inds(G["mb_env_inds_seed"], c["N"], c["mb"])[0]
    ob = gather_minibatch(bufs, G["advantages"], inds, c["T"])
    sd0 = recipe_state_dict(G["shapes"], c["seed"])
    sdr = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running_mean" not in k else v)
           for k, v in sd0.items()}
    orig = O.resnet_forward
    if emulate:
        O.resnet_forward = resnet18_q
        conv_orig = F.conv2d
    try:
        value, lp, ent, _, _, feats = O.evaluate_actions(ob["observations"], ob["recurrent_hidden_states"], ob["prev_actions"],
                                                         ob["masks"], ob["actions"], sdr, POLICY_CFG, True)
        O.ppo_loss(value, lp, ent, ob, 0.2, 0.5, 0.01, True)["total_loss"].backward()
    finally:
        O.resnet_forward = orig
    return {k: v.grad for k, v in sdr.items() if getattr(v, "grad", None) is not None}, value.detach()


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "full256"
    for k in sys.argv[2:]:
        MODE[k] = True
    torch.set_num_threads(8)
    ref, vref = run(name, False)
    got, vgot = run(name, True)
    print("mode", MODE, "values max abs diff", (vref - vgot).abs().max().item())
    worst = []
    for k in ref:
        g, r = got[k].flatten().double(), ref[k].flatten().double()
        cos = (g @ r / (g.norm() * r.norm() + 1e-30)).item()
        worst.append((cos, (g.norm() / r.norm()).item(), k))
    for cos, ratio, k in sorted(worst)[:12]:
        print(f"{k:66s} cos={cos:.5f} ratio={ratio:.4f}")


if __name__ == "__main__":
    main()
