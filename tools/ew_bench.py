"""Effective HBM bandwidth of the elementwise / normalisation kernels at the bench shapes (4096 frames).
Usage: python tools/ew_bench.py            (GPU box only)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import habitat_lab_b200 as hb  # noqa: E402
from habitat_lab_b200 import ops  # noqa: E402

DEV = torch.device("cuda", 0)
BF = torch.bfloat16
B = int(os.environ.get("EW_B", "4096"))


def t_ms(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    hb.load()
    rows = []
    for name, C, H in (("stem", 32, 64), ("layer1", 32, 32), ("layer2", 64, 16), ("layer3", 128, 8), ("layer4", 256, 4)):
        hw, G = H * H, 16
        n = B * hw * C
        y = (torch.randn(B, hw, C, device=DEV) * 1.5).to(BF)
        g = torch.randn(B, hw, C, device=DEV).to(BF)
        act = torch.randn(B, hw, C, device=DEV).to(BF)
        yf = y.float().view(B, hw, G, C // G)
        stats = torch.stack([yf.sum((1, 3)), (yf * yf).sum((1, 3))], dim=-1).double().contiguous()
        del yf
        gamma, beta = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.1
        dga, dbe = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        dy, gz, out = torch.empty_like(y), torch.empty_like(y), torch.empty_like(y)
        sums = torch.zeros(B, G, 2, device=DEV)
        for mode, nb in ((1, 6), (2, 10), (0, 6)):
            ms = t_ms(lambda: ops.gn_bwd(g, act if mode == 2 else None, y, stats, gamma, beta, dga, dbe, dy,
                                         gz if mode == 2 else None, B, hw, C, G, mode))
            rows.append((f"gn_bwd mode{mode} {name}", ms, n * nb))
        ms = t_ms(lambda: (ops.gn_bwd_reduce(g, None, y, stats, gamma, beta, sums, dga, dbe, B, hw, C, G, 1),
                           ops.gn_bwd_apply(g, None, y, stats, gamma, beta, sums, dy, None, B, hw, C, G, 1)))
        rows.append((f"gn_bwd 2-pass mode1 {name}", ms, n * 10))
        ms = t_ms(lambda: ops.gn_apply(y, stats, gamma, beta, out, B, hw, C, G, relu=True))
        rows.append((f"gn_apply {name}", ms, n * 4))
        ms = t_ms(lambda: ops.gn_residual_relu(y, stats, gamma, beta, act, out, B, hw, C, G))
        rows.append((f"gn_residual_relu {name}", ms, n * 6))
        if name == "stem":
            pooled = torch.empty(B, hw // 4, C, device=DEV, dtype=BF)
            argmax = torch.empty(B, hw // 4, C, device=DEV, dtype=torch.uint8)
            ms = t_ms(lambda: ops.gn_relu_maxpool(y, stats, gamma, beta, pooled, argmax, B, H, H, C, G))
            rows.append(("gn_relu_maxpool stem", ms, n * 2 + n // 4 * 3))
            dp = torch.randn(B, hw // 4, C, device=DEV).to(BF)
            ms = t_ms(lambda: ops.maxpool_bwd(dp, argmax, dy, B, H, H, C))
            rows.append(("maxpool_bwd stem", ms, n * 2 + n // 4 * 3))
            if hasattr(ops, "gn_relu_maxpool_bwd"):
                ms = t_ms(lambda: ops.gn_relu_maxpool_bwd(dp, argmax, y, stats, gamma, beta, dga, dbe, dy, B, H, H, C, G))
                rows.append(("gn_relu_maxpool_bwd stem (fused)", ms, n * 4 + n // 4 * 3))
        del y, g, act, dy, gz, out
    for name, ms, nbytes in rows:
        print(f"{name:38s} {ms*1e3:9.1f} us  {nbytes/ms/1e6:8.1f} GB/s  ({nbytes/1e6:.0f} MB algorithmic)")


if __name__ == "__main__":
    main()
