"""Stand-alone launches of the gather convolution (conv_igemm) for ncu: `python tools/conv_probe.py B H C Co k [stride]`.
Prints the CUDA-event time per launch (warm, 20 launches) after one profiled launch between cudaProfilerStart/Stop."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import habitat_lab_b200 as hb  # noqa: E402
from habitat_lab_b200 import ops  # noqa: E402

B, H, C, Co, k = (int(v) for v in sys.argv[1:6])
stride = int(sys.argv[6]) if len(sys.argv) > 6 else 1
dev = torch.device("cuda:0")
hb.load()
torch.manual_seed(0)
pad = k // 2
s = ops.conv_shape(B, H, H, C, Co, k, k, stride, pad)
x = torch.randn(B, H, H, C, device=dev).half()
w = torch.randn(Co, C, k, k, device=dev) * 0.05
wp, wt = ops.pack_conv_weight(w, C, want_t=True)
y = torch.empty(B, s.ho, s.wo, Co, device=dev, dtype=torch.float16)
stats = torch.zeros(B, 16, 2, device=dev, dtype=torch.float64)
dy = torch.randn(B, s.ho, s.wo, Co, device=dev).bfloat16()
dx = torch.empty(B, H, H, C, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.conv_fwd(x, wp, y, s, stats, 16)
    ops.conv_dgrad(dy, wt, dx, s, addend=None)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.conv_fwd(x, wp, y, s, stats, 16)
ops.conv_dgrad(dy, wt, dx, s, addend=None)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
for name, fn in (("fwd", lambda: ops.conv_fwd(x, wp, y, s, stats, 16)),
                 ("dgrad", lambda: ops.conv_dgrad(dy, wt, dx, s, addend=None))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nchunks = (k * k * (C if name == "fwd" else Co) + 63) // 64
    flop = 2.0 * B * s.ho * s.wo * Co * C * k * k
    print(f"{name} B={B} {H}x{H} {C}->{Co} k{k} s{stride}: {ms * 1e3:.1f} us/launch, {flop / ms / 1e9:.1f} TFLOP/s, "
          f"{nchunks} K chunks")
