"""time the persistent LSTM kernels alone at the bench shape (T=128, n=32, H=512)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import habitat_lab_b200 as hb
from habitat_lab_b200 import ops
hb.load()
dev = torch.device("cuda:0")
T, n, H = 128, 32, 512
torch.manual_seed(0)
xproj = torch.randn(T * n, 4 * H, device=dev) * 0.5
w_hh = torch.randn(4 * H, H, device=dev) * 0.04
b_hh = torch.zeros(4 * H, device=dev)
md = (torch.rand(T * n, device=dev) > 0.04).to(torch.uint8)
h0, c0 = torch.randn(n, H, device=dev) * 0.1, torch.randn(n, H, device=dev) * 0.1
hs, cs, gates = (torch.empty(T, n, H, device=dev), torch.empty(T, n, H, device=dev), torch.empty(T, n, 4 * H, device=dev))
gout = torch.randn(T, n, H, device=dev)
dg = torch.empty(T, n, 4 * H, device=dev)
ws = torch.zeros(64, dtype=torch.uint8, device=dev)


def t_ms(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


f = t_ms(lambda: ops.lstm_seq_fwd(xproj, w_hh, b_hh, md, h0, c0, hs, cs, gates, T, n, H, ws))
b = t_ms(lambda: ops.lstm_seq_bwd(gout, gates, cs, c0, w_hh, md, dg, T, n, H, ws))
print(f"lstm_seq fwd {f*1e3:.0f} us ({f*1e3/T:.2f} us/step)  bwd {b*1e3:.0f} us ({b*1e3/T:.2f} us/step)  "
      f"env V1={os.environ.get('HB200_LSTM_V1')} SLEEP={os.environ.get('HB200_LSTM_SLEEP')}")
