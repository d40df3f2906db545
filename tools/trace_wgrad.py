"""Debug: clock64 trace of the wgrad kernel (build with PVB_EXTRA_FLAGS=-DPVB_TC_TRACE)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pyroved_b200 import ops, _lib  # noqa: E402

cin, cout, k, hw, B = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (64, 64, 3, 32, 512))]
x = torch.randn(B, cin, hw, hw, device="cuda")
dpre = torch.randn(B, cout, hw, hw, device="cuda") * 1e-3
W = torch.randn(cout, cin, k, k, device="cuda")
dW = torch.zeros_like(W)
db = torch.zeros(cout, device="cuda")
for _ in range(3):
    ops.conv_tc_bwd_weight(dpre, x, W, dW, db)
torch.cuda.synchronize()
h = C.CDLL(_lib.LIB_PATH)
buf = (C.c_longlong * 128)()
h.pvb_wgrad_trace_read(buf)
v = list(buf)
prod, mma = v[:64], v[64:]
t0 = min(t for t in prod + mma if t > 0)
print("producer warp 0 (events: finish start / after wait-empty / after stores):")
print([t - t0 for t in prod if t > 0])
print("mma warp (per tile: before wait-full / after wait / after issue):")
print([t - t0 for t in mma if t > 0])
