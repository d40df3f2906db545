"""Debug: clock64 trace of CTA 300 of the tap-reuse pixel GEMM (build with -DPVB_TC_TRACE)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pyroved_b200 import ops, _lib  # noqa: E402

cin, cout, k, hw, B = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (64, 64, 3, 32, 512))]
x = torch.randn(B, cin, hw, hw, device="cuda")
W = torch.randn(cout, cin, k, k, device="cuda") * 0.05
b = torch.zeros(cout, device="cuda")
y = torch.empty(B, cout, hw, hw, device="cuda")
ws = ops.conv_tc_workspace(W)
for _ in range(3):
    ops.conv_tc_fwd(x, W, b, "lrelu", y, ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.conv_tc_fwd(x, W, b, "lrelu", y, ws)
e1.record()
torch.cuda.synchronize()
print("fwd {}->{} k{} {}x{}: {:.1f} us per call (prep + GEMM)".format(cin, cout, k, hw, hw, e0.elapsed_time(e1) / 5 * 1e3))
h = C.CDLL(_lib.LIB_PATH)
buf = (C.c_longlong * 128)()
h.pvb_wgrad_trace_read(buf)
v = list(buf)
prod, mma = v[:64], v[64:]
if any(t > 0 for t in prod + mma):      # the per-tile kernel (conv_tc_pix2_kernel) ran
    t0 = min(t for t in prod + mma if t > 0)
    print("producer warp 0: [0]=start [1]=gather done [2]=x arrived [3+i]=weights of tap i stored [40]=acc ready [41]=end")
    print({i: t - t0 for i, t in enumerate(prod) if t > 0})
    print("mma warp: [0]=start [1]=x ready [2+2i]=weights i ready [3+2i]=tap i issued")
    print({i: t - t0 for i, t in enumerate(mma) if t > 0})

if hasattr(h, "pvb_pix3_trace_read"):
    buf3 = (C.c_longlong * 192)()
    h.pvb_pix3_trace_read(buf3)
    v3 = list(buf3)
    if any(v3):
        t0 = min(t for t in v3 if t > 0)
        for name, arr in (("producer", v3[:64]), ("mma", v3[64:128]), ("epilogue", v3[128:])):
            print("pix3", name, "(per tile: start / after wait / done):", [t - t0 for t in arr[:24] if t > 0])
