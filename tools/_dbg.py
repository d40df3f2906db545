import os, sys, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pyroved_b200 import ops
B, Cin, Cout, H, W = 64, 64, 64, 32, 32
x = torch.randn(B, Cin, H, W, device="cuda"); wt = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05; b = torch.zeros(Cout, device="cuda")
y = torch.empty(B, Cout, H, W, device="cuda"); ws = ops.conv_tc_workspace(wt)
for _ in range(3): ops.conv_tc_fwd(x, wt, b, "lrelu", y, ws)
torch.cuda.synchronize()
lib = C.CDLL("/root/repo/pyroved_b200/csrc/libpvb.so")
buf = np.zeros((2, 64), dtype=np.int64)
lib.pvb_conv_trace_read(buf.ctypes.data_as(C.POINTER(C.c_longlong)))
P, M = buf[0], buf[1]
t0 = P[0]
print("producer: start", 0, "chunks (begin, gathered, signalled):")
for c in range(9): print(c, (P[1+3*c:4+3*c] - t0).tolist())
print("acc ready", P[60]-t0, "epilogue end", P[61]-t0)
print("mma: (full seen, issued)", [((M[1+2*c]-t0), (M[2+2*c]-t0)) for c in range(9)])
