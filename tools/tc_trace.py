"""Stage timeline of the fused decoder kernel (debug build: PVB_EXTRA_FLAGS=-DPVB_TC_TRACE).
Prints, for CTA 0, SM-clock deltas between trace points of epilogue warp 0 and the MMA warp."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import pyroved_b200 as pv  # noqa: E402
from pyroved_b200 import _lib  # noqa: E402

B = 512
m = pv.models.iVAE((28, 28), 2, ['r', 't'], seed=1, device="cuda:0")
tr = pv.trainers.SVItrainer(m, device="cuda:0")
tr.svi.use_graphs = False
x = (torch.rand(B, 28, 28) < 0.3).float().cuda()
for _ in range(3):
    tr.svi.step(x)
torch.cuda.synchronize()
buf = np.zeros((2, 64, 32), dtype=np.int64)
lib = C.CDLL(_lib.LIB_PATH)
rc = lib.pvb_tc_trace_read(buf.ctypes.data_as(C.POINTER(C.c_longlong)))
assert rc == 0, rc
E, M = buf[0], buf[1]
t0 = E[0, 0]
np.set_printoptions(linewidth=250)
names = ["top", "S0c0", "S0c1", "S0c2", "S0c3", "S0smem", "staged", "acc1", "S2end", "acc2", "S4A",
         "bar", "S4B", "S4end", "acc3", "S6end", "acc4", "S8end"]
print("epilogue warp 0 (cycles from tile start):", names)
for it in range(2, 7):
    print(it, (E[it, :18] - E[it, 0]).tolist(), "tile period", int(E[it + 1, 0] - E[it, 0]))
print("MMA warp: top, ready x4 for GEMM1..4, dUv operands ready (relative to the epilogue tile start)")
for it in range(2, 7):
    print(it, (M[it, :18] - E[it, 0]).tolist())

wb = np.zeros((16, 16), dtype=np.int64)
rc = lib.pvb_tc_wtrace_read(wb.ctypes.data_as(C.POINTER(C.c_longlong)))
assert rc == 0, rc
base = wb[:, 0].min()
print("per-warp S4B of tile 3: [bar-arrive, bar, dl, c0-computed, c0..c3 published, stores done, smem signalled]")
for w in range(16):
    print(w, (wb[w, :10] - base).tolist())

kb = np.zeros(8, dtype=np.int64)
if hasattr(lib, "pvb_tc_ktrace_read") and lib.pvb_tc_ktrace_read(kb.ctypes.data_as(C.POINTER(C.c_longlong))) == 0:
    print("kernel milestones of CTA 0 (cycles from entry): set-up done, first tile, tile loop done, partials "
          "written, exit | last dUv written, dW tiles stored:", (kb[1:6] - kb[0]).tolist(), (kb[6:8] - kb[0]).tolist())
