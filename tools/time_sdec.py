"""Time the fused decoder kernel alone (CUDA events) at the cfg2 and cfg3 shapes for the library in
use (PVB_LIB selects an experimental build) and, in the same process, the one-tile kernel
(default; PVB_SDEC_V2=1 selects the interleaved one).  With a -DPVB_TC2_TRACE build also prints the stage timeline of CTA 0."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from pyroved_b200 import _lib, ops  # noqa: E402


def run(I, B, H, W, iters=20):
    N = H * W
    g = torch.Generator().manual_seed(0)
    r = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).cuda()   # noqa: E731
    Uv = r(I, 3, 128, sc=0.8)
    W1, b1, W2, b2 = r(128, 128, sc=0.09), r(128, sc=0.05), r(128, 128, sc=0.09), r(128, sc=0.05)
    wo, bo = r(1, 128, sc=0.09), r(1, sc=0.05)
    x = (torch.rand(B, N, generator=g) < 0.3).float().cuda()
    w = torch.rand(I, generator=g).cuda() if I != B else None
    sz = ops.sdec_tc_sizes(I, N)
    rowll, loc = torch.empty(I * N, device="cuda"), torch.empty(I * N, device="cuda")
    gp = torch.zeros(max(sz.gUv_part_floats, 4), device="cuda")
    wp = torch.zeros(max(sz.wgrad_part_floats, 4), device="cuda")

    packed = None
    if os.environ.get("PVB_TIME_PACKED", "1") == "1":
        packed = ops.sdec_tc_pack_weights(W1, W2, ops.sdec_tc_packed_weights("cuda"))

    def once():
        ops.sdec_tc_step(Uv, x, w, W1, b1, W2, b2, wo, bo, rowll, loc, gp, wp, I, B, H, W, 2,
                         "bernoulli", True, 0.5, True, packed_w=packed)
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        once()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    tiles = (I * N + 127) // 128
    return us, tiles


name = os.path.basename(_lib.LIB_PATH)
out = [name]
for tag, (I, B) in (("cfg2", (512, 512)), ("cfg3", (10240, 1024))):
    for v1 in ("0", "1"):
        os.environ["PVB_SDEC_V2"] = "0" if v1 == "1" else "1"
        us, tiles = run(I, B, 28, 28, iters=20 if tag == "cfg2" else 5)
        per = us * 1.965e3 / (tiles / 148.0)
        out.append("{} {}: {:.1f} us ({:.0f} cyc/tile @1965MHz, {:.0f} TFLOP/s)".format(
            tag, "one-tile" if v1 == "1" else "interleaved", us, per,
            198912.0 * I * 784 / us / 1e6))
print(" | ".join(out))
lib = C.CDLL(_lib.LIB_PATH)
if hasattr(lib, "pvb_tc_trace_read"):
    # one-tile kernel (csrc/pvb_sdec_tc.cu built with -DPVB_TC_TRACE)
    os.environ["PVB_SDEC_V2"] = "0"
    run(512, 512, 28, 28, iters=1)
    buf = np.zeros((2, 64, 32), dtype=np.int64)
    assert lib.pvb_tc_trace_read(buf.ctypes.data_as(C.POINTER(C.c_longlong))) == 0
    E, M = buf[0], buf[1]
    np.set_printoptions(linewidth=250)
    names = ["top", "S0c0", "S0c1", "S0c2", "S0c3", "S0smem", "staged", "acc1", "S2end", "acc2", "S4A",
             "bar", "S4B", "S4end", "acc3", "S6end", "acc4", "S8end"]
    print("one-tile kernel, epilogue warp 0 of CTA 0 (cycles from tile start):", names)
    for it in range(3, 9):
        print(it, (E[it, :18] - E[it, 0]).tolist(), "tile period", int(E[it + 1, 0] - E[it, 0]))
    print("MMA warp: top, ready x4 for GEMM1..4, dUv operands ready (same origin)")
    for it in range(3, 9):
        print(it, (M[it, :18] - E[it, 0]).tolist())
if hasattr(lib, "pvb_tc2_trace_read"):
    os.environ["PVB_SDEC_V2"] = "1"
    run(512, 512, 28, 28, iters=1)
    buf = np.zeros((2, 64, 16), dtype=np.int64)
    assert lib.pvb_tc2_trace_read(buf.ctypes.data_as(C.POINTER(C.c_longlong))) == 0
    E, M = buf[0], buf[1]
    np.set_printoptions(linewidth=250)
    en = ["top", "S0go", "S0pub", "S6acc", "S6dw2", "S6pub", "S2acc", "S2pub", "S8acc", "S8dw1", "S8pub",
          "S4acc", "S4bar<", "S4bar>", "S4pub"]
    print("epilogue warp 0 of CTA 0, cycles from the iteration top:", en)
    for it in range(3, 9):
        print(it, (E[it, :15] - E[it, 0]).tolist(), "period", int(E[it + 1, 0] - E[it, 0]))
    mn = ["OP_S0", "G1go", "OP_S6", "G4go", "OP_S2", "G2go", "OP_S8(dUv go)", "OP_S4(G3 go)"]
    print("MMA warp, same origin:", mn)
    for it in range(3, 9):
        print(it, (M[it, :8] - E[it, 0]).tolist())
