"""Time the tensor-core convolution kernels (forward, backward data, weight gradient without / with the
scratch read-out) on VED's layer shapes at batch 512, and the 1-D decoder weight gradients."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from pyroved_b200 import ops
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
out = []
for cin, cout, hw in ((64, 64, 32), (32, 32, 64), (128, 128, 16), (64, 128, 16), (64, 64, 64)):
    B = 512
    x = torch.randn(B, cin, hw, hw, device="cuda"); W = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    b = torch.zeros(cout, device="cuda"); y = torch.empty(B, cout, hw, hw, device="cuda")
    d = torch.randn(B, cout, hw, hw, device="cuda") * 1e-3; dx = torch.empty_like(x)
    ws = ops.conv_tc_workspace(W); ws2 = ops.conv_tc_workspace(W)
    dW = torch.zeros_like(W); db = torch.zeros(cout, device="cuda")
    f = t(lambda: ops.conv_tc_fwd(x, W, b, "lrelu", y, ws))
    g = t(lambda: ops.conv_tc_bwd_data(d, W, dx, ws2, x, "lrelu"))
    w = t(lambda: ops.conv_tc_bwd_weight(d, x, W, dW, db))
    sc = ops.conv_tc_wgrad_scratch([W], "cuda")
    w2 = t(lambda: ops.conv_tc_bwd_weight(d, x, W, dW, db, sc))
    out.append("%d->%d@%d fwd %.1f bwd %.1f wgrad %.1f / %.1f" % (cin, cout, hw, f, g, w, w2))
for cin, cout, L in ((128, 128, 16), (64, 64, 64), (32, 32, 128)):       # VED decoder: 1-D layers
    B = 512
    x = torch.randn(B, cin, L, device="cuda"); W = torch.randn(cout, cin, 3, device="cuda") * 0.05
    d = torch.randn(B, cout, L, device="cuda") * 1e-3
    dW = torch.zeros_like(W); db = torch.zeros(cout, device="cuda")
    w = t(lambda: ops.conv_tc_bwd_weight(d, x, W, dW, db))
    sc = ops.conv_tc_wgrad_scratch([W], "cuda")
    w2 = t(lambda: ops.conv_tc_bwd_weight(d, x, W, dW, db, sc))
    out.append("1-D %d->%d@%d wgrad %.1f / %.1f" % (cin, cout, L, w, w2))
print(os.environ.get("PVB_LIB", "default")[-20:], os.environ.get("PVB_P3_STAGES", ""), " | ".join(out))
