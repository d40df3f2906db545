"""Isolated launches of the tcgen05 weight-gradient kernel on cfg5 layer shapes (for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pyroved_b200 import ops  # noqa: E402

B = 512
for (cin, cout, k, hw) in [(32, 32, 1, 64), (64, 32, 3, 64), (64, 64, 3, 32)]:
    x = torch.randn(B, cin, hw, hw, device="cuda")
    dpre = torch.randn(B, cout, hw, hw, device="cuda") * 1e-3
    W = torch.randn(cout, cin, k, k, device="cuda")
    dW = torch.zeros_like(W)
    db = torch.zeros(cout, device="cuda")
    for _ in range(3):
        ops.conv_tc_bwd_weight(dpre, x, W, dW, db)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv_tc_bwd_weight(dpre, x, W, dW, db)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    gb = (x.numel() + dpre.numel()) * 4 / 1e9
    print("wgrad {}->{} k{} {}x{}: {:.1f} us, unique {:.2f} GB -> {:.2f} TB/s".format(
        cin, cout, k, hw, hw, us, gb, gb / us * 1e3), flush=True)
