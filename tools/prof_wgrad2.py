"""One weight-gradient launch (64 -> 64, 3x3, 32x32, batch 512) for `ncu --set full -k regex:conv_tc_wgrad`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pyroved_b200 import ops  # noqa: E402

B, cin, cout, hw = 512, 64, 64, 32
x = torch.randn(B, cin, hw, hw, device="cuda")
W = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
d = torch.randn(B, cout, hw, hw, device="cuda") * 1e-3
dW, db = torch.zeros_like(W), torch.zeros(cout, device="cuda")
for _ in range(3):
    ops.conv_tc_bwd_weight(d, x, W, dW, db)
torch.cuda.synchronize()
