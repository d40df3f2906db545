"""One forward and one backward-data launch of the persistent pixel GEMM (64 -> 64, 3x3, 32x32, batch 512)
for `ncu --set full -k regex:conv_tc_pix3`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pyroved_b200 import ops  # noqa: E402

B, cin, cout, hw = 512, 64, 64, 32
x = torch.randn(B, cin, hw, hw, device="cuda")
W = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
b = torch.zeros(cout, device="cuda")
y = torch.empty(B, cout, hw, hw, device="cuda")
d = torch.randn(B, cout, hw, hw, device="cuda") * 1e-3
dx = torch.empty_like(x)
ws, ws2 = ops.conv_tc_workspace(W), ops.conv_tc_workspace(W)
for _ in range(3):
    ops.conv_tc_fwd(x, W, b, "lrelu", y, ws)
    ops.conv_tc_bwd_data(d, W, dx, ws2, x, "lrelu")
torch.cuda.synchronize()
