"""Time the skinny linear kernels at VED's features2latent shape (batch 512, 32768 -> 4)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyroved_b200 import ops
M, K, N = 512, 32768, 4
x = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") * 0.01; b = torch.zeros(N, device="cuda")
g = torch.randn(M, N, device="cuda"); dx = torch.empty_like(x); dW = torch.zeros_like(W); db = torch.zeros_like(b)
y = torch.empty(M, N, device="cuda")
def t(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
f = t(lambda: ops.linear_fwd(x, W, b, None, out=y))
bw = t(lambda: ops.linear_bwd(x, W, None, None, g, g, dx, False, dW, db, None))
print(os.environ.get("PVB_LIB", "default")[-16:], "fwd %.1f us   bwd (dx + dW) %.1f us" % (f, bw))
