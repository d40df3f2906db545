"""Warm, back-to-back timings (CUDA events) of the non-decoder launches of one cfg2 step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import pyroved_b200 as pv  # noqa: E402
from pyroved_b200 import ops  # noqa: E402

B = 512
m = pv.models.iVAE((28, 28), 2, ['r', 't'], seed=1, device="cuda:0")
tr = pv.trainers.SVItrainer(m, device="cuda:0")
eng = tr.svi
eng.use_graphs = False
x = (torch.rand(B, 28, 28) < 0.3).float().cuda()
for _ in range(3):
    eng.step(x)
prog = next(iter(eng.programs.values()))
flat = eng.flat
enc, dec, head = prog.enc, prog.dec, prog.head


def timeit(name, fn, n=20, reps=10):
    """n back-to-back launches captured in one CUDA graph (no host launch overhead)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("{:34s} {:8.2f} us".format(name, e0.elapsed_time(e1) * 1e3 / (n * reps)))


l0 = enc.layers[0]
timeit("memset g", lambda: flat.g.zero_())
timeit("linear fwd (wide layer)", lambda: ops.linear_fwd(prog.enc_in, l0.weight.data, l0.bias.data, "tanh", out=enc.h[0]))
timeit("mlp_tail_fwd", lambda: ops.mlp_tail_fwd(enc._tail[True] if True in enc._tail else next(iter(enc._tail.values()))))
timeit("elbo_reduce (2 launches)", lambda: ops.elbo_reduce(dec.rowll, head.kl, None, 1.0, dec.ll, flat.loss, True, dec.I, dec.N))
from pyroved_b200._lib import TC_WGRAD_FLOATS, TC_WGRAD_STRIDE
from pyroved_b200.nets.fc import linear_layers
L = linear_layers(m.decoder.fc_layers)
base = flat.offset(L[0].weight)
timeit("reduce wgrad partials", lambda: ops.reduce_partials(dec.wgrad_part, flat.g[base:base + TC_WGRAD_FLOATS], dec.tc_sizes.ctas, TC_WGRAD_FLOATS, TC_WGRAD_STRIDE, True))
cl = m.decoder.coord_latent
timeit("latent_side_bwd", lambda: ops.latent_side_bwd(dec.fold_cfg, head.z, None, cl.fc_coord.weight.data, cl.fc_latent.weight.data, None, dec.gUv_part, dec.N, dec.gz, None, dec.fold_part, head.eps, head.sigma, head.s_pre, None, 1.0, head.gmu, head.gs_pre))
timeit("reduce fold partials", lambda: dec._reduce_fold_partials())
timeit("mlp_chain_bwd", lambda: ops.mlp_chain_bwd(enc._chain))
timeit("mlp_wgrad", lambda: ops.mlp_wgrad(enc._wgrad, enc.M))
timeit("adam_flat_step", lambda: eng._update())
eng.use_graphs = True
for _ in range(3):
    eng.step(x, _sync=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    eng.step(x, _sync=False)
e1.record()
torch.cuda.synchronize()
print("{:34s} {:8.2f} us".format("whole step (graph)", e0.elapsed_time(e1) * 1e3 / 50))
