"""Kernel-level GPU tests: the fused small-batch MLP kernels, the pipelined forward layer, the
reductions and the optimizer against plain PyTorch fp32 ops (awkward sizes on purpose)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from pyroved_b200 import ops

pytestmark = pytest.mark.gpu
torch.backends.cuda.matmul.allow_tf32 = False
ACTS = {"tanh": torch.tanh, "lrelu": lambda t: F.leaky_relu(t, 0.01), "gelu": F.gelu,
        "softplus": F.softplus, "relu": F.relu}


def _lin(i, o, gen):
    l = nn.Linear(i, o)
    with torch.no_grad():
        l.weight.copy_(torch.randn(o, i, generator=gen) / i ** 0.5)
        l.bias.copy_(torch.randn(o, generator=gen) * 0.1)
    return l.cuda()


@pytest.mark.parametrize("M,w_in,widths,hdims,act", [
    (13, 96, [160, 72], [5, 5, 7], "gelu"),
    (512, 128, [128], [5, 5], "tanh"),
    (7, 256, [], [3, 3], "lrelu"),
    (33, 64, [256, 36, 128], [64], "softplus"),
])
def test_mlp_tail_chain_wgrad_vs_torch(M, w_in, widths, hdims, act):
    gen = torch.Generator().manual_seed(M + w_in)
    f = ACTS[act]
    dims = [w_in] + widths
    layers = [_lin(dims[i], dims[i + 1], gen) for i in range(len(widths))]
    heads = [_lin(dims[-1], h, gen) for h in hdims]
    h_in = torch.randn(M, w_in, generator=gen).cuda()
    gauss = len(hdims) >= 2 and hdims[0] == hdims[1]
    # reference
    hs, pres, cur = [], [], h_in
    for l in layers:
        p = F.linear(cur, l.weight, l.bias)
        pres.append(p)
        cur = f(p)
        hs.append(cur)
    outs = [F.linear(cur, hd.weight, hd.bias) for hd in heads]
    # kernels
    kh = [torch.empty(M, w, device="cuda") for w in widths]
    kp = [torch.empty(M, w, device="cuda") if act == "gelu" else None for w in widths]
    ko = [torch.empty(M, h, device="cuda") for h in hdims]
    g = None
    if gauss:
        Z = hdims[0]
        eps = torch.randn(M, Z, generator=gen).cuda()
        g = dict(eps=eps, sigma=torch.empty(M, Z, device="cuda"), z=torch.empty(M, Z, device="cuda"),
                 kl=torch.empty(M, device="cuda"), gen_eps=False, seed=1,
                 step_counter=torch.zeros(1, dtype=torch.int32, device="cuda"), first_index=0)
    args = ops.make_mlp_tail_args(M, h_in, layers, kh, kp, act, heads, ko, g, None)
    ops.mlp_tail_fwd(args)
    for a, b in zip(kh, hs):
        assert torch.allclose(a, b, atol=2e-5, rtol=1e-4)
    for a, b in zip(ko, outs):
        assert torch.allclose(a, b, atol=5e-5, rtol=1e-4)
    if gauss:
        sig = F.softplus(outs[1])
        z = outs[0] + sig * eps
        kl = (-0.5 * z * z + 0.5 * eps * eps + torch.log(sig)).sum(1)
        assert torch.allclose(g["sigma"], sig, atol=2e-5) and torch.allclose(g["z"], z, atol=5e-5)
        assert torch.allclose(g["kl"], kl, atol=2e-4, rtol=1e-4)
    if not widths:
        return
    # backward: random head gradients through heads + stack (layer 0 input gradient not needed)
    gh = [torch.randn(M, h, generator=gen).cuda() for h in hdims]
    x0 = torch.randn(M, 24, generator=gen).cuda()       # pretend input of layer 0 for its dW
    l0 = _lin(24, w_in, gen)
    params = [p for l in [l0] + layers + heads for p in (l.weight, l.bias)]
    for p in params:
        p.grad = None
    pre0 = F.linear(x0, l0.weight, l0.bias)
    cur = f(pre0)
    h0 = cur.detach()
    hs2, pres2 = [h0], [pre0.detach()]
    for l in layers:
        p = F.linear(cur, l.weight, l.bias)
        pres2.append(p.detach())
        cur = f(p)
        hs2.append(cur.detach())
    loss = sum((F.linear(cur, hd.weight, hd.bias) * gk).sum() for hd, gk in zip(heads, gh))
    loss.backward()
    all_layers = [l0] + layers
    dpre = [torch.empty(M, l.out_features, device="cuda") for l in all_layers]
    cargs = ops.make_mlp_chain_args(M, all_layers, hs2, pres2 if act == "gelu" else [None] * len(all_layers),
                                    act, dpre, heads, gh)
    ops.mlp_chain_bwd(cargs)
    items = []
    gW = [torch.zeros_like(l.weight) for l in all_layers + heads]
    gb = [torch.zeros_like(l.bias) for l in all_layers + heads]
    for k, l in enumerate(all_layers):
        items.append((dpre[k], x0 if k == 0 else hs2[k - 1], gW[k], gb[k]))
    for k, hd in enumerate(heads):
        items.append((gh[k], hs2[-1], gW[len(all_layers) + k], gb[len(all_layers) + k]))
    ops.mlp_wgrad(ops.make_wgrad_problems(items), M)
    for k, l in enumerate(all_layers + heads):
        sw = l.weight.grad.abs().max().item() + 1e-6
        assert (gW[k] - l.weight.grad).abs().max().item() <= 2e-4 * sw + 1e-5, k
        assert torch.allclose(gb[k], l.bias.grad, atol=2e-4, rtol=1e-4), k


@pytest.mark.parametrize("M,N,K,act", [(512, 128, 784, "tanh"), (100, 96, 260, "gelu"),
                                       (1024, 128, 4100, None), (64, 48, 30, "relu"),
                                       # skinny layers (N <= 8 over a long K: VED's features2latent)
                                       (96, 4, 32768, None), (33, 7, 2052, "tanh"), (512, 2, 8192, None)])
def test_linear_fwd_bwd_vs_torch(M, N, K, act):
    gen = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=gen).cuda()
    l = _lin(K, N, gen)
    f = ACTS[act] if act else (lambda t: t)
    xr = x.clone().requires_grad_(True)
    pre_r = F.linear(xr, l.weight, l.bias)
    yr = f(pre_r)
    dy = torch.randn(M, N, generator=gen).cuda()
    yr.backward(dy)
    y = torch.empty(M, N, device="cuda")
    pre = torch.empty(M, N, device="cuda")
    ops.linear_fwd(x, l.weight.data, l.bias.data, act, out=y, pre=pre)
    assert torch.allclose(y, yr.detach(), atol=1e-4, rtol=1e-4)
    assert torch.allclose(pre, pre_r.detach(), atol=1e-4, rtol=1e-4)
    dx = torch.empty(M, K, device="cuda")
    dW, db = torch.zeros_like(l.weight), torch.zeros_like(l.bias)
    ws = torch.empty(M, N, device="cuda")
    ops.linear_bwd(x, l.weight.data, y, pre, dy, ws, dx, False, dW, db, act)
    assert torch.allclose(dx, xr.grad, atol=2e-4, rtol=1e-4)
    assert (dW - l.weight.grad).abs().max().item() <= 2e-4 * l.weight.grad.abs().max().item() + 1e-5
    assert torch.allclose(db, l.bias.grad, atol=5e-4, rtol=1e-4)
    # accumulate into dx (second head sharing the input), gradients accumulate too
    ops.linear_bwd(x, l.weight.data, y, pre, dy, ws, dx, True, dW, db, act)
    assert torch.allclose(dx, 2 * xr.grad, atol=4e-4, rtol=1e-4)
    assert (dW - 2 * l.weight.grad).abs().max().item() <= 4e-4 * l.weight.grad.abs().max().item() + 2e-5


def test_reductions_adam_and_regression_terms():
    gen = torch.Generator().manual_seed(0)
    part = torch.randn(148, 1000, generator=gen).cuda()
    out = torch.ones(777, device="cuda")
    ops.reduce_partials(part, out, 148, 777, 1000, True)
    assert torch.allclose(out, 1 + part[:, :777].sum(0), atol=1e-4)
    # Adam with the folded step counter == torch.optim.Adam over three steps
    p0 = torch.randn(1003, generator=gen)
    ref = nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3)
    p = p0.clone().cuda()
    pad = torch.zeros(1004, device="cuda")
    pad[:1003] = p
    m, v = torch.zeros(1004, device="cuda"), torch.zeros(1004, device="cuda")
    ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
    ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
    for step in range(3):
        gr = torch.randn(1003, generator=gen)
        ref.grad = gr.clone()
        opt.step()
        gpad = torch.zeros(1004, device="cuda")
        gpad[:1003] = gr.cuda()
        ops.adam_flat_step(pad, gpad, m, v, 1003, 1e-3, ctr, ticket)
        assert ctr.item() == step + 1 and ticket.item() == 0
    assert torch.allclose(pad[:1003].cpu(), ref.detach(), atol=2e-6)
    # Normal log-prob sum and its gradient; input-gradient column slice
    y = torch.randn(50, 3, generator=gen).cuda()
    c = torch.randn(50, 3, generator=gen).cuda().requires_grad_(True)
    lp = torch.distributions.Normal(c, 0.4).log_prob(y).sum()
    (-7.0 * lp).backward()
    loss = torch.zeros(1, device="cuda")
    gc = torch.empty(50, 3, device="cuda")
    ops.normal_logprob(y, c.detach(), 0.4, -7.0, loss, gc)
    assert abs(loss.item() + 7.0 * lp.item()) <= 1e-4 * abs(7.0 * lp.item())
    assert torch.allclose(gc, c.grad, atol=1e-4, rtol=1e-4)
    dpre = torch.randn(37, 64, generator=gen).cuda()
    W = torch.randn(64, 103, generator=gen).cuda()
    dxc = torch.empty(37, 3, device="cuda")
    ops.linear_dx_cols(dpre, W, dxc, 100)
    assert torch.allclose(dxc, (dpre @ W)[:, 100:], atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("n", [4 * 1000, 4 * 300_000])
def test_peer_allreduce_adam_single_rank_equals_adam_flat_step(n):
    """world = 1 degenerate case of the fused NVLink exchange (csrc/pvb_peer.cu): with only its own
    buffer to read, the kernel must reproduce pvb_adam_flat_step bit for bit, publish the loss and
    advance the step counter / epoch; run repeatedly (epoch flags re-arm).  The large case (1.2 M
    parameters, the ssiVAE benchmark's size) launches the grid at its co-residency cap: the kernel waits
    grid-wide, so every CTA must be resident (a register increase once broke exactly that at 8 GPUs; the
    host now sizes the grid from the occupancy the driver reports)."""
    torch.manual_seed(0)
    dev = "cuda"
    p0 = torch.randn(n, device=dev)
    g = torch.zeros(n + 4, device=dev)
    first = torch.zeros(n, dtype=torch.int32, device=dev)
    first[100:200] = -1           # never carried a gradient
    first[200:300] = 2            # joins at step 2
    ref = dict(p=p0.clone(), m=torch.zeros(n, device=dev), v=torch.zeros(n, device=dev),
               c=torch.zeros(1, dtype=torch.int32, device=dev),
               t=torch.zeros(1, dtype=torch.int32, device=dev))
    new = dict(p=p0.clone(), m=torch.zeros(n, device=dev), v=torch.zeros(n, device=dev),
               c=torch.zeros(1, dtype=torch.int32, device=dev))
    flags = torch.zeros(max(64, ops.peer_flag_words()), dtype=torch.int32, device=dev)
    state = torch.zeros(ops.peer_state_words(), dtype=torch.int32, device=dev)
    stage = [torch.zeros(n + 4, device=dev) for _ in range(2)]
    stage_ptrs = torch.tensor([t.data_ptr() for t in stage], dtype=torch.int64, device=dev)
    peer_f = torch.tensor([flags.data_ptr()], dtype=torch.int64, device=dev)
    for it in range(4):
        g[:n] = torch.randn(n, device=dev)
        g[n] = 3.5 + it
        g2 = g.clone()
        ops.adam_flat_step(ref["p"], g, ref["m"], ref["v"], n, 1e-3, ref["c"], ref["t"], first,
                           loss_src=g[n:n + 2])
        ops.peer_allreduce_adam(new["p"], new["m"], new["v"], g2, n, stage_ptrs, peer_f, state, 0, 1,
                                1e-3, new["c"], first, two_shot=bool(it & 1))
        torch.cuda.synchronize()
        for k in ("p", "m", "v"):
            assert torch.equal(ref[k], new[k]), (it, k)
        assert int(new["c"]) == it + 1 and int(state[0]) == it + 1
        assert all(int(state[j]) == 0 for j in (1, 2, 3))
        # both optimizer kernels consume the gradients: buffer zeroed, loss moved to the next slot
        for buf in (g, g2):
            assert float(buf[:n + 1].abs().max()) == 0.0
            assert float(buf[n + 1]) == 3.5 + it
    assert torch.equal(new["p"][100:200], p0[100:200])


# ---- fused tcgen05 spatial decoder, kernel level ------------------------------------------------------
def _sdec_torch_reference(Uv, x, w, W1, b1, W2, b2, wo, bo, H, W, B):
    """Plain PyTorch fp32 (autograd) of what pvb_sdec_tc_step computes: h0 = tanh(U g + v) from the
    folded first-layer coefficients, two tanh layers, output layer, Bernoulli(sigmoid) log-lik,
    loss = -sum_i w_i sum_p ll_ip; gradients wrt Uv and the decoder weights."""
    I = Uv.shape[0]
    gx = torch.linspace(-1, 1, H, device=Uv.device)
    gy = torch.linspace(1, -1, W, device=Uv.device)
    g = torch.stack(torch.meshgrid(gx, gy, indexing="ij"), -1).reshape(-1, 2)      # [N,2]
    leaves = [t.clone().requires_grad_(True) for t in (Uv, W1, b1, W2, b2, wo, bo)]
    Uv_, W1_, b1_, W2_, b2_, wo_, bo_ = leaves
    h0 = torch.tanh(g[None, :, 0, None] * Uv_[:, None, 0] + g[None, :, 1, None] * Uv_[:, None, 1]
                    + Uv_[:, None, 2])                                           # [I,N,128]
    h1 = torch.tanh(h0 @ W1_.t() + b1_)
    h2 = torch.tanh(h1 @ W2_.t() + b2_)
    logit = (h2 @ wo_.t() + bo_).squeeze(-1)                                        # [I,N]
    xi = x[torch.arange(I, device=Uv.device) % B]
    ll = xi * logit - torch.nn.functional.softplus(logit)
    wt = w if w is not None else torch.ones(I, device=Uv.device)
    (-(wt[:, None] * ll).sum()).backward()
    return ll.detach(), torch.sigmoid(logit).detach(), [t.grad for t in leaves]


@pytest.mark.parametrize("variant", ["one-tile", "interleaved"])
@pytest.mark.parametrize("shape", [(64, 64, 28, 28, False), (12, 4, 28, 28, True), (37, 37, 6, 6, False),
                                   (300, 300, 28, 28, False)])
def test_sdec_tc_kernel_vs_torch(shape, variant, monkeypatch):
    """pvb_sdec_tc_step (training: forward + backward in one launch) against PyTorch fp32 autograd,
    both kernel variants: the one-tile-in-flight kernel (csrc/pvb_sdec_tc.cu, default) and the
    experimental interleaved two-tiles-in-flight kernel (csrc/pvb_sdec_tc2.cu, PVB_SDEC_V2=1).  Shapes: a CTA with one
    tile only (64 x 784 rows = 392 tiles over 148 CTAs: 2-3 tiles each), enumerated instances with
    weights (I = 3 B), tiny images (5 instances per tile, ragged last tile), and enough tiles that
    every CTA pipelines many (300 x 784 / 128 = 1838 tiles)."""
    from pyroved_b200._lib import TC_WGRAD_FLOATS, TC_WGRAD_STRIDE
    monkeypatch.setenv("PVB_SDEC_V2", "1" if variant == "interleaved" else "0")
    I, B, H, W, weighted = shape
    N = H * W
    gen = torch.Generator().manual_seed(I * 7 + H)
    r = lambda *s_, sc=1.0: (torch.randn(*s_, generator=gen) * sc).cuda()   # noqa: E731
    Uv = r(I, 3, 128, sc=0.8)
    W1, b1, W2, b2 = r(128, 128, sc=0.09), r(128, sc=0.05), r(128, 128, sc=0.09), r(128, sc=0.05)
    wo, bo = r(1, 128, sc=0.09), r(1, sc=0.05)
    x = (torch.rand(B, N, generator=gen) < 0.3).float().cuda()
    w = torch.rand(I, generator=gen).cuda() if weighted else None
    sz = ops.sdec_tc_sizes(I, N)
    rowll, loc = torch.empty(I * N, device="cuda"), torch.empty(I * N, device="cuda")
    gpart = torch.zeros(max(sz.gUv_part_floats, 4), device="cuda")
    wpart = torch.zeros(max(sz.wgrad_part_floats, 4), device="cuda")
    # first launch: weights converted inside the kernel; second: pre-packed tiles fetched by TMA bulk
    # copies (and no state may leak from one launch into the next)
    packed = ops.sdec_tc_pack_weights(W1, W2, ops.sdec_tc_packed_weights("cuda"))
    for pk in (None, packed):
        rowll.fill_(7.0)
        ops.sdec_tc_step(Uv, x, w, W1, b1, W2, b2, wo, bo, rowll, loc, gpart, wpart, I, B, H, W, 2,
                         "bernoulli", True, 0.5, True, packed_w=pk)
    gUv = torch.empty(I, 3, 128, device="cuda")
    ops.sdec_tc_gather_gUv(gpart, gUv, I, N)
    wsum = wpart.view(sz.ctas, TC_WGRAD_STRIDE)[:, :TC_WGRAD_FLOATS].sum(0)
    ll_ref, loc_ref, (gUv_r, gW1, gb1, gW2, gb2, gwo, gbo) = _sdec_torch_reference(
        Uv, x, w, W1, b1, W2, b2, wo, bo, H, W, B)
    assert (rowll.view(I, N) - ll_ref).abs().max().item() <= 2e-3
    assert (loc.view(I, N) - loc_ref).abs().max().item() <= 1e-3
    got = torch.split(wsum, [128 * 128, 128, 128 * 128, 128, 128, 1])
    for name, a, b_ in (("dUv", gUv, gUv_r), ("dW1", got[0], gW1), ("db1", got[1], gb1),
                        ("dW2", got[2], gW2), ("db2", got[3], gb2), ("dwo", got[4], gwo),
                        ("dbo", got[5], gbo)):
        err = (a.reshape(-1) - b_.reshape(-1)).abs().max().item() / (b_.abs().max().item() + 1e-6)
        assert err <= 3e-3, (name, err)


@pytest.mark.parametrize("M,N,K", [(5, 2, 4096), (301, 2, 4096), (512, 2, 4096), (512, 4, 32768),
                                   (301, 3, 8200), (298, 1, 8192)])
def test_skinny_linear_forward_vs_torch(M, N, K):
    """<= 8 outputs over a long K (VED's 32768 -> 4 features2latent layer): the one-row-per-block kernel
    (small M), the four-rows-per-block kernel (M >= 296, ragged tail) and, for K >= 8192, its variant with the
    K range split over a cluster of 8 CTAs (partial sums joined through distributed shared memory; a K that
    does not divide evenly, fewer than 4 outputs)."""
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) * 0.02).cuda()
    b = torch.randn(N, generator=g).cuda()
    y = ops.linear_fwd(x, W, b, None)
    ref = x @ W.t() + b
    if K >= 8192:      # deterministic: rank-ordered sums
        assert torch.equal(y, ops.linear_fwd(x, W, b, None))
    assert torch.allclose(y, ref, atol=2e-4, rtol=1e-4), (y - ref).abs().max().item()
