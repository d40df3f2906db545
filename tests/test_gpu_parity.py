"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called
through the C ABI via the drop-in Python classes, against
  (a) golden vectors produced by the unmodified reference (tests/golden), and
  (b) the CPU oracle port (oracle/svi_port.py) on fresh seeded inputs.
Tolerances: ELBO (loss) <= 1e-3 relative and reconstruction max-abs <= 1e-3,
as BASELINE.json's north_star states (fp32 reference)."""
import pytest
import torch

import pyroved_b200 as pv
from golden_util import CASES, Golden
from oracle import svi_port as sp
from parity_util import FP32_GRAD_TOL, TC_GRAD_TOL, check_w1, grad_check

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-3
LOC_ATOL = 1e-3
SEEDS = {"ivae_12_rts_cond_gauss": 2, "ivae_12_vanilla": 3, "ivae_16_s_softplus": 4,
         "ivae_12_r_cbern": 5}
IVAE_CASES = [n for n in sorted(CASES) if CASES[n][0] == "ivae"]


def build_ivae(name, g, generic):
    m = pv.models.iVAE(seed=SEEDS.get(name, 1), device="cuda:0", **g.kwargs)
    m.load_state_dict(g.group("w0"))
    tr = pv.trainers.SVItrainer(m, seed=1, device="cuda:0", force_generic=generic)
    return m, tr


@pytest.mark.parametrize("generic", [True, False], ids=["fp32-generic", "default"])
@pytest.mark.parametrize("name", IVAE_CASES)
def test_ivae_loss_recon_grads_vs_reference_golden(name, generic):
    g = Golden(name)
    m, tr = build_ivae(name, g, generic)
    x, y = g.args()
    args = (x.cuda(),) if y is None else (x.cuda(), y.cuda())
    kw = {k: float(v) for k, v in g.kw().items()}
    loss = tr.svi.loss_and_grads(*args, _eps=g.eps().cuda(), **kw)
    assert abs(loss - g.loss) <= LOSS_RTOL * abs(g.loss), (loss, g.loss)
    prog = next(iter(tr.svi.programs.values()))
    loc = prog.loc.reshape(-1).cpu()
    assert (loc - g.t("loc").reshape(-1)).abs().max().item() <= LOC_ATOL
    assert torch.allclose(prog.mu.cpu(), g.t("mu"), atol=1e-4)
    assert torch.allclose(prog.sigma.cpu(), g.t("sigma"), atol=1e-4)
    tc = getattr(prog, "use_tc", False)
    grad_check(m, g.group("grad"), TC_GRAD_TOL if tc else FP32_GRAD_TOL,
               "golden {} {}".format(name, "tc" if tc else "fp32"))


@pytest.mark.parametrize("generic", [True, False], ids=["fp32-generic", "default"])
@pytest.mark.parametrize("name", ["ivae_28_rt", "ivae_1d_t"])
def test_ivae_full_step_matches_reference_adam(name, generic):
    """loss_and_grads + Adam == reference SVI.step (weights after one step), on both paths."""
    g = Golden(name)
    m, tr = build_ivae(name, g, generic=generic)
    x, y = g.args()
    loss = tr.svi.step(x.cuda(), _eps=g.eps().cuda())
    assert abs(loss - g.loss_step) <= LOSS_RTOL * abs(g.loss_step)
    prog = next(iter(tr.svi.programs.values()))
    assert bool(getattr(prog, "use_tc", False)) == (not generic)
    check_w1(m, g, noise_floor=None if generic else TC_GRAD_TOL)


@pytest.mark.parametrize("generic", [True, False], ids=["fp32-generic", "default"])
@pytest.mark.parametrize("B", [1, 7, 64])
def test_ivae_vs_oracle_fresh_inputs(B, generic):
    """Ragged / tiny batches against the oracle port on seeded inputs, on both paths."""
    torch.manual_seed(B)
    m = pv.models.iVAE((28, 28), 2, ['r', 't', 's'], seed=5, device="cuda:0")
    tr = pv.trainers.SVItrainer(m, device="cuda:0", force_generic=generic)
    gen = torch.Generator().manual_seed(100 + B)
    x = (torch.rand(B, 28, 28, generator=gen) < 0.4).float()
    eps = torch.randn(B, m.z_dim, generator=gen)
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    loss = tr.svi.loss_and_grads(x.cuda(), _eps=eps.cuda(), scale_factor=2.0)
    cfg = sp.Cfg((28, 28), 2, ['r', 't', 's'])
    ref, grads = sp.loss_and_grads(sp.ivae_loss, sd, cfg, x, eps, None, 2.0)
    assert abs(loss - float(ref["loss"])) <= LOSS_RTOL * abs(float(ref["loss"]))
    prog = next(iter(tr.svi.programs.values()))
    assert bool(prog.use_tc) == (not generic)
    assert (prog.loc.cpu().reshape(B, -1) - ref["loc"]).abs().max().item() <= LOC_ATOL
    assert torch.allclose(prog.ll.cpu(), ref["ll"], rtol=1e-3, atol=1e-2)
    grad_check(m, grads, FP32_GRAD_TOL if generic else TC_GRAD_TOL,
               "fresh rts B={} {}".format(B, "fp32" if generic else "tc"))


def test_graph_replay_equals_eager_and_training_reduces_loss():
    """CUDA-graph replay gives the same numbers as eager launches; a few
    hundred steps on structured data reduce the per-sample loss."""
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(3)
    x = (torch.rand(128, 28, 28, generator=gen) < 0.2).float()
    x[:, 10:18, 10:18] = 1.0
    eps = torch.randn(128, 5, generator=gen).cuda()
    losses = {}
    for graphs in (False, True):
        m = pv.models.iVAE((28, 28), 2, ['r', 't'], seed=1, device="cuda:0")
        tr = pv.trainers.SVItrainer(m, device="cuda:0")
        tr.svi.use_graphs = graphs
        ls = [tr.svi.step(x.cuda(), _eps=eps) for _ in range(6)]
        losses[graphs] = ls
    for a, b in zip(losses[False], losses[True]):
        assert abs(a - b) <= 1e-4 * abs(a), (a, b)
    assert losses[True][-1] < losses[True][0]
    m = pv.models.iVAE((28, 28), 2, ['r', 't'], seed=1, device="cuda:0")
    tr = pv.trainers.SVItrainer(m, device="cuda:0")
    loader = pv.utils.init_dataloader(x, batch_size=64, shuffle=False)
    for _ in range(30):
        tr.step(loader)
    hist = tr.loss_history["training_loss"]
    assert hist[-1] < 0.9 * hist[0], hist[::5]
    assert all(h == h for h in hist)


def test_sanity_noise_loss_is_784_ln2():
    m = pv.models.iVAE((28, 28), 2, ['r', 't'], seed=1, device="cuda:0")
    tr = pv.trainers.SVItrainer(m, device="cuda:0")
    gen = torch.Generator().manual_seed(0)
    x = (torch.rand(256, 28, 28, generator=gen) < 0.5).float()
    per = tr.svi.evaluate_loss(x.cuda()) / 256
    assert 540 < per < 550, per


def test_evaluate_matches_reference_quirk_and_api():
    """evaluate(): forward-only loss; trainer bookkeeping like the reference."""
    m = pv.models.iVAE((8, 8), 2, ['r'], seed=1, device="cuda:0")
    tr = pv.trainers.SVItrainer(m, device="cuda:0")
    x = torch.rand(10, 8, 8)
    tl = pv.utils.init_dataloader(x, batch_size=2)
    w0 = {k: v.clone() for k, v in m.state_dict().items()}
    for _ in range(2):
        tr.step(tl, tl)
    assert tr.current_epoch == 2
    assert len(tr.loss_history["training_loss"]) == 2 and len(tr.loss_history["test_loss"]) == 2
    assert all(v == v for v in tr.loss_history["training_loss"])
    assert any(not torch.equal(w0[k], v) for k, v in m.state_dict().items())
    tr.print_statistics()


def test_encode_decode_manifold_shapes_and_values():
    g = Golden("ivae_28_rt")
    m = pv.models.iVAE(seed=1, device="cuda:0", **g.kwargs)
    m.load_state_dict(g.group("w0"))
    x, _ = g.args()
    mu, sd = m.encode(x)
    assert mu.shape == (16, 5) and sd.shape == (16, 5)
    assert torch.allclose(mu, g.t("mu"), atol=1e-4) and torch.allclose(sd, g.t("sigma"), atol=1e-4)
    # decode(z) with the transform latents of the golden case reproduces its `loc`
    z = g.t("z")
    loc = m.decode(z[:, 3:], angle=z[0, 0], shift=z[0, 1:3] * 0.1)
    assert loc.shape == (16, 28, 28)
    assert (loc[0].reshape(-1) - g.t("loc")[0].reshape(-1)).abs().max().item() <= LOC_ATOL
    man = m.manifold2d(4, plot=False)
    assert man.shape == (16, 28, 28)
    assert man.min() >= 0 and man.max() <= 1


def test_rng_is_counter_based_and_reproducible():
    from pyroved_b200 import ops
    ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
    a = torch.empty(100000, device="cuda")
    b = torch.empty(50000, device="cuda")
    ops.randn(a, 1234, ctr, 0)
    ops.randn(b, 1234, ctr, 50000)
    assert torch.equal(a[50000:], b)          # sharding-invariant
    assert abs(a.mean().item()) < 0.02 and abs(a.std().item() - 1) < 0.02
    ops.counter_add(ctr, 1)
    c = torch.empty(100000, device="cuda")
    ops.randn(c, 1234, ctr, 0)
    assert not torch.equal(a, c)


@pytest.mark.parametrize("dim,inv,B", [((6, 6), ['r', 't'], 37), ((33,), ['t'], 50), ((5, 9), ['s'], 3)])
def test_fused_decoder_small_images_many_slots_per_tile(dim, inv, B):
    """N = 36 / 33 / 45 pixels: a 128-row tile of the fused kernel touches up to 5 instances, tiles
    straddle instance boundaries and the last tile is partial (ragged R)."""
    m = pv.models.iVAE(dim, 2, inv, seed=7, device="cuda:0")
    tr = pv.trainers.SVItrainer(m, device="cuda:0", force_generic=False)
    gen = torch.Generator().manual_seed(B)
    x = (torch.rand(B, *dim, generator=gen) < 0.4).float()
    eps = torch.randn(B, m.z_dim, generator=gen)
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    loss = tr.svi.loss_and_grads(x.cuda(), _eps=eps.cuda())
    prog = next(iter(tr.svi.programs.values()))
    assert prog.use_tc
    cfg = sp.Cfg(dim, 2, inv)
    ref, grads = sp.loss_and_grads(sp.ivae_loss, sd, cfg, x, eps)
    assert abs(loss - float(ref["loss"])) <= LOSS_RTOL * abs(float(ref["loss"]))
    assert (prog.loc.cpu().reshape(B, -1) - ref["loc"]).abs().max().item() <= LOC_ATOL
    # B = 3 images of 45 pixels: every gradient is a sum over only 135 rows of terms of both signs
    # (sum |dl| ~ 67 against |sum dl| = 0.27), so the 2.4e-4 operand rounding and the ~5e-4 coherent
    # error of tanh.approx are amplified ~100x relative to the tensor's largest entry: measured 1.27e-2
    # on decoder.fc_layers.2.bias (max 0.047), against <= 1.4e-3 at every realistic shape
    # (gpurun_out/margins.tsv).  Bound: 1.6x the measurement for this case, TC_GRAD_TOL for the others.
    tol = 2e-2 if B == 3 else TC_GRAD_TOL
    grad_check(m, grads, tol, "small images {} B={}".format(dim, B))


def test_epoch_loop_equals_step_by_step():
    """SVItrainer.train (pipelined: staging-slot copies inside the step graphs, losses delivered through
    the pinned host ring) == the same batches fed one by one through svi.step: same epoch loss, same
    weights (to 1e-6).  5 batches of 512 + a ragged batch of 256: the 4-slot ring wraps and a second
    program / graph pair serves the last batch; two epochs so every graph is replayed.  (Batches this
    size run the fixed-order kernels throughout; tiny batches take the split-K GEMM whose atomics make
    the last bit of the encoder output run-dependent, which Adam's g / sqrt(v) turns into lr-sized
    differences on weights whose gradient is rounding noise.)"""
    gen = torch.Generator().manual_seed(9)
    x = (torch.rand(5 * 512 + 256, 28, 28, generator=gen) < 0.25).float()
    loader = pv.utils.init_dataloader(x, batch_size=512, shuffle=False)
    ma = pv.models.iVAE((28, 28), 2, ['r', 't'], seed=1, device="cuda:0")
    ta = pv.trainers.SVItrainer(ma, seed=1, device="cuda:0")
    mb = pv.models.iVAE((28, 28), 2, ['r', 't'], seed=1, device="cuda:0")
    tb = pv.trainers.SVItrainer(mb, seed=1, device="cuda:0")
    for epoch in range(2):
        got = ta.train(loader) * len(loader.dataset)
        ref = sum(tb.svi.step(xb.cuda()) for (xb,) in loader)
        assert abs(got - ref) <= 1e-6 * abs(ref), (epoch, got, ref)
    for (k, a), (_, b) in zip(ma.state_dict().items(), mb.state_dict().items()):
        assert torch.allclose(a, b, atol=1e-6, rtol=0), (k, (a - b).abs().max().item())


@pytest.mark.parametrize("generic", [True, False], ids=["fp32-generic", "default"])
def test_ivae_with_conv_encoder_via_set_encoder(generic):
    """iVAE.set_encoder(convEncoderNet(...)) (reference models/base.py:173-176; BASELINE configs[1]
    words the model as "conv encoder / fc decoder"): convolutional guide on the conv kernels, fused
    spatial decoder behind it; loss / reconstruction / gradients against the port, then training."""
    from pyroved_b200.nets import convEncoderNet
    B = 48
    m = pv.models.iVAE((28, 28), 2, ['r', 't'], seed=1, device="cuda:0")
    torch.manual_seed(11)
    m.set_encoder(convEncoderNet((28, 28), latent_dim=m.z_dim, hidden_dim=[(16,), (32, 32)]))
    tr = pv.trainers.SVItrainer(m, device="cuda:0", force_generic=generic)
    gen = torch.Generator().manual_seed(12)
    x = (torch.rand(B, 28, 28, generator=gen) < 0.3).float()
    eps = torch.randn(B, m.z_dim, generator=gen)
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    loss = tr.svi.loss_and_grads(x.cuda(), _eps=eps.cuda(), scale_factor=2.0)
    prog = next(iter(tr.svi.programs.values()))
    assert prog.conv_enc and bool(prog.use_tc) == (not generic)
    cfg = sp.Cfg((28, 28), 2, ['r', 't'])
    vcfg = sp.VedCfg((28, 28), (1,), latent_dim=m.z_dim, hidden_dim_e=[(16,), (32, 32)])
    ref, grads = sp.loss_and_grads(sp.ivae_loss, sd, cfg, x, eps, None, 2.0, conv_encoder=vcfg)
    assert abs(loss - float(ref["loss"])) <= LOSS_RTOL * abs(float(ref["loss"]))
    assert (prog.loc.cpu().reshape(B, -1) - ref["loc"]).abs().max().item() <= LOC_ATOL
    assert torch.allclose(prog.mu.cpu(), ref["mu"], atol=1e-3 if not generic else 1e-4)
    # decoder gradients at the decoder's tolerance; encoder (leaky-ReLU convolutions on fp16
    # tensor-core operands on the default path) at the VED bound
    dec_grads = {k: v for k, v in grads.items() if k.startswith("decoder.")}
    sub = torch.nn.Module()
    sub.decoder = m.decoder
    grad_check(sub, dec_grads, FP32_GRAD_TOL if generic else TC_GRAD_TOL,
               "conv-encoder iVAE decoder {}".format("fp32" if generic else "tc"))
    for k, p in m.encoder_z.named_parameters():
        r = grads["encoder_z." + k].cuda()
        err = (p.grad - r).norm().item() / (r.norm().item() + 1e-9)
        assert err <= (2e-3 if generic else 6e-2), (k, err)
    xc, ec = x.cuda(), eps.cuda()
    ls = [tr.svi.step(xc, _eps=ec) for _ in range(20)]
    assert all(v == v for v in ls) and ls[-1] < ls[0], ls[::5]
    mu, sig = m.encode(x)
    assert mu.shape == (B, m.z_dim) and (sig > 0).all()
