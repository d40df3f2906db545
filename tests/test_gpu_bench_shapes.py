"""GPU parity at the BENCHMARKED shapes (BASELINE.json configs[1..4]): the default
(tcgen05) path of every benchmark workload, called through the drop-in trainers / the C ABI, against
the CPU oracle port on the same seeded inputs, weights and noise.

  cfg2  iVAE 28x28 r+t, B = 512                       (sdec_tc_kernel: 3,136 tiles, persistent)
  cfg3  jiVAE 28x28, K = 10, B = 1024, scale_factor=[3,3]   (8.0 M decoder rows)
  cfg4  ssiVAE 64x64, K = 4, B = 256: unsupervised step, labelled step, auxiliary step
  cfg5  VED 64x64 -> 128, default filters, B = 512    (conv_tc_pix3 / two-launch 128->128 / wgrad)

The oracle is evaluated in chunks of the batch (the loss is a sum over samples; benchlib.chunked_oracle,
checked on CPU by tests/test_oracle.py::test_port_is_separable_over_the_batch).
Tolerances: ELBO <= 1e-3 relative, reconstruction max-abs <= 1e-3 (north_star); gradients per
parity_util / measured margins (gpurun_out/margins.tsv)."""
import pytest
import torch

import benchlib as bl
from conftest import record_margin
from parity_util import TC_GRAD_TOL, grad_check

pytestmark = pytest.mark.gpu
LOSS_RTOL = 1e-3
LOC_ATOL = 1e-3
torch.set_num_threads(max(1, torch.get_num_threads()))


def _sd(m):
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


def _check(tag, loss, ref, prog_loc, m, grads, gtol=TC_GRAD_TOL):
    rel = abs(loss - ref["loss"]) / abs(ref["loss"])
    record_margin(tag, "loss rel err", rel, LOSS_RTOL)
    assert rel <= LOSS_RTOL, (tag, loss, ref["loss"])
    if prog_loc is not None:
        err = (prog_loc.cpu().reshape(-1) - ref["loc"].reshape(-1)).abs().max().item()
        record_margin(tag, "loc max-abs err", err, LOC_ATOL)
        assert err <= LOC_ATOL, (tag, err)
    grad_check(m, grads, gtol, tag, allow_missing=True)


def test_cfg1_shift_ivae_1d_b64():
    """BASELINE configs[0]: 1-D shift-invariant iVAE on 1x64 spectra, latent_dim=2, batch 64 (the
    reference's own CPU-runnable case), data as examples/shiftVAE.ipynb: noisy shifted Gaussian peaks."""
    import pyroved_b200 as pv
    from oracle import svi_port as sp
    B, L = 64, 64
    g = torch.Generator().manual_seed(1)
    xs = torch.linspace(-12, 12, L)
    mu = (torch.rand(B, generator=g) - 0.5) * 14
    sig = 1.0 + torch.rand(B, generator=g)
    x = torch.exp(-0.5 * ((xs[None] - mu[:, None]) / sig[:, None]) ** 2) + 0.05 * torch.randn(B, L, generator=g)
    x = (x - x.min()) / (x.max() - x.min())                        # [B, 64], scaled to [0, 1]
    eps = torch.randn(B, 3, generator=g)
    for generic in (True, False):
        m = pv.models.iVAE((L,), latent_dim=2, invariances=['t'], seed=1, device="cuda:0")
        tr = pv.trainers.SVItrainer(m, seed=1, device="cuda:0", force_generic=generic)
        sd = _sd(m)
        loss = tr.svi.loss_and_grads(x.cuda(), _eps=eps.cuda())
        prog = next(iter(tr.svi.programs.values()))
        assert bool(prog.use_tc) == (not generic)
        ref, grads = sp.loss_and_grads(sp.ivae_loss, sd, sp.Cfg((L,), 2, ['t']), x, eps)
        ref = dict(ref, loss=float(ref["loss"]))
        _check("cfg1 1-D shift B=64 {}".format("fp32" if generic else "tc"), loss, ref, prog.loc, m, grads,
               2e-3 if generic else TC_GRAD_TOL)


def test_cfg2_ivae_b512():
    name = "cfg2"
    m, tr = bl.build(name, "cuda:0")
    B = bl.WORKLOADS[name]["batch"]
    (x,) = bl.synth(name, B, seed=3)
    eps = torch.randn(B, m.z_dim, generator=torch.Generator().manual_seed(4))
    sd = _sd(m)
    loss = tr.svi.loss_and_grads(x.cuda(), _eps=eps.cuda())
    prog = next(iter(tr.svi.programs.values()))
    assert prog.use_tc
    ref, grads = bl.chunked_oracle("ivae", sd, bl.oracle_cfg(name), (x,), eps, 1.0, chunk=B)
    _check("cfg2 B=512 tc", loss, ref, prog.loc, m, grads)
    assert torch.allclose(prog.mu.cpu(), ref["mu"], atol=1e-4)


def test_cfg3_jivae_b1024_k10():
    name = "cfg3"
    m, tr = bl.build(name, "cuda:0")
    B = bl.WORKLOADS[name]["batch"]
    (x,) = bl.synth(name, B, seed=5)
    eps = torch.randn(B, m.z_dim, generator=torch.Generator().manual_seed(6))
    sd = _sd(m)
    loss = tr.svi.loss_and_grads(x.cuda(), _eps=eps.cuda(), scale_factor=[3.0, 3.0])
    prog = next(iter(tr.svi.programs.values()))
    assert prog.use_tc
    ref, grads = bl.chunked_oracle("jivae", sd, bl.oracle_cfg(name), (x,), eps, (3.0, 3.0), chunk=64)
    _check("cfg3 B=1024 K=10 tc", loss, ref, prog.loc, m, grads)
    assert torch.allclose(prog.alpha.cpu(), ref["alpha"], atol=1e-5)


@pytest.mark.parametrize("labelled", [False, True], ids=["unsup", "sup"])
def test_cfg4_ssivae_64x64_k4_b256(labelled):
    name = "cfg4"
    m, tr = bl.build(name, "cuda:0")
    B = bl.WORKLOADS[name]["batch"]
    data = bl.synth(name, B, seed=7, labelled=labelled)
    gen = torch.Generator().manual_seed(8)
    eps = torch.randn(B, m.z_dim, generator=gen) if labelled else torch.randn(4, B, m.z_dim, generator=gen)
    sd = _sd(m)
    args = tuple(t.cuda() for t in data)
    loss = tr.svi.loss_and_grads(*args, _eps=eps.cuda())
    prog = tr.svi.programs[(B, labelled, "main")]
    assert prog.use_tc
    cfg = bl.oracle_cfg(name)
    ref, grads = bl.chunked_oracle("ssivae", sd, cfg, data, eps, 1.0, chunk=16 if not labelled else 64)
    tag = "cfg4 64x64 K=4 B=256 {} tc".format("sup" if labelled else "unsup")
    _check(tag, loss, ref, prog.loc, m, grads)
    # the auxiliary step of auxSVItrainer.compute_loss at this shape (classifier only)
    la = tr.svi._step(args, {"aux_loss_multiplier": 50.0}, train=True, update=False, mode="aux")
    ref_a, grads_a = bl.chunked_oracle("ssivae_aux", sd, cfg, data, None, 50.0, chunk=B)
    if labelled:
        assert abs(la - ref_a["loss"]) <= LOSS_RTOL * abs(ref_a["loss"]), (la, ref_a["loss"])
        ga = {k: v for k, v in grads_a.items() if v is not None}
        sub = torch.nn.Module()
        sub.encoder_y = m.encoder_y
        grad_check(sub, ga, 2e-3, tag + " aux", allow_missing=True)
    else:
        assert la == 0.0 and ref_a["loss"] == 0.0


def test_cfg5_ved_b512():
    name = "cfg5"
    m, tr = bl.build(name, "cuda:0")
    B = bl.WORKLOADS[name]["batch"]
    x, y = bl.synth(name, B, seed=9)
    eps = torch.randn(B, 2, generator=torch.Generator().manual_seed(10))
    sd = _sd(m)
    loss = tr.svi.loss_and_grads(x.cuda(), y.cuda(), _eps=eps.cuda(), scale_factor=4.0)
    prog = next(iter(tr.svi.programs.values()))
    assert prog.use_tc
    ref, grads = bl.chunked_oracle("ved", sd, bl.oracle_cfg(name), (x, y), eps, 4.0, chunk=128)
    rel = abs(loss - ref["loss"]) / abs(ref["loss"])
    record_margin("cfg5 B=512 tc", "loss rel err", rel, LOSS_RTOL)
    assert rel <= LOSS_RTOL, (loss, ref["loss"])
    err = (prog.loc.cpu().reshape(B, -1) - ref["loc"]).abs().max().item()
    record_margin("cfg5 B=512 tc", "loc max-abs err", err, LOC_ATOL)
    assert err <= LOC_ATOL, err
    # gradients: per-tensor max-norm / L2 and whole-vector L2, bounds as test_gpu_ved.grad_check
    from test_gpu_ved import grad_check as ved_grad_check
    ved_grad_check(m, {k: v.cuda() for k, v in grads.items()}, False, "cfg5 B=512")
