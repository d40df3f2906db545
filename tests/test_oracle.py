"""CPU tests: the oracle port (oracle/svi_port.py) against golden vectors
produced by the unmodified reference (oracle/make_golden.py)."""
import math

import numpy as np
import pytest
import torch

from golden_util import CASES, Golden
from oracle import svi_port as sp

torch.set_num_threads(4)


def port_loss_fn(g, dtype=torch.float32):
    """Returns (loss_fn, args, kwargs) for svi_port.loss_and_grads."""
    kw = dict(g.kwargs)
    if g.kind == "ved":
        cfg = sp.VedCfg(**kw)
        x, y = g.args(dtype)
        return sp.ved_loss, (cfg, x, y, g.eps(dtype), float(g.kw().get("scale_factor", 1.0))), cfg
    if g.kind == "ssreg":
        cfg = sp.Cfg(kw["data_dim"], kw["latent_dim"], kw["invariances"], c_dim=kw["reg_dim"])
        x, y = g.args(dtype)
        e = g.eps(dtype)
        ez, ey = (e["z"], e["y"]) if isinstance(e, dict) else (e, None)
        return sp.ss_reg_loss, (cfg, x, ez, y, float(g.kw().get("scale_factor", 1.0)), ey), cfg
    kw.pop("hidden_dim_e", None)
    kw.pop("hidden_dim_d", None)
    cfg = sp.Cfg(**kw)
    x, y = g.args(dtype)
    eps = g.eps(dtype)
    k = g.kw()
    if g.kind == "ivae":
        return sp.ivae_loss, (cfg, x, eps, y, float(k.get("scale_factor", 1.0))), cfg
    if g.kind == "jivae":
        return sp.jivae_loss, (cfg, x, eps, tuple(k["scale_factor"])), cfg
    if g.kind == "ssivae":
        return sp.ssivae_loss, (cfg, x, eps, y, float(k.get("scale_factor", 1.0))), cfg
    raise KeyError(g.kind)


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_matches_reference_fp32(name):
    g = Golden(name)
    fn, args, cfg = port_loss_fn(g)
    sd = g.group("w0")
    out, grads = sp.loss_and_grads(fn, sd, *args)
    # loss: same ops in the same order on the same CPU -> tight tolerance
    assert abs(float(out["loss"]) - g.loss) <= 2e-6 * abs(g.loss)
    if g.has("loc"):
        ref = g.t("loc").reshape(out["loc"].shape)
        assert torch.allclose(out["loc"], ref, atol=2e-6)
    if g.has("mu"):
        assert torch.allclose(out["mu"], g.t("mu").reshape(out["mu"].shape), atol=1e-6)
        assert torch.allclose(out["sigma"], g.t("sigma").reshape(out["sigma"].shape), atol=1e-6)
    gref = g.group("grad")
    assert len(gref) > 0
    for k, v in gref.items():
        scale = v.abs().max().item() + 1e-12
        assert grads[k] is not None, k
        err = (grads[k] - v).abs().max().item()
        assert err <= 2e-4 * scale + 1e-6, (k, err, scale)


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_fp64_close_to_reference(name):
    """fp64 port vs fp32 reference: bounds the fp32 rounding of the oracle."""
    g = Golden(name)
    fn, args, cfg = port_loss_fn(g, torch.float64)
    sd = g.group("w0", torch.float64)
    out, _ = sp.loss_and_grads(fn, sd, *args)
    assert abs(float(out["loss"]) - g.loss) <= 1e-5 * abs(g.loss)


@pytest.mark.parametrize("name", [n for n in sorted(CASES) if CASES[n][0] in ("ivae", "jivae", "ved")])
def test_port_adam_step_matches_reference(name):
    g = Golden(name)
    fn, args, cfg = port_loss_fn(g)
    sd = g.group("w0")
    stats = {}
    kw = {"stats": stats} if getattr(cfg, "batchnorm", False) else {}
    out, grads = sp.loss_and_grads(fn, sd, *args, **kw)
    new = sp.AdamState(lr=1e-3).step(dict(sd), grads)
    new.update(stats)        # batch-norm running statistics after the step
    assert not kw or len(stats) > 0
    for k, v in g.group("w1").items():
        assert torch.allclose(new[k], v, atol=2e-5), k
    for k, idx in g.group("w1idx", torch.int64).items():
        assert torch.allclose(new[k].reshape(-1)[idx], g.t("w1sub." + k), atol=2e-5), k


def test_ssivae_aux_and_two_optimizer_steps():
    """auxSVItrainer.compute_loss = basic step then aux step, two Adam updates
    sharing per-parameter state (trainers/auxsvi.py:88-100)."""
    g = Golden("ssivae_16_r_sup")
    fn, args, cfg = port_loss_fn(g)
    sd = g.group("w0")
    x, y = g.args()
    opt = sp.AdamState(lr=5e-4)
    out, grads = sp.loss_and_grads(fn, sd, *args)
    sd1 = opt.step(dict(sd), grads)
    out2, grads2 = sp.loss_and_grads(sp.ssivae_aux_loss, sd1, cfg, x, y, 50.0)
    # reference zeroes (not None) grads of params unused by the aux loss, so
    # Adam still moves them with g=0
    grads2 = {k: (v if v is not None else torch.zeros_like(sd1[k])) for k, v in grads2.items()}
    sd2 = opt.step(sd1, grads2)
    assert abs(float(out["loss"]) + float(out2["loss"]) - g.loss_step) <= 1e-5 * abs(g.loss_step)
    for k, v in g.group("w1").items():
        assert torch.allclose(sd2[k], v, atol=2e-5), k


def test_closed_form_folded_first_layer():
    """W_c (A g + t) + b_c + W_z z == U g + v."""
    torch.manual_seed(0)
    b, hd = 5, 128
    grid = sp.generate_grid((9, 7), torch.float64)
    phi, dx, sc = torch.randn(b, dtype=torch.float64), torch.randn(b, 1, 2, dtype=torch.float64), \
        1 + 0.1 * torch.randn(b, dtype=torch.float64)
    wc, bc = torch.randn(hd, 2, dtype=torch.float64), torch.randn(hd, dtype=torch.float64)
    xc = sp.transform_coordinates(grid.expand(b, *grid.shape), phi, dx, sc)
    ref = xc @ wc.t() + bc
    c, s = torch.cos(phi), torch.sin(phi)
    # row-vector convention: x' = s*(gx cos - gy sin) + dx ; y' = s*(gx sin + gy cos) + dy
    a = torch.stack([torch.stack([sc * c, -sc * s], 1), torch.stack([sc * s, sc * c], 1)], 1)  # [b,2,2]
    u = torch.einsum("hk,bkj->bhj", wc, a)
    v = torch.einsum("hk,bk->bh", wc, dx[:, 0]) + bc
    alt = torch.einsum("bhj,nj->bnh", u, grid) + v[:, None]
    assert torch.allclose(ref, alt, atol=1e-12)


def test_sanity_noise_loss_is_784_ln2():
    cfg = sp.Cfg((28, 28), 2, ["r", "t"])
    sd = sp.init_ivae_state(cfg, seed=1)
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(64, 28, 28, generator=g) < 0.5).float()
    eps = torch.randn(64, 5, generator=g)
    out = sp.ivae_loss(sd, cfg, x, eps)
    per = float(out["loss"]) / 64
    assert 540 < per < 550, per  # 784 ln 2 = 543.4 (+KL), SURVEY 8c(3)


def test_init_matches_reference_weights():
    """init_ivae_state reproduces the reference constructor's weights."""
    g = Golden("ivae_28_rt")
    cfg = sp.Cfg((28, 28), 2, ["r", "t"])
    sd = sp.init_ivae_state(cfg, seed=1)
    w0 = g.group("w0")
    assert list(sd.keys()) == list(w0.keys())
    for k in sd:
        assert torch.equal(sd[k], w0[k]), k


def test_ss_reg_aux_and_two_optimizer_steps():
    """auxSVItrainer(task="regression").compute_loss on a labeled batch: ELBO step, then the
    auxiliary regression step, two Adam updates sharing per-parameter state."""
    g = Golden("ssreg_16_rt_sup")
    fn, args, cfg = port_loss_fn(g)
    sd = g.group("w0")
    x, y = g.args()
    opt = sp.AdamState(lr=5e-4)
    out, grads = sp.loss_and_grads(fn, sd, *args)
    sd1 = opt.step(dict(sd), grads)
    out2, grads2 = sp.loss_and_grads(sp.ss_reg_aux_loss, sd1, cfg, x, y, 30.0)
    grads2 = {k: (v if v is not None else torch.zeros_like(sd1[k])) for k, v in grads2.items()}
    sd2 = opt.step(sd1, grads2)
    assert abs(float(out["loss"]) + float(out2["loss"]) - g.loss_step) <= 1e-5 * abs(g.loss_step)
    for k, v in g.group("w1").items():
        assert torch.allclose(sd2[k], v, atol=2e-5), k


@pytest.mark.parametrize("name", ["jivae_28_r", "jivae_28_r_beta1", "ssivae_16_r_unsup"])
def test_enumerated_elbo_independent_route(name):
    """The enumerated goldens come from the restated TraceEnum_ELBO (oracle/pyro_min).  An
    independent route that never touches it -- K*B single-sample supervised Trace_ELBO runs of the
    unmodified reference for ssiVAE, direct calls of the reference's nets + torch.distributions for
    jiVAE (oracle/make_golden.py::main_enum) -- gives the same losses; so does the port."""
    import os
    from golden_util import GOLDEN_DIR
    indep = float(np.load(os.path.join(GOLDEN_DIR, "enum_indep.npz"))[name])
    g = Golden(name)
    assert abs(indep - g.loss) <= 2e-6 * abs(g.loss), (indep, g.loss)
    fn, args, cfg = port_loss_fn(g)
    out, _ = sp.loss_and_grads(fn, g.group("w0"), *args)
    assert abs(float(out["loss"]) - indep) <= 2e-6 * abs(indep)


def test_port_is_separable_over_the_batch():
    """The SVI loss is a SUM over samples (plate "data"): evaluating the port on chunks of the
    batch and adding losses / gradients equals one evaluation -- the GPU parity tests at the
    benchmark shapes (tests/test_gpu_bench_shapes.py) rely on this to bound the oracle's memory."""
    from benchlib import chunked_oracle
    g = Golden("jivae_28_r")
    fn, args, cfg = port_loss_fn(g)
    sd = g.group("w0")
    out, grads = sp.loss_and_grads(fn, sd, *args)
    x, _ = g.args()
    out_c, grads_c = chunked_oracle("jivae", sd, cfg, (x,), g.eps(), tuple(g.kw()["scale_factor"]),
                                    chunk=3)
    assert abs(float(out["loss"]) - out_c["loss"]) <= 1e-6 * abs(out_c["loss"])
    assert torch.allclose(out["loc"].reshape(3, 8, -1), out_c["loc"], atol=1e-6)
    for k, v in grads.items():
        assert torch.allclose(v, grads_c[k], atol=1e-4 * v.abs().max().item() + 1e-7), k
