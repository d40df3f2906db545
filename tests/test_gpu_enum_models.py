"""GPU parity tests for the enumerated models (jiVAE, ssiVAE + auxSVItrainer)
against golden vectors of the unmodified reference run under oracle/pyro_min.
NOTE: the TraceEnum_ELBO expectation is restated from Pyro's documentation in
the oracle ("parity unpinned" against real Pyro)."""
import pytest
import torch

import pyroved_b200 as pv
from golden_util import Golden
from parity_util import FP32_GRAD_TOL, TC_GRAD_TOL, check_w1, grad_check

pytestmark = pytest.mark.gpu
LOSS_RTOL = 1e-3
LOC_ATOL = 1e-3


def check_grads(m, g, generic, tag):
    gref = g.group("grad")
    assert len(gref) > 4
    grad_check(m, gref, FP32_GRAD_TOL if generic else TC_GRAD_TOL,
               "{} {}".format(tag, "fp32" if generic else "tc"), allow_missing=True)


@pytest.mark.parametrize("generic", [True, False], ids=["fp32-generic", "default"])
@pytest.mark.parametrize("name", ["jivae_28_r", "jivae_28_r_beta1"])
def test_jivae_vs_reference_golden(name, generic):
    g = Golden(name)
    m = pv.models.jiVAE(seed=1, device="cuda:0", **g.kwargs)
    m.load_state_dict(g.group("w0"))
    tr = pv.trainers.SVItrainer(m, enumerate_parallel=True, device="cuda:0", force_generic=generic)
    x, _ = g.args()
    sf = [float(v) for v in g.kw()["scale_factor"]]
    loss = tr.svi.loss_and_grads(x.cuda(), _eps=g.eps().cuda(), scale_factor=sf)
    assert abs(loss - g.loss) <= LOSS_RTOL * abs(g.loss), (loss, g.loss)
    prog = next(iter(tr.svi.programs.values()))
    assert (prog.loc.cpu() - g.t("loc").reshape(-1)).abs().max().item() <= LOC_ATOL
    assert torch.allclose(prog.alpha.cpu(), g.t("alpha"), atol=1e-5)
    check_grads(m, g, generic, "golden " + name)
    # full step == reference SVI.step
    m.load_state_dict(g.group("w0"))
    loss = tr.svi.step(x.cuda(), _eps=g.eps().cuda(), scale_factor=sf)
    assert abs(loss - g.loss_step) <= LOSS_RTOL * abs(g.loss_step)
    check_w1(m, g, noise_floor=None if generic else TC_GRAD_TOL)


def test_jivae_requires_enumeration():
    m = pv.models.jiVAE((8, 8), 2, 3, ['r'], device="cuda:0")
    tr = pv.trainers.SVItrainer(m, device="cuda:0")
    with pytest.raises(ValueError):
        tr.svi.step(torch.rand(4, 8, 8).cuda())


@pytest.mark.parametrize("generic", [True, False], ids=["fp32-generic", "default"])
@pytest.mark.parametrize("name", ["ssivae_16_r_unsup", "ssivae_16_r_sup"])
def test_ssivae_vs_reference_golden(name, generic):
    g = Golden(name)
    m = pv.models.ssiVAE(seed=1, device="cuda:0", **g.kwargs)
    m.load_state_dict(g.group("w0"))
    tr = pv.trainers.auxSVItrainer(m, device="cuda:0", force_generic=generic)
    x, y = g.args()
    args = (x.cuda(),) if y is None else (x.cuda(), y.cuda())
    loss = tr.svi.loss_and_grads(*args, _eps=g.eps().cuda())
    assert abs(loss - g.loss) <= LOSS_RTOL * abs(g.loss), (loss, g.loss)
    prog = tr.svi.programs[(x.shape[0], y is not None, "main")]
    assert (prog.loc.cpu() - g.t("loc").reshape(-1)).abs().max().item() <= LOC_ATOL
    check_grads(m, g, generic, "golden " + name)
    # compute_loss = ELBO step + auxiliary step, two Adam updates (auxsvi.py:88-100)
    m.load_state_dict(g.group("w0"))
    tr = pv.trainers.auxSVItrainer(m, device="cuda:0", force_generic=generic)
    total = tr.compute_loss(x, y, _eps=g.eps().cuda(), aux_loss_multiplier=50.0)
    assert abs(total - g.loss_step) <= LOSS_RTOL * abs(g.loss_step), (total, g.loss_step)
    if generic:
        check_w1(m, g)     # two Adam updates: the sign-noise argument of check_w1 covers one


@pytest.mark.parametrize("inv", [None, ['r'], ['t'], ['r', 't', 's']])
def test_aux_trainer_runs_like_reference_tests(inv):
    """reference tests/test_trainers.py:56-73: no NaN, weights change."""
    torch.manual_seed(0)
    xu = torch.rand(10, 64)
    xs = xu + .1 * torch.rand_like(xu)
    labels = pv.utils.to_onehot(torch.randint(0, 3, (10,)), 3)
    lu, ls, lv = pv.utils.init_ssvae_dataloaders(xu, (xs, labels), (xs, labels), batch_size=2)
    vae = pv.models.ssiVAE((8, 8), 2, 3, inv, device="cuda:0")
    tr = pv.trainers.auxSVItrainer(vae, device="cuda:0")
    w0 = {k: v.clone() for k, v in vae.state_dict().items()}
    for _ in range(2):
        tr.step(lu, ls, lv)
        tr.save_running_weights("encoder_y")
    assert all(v == v for v in tr.history["training_loss"])
    assert len(tr.history["test"]) == 2
    assert any(not torch.equal(w0[k], v) for k, v in vae.state_dict().items())
    tr.average_weights("encoder_y")
    tr.print_statistics()


@pytest.mark.parametrize("inv", [None, ['r'], ['r', 't', 's']])
def test_jivae_trainer_and_inference_api(inv):
    torch.manual_seed(0)
    x = torch.rand(6, 8, 8)
    loader = pv.utils.init_dataloader(x, batch_size=2)
    vae = pv.models.jiVAE((8, 8), 2, 3, inv, device="cuda:0")
    tr = pv.trainers.SVItrainer(vae, enumerate_parallel=True, device="cuda:0")
    for _ in range(2):
        tr.step(loader, scale_factor=[2., 3.])
    assert all(v == v for v in tr.loss_history["training_loss"])
    zl, zs, cls = vae.encode(x)
    assert zl.shape == (6, vae.z_dim) and zs.shape == (6, vae.z_dim) and cls.shape == (6,)
    _, _, probs = vae.encode(x, logits=True)
    assert probs.shape == (6, 3) and torch.allclose(probs.sum(1), torch.ones(6), atol=1e-5)
    man = vae.manifold2d(3, disc_idx=1, plot=False)
    assert man.shape == (9, 8, 8)
    trav = vae.manifold_traversal(3, 0, plot=False)
    assert trav.shape == (9, 8, 8)


# ---- ss_reg_iVAE (regression variant, Trace_ELBO with a reparameterised y) ---------------------
@pytest.mark.parametrize("generic", [True, False], ids=["fp32-generic", "default"])
@pytest.mark.parametrize("name", ["ssreg_16_rt_unsup", "ssreg_16_rt_sup"])
def test_ss_reg_ivae_vs_reference_golden(name, generic):
    g = Golden(name)
    m = pv.models.ss_reg_iVAE(seed=1, device="cuda:0", **g.kwargs)
    m.load_state_dict(g.group("w0"))
    tr = pv.trainers.auxSVItrainer(m, task="regression", device="cuda:0", force_generic=generic)
    x, y = g.args()
    args = (x.cuda(),) if y is None else (x.cuda(), y.cuda())
    e = g.eps()
    eps = {k: v.cuda() for k, v in e.items()} if isinstance(e, dict) else e.cuda()
    kw = {k: float(v) for k, v in g.kw().items()}
    loss = tr.svi.loss_and_grads(*args, _eps=eps, **kw)
    assert abs(loss - g.loss) <= LOSS_RTOL * abs(g.loss), (loss, g.loss)
    prog = next(iter(tr.svi.programs.values()))
    assert (prog.loc.cpu() - g.t("loc").reshape(-1)).abs().max().item() <= LOC_ATOL
    assert torch.allclose(prog.mu.cpu(), g.t("mu"), atol=1e-4)
    check_grads(m, g, generic, "golden " + name)
    if y is not None and generic:
        # compute_loss = ELBO step + auxiliary regression step (two Adam updates)
        m.load_state_dict(g.group("w0"))
        tr2 = pv.trainers.auxSVItrainer(m, task="regression", device="cuda:0", force_generic=True)
        tr2.svi.step(x.cuda(), y.cuda(), _eps=eps, **kw)
        tr2.svi.step_aux(x.cuda(), y.cuda(), **kw)
        check_w1(m, g)


def test_ss_reg_trainer_loop_and_inference_api():
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(1)
    xu = (torch.rand(96, 16, 16, generator=gen) < 0.3).float().flatten(1)
    xs = (torch.rand(32, 16, 16, generator=gen) < 0.3).float().flatten(1)
    ys = xs.mean(1, keepdim=True) * 4 - 1
    m = pv.models.ss_reg_iVAE((16, 16), latent_dim=2, reg_dim=1, invariances=['r'], seed=1,
                              device="cuda:0")
    tr = pv.trainers.auxSVItrainer(m, task="regression", device="cuda:0")
    lu, ls, lv = pv.utils.init_ssvae_dataloaders(xu, (xs, ys), (xs, ys), batch_size=16)
    for _ in range(3):
        tr.step(lu, ls, lv, aux_loss_multiplier=20)
    assert len(tr.history["training_loss"]) == 3 and all(v == v for v in tr.history["training_loss"])
    assert len(tr.history["test"]) == 3 and tr.history["test"][-1] >= 0
    tr.print_statistics()
    pred = m.regressor(xs)
    assert pred.shape == (32, 1)
    zm, zs, yy = m.encode(xs)
    assert zm.shape == (32, 3) and zs.shape == (32, 3) and yy.shape == (32, 1)
    img = m.decode(torch.randn(4, 2), torch.zeros(4, 1))
    assert img.shape == (4, 16, 16)
    assert m.manifold2d(3, torch.zeros(1), plot=False).shape == (9, 16, 16)
