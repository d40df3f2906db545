"""Generate golden vectors by running the UNMODIFIED reference
(/root/reference/pyroved) on CPU under oracle/pyro_min, with epsilon injected.

Run in the authoring container only (the reference does not travel):
    python oracle/make_golden.py
Writes tests/golden/<case>.npz (inputs, weights with the reference's
state_dict keys, and outputs: loss, per-parameter gradients, reconstruction
`loc`, encoder mu/sigma and the weights after one SVI step with Adam).
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "pyro_min"))
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
import torch  # noqa: E402
import pyro  # noqa: E402
import pyro.poutine as poutine  # noqa: E402
from pyro.poutine import runtime  # noqa: E402
import pyroved as pv  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(4)


def blobs(n, h, w, seed=0, binary=True):
    """SURVEY 8(d) cfg2 synthetic data: rotated/shifted anisotropic blobs."""
    g = torch.Generator().manual_seed(seed)
    th = (torch.rand(n, generator=g) * 2 - 1) * np.pi / 3
    t = (torch.rand(n, 2, generator=g) * 2 - 1) * 0.1
    xx = torch.linspace(-1, 1, h)
    yy = torch.linspace(1, -1, w)
    gx, gy = torch.meshgrid(xx, yy, indexing="ij")
    gx = gx[None] - t[:, 0, None, None]
    gy = gy[None] - t[:, 1, None, None]
    c, s = torch.cos(th)[:, None, None], torch.sin(th)[:, None, None]
    u = c * gx + s * gy
    v = -s * gx + c * gy
    p = torch.exp(-(u ** 2 / (2 * 0.15 ** 2) + v ** 2 / (2 * 0.45 ** 2)))
    if binary:
        return (torch.rand(n, h, w, generator=g) < p).float()
    return p.float()


def spectra(n, length, seed=1):
    """examples/shiftVAE.ipynb cell 7 style: noisy shifted Gaussian peaks."""
    g = torch.Generator().manual_seed(seed)
    x = torch.linspace(-12, 12, length)
    mu = (torch.rand(n, generator=g) - 0.5) * 14
    sig = 1.0 + torch.rand(n, generator=g)
    y = torch.exp(-0.5 * ((x[None] - mu[:, None]) / sig[:, None]) ** 2)
    y = y + 0.05 * torch.randn(n, length, generator=g)
    y = (y - y.min()) / (y.max() - y.min())
    return y.float()


def run_case(name, model, trainer_kw, args, eps_by_site, step_kw, aux=False,
             model_fns=None):
    """One SVI step of the reference with injected eps; saves everything."""
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    enum = trainer_kw.get("enumerate_parallel", False)

    def hook(site, fn):
        e = eps_by_site.get(site)
        return e

    runtime.EPS_HOOK[0] = hook
    try:
        # 1. loss + grads without optimizer (fresh autograd)
        if aux:
            tr = pv.trainers.auxSVItrainer(model, **trainer_kw)
            svi = tr.loss_basic
        else:
            tr = pv.trainers.SVItrainer(model, **trainer_kw)
            svi = tr.svi
        model.load_state_dict(sd0)
        for p in model.parameters():
            p.grad = None
        loss = svi.loss.loss_and_grads(svi.model, svi.guide, *args, **step_kw)
        grads = {k: (p.grad.detach().clone() if p.grad is not None else None)
                 for k, p in model.named_parameters()}
        # reconstruction / encoder outputs from the traces
        with torch.no_grad():
            mt, gt = svi.loss._traces(svi.model, svi.guide, args, step_kw)
        extra = {}
        for n, site in mt.nodes.items():
            if site["type"] == "sample" and site["is_observed"] and n in ("obs", "x"):
                base = getattr(site["fn"], "base_dist", site["fn"])
                extra["loc"] = (base.probs if hasattr(base, "probs") and not hasattr(base, "loc")
                                else base.loc).detach().clone()
        for n, site in gt.nodes.items():
            if site["type"] == "sample":
                base = getattr(site["fn"], "base_dist", site["fn"])
                if hasattr(base, "loc") and hasattr(base, "scale"):
                    extra["mu"] = base.loc.detach().clone()
                    extra["sigma"] = base.scale.detach().clone()
                    extra["z"] = site["value"].detach().clone()
                elif hasattr(base, "probs"):
                    extra["alpha"] = base.probs.detach().clone()
        # 2. a real SVI step (loss_and_grads + Adam) from the same weights
        model.load_state_dict(sd0)
        for p in model.parameters():
            p.grad = None
        if aux:
            tr = pv.trainers.auxSVItrainer(model, **trainer_kw)
            model.load_state_dict(sd0)
            loss_step = tr.compute_loss(*args, **step_kw)
        else:
            tr = pv.trainers.SVItrainer(model, **trainer_kw)
            model.load_state_dict(sd0)
            loss_step = tr.svi.step(*args, **step_kw)
        sd1 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    finally:
        runtime.EPS_HOOK[0] = None
    out = {"loss": np.float64(loss), "loss_step": np.float64(loss_step)}
    for i, a in enumerate(args):
        if a is not None:
            out["arg{}".format(i)] = a.numpy()
    for k, v in eps_by_site.items():
        out["eps." + k] = v.numpy()
    for k, v in sd0.items():
        out["w0." + k] = v.numpy()
    # weights after one SVI step: small tensors whole, large ones as a fixed
    # random subsample (keeps the fixtures small)
    for k, v in sd1.items():
        flat = v.reshape(-1)
        if flat.numel() <= 4096:
            out["w1." + k] = v.numpy()
        else:
            idx = torch.randperm(flat.numel(), generator=torch.Generator().manual_seed(99))[:512]
            out["w1idx." + k] = idx.numpy()
            out["w1sub." + k] = flat[idx].numpy()
    for k, v in grads.items():
        if v is not None:
            out["grad." + k] = v.numpy()
    for k, v in extra.items():
        out[k] = v.numpy()
    for k, v in step_kw.items():
        out["kw." + k] = np.asarray(v, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("{:28s} loss {:.6f}  loss_step {:.6f}  ({} arrays)".format(
        name, loss, loss_step, len(out)))


def gen(seed):
    return torch.Generator().manual_seed(seed)


def main_ivae():
    dev = dict(device="cpu")
    # cfg1: 1-D shift-invariant iVAE (BASELINE configs[0], reduced batch)
    x = spectra(16, 64)[:, None, :]  # [B,1,64] as in the notebook
    m = pv.models.iVAE((64,), latent_dim=2, invariances=["t"], seed=1, **dev)
    run_case("ivae_1d_t", m, dev, (x,), {"latent": torch.randn(16, 3, generator=gen(1234))}, {})

    # cfg2 (reduced batch): 28x28 rot+trans Bernoulli
    x = blobs(16, 28, 28, seed=0)
    m = pv.models.iVAE((28, 28), latent_dim=2, invariances=["r", "t"], seed=1, **dev)
    eps = torch.randn(16, 5, generator=gen(1234))
    run_case("ivae_28_rt", m, dev, (x,), {"latent": eps}, {})
    m = pv.models.iVAE((28, 28), latent_dim=2, invariances=["r", "t"], seed=1, **dev)
    run_case("ivae_28_rt_beta3", m, dev, (x,), {"latent": eps}, {"scale_factor": 3.0})

    # r+t+s, class-conditional, gaussian sampler, relu (option coverage)
    x = blobs(8, 12, 12, seed=3, binary=False)[:, None]  # [B,1,12,12] (Concat quirk)
    y = pv.utils.to_onehot(torch.randint(0, 3, (8,), generator=gen(5)), 3)
    m = pv.models.iVAE((12, 12), latent_dim=3, invariances=["r", "t", "s"], c_dim=3,
                       activation="relu", sampler_d="gaussian", sigmoid_d=False,
                       seed=2, sc_prior=0.2, dx_prior=0.15, dy_prior=0.05, **dev)
    run_case("ivae_12_rts_cond_gauss", m, dev, (x, y),
             {"latent": torch.randn(8, 7, generator=gen(7))}, {"scale_factor": 2.0})

    # vanilla VAE (no invariances -> fcDecoderNet)
    x = blobs(8, 12, 12, seed=4)
    m = pv.models.iVAE((12, 12), latent_dim=2, invariances=None, seed=3, **dev)
    run_case("ivae_12_vanilla", m, dev, (x,),
             {"latent": torch.randn(8, 2, generator=gen(8))}, {})

    # scale only, softplus activation, non-default hidden dims
    x = blobs(8, 16, 16, seed=6)
    m = pv.models.iVAE((16, 16), latent_dim=2, invariances=["s"], seed=4,
                       hidden_dim_e=[64, 32], hidden_dim_d=[64, 64, 64],
                       activation="softplus", **dev)
    run_case("ivae_16_s_softplus", m, dev, (x,),
             {"latent": torch.randn(8, 3, generator=gen(9))}, {})

    # cfg3 (reduced): jiVAE rot-invariant, enumerated discrete latent
    x = blobs(8, 28, 28, seed=10)
    m = pv.models.jiVAE((28, 28), latent_dim=2, discrete_dim=3, invariances=["r"], seed=1, **dev)
    run_case("jivae_28_r", m, dict(enumerate_parallel=True, **dev), (x,),
             {"latent_cont": torch.randn(8, 3, generator=gen(11))},
             {"scale_factor": [3.0, 2.0]})

    # cfg4 (reduced): ssiVAE, unsupervised (enumerated y) and supervised
    x = blobs(8, 16, 16, seed=12).flatten(1)  # [B,N]: Concat only flattens >=4-D (utils/nn.py:69)
    m = pv.models.ssiVAE((16, 16), latent_dim=2, num_classes=4, invariances=["r"], seed=1, **dev)
    run_case("ssivae_16_r_unsup", m, dev, (x, None),
             {"z": torch.randn(4, 8, 3, generator=gen(13))},
             {"aux_loss_multiplier": 50.0}, aux=True)
    ys = pv.utils.to_onehot(torch.randint(0, 4, (8,), generator=gen(14)), 4)
    m = pv.models.ssiVAE((16, 16), latent_dim=2, num_classes=4, invariances=["r"], seed=1, **dev)
    run_case("ssivae_16_r_sup", m, dev, (x, ys),
             {"z": torch.randn(8, 3, generator=gen(15))},
             {"aux_loss_multiplier": 50.0}, aux=True)



def main_ved():
    """cfg5 (reduced channels / sizes so the fixtures stay small): VED, reference
    models/ved.py:122-163, nets/conv.py."""
    dev = dict(device="cpu")
    # image -> spectrum: conv2d encoder (two max-pools), conv1d decoder (nearest upsampling)
    x = blobs(6, 32, 32, seed=16, binary=False)[:, None]
    ysp = spectra(6, 64, seed=17)[:, None]
    m = pv.models.VED((32, 32), (64,), latent_dim=2, seed=1,
                      hidden_dim_e=[(8,), (16, 16), (32, 32)],
                      hidden_dim_d=[(32, 32), (16, 16), (8,)])
    m.to("cpu")
    run_case("ved_im2spec_32_64", m, dev, (x, ysp),
             {"z": torch.randn(6, 2, generator=gen(18))}, {"scale_factor": 4.0})
    # spectrum -> image: conv1d encoder, conv2d decoder (bilinear upsampling), tanh, gaussian
    xs = spectra(5, 32, seed=19)[:, None]
    yim = blobs(5, 16, 16, seed=20, binary=False)[:, None]
    m = pv.models.VED((32,), (16, 16), latent_dim=3, seed=2, activation="tanh",
                      sampler_d="gaussian", sigmoid_d=False, decoder_sig=0.3,
                      hidden_dim_e=[(8,), (16, 16)], hidden_dim_d=[(16, 16), (8,)])
    m.to("cpu")
    run_case("ved_spec2im_32_16", m, dev, (xs, yim),
             {"z": torch.randn(5, 3, generator=gen(21))}, {})
    # batchnorm=True (nets/conv.py:186-187, 239-240): BatchNorm2d encoder / BatchNorm1d decoder and
    # the other way round; w1.* then also pins the running statistics after one step
    x = blobs(6, 16, 16, seed=25, binary=False)[:, None]
    ysp = spectra(6, 32, seed=26)[:, None]
    m = pv.models.VED((16, 16), (32,), latent_dim=2, seed=3, batchnorm=True,
                      hidden_dim_e=[(8,), (16, 16)], hidden_dim_d=[(16, 16), (8,)])
    m.to("cpu")
    run_case("ved_bn_im2spec_16_32", m, dev, (x, ysp),
             {"z": torch.randn(6, 2, generator=gen(27))}, {"scale_factor": 2.0})
    xs = spectra(5, 32, seed=28)[:, None]
    yim = blobs(5, 16, 16, seed=29, binary=False)[:, None]
    m = pv.models.VED((32,), (16, 16), latent_dim=2, seed=4, batchnorm=True, activation="tanh",
                      hidden_dim_e=[(8,), (16, 16)], hidden_dim_d=[(16, 16), (8,)])
    m.to("cpu")
    run_case("ved_bn_spec2im_32_16", m, dev, (xs, yim),
             {"z": torch.randn(5, 2, generator=gen(30))}, {})
    # volumetric data: Conv3d / MaxPool3d / nearest x2 up-sampling (nets/conv.py with ndim = 3)
    g3 = gen(32)
    xv = torch.rand(4, 1, 8, 8, 8, generator=g3)
    yv = (torch.rand(4, 1, 8, 8, 8, generator=g3) < 0.4).float()
    m = pv.models.VED((8, 8, 8), (8, 8, 8), latent_dim=2, seed=5,
                      hidden_dim_e=[(4,), (8, 8)], hidden_dim_d=[(8, 8), (4,)])
    m.to("cpu")
    run_case("ved_vol_8", m, dev, (xv, yv), {"z": torch.randn(4, 2, generator=gen(33))}, {})
    # default architecture, two input channels, small spatial size (weights subsampled? no:
    # kept whole, the fixture is ~2 MB) -- skipped by default, enable with --ved-full
    if "--ved-full" in sys.argv:
        x = blobs(2, 16, 16, seed=22, binary=False)[:, None]
        ysp = spectra(2, 32, seed=23)[:, None]
        m = pv.models.VED((16, 16), (32,), latent_dim=2, seed=1)
        m.to("cpu")
        run_case("ved_default_16_32", m, dev, (x, ysp),
                 {"z": torch.randn(2, 2, generator=gen(24))}, {})


def main_cb():
    """continuous_bernoulli sampler (utils/prob.py:27) on real-valued targets in [0, 1]"""
    dev = dict(device="cpu")
    x = blobs(8, 12, 12, seed=30, binary=False).clamp(0.02, 0.98)
    x[0, 0, :6] = 0.5            # exercises the Taylor branch of the log-normaliser
    m = pv.models.iVAE((12, 12), latent_dim=2, invariances=["r"], seed=5,
                       sampler_d="continuous_bernoulli", **dev)
    run_case("ivae_12_r_cbern", m, dev, (x,),
             {"latent": torch.randn(8, 3, generator=gen(31))}, {})


def main_ssreg():
    """ss_reg_iVAE (regression variant, Trace_ELBO): unsupervised (sampled y) and supervised"""
    dev = dict(device="cpu")
    tkw = dict(task="regression", **dev)
    x = blobs(8, 16, 16, seed=40).flatten(1)
    m = pv.models.ss_reg_iVAE((16, 16), latent_dim=2, reg_dim=2, invariances=["r", "t"], seed=1, **dev)
    run_case("ssreg_16_rt_unsup", m, tkw, (x, None),
             {"z": torch.randn(8, 5, generator=gen(41)), "y": torch.randn(8, 2, generator=gen(42))},
             {"aux_loss_multiplier": 30.0, "scale_factor": 2.0}, aux=True)
    ys = torch.randn(8, 2, generator=gen(43)) * 0.7
    m = pv.models.ss_reg_iVAE((16, 16), latent_dim=2, reg_dim=2, invariances=["r", "t"], seed=1, **dev)
    run_case("ssreg_16_rt_sup", m, tkw, (x, ys),
             {"z": torch.randn(8, 5, generator=gen(44))},
             {"aux_loss_multiplier": 30.0, "scale_factor": 2.0}, aux=True)


def _trace_elbo_loss(model, guide, args, kw, eps_by_site):
    """-ELBO of (model, guide) under the restated Trace_ELBO (the pinned estimator) with eps injected."""
    runtime.EPS_HOOK[0] = lambda site, fn: eps_by_site.get(site)
    try:
        with torch.no_grad():
            return float(pyro.infer.Trace_ELBO().differentiable_loss(model, guide, *args, **kw))
    finally:
        runtime.EPS_HOOK[0] = None


def main_enum():
    """Independent routes to the ENUMERATED ELBOs (jiVAE, unsupervised ssiVAE) that do not go through
    the restated TraceEnum_ELBO (oracle/pyro_min/pyro/infer): only Trace_ELBO runs of the unmodified
    reference (pinned estimator) and forward calls of the reference's own nets are used.

      ssiVAE, unsupervised:  loss = -sum_b sum_k a_bk [ ELBO_sup(x_b, y = k; eps_kb) - log a_bk ],
          ELBO_sup = supervised Trace_ELBO of the reference on the single sample (x_b, onehot_k)
          (models/ssivae.py:153-215; it already contains log p(y) = log 1/K), a = encoder_y(x).
      jiVAE:  loss = -sum_b [ b0 (log p(z_b) - log q(z_b)) + sum_k a_bk ( ll_kb + b1 (log 1/K - log a_bk) ) ]
          with every term evaluated by calling the reference's encoder / decoder modules and
          torch.distributions directly (models/jivae.py:152-220).  Dice weights are the UNSCALED
          a_bk (Pyro builds them from score_parts.score_function, which poutine.scale does not
          touch); with b1 = 1 that assumption is immaterial (golden jivae_28_r_beta1).
    Writes tests/golden/jivae_28_r_beta1.npz and tests/golden/enum_indep.npz."""
    import math
    import torch.distributions as td
    from pyroved.utils import transform_coordinates
    dev = dict(device="cpu")
    out = {}
    # ---- jiVAE: the existing golden's inputs, and a beta1 = 1 variant ----------------------------
    x = blobs(8, 28, 28, seed=10)
    eps = torch.randn(8, 3, generator=gen(11))
    for tag, sf in (("jivae_28_r", [3.0, 2.0]), ("jivae_28_r_beta1", [3.0, 1.0])):
        m = pv.models.jiVAE((28, 28), latent_dim=2, discrete_dim=3, invariances=["r"], seed=1, **dev)
        if tag.endswith("beta1"):
            run_case(tag, m, dict(enumerate_parallel=True, **dev), (x,), {"latent_cont": eps},
                     {"scale_factor": sf})
            m = pv.models.jiVAE((28, 28), latent_dim=2, discrete_dim=3, invariances=["r"], seed=1,
                                **dev)
        K, B = 3, 8
        with torch.no_grad():
            mu, sig, alpha = m.encoder_z(x)
            z = mu + sig * eps
            kl = (td.Normal(0., 1.).log_prob(z) - td.Normal(mu, sig).log_prob(z)).sum(-1)   # [B]
            tot = sf[0] * kl.sum()
            for k in range(K):
                onehot = torch.zeros(B, K)
                onehot[:, k] = 1.
                phi, dx, sc, zc = m.split_latent(z)
                grid = m.grid.expand(B, *m.grid.shape)
                xc = transform_coordinates(grid, phi, dx, sc)
                loc = m.decoder(xc, [zc, onehot]).reshape(B, -1)
                ll = td.Bernoulli(probs=loc, validate_args=False).log_prob(x.reshape(B, -1)).sum(-1)
                a = alpha[:, k]
                tot = tot + (a * (ll + sf[1] * (math.log(1.0 / K) - torch.log(a)))).sum()
        out[tag] = np.float64(-float(tot))
    # ---- ssiVAE unsupervised: K*B supervised single-sample Trace_ELBO runs of the reference -------
    x = blobs(8, 16, 16, seed=12).flatten(1)
    eps = torch.randn(4, 8, 3, generator=gen(13))
    m = pv.models.ssiVAE((16, 16), latent_dim=2, num_classes=4, invariances=["r"], seed=1, **dev)
    K, B = 4, 8
    with torch.no_grad():
        alpha = m.encoder_y(x)
    tot = 0.0
    for b in range(B):
        for k in range(K):
            y1 = torch.zeros(1, K)
            y1[0, k] = 1.
            elbo_sup = -_trace_elbo_loss(m.model, m.guide, (x[b:b + 1], y1), {},
                                         {"z": eps[k, b:b + 1]})
            a = float(alpha[b, k])
            tot += a * (elbo_sup - math.log(a))
    out["ssivae_16_r_unsup"] = np.float64(-tot)
    np.savez_compressed(os.path.join(OUT, "enum_indep.npz"), **out)
    for k, v in out.items():
        ref = float(np.load(os.path.join(OUT, k + ".npz"))["loss"])
        print("{:24s} independent route {:.6f}   TraceEnum golden {:.6f}   rel diff {:.2e}".format(
            k, float(v), ref, abs(float(v) - ref) / abs(ref)))
        assert abs(float(v) - ref) <= 2e-6 * abs(ref)


def main_ved_eval():
    """batchnorm=True VED in eval mode (reference models/ved.py:178,193,230: encode / decode /
    manifold2d call self.eval(), so BatchNorm normalises with the RUNNING statistics): one training
    step first so the running statistics are not the initial (0, 1), then the inference calls."""
    dev = dict(device="cpu")
    x = blobs(6, 16, 16, seed=25, binary=False)[:, None]
    ysp = spectra(6, 32, seed=26)[:, None]
    m = pv.models.VED((16, 16), (32,), latent_dim=2, seed=3, batchnorm=True,
                      hidden_dim_e=[(8,), (16, 16)], hidden_dim_d=[(16, 16), (8,)])
    m.to("cpu")
    runtime.EPS_HOOK[0] = lambda site, fn: {"z": torch.randn(6, 2, generator=gen(27))}.get(site)
    try:
        tr = pv.trainers.SVItrainer(m, **dev)
        for _ in range(3):
            tr.svi.step(x, ysp, scale_factor=2.0)
    finally:
        runtime.EPS_HOOK[0] = None
    out = {"w." + k: v.detach().clone().numpy() for k, v in m.state_dict().items()}
    xn = blobs(3, 16, 16, seed=35, binary=False)[:, None]
    zn = torch.randn(4, 2, generator=gen(36))
    mu, sd = m.encode(xn)
    out.update(x=xn.numpy(), z=zn.numpy(), mu=mu.numpy(), sigma=sd.numpy(),
               dec=m.decode(zn).numpy(), man=m.manifold2d(3, plot=False).numpy())
    assert not m.training
    np.savez_compressed(os.path.join(OUT, "ved_bn_eval_16_32.npz"), **out)
    print("ved_bn_eval_16_32: mu", mu[0].tolist(), "dec mean", float(out["dec"].mean()))


def main():
    if "--enum" in sys.argv:
        main_enum()
        return
    if "--ved-eval" in sys.argv:
        main_ved_eval()
        return
    if "--ssreg" in sys.argv:
        main_ssreg()
        return
    if "--cb" in sys.argv:
        main_cb()
        return
    if "--ved" in sys.argv or "--ved-full" in sys.argv:
        main_ved()
        if "--all" not in sys.argv:
            return
    main_ivae()


if __name__ == "__main__":
    main()
