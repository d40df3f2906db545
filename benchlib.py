"""Benchmark workloads of BASELINE.json `configs`: the model / trainer each one
names, synthetic data of its shape, its algorithmic work, and the CPU oracle's evaluation of the
same step.  Shared by bench.py (measurement) and tests/test_gpu_bench_shapes.py (parity at the
benchmarked shapes) so both look at exactly the same workload.

  cfg2  iVAE 2-D rot+trans, 28x28 Bernoulli, latent_dim=2, batch 512 per GPU          (headline)
  cfg3  jiVAE 28x28, 10 classes + latent_dim=2, rot-invariant, batch 1024, scale_factor=[3,3]
  cfg4  ssiVAE 64x64, 4 classes, rot-invariant, batch 256 per GPU (2048 over 8 GPUs), one labelled
        batch per 19 unlabelled ones, auxSVItrainer(aux_loss_multiplier=50)
  cfg5  VED 64x64 image -> 128-point spectrum, default filters, batch 512 per GPU (4096 over 8),
        scale_factor=4

Only bench.py's cpu_baseline / --impl reference legs and the tests call the oracle functions here.
"""
import math
from collections import OrderedDict

import torch

WORKLOADS = OrderedDict([
    ("cfg2", dict(
        workload="iVAE 2D rot+trans, 28x28 Bernoulli, latent_dim=2, fc enc / spatial fc dec, "
                 "batch=512 per GPU (BASELINE configs[1])",
        kind="ivae", data_dim=(28, 28), batch=512, step_kw={}, lr=1e-3,
        model_kw=dict(latent_dim=2, invariances=['r', 't']))),
    ("cfg3", dict(
        workload="jiVAE 28x28, 10 classes + latent_dim=2, rot-invariant, enumerated, "
                 "scale_factor=[3,3], batch=1024 per GPU (BASELINE configs[2])",
        kind="jivae", data_dim=(28, 28), batch=1024, step_kw={"scale_factor": [3.0, 3.0]}, lr=1e-3,
        model_kw=dict(latent_dim=2, discrete_dim=10, invariances=['r']))),
    ("cfg4", dict(
        workload="ssiVAE 64x64, 4 classes, rot-invariant, auxSVItrainer, aux_loss_multiplier=50, one "
                 "labelled batch per 19 unlabelled, batch=256 per GPU = 2048 over 8 GPUs "
                 "(BASELINE configs[3])",
        kind="ssivae", data_dim=(64, 64), batch=256, step_kw={"aux_loss_multiplier": 50.0}, lr=5e-4,
        model_kw=dict(latent_dim=2, num_classes=4, invariances=['r']))),
    ("cfg5", dict(
        workload="VED im2spec 64x64 -> 1x128, default conv filters, scale_factor=4, batch=512 per "
                 "GPU = 4096 over 8 GPUs (BASELINE configs[4])",
        kind="ved", data_dim=(64, 64), batch=512, step_kw={"scale_factor": 4.0}, lr=1e-3,
        model_kw=dict(latent_dim=2))),
])

# ---- algorithmic work ---------------------------------------------------------------
FLOP_PER_ROW_FWD = 66304            # spatial decoder: 2*(2*128 + 2*128*128 + 128) per pixel row
FLOP_PER_ROW_STEP = 3 * FLOP_PER_ROW_FWD     # backward = 2 x forward


def decoder_rows_per_sample(name):
    """Pixel rows through the spatial decoder per sample and step (K-fold under enumeration)."""
    w = WORKLOADS[name]
    n = w["data_dim"][0] * w["data_dim"][1]
    if name == "cfg3":
        return 10 * n
    if name == "cfg4":
        # 19 of 20 batches are unlabelled (4 enumerated classes), 1 of 20 labelled (1 instance)
        return (19 * 4 + 1) * n / 20.0
    if name == "cfg2":
        return n
    return 0


def flop_per_sample_step(name):
    """Algorithmic FLOPs of one SVI step per sample (forward + backward = 3 x forward)."""
    if name == "cfg5":
        # 3x3 convolutions of the default encoder (32@64x64, 64,64@32x32, 128,128@16x16) and the
        # 1-D decoder (128,128@16, 64,64@32, 32@64, 1x1 convs), features2latent / latent2features
        enc = 2 * 9 * (1 * 32 * 4096 + 32 * 64 * 1024 + 64 * 64 * 1024 + 64 * 128 * 256 + 128 * 128 * 256)
        dec = 2 * 3 * (128 * 128 * 16 * 2 + 128 * 64 * 32 + 64 * 64 * 32 + 64 * 32 * 64) + \
            2 * (128 * 128 * 32 + 64 * 64 * 64 + 32 * 32 * 128 + 32 * 128)
        fc = 2 * (32768 * 4 + 2 * 2048)
        return 3.0 * (enc + dec + fc)
    w = WORKLOADS[name]
    n = w["data_dim"][0] * w["data_dim"][1]
    enc = 2.0 * (n * 128 + 128 * 128 + 2 * 128 * 8)
    if name == "cfg4":
        enc = enc * (19 * 4 + 1) / 20.0 + 2.0 * 2 * (n * 128 + 128 * 128)   # + classifier (x2: aux step)
    return 3.0 * (decoder_rows_per_sample(name) * FLOP_PER_ROW_FWD + enc)


# ---- synthetic data ------------------------------------------------------------------------------
def blob_images(n, h, w, n_classes=1, seed=0, binary=True):
    """Rotated / shifted anisotropic Gaussian blobs; class c changes the blob's
    two widths.  Returns (x [n,h,w], cls [n], theta [n])."""
    g = torch.Generator().manual_seed(seed)
    th = (torch.rand(n, generator=g) * 2 - 1) * math.pi / 3
    t = (torch.rand(n, 2, generator=g) * 2 - 1) * 0.1
    xx = torch.linspace(-1, 1, h)
    yy = torch.linspace(1, -1, w)
    gx, gy = torch.meshgrid(xx, yy, indexing="ij")
    gx = gx[None] - t[:, 0, None, None]
    gy = gy[None] - t[:, 1, None, None]
    c, s = torch.cos(th)[:, None, None], torch.sin(th)[:, None, None]
    u = c * gx + s * gy
    v = -s * gx + c * gy
    if n_classes > 1:
        cls = torch.randint(0, n_classes, (n,), generator=g)
        su = (0.10 + 0.25 * cls.float() / n_classes)[:, None, None]
        sv = (0.55 - 0.30 * cls.float() / n_classes)[:, None, None]
    else:
        cls = torch.zeros(n, dtype=torch.long)
        su, sv = 0.15, 0.45
    p = torch.exp(-(u ** 2 / (2 * su ** 2) + v ** 2 / (2 * sv ** 2)))
    x = (torch.rand(n, h, w, generator=g) < p).float() if binary else p.float()
    return x, cls, th


def synth(name, n, seed=0, labelled=False):
    """n samples of workload `name` as the tuple a loader of the reference yields:
    cfg2 / cfg3: (x [n,28,28],); cfg4: (x [n,4096],) or (x, onehot y [n,4]) when labelled;
    cfg5: (x [n,1,64,64], y [n,1,128])."""
    if name == "cfg2":
        return (blob_images(n, 28, 28, 1, seed)[0],)
    if name == "cfg3":
        return (blob_images(n, 28, 28, 10, seed)[0],)
    if name == "cfg4":
        x, cls, _ = blob_images(n, 64, 64, 4, seed)
        x = x.reshape(n, -1)      # [B,N]: the reference's Concat only flattens >= 4-D inputs
        if labelled:
            return x, torch.nn.functional.one_hot(cls, 4).float()
        return (x,)
    if name == "cfg5":
        x, cls, th = blob_images(n, 64, 64, 4, seed, binary=False)
        j = torch.arange(128, dtype=torch.float32)[None]
        mu = (64 + 40 * th / (math.pi / 3))[:, None]
        sg = (4 + 2 * cls.float())[:, None]
        y = torch.exp(-0.5 * ((j - mu) / sg) ** 2)
        return x[:, None].contiguous(), y[:, None].contiguous()
    raise KeyError(name)


# ---- the product side ----------------------------------------------------------------------------
def build(name, device, **trainer_kw):
    """(model, trainer) of workload `name` from the drop-in classes."""
    import pyroved_b200 as pv
    w = WORKLOADS[name]
    kw = dict(w["model_kw"])
    if w["kind"] == "ivae":
        m = pv.models.iVAE(w["data_dim"], seed=1, device=device, **kw)
        tr = pv.trainers.SVItrainer(m, seed=1, device=device, **trainer_kw)
    elif w["kind"] == "jivae":
        m = pv.models.jiVAE(w["data_dim"], seed=1, device=device, **kw)
        tr = pv.trainers.SVItrainer(m, enumerate_parallel=True, seed=1, device=device, **trainer_kw)
    elif w["kind"] == "ssivae":
        m = pv.models.ssiVAE(w["data_dim"], seed=1, device=device, **kw)
        tr = pv.trainers.auxSVItrainer(m, seed=1, device=device, **trainer_kw)
    else:
        m = pv.models.VED(w["data_dim"], (128,), seed=1, device=device, **kw)
        tr = pv.trainers.SVItrainer(m, seed=1, device=device, **trainer_kw)
    return m, tr


# ---- the oracle side (tests, bench cpu_baseline / --impl reference only) ------------------------------
def oracle_cfg(name):
    from oracle import svi_port as sp
    w = WORKLOADS[name]
    kw = w["model_kw"]
    if w["kind"] == "ved":
        return sp.VedCfg(w["data_dim"], (128,), kw["latent_dim"])
    return sp.Cfg(w["data_dim"], kw["latent_dim"], kw["invariances"],
                  discrete_dim=kw.get("discrete_dim", 0), num_classes=kw.get("num_classes", 0))


def chunked_oracle(kind, sd, cfg, args, eps, beta, chunk=64):
    """The oracle port's loss / reconstruction / gradients for one batch, evaluated on chunks of
    `chunk` samples and summed (the SVI loss is a sum over the samples of plate "data", so this is
    exact; it bounds the CPU memory of the enumerated models at the benchmark shapes).
    kind: ivae | jivae | ssivae | ssivae_aux | ved.  eps: [B,Z], or [K,B,Z] for unsupervised ssiVAE.
    Returns ({"loss": float, "loc": tensor, ...}, grads)."""
    from oracle import svi_port as sp
    x = args[0]
    y = args[1] if len(args) > 1 else None
    B = x.shape[0]
    total, locs, grads = 0.0, [], None
    mus, sigs, alphas = [], [], []
    for lo in range(0, B, chunk):
        sl = slice(lo, min(B, lo + chunk))
        ys = y[sl] if y is not None else None
        if kind == "ivae":
            out, g = sp.loss_and_grads(sp.ivae_loss, sd, cfg, x[sl], eps[sl], ys, beta)
        elif kind == "jivae":
            out, g = sp.loss_and_grads(sp.jivae_loss, sd, cfg, x[sl], eps[sl], beta)
        elif kind == "ssivae":
            e = eps[:, sl] if ys is None else eps[sl]
            out, g = sp.loss_and_grads(sp.ssivae_loss, sd, cfg, x[sl], e, ys, beta)
        elif kind == "ssivae_aux":
            out, g = sp.loss_and_grads(sp.ssivae_aux_loss, sd, cfg, x[sl], ys, beta)
        elif kind == "ved":
            out, g = sp.loss_and_grads(sp.ved_loss, sd, cfg, x[sl], ys, eps[sl], beta)
        else:
            raise KeyError(kind)
        total += float(out["loss"])
        if "loc" in out:
            locs.append(out["loc"])
        if "mu" in out:
            mus.append(out["mu"])
            sigs.append(out["sigma"])
        if "alpha" in out:
            alphas.append(out["alpha"])
        if grads is None:
            grads = OrderedDict((k, (v.clone() if v is not None else None)) for k, v in g.items())
        else:
            for k, v in g.items():
                if v is not None:
                    grads[k] = v.clone() if grads[k] is None else grads[k].add_(v)
    res = {"loss": total}
    if locs:
        # enumerated models: loc is [K, b, N] per chunk -> concatenate along the batch dim
        res["loc"] = torch.cat(locs, dim=1 if locs[0].dim() == 3 else 0)
    if mus:
        d = 1 if mus[0].dim() == 3 else 0
        res["mu"], res["sigma"] = torch.cat(mus, d), torch.cat(sigs, d)
    if alphas:
        res["alpha"] = torch.cat(alphas, 0)
    return res, grads


def oracle_step_seconds(name, n_samples, sd=None, seed=0):
    """Wall time of the oracle port's loss + gradients (+ Adam for cfg2) on `n_samples` samples of
    workload `name` with all host threads: the bounded CPU sample of bench.py's cpu_baseline."""
    import time
    from oracle import svi_port as sp
    w = WORKLOADS[name]
    cfg = oracle_cfg(name)
    g = torch.Generator().manual_seed(seed)
    data = synth(name, n_samples, seed=seed + 17, labelled=False)
    if sd is None:
        raise ValueError("pass the model's state_dict")
    kind = w["kind"]
    beta = w["step_kw"].get("scale_factor", 1.0)
    if kind == "jivae":
        eps = torch.randn(n_samples, cfg.z_dim, generator=g)
        beta = tuple(beta)
    elif kind == "ssivae":
        eps = torch.randn(4, n_samples, cfg.z_dim, generator=g)
    elif kind == "ved":
        eps = torch.randn(n_samples, cfg.latent_dim, generator=g)
    else:
        eps = torch.randn(n_samples, cfg.z_dim, generator=g)
    chunk = {"jivae": 64, "ssivae": 16}.get(kind, n_samples)
    t0 = time.perf_counter()
    out, grads = chunked_oracle(kind, sd, cfg, data, eps, beta, chunk=chunk)
    if kind == "ssivae":     # the auxiliary step of auxSVItrainer.compute_loss (no labels: no terms)
        chunked_oracle("ssivae_aux", sd, cfg, data, None, 50.0, chunk=n_samples)
    sp.AdamState(lr=w["lr"]).step(dict(sd), grads)
    return time.perf_counter() - t0, out["loss"]
