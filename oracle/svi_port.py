"""svi_port -- CPU restatement ("port") of pyroVED's SVI hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `pyroved_b200/` may import this module;
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs use it, as the checker / the timed CPU baseline.

It restates, with plain torch CPU ops (fp32 or fp64), exactly the arithmetic
the reference executes for one `SVItrainer.step` mini-batch, following the
reference's own op order (materialised grid, bmm rotations, per-layer
Linear+activation), so that it is both the parity oracle and an honest CPU
baseline.  It does NOT need /root/reference or Pyro at run time (neither
exists on the GPU box).  Every function cites the reference lines it follows.

Parity pin: the reference holds NO golden vectors / known-answer tests for
this path (tests pin shapes only), so the port is pinned against
the UNMODIFIED reference executed in the authoring container under
`oracle/pyro_min` (a restatement of the Pyro calls; real pyro-ppl is not
installable offline) -- `oracle/make_golden.py` writes those outputs to
`tests/golden/*.npz` and `tests/test_oracle.py` checks this port against them.
For the enumerated models (jiVAE, unsupervised ssiVAE) the TraceEnum_ELBO
expectation itself is restated from Pyro's documentation: that part is
"parity unpinned" against real Pyro.

Weights are passed as a dict with the reference's `state_dict` key names
(SURVEY 8b), e.g. 'encoder_z.fc_layers.0.weight'.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

EPS_PROBS = torch.finfo(torch.float32).eps  # torch clamp_probs eps (fp32)


# --------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------
def activation_fn(name):
    """reference utils/nn.py:118-124 (torch module defaults)."""
    return {"tanh": torch.tanh, "relu": F.relu, "softplus": F.softplus,
            "lrelu": lambda t: F.leaky_relu(t, 0.01), "gelu": F.gelu}[name]


def fc_stack(sd, prefix, x, act):
    """make_fc_layers (nets/fc.py:307-324): Linear+act at even indices."""
    i = 0
    f = activation_fn(act)
    while "{}.fc_layers.{}.weight".format(prefix, i) in sd:
        x = f(F.linear(x, sd["{}.fc_layers.{}.weight".format(prefix, i)],
                       sd["{}.fc_layers.{}.bias".format(prefix, i)]))
        i += 2
    return x


def concat_bcast(args):
    """utils/nn.py:62-74 Concat with broadcasting."""
    if torch.is_tensor(args):
        return args
    args = [a.flatten(1) if a.ndim >= 4 else a for a in args]
    shape = torch.broadcast_shapes(*[a.shape[:-1] for a in args]) + (-1,)
    return torch.cat([a.expand(shape) for a in args], dim=-1)


def generate_grid(data_dim, dtype=torch.float32):
    """utils/coord.py:7-44."""
    if len(data_dim) == 1:
        return torch.linspace(1, -1, data_dim[0], dtype=dtype)[:, None]
    xx = torch.linspace(-1, 1, data_dim[0], dtype=dtype)
    yy = torch.linspace(1, -1, data_dim[1], dtype=dtype)
    x0, x1 = torch.meshgrid(xx, yy, indexing="ij")
    return torch.stack([x0.reshape(-1), x1.reshape(-1)], 1)


def transform_coordinates(coord, phi, dx, scale):
    """utils/coord.py:47-88 (rotate via bmm, scale via bmm, translate)."""
    if coord.shape[-1] == 1:
        return coord + dx
    b = coord.shape[0]
    if not torch.is_tensor(phi) or phi.ndim == 0:
        phi = coord.new_zeros(b) + phi
    if torch.sum(phi) == 0:  # coord.py:69-70 quirk
        phi = coord.new_zeros(b)
    r1 = torch.stack([torch.cos(phi), torch.sin(phi)], 1)
    r2 = torch.stack([-torch.sin(phi), torch.cos(phi)], 1)
    rot = torch.stack([r1, r2], dim=1)
    coord = torch.bmm(coord, rot)
    sm = coord.new_zeros(b, 2, 2)
    sm[:, 0, 0] = scale
    sm[:, 1, 1] = scale
    coord = torch.bmm(coord, sm)
    return coord + dx


def coord_count(ndim, invariances):
    """models/base.py:54-67."""
    if invariances is None:
        return 0
    c = len(invariances)
    if ndim == 1:
        if c > 1 or invariances[0] != "t":
            raise ValueError("For 1D data, the only invariance to enforce "
                             "is translation ('t')")
    if "t" in invariances and ndim == 2:
        c += 1
    return c


def split_latent(z, ndim, invariances, sc_prior):
    """models/base.py:97-119 (order r, t, s)."""
    if ndim == 1:
        return None, z[:, 0:1], None, z[:, 1:]
    phi = z.new_zeros(())
    dx = z.new_zeros(())
    sc = z.new_ones(())
    if "r" in invariances:
        phi, z = z[:, 0], z[:, 1:]
    if "t" in invariances:
        dx, z = z[:, :2], z[:, 2:]
    if "s" in invariances:
        sc = sc + sc_prior * z[:, 0]
        z = z[:, 1:]
    return phi, dx, sc, z


def log_lik(loc, x, sampler, decoder_sig=0.5):
    """utils/prob.py:25-29 + torch.distributions log_prob, summed over the
    event dim (to_event(1)).  loc/x: [..., N] -> [...]."""
    if sampler == "bernoulli":
        # Bernoulli(probs).log_prob = -BCEWithLogits(probs_to_logits(probs), x)
        eps = torch.finfo(loc.dtype).eps
        p = loc.clamp(eps, 1 - eps)
        logits = torch.log(p) - torch.log1p(-p)
        return -(F.binary_cross_entropy_with_logits(
            logits, x.expand_as(logits), reduction="none")).sum(-1)
    if sampler == "gaussian":
        var = decoder_sig ** 2
        return (-((x - loc) ** 2) / (2 * var) - math.log(decoder_sig)
                - 0.5 * math.log(2 * math.pi)).sum(-1)
    if sampler == "continuous_bernoulli":
        return torch.distributions.ContinuousBernoulli(
            probs=loc).log_prob(x.expand_as(loc)).sum(-1)
    raise KeyError(sampler)


def normal_logprob(z, loc, scale):
    return (-((z - loc) ** 2) / (2 * scale ** 2) - torch.log(scale)
            - 0.5 * math.log(2 * math.pi)).sum(-1)


# --------------------------------------------------------------------------
# nets
# --------------------------------------------------------------------------
def fc_encoder(sd, x, act, in_dim, flat=True, prefix="encoder_z"):
    """fcEncoderNet.forward (nets/fc.py:51-61)."""
    x = concat_bcast(x)
    if flat:
        x = x.reshape(-1, in_dim)
    h = fc_stack(sd, prefix, x, act)
    mu = F.linear(h, sd[prefix + ".fc11.weight"], sd[prefix + ".fc11.bias"])
    sig = F.softplus(F.linear(h, sd[prefix + ".fc12.weight"], sd[prefix + ".fc12.bias"]))
    return mu, sig, h


def s_decoder(sd, x_coord, z, act, sigmoid_out=True, prefix="decoder"):
    """sDecoderNet.forward + coord_latent.forward (nets/fc.py:189-237).
    Returns flat [B*N, 1] probabilities (or logits if not sigmoid_out)."""
    z = concat_bcast(z)
    b, n = x_coord.shape[:2]
    h_x = F.linear(x_coord.reshape(b * n, -1),
                   sd[prefix + ".coord_latent.fc_coord.weight"],
                   sd[prefix + ".coord_latent.fc_coord.bias"]).reshape(b, n, -1)
    h_z = F.linear(z, sd[prefix + ".coord_latent.fc_latent.weight"])
    h_z = h_z.reshape(-1, h_z.shape[-1])
    h = torch.tanh((h_x + h_z.unsqueeze(1)).reshape(b * n, -1))
    h = fc_stack(sd, prefix, h, act)
    out = F.linear(h, sd[prefix + ".out.weight"], sd[prefix + ".out.bias"])
    return torch.sigmoid(out) if sigmoid_out else out


def fc_decoder(sd, z, act, sigmoid_out=True, prefix="decoder"):
    """fcDecoderNet.forward (nets/fc.py:143-152), flat output."""
    z = concat_bcast(z)
    h = fc_stack(sd, prefix, z, act)
    out = F.linear(h, sd[prefix + ".out.weight"], sd[prefix + ".out.bias"])
    return torch.sigmoid(out) if sigmoid_out else out


# --------------------------------------------------------------------------
# model configuration
# --------------------------------------------------------------------------
class Cfg:
    def __init__(self, data_dim, latent_dim=2, invariances=None, c_dim=0,
                 activation="tanh", sampler_d="bernoulli", sigmoid_d=True,
                 dx_prior=0.1, dy_prior=None, sc_prior=0.1, decoder_sig=0.5,
                 discrete_dim=0, num_classes=0):
        self.data_dim = tuple(data_dim)
        self.ndim = len(self.data_dim)
        self.n_pix = int(math.prod(self.data_dim))
        self.latent_dim = latent_dim
        self.invariances = invariances
        self.coord = coord_count(self.ndim, invariances)
        self.z_dim = latent_dim + self.coord
        self.c_dim = c_dim
        self.activation = activation
        self.sampler_d = sampler_d
        self.sigmoid_d = sigmoid_d
        self.dx_prior = dx_prior
        self.dy_prior = dx_prior if dy_prior is None else dy_prior
        self.sc_prior = sc_prior
        self.decoder_sig = decoder_sig
        self.discrete_dim = discrete_dim
        self.num_classes = num_classes

    def t_prior(self, ref):
        if self.ndim == 2:
            return ref.new_tensor([self.dx_prior, self.dy_prior])
        return ref.new_tensor(self.dx_prior)


def _decode_spatial(sd, cfg, z, extra, grid):
    """shared tail of iVAE/jiVAE/ssiVAE.model: split, transform grid, decode.
    z: [I, z_dim]; extra: tensor or None to concatenate to the content code."""
    if cfg.coord > 0:
        phi, dx, sc, zc = split_latent(z, cfg.ndim, cfg.invariances, cfg.sc_prior)
        if "t" in cfg.invariances:
            dx = (dx * cfg.t_prior(z)).unsqueeze(1)
        g = grid.to(z.dtype).expand(z.shape[0], *grid.shape)
        xc = transform_coordinates(g, phi, dx, sc)
        zz = zc if extra is None else torch.cat([zc, extra], -1)
        return s_decoder(sd, xc, zz, cfg.activation, cfg.sigmoid_d)
    zz = z if extra is None else torch.cat([z, extra], -1)
    return fc_decoder(sd, zz, cfg.activation, cfg.sigmoid_d)


# --------------------------------------------------------------------------
# iVAE  (models/ivae.py:165-221 under Trace_ELBO)
# --------------------------------------------------------------------------
def ivae_loss(sd, cfg, x, eps, y=None, beta=1.0, conv_encoder=None):
    """Returns dict(loss, ll[B], kl_term[B], loc[B,N], z, mu, sigma).
    loss = -( sum_b ll_b + beta * sum_b (log p(z_b) - log q(z_b)) ).
    conv_encoder: a VedCfg describing a convEncoderNet installed with iVAE.set_encoder
    (models/base.py:173-176; its latent_dim is the model's full latent width)."""
    b = x.shape[0]
    grid = generate_grid(cfg.data_dim, x.dtype) if cfg.coord > 0 else None
    enc_in = [x, y] if y is not None else x
    if conv_encoder is not None:
        mu, sig = ved_encoder(sd, conv_encoder, x)
    else:
        mu, sig, _ = fc_encoder(sd, enc_in, cfg.activation, cfg.n_pix + cfg.c_dim)
    z = mu + sig * eps
    log_q = normal_logprob(z, mu, sig)
    log_p = normal_logprob(z, torch.zeros_like(z), torch.ones_like(z))
    loc = _decode_spatial(sd, cfg, z, y, grid).reshape(b, cfg.n_pix)
    ll = log_lik(loc, x.reshape(b, cfg.n_pix), cfg.sampler_d, cfg.decoder_sig)
    elbo = ll.sum() + beta * (log_p - log_q).sum()
    return {"loss": -elbo, "ll": ll, "kl_term": log_p - log_q, "loc": loc,
            "z": z, "mu": mu, "sigma": sig}


# --------------------------------------------------------------------------
# jiVAE  (models/jivae.py:152-220 under TraceEnum_ELBO)
# --------------------------------------------------------------------------
def _beta2(beta):
    if isinstance(beta, (float, int)):
        return float(beta), float(beta)
    beta = torch.as_tensor(beta)
    if beta.ndim == 0:
        return float(beta), float(beta)
    return float(beta[0]), float(beta[1])


def jivae_loss(sd, cfg, x, eps, beta=(1.0, 1.0)):
    b0, b1 = _beta2(beta)
    bsz, k = x.shape[0], cfg.discrete_dim
    grid = generate_grid(cfg.data_dim, x.dtype) if cfg.coord > 0 else None
    mu, sig, h = fc_encoder(sd, x, cfg.activation, cfg.n_pix)
    alpha = torch.softmax(F.linear(h, sd["encoder_z.fc13.weight"],
                                   sd["encoder_z.fc13.bias"]), -1)
    z = mu + sig * eps
    log_q = normal_logprob(z, mu, sig)
    log_p = normal_logprob(z, torch.zeros_like(z), torch.ones_like(z))
    onehot = torch.eye(k, dtype=x.dtype)[:, None, :].expand(k, bsz, k)
    if cfg.coord > 0:
        zr = z.repeat(k, 1)                               # jivae.py:182
        loc = _decode_spatial(sd, cfg, zr, onehot.reshape(-1, k), grid)
    else:
        loc = fc_decoder(sd, [z, onehot], cfg.activation, cfg.sigmoid_d)
    loc = loc.reshape(k, bsz, cfg.n_pix)
    ll = log_lik(loc, x.reshape(bsz, cfg.n_pix), cfg.sampler_d, cfg.decoder_sig)  # [K,B]
    log_alpha = torch.log(alpha.t())                      # [K,B] = log q(c=k)
    w = alpha.t()
    elbo = (b0 * (log_p - log_q)).sum() + (
        w * (ll + b1 * (math.log(1.0 / k) - log_alpha))).sum()
    return {"loss": -elbo, "ll": ll, "alpha": alpha, "loc": loc, "z": z,
            "mu": mu, "sigma": sig}


# --------------------------------------------------------------------------
# ssiVAE  (models/ssivae.py:153-248, trainers/auxsvi.py:88-100)
# --------------------------------------------------------------------------
def classifier(sd, cfg, x, prefix="encoder_y"):
    """fcClassifierNet.forward (nets/fc.py:264-271)."""
    h = fc_stack(sd, prefix, x.reshape(-1, cfg.n_pix), cfg.activation)
    return torch.softmax(F.linear(h, sd[prefix + ".out.weight"],
                                  sd[prefix + ".out.bias"]), -1)


def ssivae_loss(sd, cfg, xs, eps, ys=None, beta=1.0):
    """eps: [B,Z] (supervised) or [K,B,Z] (unsupervised, enumerated y)."""
    bsz, k = xs.shape[0], cfg.num_classes
    xf = xs.reshape(bsz, cfg.n_pix)
    grid = generate_grid(cfg.data_dim, xs.dtype) if cfg.coord > 0 else None
    if ys is None:
        alpha = classifier(sd, cfg, xs)                   # [B,K]
        y_en = torch.eye(k, dtype=xs.dtype)[:, None, :].expand(k, bsz, k)
        mu, sig, _ = fc_encoder(sd, [xf, y_en], cfg.activation, None, flat=False)
        z = mu + sig * eps                                # [K,B,Z]
        log_q = normal_logprob(z, mu, sig)
        log_p = normal_logprob(z, torch.zeros_like(z), torch.ones_like(z))
        loc = _decode_spatial(sd, cfg, z.reshape(k * bsz, -1),
                              y_en.reshape(k * bsz, k), grid).reshape(k, bsz, -1)
        ll = log_lik(loc, xf, cfg.sampler_d, cfg.decoder_sig)
        w = alpha.t()
        elbo = (w * (ll + beta * (log_p - log_q) + math.log(1.0 / k)
                     - torch.log(w))).sum()
        return {"loss": -elbo, "ll": ll, "alpha": alpha, "loc": loc, "z": z,
                "mu": mu, "sigma": sig}
    mu, sig, _ = fc_encoder(sd, [xf, ys], cfg.activation, None, flat=False)
    z = mu + sig * eps
    log_q = normal_logprob(z, mu, sig)
    log_p = normal_logprob(z, torch.zeros_like(z), torch.ones_like(z))
    loc = _decode_spatial(sd, cfg, z, ys, grid).reshape(bsz, -1)
    ll = log_lik(loc, xf, cfg.sampler_d, cfg.decoder_sig)
    elbo = (ll + beta * (log_p - log_q)).sum() + bsz * math.log(1.0 / k)
    return {"loss": -elbo, "ll": ll, "loc": loc, "z": z, "mu": mu, "sigma": sig}


def ssivae_aux_loss(sd, cfg, xs, ys=None, aux_loss_multiplier=20.0):
    """model_aux (ssivae.py:229-242): -mult * sum_b log Cat(y_b | alpha_b)."""
    if ys is None:
        return {"loss": xs.new_zeros(())}
    alpha = classifier(sd, cfg, xs)
    lp = torch.log((alpha * ys).sum(-1))
    return {"loss": -(aux_loss_multiplier * lp).sum(), "alpha": alpha}


# --------------------------------------------------------------------------
# ss_reg_iVAE  (models/ss_reg_ivae.py:152-242 under Trace_ELBO)
# --------------------------------------------------------------------------
def regressor(sd, cfg, x, prefix="encoder_y"):
    """fcRegressorNet.forward (nets/fc.py:298-304)."""
    h = fc_stack(sd, prefix, x.reshape(-1, cfg.n_pix), cfg.activation)
    return F.linear(h, sd[prefix + ".out.weight"], sd[prefix + ".out.bias"])


def ss_reg_loss(sd, cfg, xs, eps, ys=None, beta=1.0, eps_y=None, reg_sig=0.5):
    """cfg.c_dim = reg_dim.  Supervised: ys observed under p(y) = N(0, reg_sig).
    Unsupervised: y = c(x) + reg_sig * eps_y sampled from q(y|x) = N(c(x), reg_sig)
    (ss_reg_ivae.py:205-207), scored under the prior (:172-175); the "y" sites are NOT scaled."""
    bsz = xs.shape[0]
    xf = xs.reshape(bsz, cfg.n_pix)
    grid = generate_grid(cfg.data_dim, xs.dtype) if cfg.coord > 0 else None
    sig_t = xs.new_full((bsz, cfg.c_dim), reg_sig)
    out = {}
    if ys is None:
        c = regressor(sd, cfg, xs)
        y = c + reg_sig * eps_y
        log_qy = normal_logprob(y, c, sig_t).sum()
        out["c"] = c
    else:
        y = ys
        log_qy = 0.0
    mu, sig, _ = fc_encoder(sd, [xf, y], cfg.activation, None, flat=False)
    z = mu + sig * eps
    log_q = normal_logprob(z, mu, sig)
    log_p = normal_logprob(z, torch.zeros_like(z), torch.ones_like(z))
    log_py = normal_logprob(y, torch.zeros_like(y), sig_t).sum()
    loc = _decode_spatial(sd, cfg, z, y, grid).reshape(bsz, -1)
    ll = log_lik(loc, xf, cfg.sampler_d, cfg.decoder_sig)
    elbo = (ll + beta * (log_p - log_q)).sum() + log_py - log_qy
    out.update({"loss": -elbo, "ll": ll, "loc": loc, "z": z, "mu": mu, "sigma": sig, "y": y})
    return out


def ss_reg_aux_loss(sd, cfg, xs, ys=None, aux_loss_multiplier=20.0, reg_sig=0.5):
    """model_aux (ss_reg_ivae.py:229-242): -mult * sum_b log N(y_b; c(x_b), reg_sig)."""
    if ys is None:
        return {"loss": xs.new_zeros(())}
    c = regressor(sd, cfg, xs)
    lp = normal_logprob(ys, c, torch.full_like(c, reg_sig))
    return {"loss": -(aux_loss_multiplier * lp).sum(), "c": c}


# --------------------------------------------------------------------------
# VED  (models/ved.py:122-163 under Trace_ELBO; nets/conv.py)
# --------------------------------------------------------------------------
class VedCfg:
    """Shape / option record of a VED (constructor arguments of models/ved.py:89-103)."""

    def __init__(self, input_dim, output_dim, latent_dim=2, hidden_dim_e=None, hidden_dim_d=None,
                 activation="lrelu", sampler_d="bernoulli", sigmoid_d=True, decoder_sig=0.5,
                 input_channels=1, output_channels=1, batchnorm=False):
        self.batchnorm = batchnorm
        self.input_dim, self.output_dim = tuple(input_dim), tuple(output_dim)
        self.latent_dim = latent_dim
        self.hidden_e = hidden_dim_e or [(32,), (64, 64), (128, 128)]
        self.hidden_d = hidden_dim_d or [(128, 128), (64, 64), (32,)]
        self.activation, self.sampler_d = activation, sampler_d
        self.sigmoid_d, self.decoder_sig = sigmoid_d, decoder_sig
        self.input_channels, self.output_channels = input_channels, output_channels


def _conv_nd(ndim):
    return {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[ndim]


def _bnorm(sd, p, h, stats):
    """nn.BatchNorm{1,2}d in training mode (the SVI step runs the nets in their default mode; only
    VED.encode / decode / manifold2d call .eval(), reference models/ved.py:178,193,230): batch statistics,
    biased variance, eps 1e-5.  `stats` (optional dict) receives the running statistics after
    torch's momentum-0.1 update (unbiased variance), keyed like the state_dict."""
    red = [0] + list(range(2, h.dim()))
    mean = h.mean(red, keepdim=True)
    var = ((h - mean) ** 2).mean(red, keepdim=True)
    if stats is not None:
        n = h.numel() // h.shape[1]
        with torch.no_grad():
            stats[p + ".running_mean"] = 0.9 * sd[p + ".running_mean"] + 0.1 * mean.flatten()
            stats[p + ".running_var"] = (0.9 * sd[p + ".running_var"]
                                         + 0.1 * var.flatten() * n / max(n - 1, 1))
            stats[p + ".num_batches_tracked"] = sd[p + ".num_batches_tracked"] + 1
    shape = [1, -1] + [1] * (h.dim() - 2)
    return (h - mean) / torch.sqrt(var + 1e-5) * sd[p + ".weight"].view(shape) + sd[p + ".bias"].view(shape)


def ved_encoder(sd, cfg, x, stats=None):
    """convEncoderNet.forward (nets/conv.py:56-64): FeatureExtractor (conv+act blocks, a 2x
    max-pool after a block while more convolutions remain, conv.py:173-195) -> flatten ->
    fc_latent -> split -> softplus on the second half."""
    nd = len(cfg.input_dim)
    conv, pool = _conv_nd(nd), {1: F.max_pool1d, 2: F.max_pool2d, 3: F.max_pool3d}[nd]
    act = activation_fn(cfg.activation)
    h = x.reshape(x.shape[0], cfg.input_channels, *cfg.input_dim)
    total = sum(len(b) for b in cfg.hidden_e)
    idx, done = 0, 0
    for block in cfg.hidden_e:
        for _ in block:
            p = "encoder_z.feature_extractor.layers.{}".format(idx)
            h = act(conv(h, sd[p + ".weight"], sd[p + ".bias"], padding=1))
            idx += 2
            if cfg.batchnorm:
                h = _bnorm(sd, "encoder_z.feature_extractor.layers.{}".format(idx), h, stats)
                idx += 1
            done += 1
        if done + 1 < total:
            h = pool(h, 2, 2)
            idx += 1
    enc = F.linear(h.reshape(h.shape[0], -1), sd["encoder_z.features2latent.fc_latent.weight"],
                   sd["encoder_z.features2latent.fc_latent.bias"])
    mu, s = enc.split(cfg.latent_dim, 1)
    return mu, F.softplus(s)


def ved_decoder(sd, cfg, z, stats=None):
    """convDecoderNet.forward (nets/conv.py:96-102): latent2features -> Upsampler (conv+act
    blocks, each closed by an UpsampleBlock = x2 interpolation ('bilinear' in 2-D, 'nearest'
    in 1-D, conv.py:127-143) + 1x1 conv; final 1x1 conv, conv.py:228-246) -> sigmoid."""
    nd = len(cfg.output_dim)
    conv = _conv_nd(nd)
    act = activation_fn(cfg.activation)
    in_dim = [int(d) // 2 ** len(cfg.hidden_d) for d in cfg.output_dim]
    h = F.linear(z, sd["decoder.latent2features.fc.weight"], sd["decoder.latent2features.fc.bias"])
    h = h.view(-1, cfg.hidden_d[0][0], *in_dim)
    idx = 0
    for block in cfg.hidden_d:
        for _ in block:
            p = "decoder.upsampler.layers.{}".format(idx)
            h = act(conv(h, sd[p + ".weight"], sd[p + ".bias"], padding=1))
            idx += 2
            if cfg.batchnorm:
                h = _bnorm(sd, "decoder.upsampler.layers.{}".format(idx), h, stats)
                idx += 1
        p = "decoder.upsampler.layers.{}.conv".format(idx)
        h = F.interpolate(h, scale_factor=2, mode="bilinear" if nd == 2 else "nearest")
        h = conv(h, sd[p + ".weight"], sd[p + ".bias"])
        idx += 1
    p = "decoder.upsampler.layers.{}".format(idx)
    h = conv(h, sd[p + ".weight"], sd[p + ".bias"])
    return torch.sigmoid(h) if cfg.sigmoid_d else h


def ved_loss(sd, cfg, x, y, eps, beta=1.0, stats=None):
    """Returns dict(loss, ll[B], loc[B,N], z, mu, sigma);
    loss = -( sum_b log p(y_b | decoder(z_b)) + beta sum_b (log p(z_b) - log q(z_b)) ).
    The model replays the guide's z, so each net runs once per step (one batch-norm update)."""
    mu, sig = ved_encoder(sd, cfg, x, stats)
    z = mu + sig * eps
    log_q = normal_logprob(z, mu, sig)
    log_p = normal_logprob(z, torch.zeros_like(z), torch.ones_like(z))
    loc = ved_decoder(sd, cfg, z, stats).flatten(1)
    ll = log_lik(loc, y.flatten(1), cfg.sampler_d, cfg.decoder_sig)
    elbo = ll.sum() + beta * (log_p - log_q).sum()
    return {"loss": -elbo, "ll": ll, "loc": loc, "z": z, "mu": mu, "sigma": sig}


# --------------------------------------------------------------------------
# gradients + Adam (Pyro optim.Adam == torch.optim.Adam defaults per param)
# --------------------------------------------------------------------------
def loss_and_grads(loss_fn, sd, *args, **kwargs):
    """Runs loss_fn with autograd over every tensor of `sd`; returns
    (outputs dict, grads dict keyed like sd; params unused by the loss -> None)."""
    leaf = OrderedDict((k, v.detach().clone().requires_grad_(True) if v.is_floating_point()
                        else v.detach().clone()) for k, v in sd.items())
    out = loss_fn(leaf, *args, **kwargs)
    names = [k for k, v in leaf.items() if v.requires_grad]
    if out["loss"].requires_grad:
        g = torch.autograd.grad(out["loss"], [leaf[n] for n in names], allow_unused=True)
    else:
        g = [None] * len(names)
    out = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
    return out, OrderedDict(zip(names, g))


class AdamState:
    """torch.optim.Adam defaults (betas .9/.999, eps 1e-8, no weight decay),
    one independent step counter per parameter as with Pyro's per-parameter
    optimizers (a parameter first seen at step t starts its own count)."""

    def __init__(self, lr=1e-3):
        self.lr = lr
        self.m, self.v, self.t = {}, {}, {}

    def step(self, sd, grads):
        for k, g in grads.items():
            if g is None:
                continue
            if k not in self.t:
                self.m[k] = torch.zeros_like(sd[k])
                self.v[k] = torch.zeros_like(sd[k])
                self.t[k] = 0
            self.t[k] += 1
            t = self.t[k]
            self.m[k].mul_(0.9).add_(g, alpha=0.1)
            self.v[k].mul_(0.999).addcmul_(g, g, value=0.001)
            bc1 = 1 - 0.9 ** t
            bc2 = 1 - 0.999 ** t
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(1e-8)
            sd[k] = sd[k] - (self.lr / bc1) * self.m[k] / denom
        return sd


# --------------------------------------------------------------------------
# weight init identical to the reference constructors
# --------------------------------------------------------------------------
def _linear(sd, name, fan_in, fan_out, bias=True):
    lin = torch.nn.Linear(fan_in, fan_out, bias=bias)
    sd[name + ".weight"] = lin.weight.detach().clone()
    if bias:
        sd[name + ".bias"] = lin.bias.detach().clone()


def init_ivae_state(cfg, hidden_e=(128, 128), hidden_d=(128, 128), seed=1):
    """Same construction order as iVAE.__init__ (ivae.py:142-154) under
    torch.manual_seed(seed), so the draw sequence matches nn.Linear defaults."""
    torch.manual_seed(seed)
    sd = OrderedDict()
    dims = [cfg.n_pix + cfg.c_dim] + list(hidden_e)
    for i in range(len(hidden_e)):
        _linear(sd, "encoder_z.fc_layers.{}".format(2 * i), dims[i], dims[i + 1])
    _linear(sd, "encoder_z.fc11", dims[-1], cfg.z_dim)
    _linear(sd, "encoder_z.fc12", dims[-1], cfg.z_dim)
    if 0 < cfg.coord < 5:
        _linear(sd, "decoder.coord_latent.fc_coord", 1 if cfg.ndim == 1 else 2, hidden_d[0])
        _linear(sd, "decoder.coord_latent.fc_latent", cfg.latent_dim + cfg.c_dim,
                hidden_d[0], bias=False)
        dims = [hidden_d[0]] + list(hidden_d)
        for i in range(len(hidden_d)):
            _linear(sd, "decoder.fc_layers.{}".format(2 * i), dims[i], dims[i + 1])
        _linear(sd, "decoder.out", dims[-1], 1)
    else:
        dims = [cfg.latent_dim + cfg.c_dim] + list(hidden_d)
        for i in range(len(hidden_d)):
            _linear(sd, "decoder.fc_layers.{}".format(2 * i), dims[i], dims[i + 1])
        _linear(sd, "decoder.out", dims[-1], cfg.n_pix)
    return sd


class SVIPort:
    """One-object CPU baseline: `.step(x[,y])` = loss_and_grads + Adam, the
    port of pyro.infer.SVI.step for the iVAE (Trace_ELBO) path
    (trainers/svi.py:104-113)."""

    def __init__(self, cfg, sd=None, lr=1e-3, seed=1, dtype=torch.float32):
        self.cfg = cfg
        self.sd = sd if sd is not None else init_ivae_state(cfg, seed=seed)
        self.sd = OrderedDict((k, v.to(dtype)) for k, v in self.sd.items())
        self.opt = AdamState(lr)
        self.gen = torch.Generator().manual_seed(seed)
        self.dtype = dtype

    def step(self, x, y=None, beta=1.0, eps=None):
        x = x.to(self.dtype)
        if eps is None:
            eps = torch.randn(x.shape[0], self.cfg.z_dim, generator=self.gen, dtype=self.dtype)
        out, grads = loss_and_grads(ivae_loss, self.sd, self.cfg, x, eps, y, beta)
        self.sd = self.opt.step(self.sd, grads)
        return float(out["loss"])
