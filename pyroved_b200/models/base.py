"""baseVAE: shared bookkeeping of the invariant VAE family
(reference models/base.py:21-192)."""
from typing import List, Tuple

import torch
import torch.nn as nn

from .. import ops
from ..utils import generate_grid, init_dataloader


def _prod(t):
    n = 1
    for v in t:
        n *= int(v)
    return n


class baseVAE(nn.Module):
    """Parses `invariances`, builds the (constant) coordinate grid and the
    translation / scale priors, and provides batched encode/decode plus the
    set_encoder/set_decoder/save_weights/load_weights hooks.

    Args:
        data_dim: (height, width) for images or (length,) for spectra.
        invariances: list drawn from 'r', 't', 's' (1-D: only ['t']) or None.
    Keyword Args:
        device, dx_prior (0.1), dy_prior (= dx_prior), sc_prior (0.1)
    """

    def __init__(self, *args, **kwargs):
        super().__init__()
        data_dim, invariances = args
        self.device = kwargs.get("device", 'cuda' if torch.cuda.is_available() else 'cpu')
        self.ndim = len(data_dim)
        if invariances is None:
            coord = 0
        else:
            coord = len(invariances)
            if self.ndim == 1:
                if coord > 1 or invariances[0] != 't':
                    raise ValueError(
                        "For 1D data, the only invariance to enforce "
                        "is translation ('t')")
            if 't' in invariances and self.ndim == 2:
                coord = coord + 1
        self.coord = coord
        self.invariances = invariances
        self._data_dim = tuple(int(d) for d in data_dim)
        self._n_pix = _prod(data_dim)
        self._H = self._data_dim[0]
        self._W = self._data_dim[1] if self.ndim > 1 else 1
        self._dx_prior = self._dy_prior = 0.0
        self._sc_prior = 0.0
        if self.coord > 0:
            self.grid = generate_grid(data_dim).to(self.device)
        if self.coord > 0 and 't' in self.invariances:
            dx_pri = torch.tensor(kwargs.get("dx_prior", 0.1))
            dy_pri = kwargs.get("dy_prior", dx_pri.clone())
            self.t_prior = (torch.tensor([dx_pri, dy_pri]) if self.ndim == 2
                            else dx_pri).to(self.device)
            self._dx_prior = float(dx_pri)
            self._dy_prior = float(dy_pri)
        if self.coord > 0 and 's' in self.invariances:
            self.sc_prior = torch.tensor(kwargs.get("sc_prior", 0.1)).to(self.device)
            self._sc_prior = float(self.sc_prior)
        self.encoder_z = None
        self.decoder = None

    # ---- Pyro-style entry points: the reference exposes model()/guide() for
    # pyro.infer.SVI; here the whole trace is one fused CUDA step (engine.py).
    def model(self, *args, **kwargs):
        raise NotImplementedError(
            "pyroved_b200 evaluates model+guide as one fused CUDA step; use "
            "trainers.SVItrainer (or engine.SVIEngine) instead of pyro.infer.SVI")

    def guide(self, *args, **kwargs):
        raise NotImplementedError(
            "pyroved_b200 evaluates model+guide as one fused CUDA step; use "
            "trainers.SVItrainer (or engine.SVIEngine) instead of pyro.infer.SVI")

    def _split_latent(self, z: torch.Tensor) -> Tuple[torch.Tensor]:
        """(phi, dx, scale, content) views of z in the fixed order r, t, s
        (reference models/base.py:97-119)."""
        if self.ndim == 1:
            return None, z[:, 0:1], None, z[:, 1:]
        phi = torch.tensor(0).to(z.device)
        dx = torch.tensor(0).to(z.device)
        sc = torch.tensor(1).to(z.device)
        if 'r' in self.invariances:
            phi, z = z[:, 0], z[:, 1:]
        if 't' in self.invariances:
            dx, z = z[:, :2], z[:, 2:]
        if 's' in self.invariances:
            sc = sc + self.sc_prior.to(z.device) * z[:, 0]
            z = z[:, 1:]
        return phi, dx, sc, z

    # ---- batched inference ---------------------------------------------------
    def _encode(self, *input_args, **kwargs) -> torch.Tensor:
        """Encoder outputs concatenated along the last dim, batch by batch
        (reference models/base.py:121-143)."""
        loader = init_dataloader(*input_args, shuffle=False,
                                 batch_size=kwargs.get("batch_size", 100))
        out = []
        with self._on_device():
            for batch in loader:
                batch = [t.to(self.device) for t in batch]
                enc = self.encoder_z(batch[0] if len(batch) == 1 else batch)
                out.append(torch.cat(enc, -1).cpu())
        return torch.cat(out)

    def _decode(self, z_new: torch.Tensor, **kwargs) -> torch.Tensor:
        """Decode latent codes batch by batch; the optional `angle`, `shift`,
        `scale` kwargs condition the coordinate grid exactly like
        reference models/base.py:145-171 (absolute shift / scale)."""
        bs = kwargs.get("batch_size", 100)
        dev = self.device
        out = []
        spatial = bool(self.invariances) and self.coord > 0
        with self._on_device():
            for s in range(0, z_new.shape[0], bs):
                z = z_new[s:s + bs].to(dev).float().contiguous()
                if spatial:
                    out.append(self._decode_spatial_batch(z, **kwargs).cpu())
                else:
                    out.append(self.decoder(z).cpu())
        return torch.cat(out)

    def _on_device(self):
        """Context in which the kernels of an inference call launch: the model's own CUDA device,
        whatever the caller's current device is (the kernels take torch's current stream)."""
        dev = torch.device(self.device)
        if dev.type != "cuda":
            import contextlib
            return contextlib.nullcontext()
        return torch.cuda.device(dev)

    def _decode_spatial_batch(self, z, **kwargs):
        dec = self.decoder
        n = z.shape[0]
        dev = z.device
        a = torch.as_tensor(kwargs.get("angle", 0.), dtype=torch.float32).reshape(-1)
        t = torch.as_tensor(kwargs.get("shift", 0.), dtype=torch.float32).reshape(-1)
        s = torch.as_tensor(kwargs.get("scale", 1.), dtype=torch.float32).reshape(-1)
        cl = dec.coord_latent
        hd = cl.fc_coord.out_features
        lc = z.shape[1]
        if self.ndim == 2:
            t2 = t.expand(2) if t.numel() == 1 else t[:2]
            head = torch.stack([a[0], t2[0], t2[1], s[0] - 1.]).to(dev)
            cfg = ops.make_fold_cfg(2, ['r', 't', 's'], lc, 0, hd, 1.0, 1.0, 1.0)
        else:
            head = t[:1].to(dev)
            cfg = ops.make_fold_cfg(1, ['t'], lc, 0, hd, 1.0, 1.0, 1.0)
        zf = torch.cat([head.unsqueeze(0).expand(n, -1), z], dim=1).contiguous()
        Uv = torch.empty(n, 3, hd, device=dev)
        ops.fold_fwd(cfg, zf, None, cl.fc_coord.weight.data, cl.fc_coord.bias.data,
                     cl.fc_latent.weight.data, Uv)
        loc = self._run_sdecoder(Uv, n)
        if getattr(dec, "unflat", True):
            return loc.view(-1, *self._data_dim)
        return loc.view(-1, 1)

    def _run_sdecoder(self, Uv, n):
        """Forward-only spatial decoder on folded first-layer coefficients."""
        from ..nets.fc import fc_stack_forward, linear_layers
        dec = self.decoder
        dev = Uv.device
        N = self._n_pix
        layers = linear_layers(dec.fc_layers)
        tc = (ops.has_tcgen05() and len(layers) == 2 and dec.activation == "tanh" and N >= 32
              and all(l.in_features == 128 and l.out_features == 128 for l in layers)
              and Uv.shape[2] == 128)
        if tc:
            loc = torch.empty(n * N, device=dev)
            rowll = torch.empty(n * N, device=dev)
            ops.sdec_tc_step(Uv, None, None, layers[0].weight.data, layers[0].bias.data,
                             layers[1].weight.data, layers[1].bias.data, dec.out.weight.data,
                             dec.out.bias.data, rowll, loc, None, None, n, n, self._H, self._W,
                             self.ndim, "bernoulli", dec.sigmoid_out, 0.5, False)
            return loc
        h0 = torch.empty(n * N, Uv.shape[2], device=dev)
        ops.sdec_h0_fwd(Uv, h0, self._H, self._W, self.ndim)
        h = fc_stack_forward(dec.fc_layers, h0, dec.activation)
        return ops.linear_fwd(h, dec.out.weight.data, dec.out.bias.data,
                              "sigmoid" if dec.sigmoid_out else None).reshape(-1)

    # ---- hooks ----------------------------------------------------------------
    def set_encoder(self, encoder_net: nn.Module) -> None:
        """Replace the encoder (reference models/base.py:173-176).  The SVI step is a sequence of
        hand-written kernels chosen from the structure of the nets, so the replacement must be one
        of this package's net classes of a kind the model's step knows how to run: an
        `fcEncoderNet`-family net of the model's own kind, or -- for iVAE -- a `convEncoderNet`
        whose latent_dim equals the model's full latent width (transform + content latents).
        Anything else raises TypeError (there is no eager / autograd fallback)."""
        self._check_replacement("encoder_z", encoder_net)
        self.encoder_z = encoder_net.to(self.device)

    def set_decoder(self, decoder_net: nn.Module) -> None:
        """Replace the decoder (reference models/base.py:178-181); see set_encoder."""
        self._check_replacement("decoder", decoder_net)
        self.decoder = decoder_net.to(self.device)

    def _check_replacement(self, slot, net):
        from ..nets import conv as convnets
        from ..nets import fc as fcnets
        cur = getattr(self, slot)
        own = tuple(c for c in list(vars(fcnets).values()) + list(vars(convnets).values())
                    if isinstance(c, type) and issubclass(c, nn.Module))
        if not isinstance(net, own):
            raise TypeError(
                "pyroved_b200 runs the SVI step as fused CUDA kernels selected from the structure "
                "of the nets: set_{} accepts this package's net classes (pyroved_b200.nets), not {}"
                .format("encoder" if slot == "encoder_z" else "decoder", type(net).__name__))
        same_kind = cur is not None and type(net) is type(cur)
        conv_for_ivae = (slot == "encoder_z" and isinstance(net, convnets.convEncoderNet)
                         and type(self).__name__ == "iVAE")
        if not (same_kind or conv_for_ivae):
            raise TypeError("{} cannot replace the {} of a {} (supported: a net of the same class{})"
                            .format(type(net).__name__, slot, type(self).__name__,
                                    ", or convEncoderNet for iVAE" if slot == "encoder_z" else ""))
        if conv_for_ivae:
            if getattr(self, "c_dim", 0) > 0:
                raise TypeError("a convolutional encoder cannot take the conditioning vector y")
            if net.latent_dim != self.z_dim:
                raise ValueError("convEncoderNet(latent_dim={}) must produce the model's full latent "
                                 "vector: {} (= {} transform + {} content latents)".format(
                                     net.latent_dim, self.z_dim, self.coord, self.z_dim - self.coord))
            if tuple(net.input_dim) != tuple(self._data_dim):
                raise ValueError("encoder input_dim {} does not match data_dim {}".format(
                    net.input_dim, self._data_dim))
        elif same_kind:
            a = [(k, tuple(v.shape)) for k, v in cur.state_dict().items()]
            b = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
            # hidden sizes may differ; input / output widths must fit the model
            if a[0][1][-1] != b[0][1][-1] or a[-1][1][0] != b[-1][1][0]:
                raise ValueError("replacement {} has input / output widths {} / {}, the model needs "
                                 "{} / {}".format(type(net).__name__, b[0][1][-1], b[-1][1][0],
                                                  a[0][1][-1], a[-1][1][0]))

    def save_weights(self, filepath: str) -> None:
        torch.save(self.state_dict(), filepath + '.pt')

    def load_weights(self, filepath: str) -> None:
        weights = torch.load(filepath, map_location=self.device)
        self.load_state_dict(weights)

    # ---- engine hooks -----------------------------------------------------------
    def _beta(self, kwargs):
        b = kwargs.get("scale_factor", 1.)
        return float(b)

    def _aux_scale(self, kwargs):
        return 1.0

    def _make_program(self, engine, B, has_y, mode="main"):
        raise NotImplementedError
