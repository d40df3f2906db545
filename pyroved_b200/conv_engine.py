"""SVI step program for VED (reference models/ved.py:122-163 under Trace_ELBO): convolutional
encoder -> reparameterised sample -> convolutional decoder -> per-pixel log-likelihood + KL,
forward and hand-written backward on the kernels of csrc/pvb_conv.cu, sequenced layer by layer
with preallocated activations (replayed as a CUDA graph by engine.SVIEngine)."""
import torch

from . import ops
from .engine import StepProgram
from .nets.conv import layer_plan, out_shape


class ConvStack:
    """A conv / pool / upsample sequence with saved activations and its backward."""

    def __init__(self, engine, layers, activation, B, in_shape):
        self.engine, self.act, self.B = engine, activation, B
        dev = engine.device
        f32 = dict(device=dev, dtype=torch.float32)
        self.steps = []
        shape = tuple(in_shape)
        self.in_shape = shape
        biggest = B * _numel(shape)
        for kind, mod, act in layer_plan(layers, activation):
            oshape = out_shape(kind, mod, shape)
            st = dict(kind=kind, mod=mod, act=act, y=torch.empty((B, *oshape), **f32),
                      pre=(torch.empty((B, *oshape), **f32) if act == "gelu" else None))
            # tensor-core path for layers with >= 16 channels on both sides (tcgen05, fp16 / bf16
            # operands); the others (first / last layer) stay on the fp32 kernels
            st["tc"] = (kind == "conv" and not engine.force_generic
                        and ops.conv_tc_supported(mod.weight))
            if st["tc"]:
                # repacked weights, one workspace per direction: both are refreshed once per step
                # on the side stream (prepare()), off the critical path
                st["ws"] = ops.conv_tc_workspace(mod.weight)
                st["ws_bwd"] = ops.conv_tc_workspace(mod.weight)
            st["tc_wgrad"] = (kind == "conv" and not engine.force_generic
                              and ops.conv_tc_wgrad_supported(mod.weight))
            if kind == "bn":
                st["stats"] = torch.empty(2, oshape[0], **f32)       # batch mean, 1/std
                st["ws"] = ops.bn_workspace(oshape[0], dev)
            self.steps.append(st)
            shape = oshape
            biggest = max(biggest, B * _numel(shape))
        self.out_shape = shape
        # zeroed scratch of the tensor-core weight-gradient kernels (coalesced accumulator read-out)
        # (one slice per layer: the layers of a backward pass are folded into the gradients by ONE launch)
        wg = [st for st in self.steps if st["kind"] == "conv" and st["tc_wgrad"]]
        for st, sc in zip(wg, ops.conv_tc_wgrad_scratch([st["mod"].weight for st in wg], dev, shared=False)
                          if wg else []):
            st["wg_scratch"] = sc
        self.gbuf = [torch.empty(biggest, **f32) for _ in range(2)]
        self.x = None

    def prepare(self):
        """Repack the tensor-core layers' weights for both directions (they changed in the last
        optimizer step).  Called at the start of a step inside engine.fork_side(): the repacks run
        concurrently with the first layers of the forward pass."""
        for st in self.steps:
            if st["kind"] == "conv" and st["tc"]:
                ops.conv_tc_prep(st["mod"].weight.data, st["ws"], 0)
                ops.conv_tc_prep(st["mod"].weight.data, st["ws_bwd"], 1)
        self.prepped = True

    def forward(self, x):
        self.x = x
        cur = x
        prepped = getattr(self, "prepped", False)
        for st in self.steps:
            m = st["mod"]
            if st["kind"] == "conv":
                bias = m.bias.data if m.bias is not None else None
                if st["tc"]:
                    if prepped and self.engine._side_used:
                        self.engine.join_side()        # the repacked weights are ready
                    ops.conv_tc_fwd(cur, m.weight.data, bias, st["act"], st["y"], st["ws"], st["pre"],
                                    prepped=prepped)
                else:
                    ops.conv_fwd(cur, m.weight.data, bias, st["act"], st["y"], st["pre"])
            elif st["kind"] == "bn":
                # the module's own mode, like the reference: batch statistics + running-statistics
                # update while training; the running statistics once the user has called
                # encode / decode / manifold2d, which put the model in eval mode for good
                # (reference models/ved.py:178,193,230) -- the step graph is keyed on the mode
                ops.bn_fwd(cur, m, st["y"], st["stats"][0], st["stats"][1], st["ws"])
            elif st["kind"] == "pool":
                ops.maxpool2_fwd(cur, st["y"])
            else:
                ops.upsample2_fwd(cur, st["y"], m.mode == "bilinear")
            cur = st["y"]
        return cur

    def backward(self, dy, need_dx):
        """dy: gradient wrt the stack output (clobbered).  Accumulates weight / bias gradients
        into the flat gradient buffer; returns the gradient wrt the stack input (or None)."""
        flat = self.engine.flat
        d = dy
        folds = []          # tensor-core weight gradients still in their scratch copies
        for k in range(len(self.steps) - 1, -1, -1):
            st = self.steps[k]
            m = st["mod"]
            xin = self.steps[k - 1]["y"] if k > 0 else self.x
            want_dx = k > 0 or need_dx
            dx = None
            if want_dx:
                buf = self.gbuf[0] if d.data_ptr() != self.gbuf[0].data_ptr() else self.gbuf[1]
                dx = buf[:xin.numel()].view_as(xin)
            # activation derivative of the layer BELOW folded into the kernel that produces dx
            # (backward-data epilogue, max-pool / upsample backward): saves a pass over dx
            below = self.steps[k - 1] if k > 0 else None
            fuse = (below is not None and below["kind"] == "conv"
                    and below["act"] not in (None, "gelu") and not self.engine.force_generic
                    and (st["kind"] == "conv" or (st["kind"] in ("pool", "up") and xin.dim() < 5)))
            if st["kind"] == "conv":
                if st["act"] is not None and not st.get("dpre_ready", False):
                    ops.act_bwd(d, st["y"], st["pre"], d, st["act"])     # in place: d = dpre
                st["dpre_ready"] = False
                gb = flat.gv(m.bias) if m.bias is not None else None
                if st["tc_wgrad"]:
                    ops.conv_tc_bwd_weight(d, xin, m.weight.data, flat.gv(m.weight), gb, st["wg_scratch"],
                                           fold=False)
                    folds.append((st["wg_scratch"], m.weight.data, flat.gv(m.weight), gb))
                else:
                    ops.conv_bwd_weight(d, xin, m.weight.data, flat.gv(m.weight), gb)
                if want_dx:
                    if st["tc"]:
                        pp = getattr(self, "prepped", False)
                        if fuse:
                            ops.conv_tc_bwd_data(d, m.weight.data, dx, st["ws_bwd"], xin, below["act"],
                                                 prepped=pp)
                            below["dpre_ready"] = True
                        else:
                            ops.conv_tc_bwd_data(d, m.weight.data, dx, st["ws_bwd"], prepped=pp)
                    elif fuse:
                        ops.conv_bwd_data(d, m.weight.data, dx, xin, below["act"])
                        below["dpre_ready"] = True
                    else:
                        ops.conv_bwd_data(d, m.weight.data, dx)
            elif st["kind"] == "bn":
                # the step below is always the convolution whose output feeds this layer
                gg = flat.gv(m.weight) if m.affine else None
                gb = flat.gv(m.bias) if m.affine else None
                ops.bn_bwd(d, xin, m, st["stats"][0], st["stats"][1], dx, gg, gb, st["ws"])
            elif st["kind"] == "pool":
                if want_dx:
                    ops.maxpool2_bwd(xin, d, dx, below["act"] if fuse else None)
                    if fuse:
                        below["dpre_ready"] = True
            else:
                if want_dx:
                    if fuse:
                        ops.upsample2_bwd(d, dx, m.mode == "bilinear", xin, below["act"])
                        below["dpre_ready"] = True
                    else:
                        ops.upsample2_bwd(d, dx, m.mode == "bilinear")
            d = dx
        ops.conv_tc_wgrad_fold(folds)
        return d


def _numel(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


class ConvGaussEncoder:
    """`convEncoderNet` as the guide q(z|x) of a Trace_ELBO step (reference nets/conv.py:24-64):
    FeatureExtractor (ConvStack) -> flatten -> fc_latent, whose outputs [0, L) give mu and [L, 2L) the
    pre-softplus sigma -> reparameterised sample + KL terms; and the
    backward of all of it.  Buffer names follow engine.GaussHead (eps, mu, s_pre, sigma, z, kl)."""

    def __init__(self, engine, enc, B, L):
        dev = engine.device
        f32 = dict(device=dev, dtype=torch.float32)
        self.engine, self.enc, self.B, self.L = engine, enc, B, L
        self.in_shape = (enc.input_channels, *enc.input_dim)
        fe = enc.feature_extractor
        self.stack = ConvStack(engine, fe.layers, fe.activation, B, self.in_shape)
        self.feat_dim = _numel(self.stack.out_shape)
        fcl = enc.features2latent.fc_latent
        if fcl.in_features != self.feat_dim:
            raise ValueError("encoder feature map has {} elements but features2latent expects {} "
                             "(reference nets/conv.py:44-45 assumes len(hidden_dim)-1 poolings)"
                             .format(self.feat_dim, fcl.in_features))
        if fcl.out_features != 2 * L:
            raise ValueError("convEncoderNet(latent_dim={}) does not match the model's latent width {}"
                             .format(fcl.out_features // 2, L))
        if not getattr(enc, "softplus_out", True):
            raise NotImplementedError("pyroved_b200: convEncoderNet(softplus_out=False) as a guide")
        # both heads in ONE pass over the feature map (it is the large operand: 32768 features per sample in
        # the default VED, three passes of 67 MB per head otherwise); split / joined by [B, L] copies
        self.ms = torch.empty(B, 2 * L, **f32)
        self.gms = torch.empty(B, 2 * L, **f32)
        self.eps = torch.zeros(B, L, **f32)
        self.mu = torch.empty(B, L, **f32)
        self.s_pre = torch.empty(B, L, **f32)
        self.sigma = torch.empty(B, L, **f32)
        self.z = torch.empty(B, L, **f32)
        self.kl = torch.empty(B, **f32)
        self.gmu = torch.empty(B, L, **f32)
        self.gs_pre = torch.empty(B, L, **f32)
        self.dfeat = torch.empty(B, self.feat_dim, **f32)

    @property
    def uses_tc(self):
        return any(st["tc"] for st in self.stack.steps)

    def forward(self, x, gen_eps):
        eng = self.engine
        feat = self.stack.forward(x.view(self.B, *self.in_shape)).view(self.B, self.feat_dim)
        if gen_eps:
            ops.randn(self.eps, eng.seed, eng.step_counter, eng.eps_first_index(self.eps.numel()))
        fcl = self.enc.features2latent.fc_latent
        ops.linear_fwd(feat, fcl.weight.data, fcl.bias.data, None, out=self.ms)
        self.mu.copy_(self.ms[:, :self.L])
        self.s_pre.copy_(self.ms[:, self.L:])
        ops.latent_fwd(self.mu, self.s_pre, self.eps, self.sigma, self.z, self.kl)

    def backward(self, gz, w, beta):
        """gz = dloss/dz [B, L] from the decoder side; adds the KL terms (times beta) and
        propagates through the heads and the convolutional stack."""
        flat = self.engine.flat
        ops.latent_bwd(gz, self.eps, self.sigma, self.s_pre, self.z, w, beta, self.gmu, self.gs_pre)
        fcl = self.enc.features2latent.fc_latent
        gW, gb = flat.gv(fcl.weight), flat.gv(fcl.bias)
        L = self.L
        feat = self.stack.steps[-1]["y"].view(self.B, self.feat_dim)
        self.gms[:, :L].copy_(self.gmu)
        self.gms[:, L:].copy_(self.gs_pre)
        ops.linear_bwd(feat, fcl.weight.data, None, None, self.gms, self.gms, self.dfeat, False, gW, gb, None)
        self.stack.backward(self.dfeat.view(self.B, *self.stack.out_shape), False)


class VEDProgram(StepProgram):
    """One Trace_ELBO step of VED on a batch (x, y):
    loss = -sum_b [ log p(y_b | decoder(z_b)) + beta (log p(z_b) - log q(z_b | x_b)) ]."""

    def __init__(self, engine, B):
        super().__init__(engine, B, True)
        m = engine.model
        dev = engine.device
        f32 = dict(device=dev, dtype=torch.float32)
        enc, dec = m.encoder_z, m.decoder
        L = m.z_dim
        self.L = L
        self.N_out = m.output_channels * _numel(m.output_dim)
        self.genc = ConvGaussEncoder(engine, enc, B, L)
        self.in_shape = self.genc.in_shape
        self.enc = self.genc.stack
        self.x = torch.zeros((B, *self.in_shape), **f32)
        self.y = torch.zeros(B, self.N_out, **f32)
        self.gz = torch.empty(B, L, **f32)
        l2f = dec.latent2features
        self.dec_in_shape = tuple(l2f.reshape_)
        self.feat0 = torch.empty(B, _numel(self.dec_in_shape), **f32)
        self.dec = ConvStack(engine, dec.upsampler.layers, dec.upsampler.activation, B,
                             self.dec_in_shape)
        if _numel(self.dec.out_shape) != self.N_out:
            raise ValueError("decoder output {} does not match output_dim {}".format(
                self.dec.out_shape, m.output_dim))
        self.rowll = torch.empty(B * self.N_out, **f32)
        self.dlogit = torch.empty(B * self.N_out, **f32)
        self._loc = torch.empty(B * self.N_out, **f32)
        self.ll = torch.empty(B, **f32)

    loc = property(lambda s: s._loc)
    eps = property(lambda s: s.genc.eps)
    mu = property(lambda s: s.genc.mu)
    sigma = property(lambda s: s.genc.sigma)
    z = property(lambda s: s.genc.z)
    kl = property(lambda s: s.genc.kl)
    use_tc = property(lambda s: any(st["tc"] for st in s.enc.steps + s.dec.steps))

    def load(self, x, y):
        B = self.B
        self.x.copy_(x.reshape(self.x.shape), non_blocking=True)
        self.y.copy_(y.reshape(B, -1), non_blocking=True)

    def forward(self, beta, want_grad, gen_eps):
        eng = self.engine
        m = eng.model
        samp = m.sampler_d
        # weights of the tensor-core convolutions -> fp16 operand layout, both directions, once per
        # step on the side stream (24 small launches that used to sit on the critical path)
        with eng.fork_side():
            self.enc.prepare()
            self.dec.prepare()
        self.genc.forward(self.x, gen_eps)
        fc = m.decoder.latent2features.fc
        ops.linear_fwd(self.z, fc.weight.data, fc.bias.data, None, out=self.feat0)
        logit = self.dec.forward(self.feat0.view(self.B, *self.dec_in_shape))
        ops.obs_loglik(logit.view(-1), self.y, None, self.rowll,
                       self.dlogit if want_grad else None, self._loc, self.B, self.B, self.N_out,
                       samp.name, m.decoder.sigmoid_out, samp.decoder_sig)
        ops.elbo_reduce(self.rowll, self.kl, None, float(beta), self.ll, eng.flat.loss, True,
                        self.B, self.N_out)

    def backward(self, beta):
        eng = self.engine
        m, flat = eng.model, eng.flat
        dfeat0 = self.dec.backward(self.dlogit.view(self.B, *self.dec.out_shape), True)
        fc = m.decoder.latent2features.fc
        ops.linear_bwd(self.z, fc.weight.data, None, None, dfeat0.reshape(self.B, -1),
                       dfeat0.reshape(self.B, -1), self.gz, False, flat.gv(fc.weight),
                       flat.gv(fc.bias), None)
        self.genc.backward(self.gz, None, beta)
