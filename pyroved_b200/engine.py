"""SVI step engine: the B200-native replacement for what the reference gets
from `pyro.infer.SVI(model, guide, Adam, Trace_ELBO/TraceEnum_ELBO).step`
(reference trainers/svi.py:81-91,107; trainers/auxsvi.py:67-81,88-100).

One `step` = encoder forward -> reparameterised latent sample -> affine fold
-> spatial decoder forward+backward -> latent/encoder backward -> (NCCL
all-reduce of the flat gradient) -> fused Adam, all as hand-written CUDA
kernels launched through the C ABI on torch's current stream and replayed as
CUDA graphs.  PyTorch supplies device memory, streams and torch.distributed.

Layout in HBM
  * all parameters live in ONE flat fp32 buffer (`FlatParams.p`); each
    nn.Parameter is a view into it, so state_dict()/load_state_dict() keep the
    reference's keys; gradients, Adam m and v are flat buffers of the same
    layout; the last slot of the gradient buffer holds the loss so that one
    all-reduce(SUM) moves both.
  * activations/workspaces are allocated once per (batch shape) and reused.
"""
import gc
import math
import os
import warnings

import torch
import torch.nn as nn

from . import ops, parallel
from ._lib import TC_WGRAD_FLOATS, TC_WGRAD_STRIDE
from .nets.fc import linear_layers


def _align4(n):
    return (n + 3) // 4 * 4


def _numel_shape(shape):
    n = 1
    for v in shape:
        n *= int(v)
    return n


class FlatParams:
    """Flat parameter / gradient / Adam-state buffers with per-parameter views."""

    def __init__(self, module: nn.Module, device):
        self.module = module
        self.device = torch.device(device)
        self._build()

    def _build(self):
        named = [(n, p) for n, p in self.module.named_parameters()]
        self.names = [n for n, _ in named]
        self.offsets = {}
        off = 0
        for n, p in named:
            self.offsets[n] = (off, p.numel(), tuple(p.shape))
            off += _align4(p.numel())
        self.total = off
        dev = self.device
        new_p = torch.zeros(self.total, device=dev, dtype=torch.float32)
        for n, p in named:
            o, k, shp = self.offsets[n]
            new_p[o:o + k].copy_(p.data.reshape(-1).to(dev, torch.float32))
        self.p = new_p
        # gradient buffer: [grads | loss accumulator, last step's loss, pad2]
        self.g = torch.zeros(self.total + 4, device=dev, dtype=torch.float32)
        self.m = torch.zeros(self.total, device=dev, dtype=torch.float32)
        self.v = torch.zeros(self.total, device=dev, dtype=torch.float32)
        # optimizer step at which each parameter first carried a gradient (-1: never);
        # Pyro creates one Adam per parameter on first sight (oracle/pyro_min/pyro/optim)
        self.first_step = torch.full((self.total,), -1, device=dev, dtype=torch.int32)
        self.active = set()
        self._views = {}
        for n, p in named:
            o, k, shp = self.offsets[n]
            p.data = self.p[o:o + k].view(shp)
            p.grad = self.g[o:o + k].view(shp)
            self._views[id(p)] = n
        self.loss = self.g[self.total:self.total + 1]            # accumulator of the running step
        self.last_loss = self.g[self.total + 1:self.total + 2]   # written by the optimizer kernels
        self.loss_pair = self.g[self.total:self.total + 2]
        self._ptrs = [(p, p.data_ptr()) for _, p in named]

    def intact(self):
        """False if the module's parameters were re-allocated (e.g. .to())."""
        named = list(self.module.named_parameters())
        if len(named) != len(self._ptrs):
            return False
        return all(p is q and p.data_ptr() == ptr for (_, p), (q, ptr) in zip(named, self._ptrs))

    def ensure(self):
        if not self.intact():
            m, v, fs, act = self.m, self.v, self.first_step, self.active
            old_total = self.total
            self._build()
            if self.total == old_total:   # keep optimizer state across a rebuild
                self.m.copy_(m)
                self.v.copy_(v)
                self.first_step.copy_(fs)
                self.active = act
            return True
        return False

    def pv(self, param):
        return param.data

    def gv(self, param):
        n = self._views[id(param)]
        o, k, shp = self.offsets[n]
        return self.g[o:o + k].view(shp)

    def offset(self, param):
        return self.offsets[self._views[id(param)]][0]

    def activate(self, modules, step):
        """Mark the parameters of `modules` as carrying gradients from optimizer
        step `step` (0-based count of updates done so far) onwards."""
        for mod in modules:
            for p in mod.parameters():
                n = self._views[id(p)]
                if n not in self.active:
                    o, k, _ = self.offsets[n]
                    self.first_step[o:o + k] = step
                    self.active.add(n)


class MLP:
    """A stack of Linear+activation layers with preallocated activations."""

    def __init__(self, layers, act, M, device, flat: FlatParams):
        self.layers = layers
        self.act = act
        self.M = M
        self.flat = flat
        self.h = [torch.empty(M, l.out_features, device=device) for l in layers]
        self.pre = ([torch.empty(M, l.out_features, device=device) for l in layers]
                    if act == "gelu" else [None] * len(layers))
        self.x = None

    def forward(self, x):
        self.x = x
        cur = x
        for l, h, pre in zip(self.layers, self.h, self.pre):
            ops.linear_fwd(cur, l.weight.data, l.bias.data if l.bias is not None else None,
                           self.act, out=h, pre=pre)
            cur = h
        return cur

    def backward(self, d_last, scratch, need_dx):
        """d_last: gradient wrt the last activation (overwritten). `scratch`:
        two flat buffers of >= M*max_width floats used ping-pong for dx.
        Returns gradient wrt the stack input (or None)."""
        d = d_last
        n = len(self.layers)
        for k in range(n - 1, -1, -1):
            l = self.layers[k]
            xin = self.h[k - 1] if k > 0 else self.x
            want_dx = (k > 0) or need_dx
            dx = None
            if want_dx:
                dx = scratch[k % 2][:self.M * l.in_features].view(self.M, l.in_features)
            ops.linear_bwd(xin, l.weight.data, self.h[k], self.pre[k], d, d, dx, False,
                           self.flat.gv(l.weight),
                           self.flat.gv(l.bias) if l.bias is not None else None, self.act)
            d = dx
        return d


class FusedStack:
    """A Linear+activation stack with linear heads on the fused small-batch kernels
    (pvb_mlp_tail_fwd / pvb_mlp_chain_bwd / pvb_mlp_wgrad): the first, wide layer is one GEMM
    with fused bias + activation, everything behind it (hidden layers, heads, the
    reparameterised sample and the coordinate-transform fold) is one launch; backward is two.
    Same arithmetic as `MLP` + `GaussHead` (+ `ops.fold_fwd`)."""
    MAX_ROWS = 4096

    @staticmethod
    def eligible(layers, heads, M):
        return (1 <= len(layers) <= 4 and 1 <= len(heads) <= 3 and M <= FusedStack.MAX_ROWS
                and all(l.out_features <= 256 for l in layers)
                and all(h.out_features <= 64 and h.in_features == layers[-1].out_features
                        for h in heads))

    def __init__(self, engine, layers, act, heads, M, gauss_head=None, fold=None):
        dev, flat = engine.device, engine.flat
        f32 = dict(device=dev, dtype=torch.float32)
        self.engine, self.layers, self.act, self.heads, self.M = engine, layers, act, heads, M
        self.h = [torch.empty(M, l.out_features, **f32) for l in layers]
        self.pre = ([torch.empty(M, l.out_features, **f32) for l in layers]
                    if act == "gelu" else [None] * len(layers))
        self.dpre = [torch.empty(M, l.out_features, **f32) for l in layers]
        self.gauss = gauss_head            # GaussHead: buffers mu/s_pre/eps/sigma/z/kl/gmu/gs_pre
        if gauss_head is not None:
            self.hout = [gauss_head.mu, gauss_head.s_pre] + [
                torch.empty(M, h.out_features, **f32) for h in heads[2:]]
        else:
            self.hout = [torch.empty(M, h.out_features, **f32) for h in heads]
        self.fold = fold
        self.x = None
        self._tail = {}
        self._chain = None
        self._wgrad = None

    def forward(self, x, gen_eps=False):
        eng = self.engine
        self.x = x
        l0 = self.layers[0]
        ops.linear_fwd(x, l0.weight.data, l0.bias.data if l0.bias is not None else None, self.act,
                       out=self.h[0], pre=self.pre[0])
        key = bool(gen_eps)
        args = self._tail.get(key)
        if args is None:
            gauss = None
            if self.gauss is not None:
                g = self.gauss
                gauss = dict(eps=g.eps, sigma=g.sigma, z=g.z, kl=g.kl, gen_eps=gen_eps,
                             seed=eng.seed, step_counter=eng.step_counter,
                             first_index=eng.eps_first_index(g.eps.numel()))
            args = ops.make_mlp_tail_args(self.M, self.h[0], self.layers[1:], self.h[1:],
                                          self.pre[1:], self.act, self.heads, self.hout, gauss,
                                          self.fold)
            self._tail[key] = args
        ops.mlp_tail_fwd(args)
        return self.h[-1]

    def backward(self, head_grads):
        """head_grads[k]: gradient wrt the output of head k, [M, hdim_k].  Accumulates every
        weight / bias gradient of the stack and its heads into the flat gradient buffer."""
        flat = self.engine.flat
        if self._chain is None:
            self._chain = ops.make_mlp_chain_args(self.M, self.layers, self.h, self.pre, self.act,
                                                  self.dpre, self.heads, head_grads)
            items = []
            for k, l in enumerate(self.layers):
                xin = self.x if k == 0 else self.h[k - 1]
                items.append((self.dpre[k], xin, flat.gv(l.weight),
                              flat.gv(l.bias) if l.bias is not None else None))
            for k, hd in enumerate(self.heads):
                items.append((head_grads[k], self.h[-1], flat.gv(hd.weight),
                              flat.gv(hd.bias) if hd.bias is not None else None))
            self._wgrad = ops.make_wgrad_problems(items)
            self._grads_id = [g.data_ptr() for g in head_grads] + [self.x.data_ptr()]
        assert self._grads_id == [g.data_ptr() for g in head_grads] + [self.x.data_ptr()]
        ops.mlp_chain_bwd(self._chain)
        ops.mlp_wgrad(self._wgrad, self.M)


class StepProgram:
    """Static buffers + kernel sequence for one (model, batch shape)."""
    loss_const = 0.0   # host-side constant added to the returned loss (per rank)

    def __init__(self, engine, B, has_y):
        self.engine = engine
        self.B = B
        self.has_y = has_y

    def grad_modules(self):
        """Modules whose parameters receive gradients from this program."""
        m = self.engine.model
        return [m.encoder_z, m.decoder]

    def set_eps(self, eps):
        """Inject the noise of the reparameterised sites (parity tests)."""
        self.eps.copy_(eps.reshape(self.eps.shape), non_blocking=True)


class DecoderOps:
    """Decoder forward/backward over I instances (I = B, or K*B when a discrete
    latent is enumerated).  Spatial decoder: fused tcgen05 kernel when the net
    is the default 128-128 tanh MLP, generic fp32 kernels otherwise; non-spatial
    (`fcDecoderNet`, coord == 0): generic GEMMs on the latent code."""

    def __init__(self, engine, I, B, cond_dim):
        m = engine.model
        self.engine = engine
        self.I, self.Bx, self.Cd = I, B, cond_dim
        self.N = m._n_pix
        self.Zf = m.z_dim            # full latent width (transform + content)
        dev, flat = engine.device, engine.flat
        f32 = dict(device=dev, dtype=torch.float32)
        dec = m.decoder
        self.spatial = m.coord > 0
        N = self.N
        R = I * N
        self.rowll = torch.empty(R, **f32)
        self.loc = torch.empty(R, **f32)
        self.ll = torch.empty(I, **f32)
        self.gz = torch.zeros(I, self.Zf, **f32)
        self.gcond = None
        layers = linear_layers(dec.fc_layers)
        self.use_tc = False
        if self.spatial:
            cl = dec.coord_latent
            Hd0 = cl.fc_coord.out_features
            self.fold_cfg = ops.make_fold_cfg(m.ndim, m.invariances, m._latent_dim, cond_dim, Hd0,
                                              m._dx_prior, m._dy_prior, m._sc_prior)
            self.Uv = torch.empty(I, 3, Hd0, **f32)
            self.gUv = torch.empty(I, 3, Hd0, **f32)
            self.G_fold = max(ops.fold_bwd_num_partials(), ops.latent_side_num_partials(I))
            self.G_side = ops.latent_side_num_partials(I)
            self.fold_per = Hd0 * (m.ndim + 1 + m._latent_dim + cond_dim)
            self.fold_part = torch.empty(self.G_fold, self.fold_per, **f32)
            self.use_tc = engine.tc_eligible(dec, N)
            if self.use_tc:
                s = ops.sdec_tc_sizes(I, N)
                self.tc_sizes = s
                self.gUv_part = torch.empty(max(s.gUv_part_floats, 4), **f32)
                self.wgrad_part = torch.empty(max(s.wgrad_part_floats, 4), **f32)
                # the two 128x128 layers as fp16 operand tiles: repacked once per step on the side
                # stream (prepare()), fetched by every CTA with two TMA bulk copies
                self.w_packed = ops.sdec_tc_packed_weights(dev)
            else:
                self.h0 = torch.empty(R, Hd0, **f32)
                self.dmlp = MLP(layers, dec.activation, R, dev, flat)
                Hl = layers[-1].out_features if layers else Hd0
                self.logit = torch.empty(R, 1, **f32)
                self.dlogit = torch.empty(R, 1, **f32)
                wmax = max([Hd0, Hl] + [l.in_features for l in layers])
                self.dec_scratch = [torch.empty(R * wmax, **f32) for _ in range(3)]
        else:
            self.dec_in = torch.zeros(I, self.Zf + cond_dim, **f32)
            self.dmlp = MLP(layers, dec.activation, I, dev, flat)
            Hl = layers[-1].out_features
            self.logit = torch.empty(I, N, **f32)
            self.dlogit = torch.empty(I, N, **f32)
            wmax = max([Hl, self.Zf + cond_dim] + [l.in_features for l in layers])
            self.dec_scratch = [torch.empty(I * wmax, **f32) for _ in range(3)]

    def prepare(self):
        """Start-of-step work that does not depend on the batch: repack the fused kernel's weights
        (they changed in the last optimizer step) on the side stream, under the encoder forward."""
        if self.spatial and self.use_tc:
            eng = self.engine
            L = linear_layers(eng.model.decoder.fc_layers)
            if eng.overlap:
                with eng.fork_side():
                    ops.sdec_tc_pack_weights(L[0].weight.data, L[1].weight.data, self.w_packed)
                self._pack_pending = True
            else:
                ops.sdec_tc_pack_weights(L[0].weight.data, L[1].weight.data, self.w_packed)

    def fold_ctx(self, cond):
        """Arguments of the coordinate-transform fold, for kernels that fuse it."""
        cl = self.engine.model.decoder.coord_latent
        return dict(cfg=self.fold_cfg, cond=cond, Wc=cl.fc_coord.weight.data,
                    bc=cl.fc_coord.bias.data, Wz=cl.fc_latent.weight.data, Uv=self.Uv)

    def forward(self, z, cond, x, w, want_grad, kl=None, beta=0.0, loss_out=None, uv_ready=False,
                side_loss=False):
        """z [I,Zf], cond [I,Cd] or None, x [B,N], w [I] or None -> fills rowll/loc/ll.
        With loss_out: also loss_out += -sum_i (ll_i + beta kl_i) in the same reduction.
        side_loss: nothing later in the step reads ll / writes loss_out, so that reduction runs
        on the engine's side stream, off the critical path of the backward pass."""
        m = self.engine.model
        dec, samp = m.decoder, m.sampler_d
        if self.spatial:
            cl = dec.coord_latent
            if not uv_ready:
                ops.fold_fwd(self.fold_cfg, z, cond, cl.fc_coord.weight.data,
                             cl.fc_coord.bias.data, cl.fc_latent.weight.data, self.Uv)
            if self.use_tc:
                L = linear_layers(dec.fc_layers)
                if getattr(self, "_pack_pending", False):
                    self.engine.join_side()          # the repacked weights are ready
                    self._pack_pending = False
                ops.sdec_tc_step(self.Uv, x, w, L[0].weight.data, L[0].bias.data,
                                 L[1].weight.data, L[1].bias.data, dec.out.weight.data,
                                 dec.out.bias.data, self.rowll, self.loc, self.gUv_part,
                                 self.wgrad_part, self.I, self.Bx, m._H, m._W, m.ndim, samp.name,
                                 dec.sigmoid_out, samp.decoder_sig, want_grad,
                                 packed_w=self.w_packed)
            else:
                ops.sdec_h0_fwd(self.Uv, self.h0, m._H, m._W, m.ndim)
                hl = self.dmlp.forward(self.h0)
                ops.linear_fwd(hl, dec.out.weight.data, dec.out.bias.data, None, out=self.logit)
                ops.obs_loglik(self.logit, x, w, self.rowll, self.dlogit if want_grad else None,
                               self.loc, self.I, self.Bx, self.N, samp.name, dec.sigmoid_out,
                               samp.decoder_sig)
        else:
            self.dec_in[:, :self.Zf].copy_(z)
            if self.Cd > 0:
                self.dec_in[:, self.Zf:].copy_(cond)
            hl = self.dmlp.forward(self.dec_in)
            ops.linear_fwd(hl, dec.out.weight.data, dec.out.bias.data, None, out=self.logit)
            ops.obs_loglik(self.logit, x, w, self.rowll, self.dlogit if want_grad else None,
                           self.loc, self.I, self.Bx, self.N, samp.name, dec.sigmoid_out,
                           samp.decoder_sig)
        if loss_out is not None and side_loss and self.engine.overlap:
            with self.engine.fork_side():
                ops.elbo_reduce(self.rowll, kl, w, beta, self.ll, loss_out, True, self.I, self.N)
        elif loss_out is not None:
            ops.elbo_reduce(self.rowll, kl, w, beta, self.ll, loss_out, True, self.I, self.N)
        else:
            ops.elbo_reduce(self.rowll, None, None, 0.0, self.ll, None, False, self.I, self.N)

    def _reduce_fold_partials(self, G=None):
        eng = self.engine
        m, flat = eng.model, eng.flat
        cl = m.decoder.coord_latent
        Hd0 = cl.fc_coord.out_features
        nd = m.ndim
        LC = m._latent_dim + self.Cd
        G = ops.fold_bwd_num_partials() if G is None else G
        per = self.fold_per
        o_w, o_b = flat.offset(cl.fc_coord.weight), flat.offset(cl.fc_coord.bias)
        o_z = flat.offset(cl.fc_latent.weight) if LC > 0 else o_b + Hd0
        if o_b == o_w + Hd0 * nd and o_z == o_b + Hd0:
            # the three gradients are contiguous in the flat buffer, in partial order: one launch
            ops.reduce_partials(self.fold_part, flat.g[o_w:o_w + per], G, per, per, True, 0)
            return
        ops.reduce_partials(self.fold_part, flat.gv(cl.fc_coord.weight), G, Hd0 * nd, per, True, 0)
        ops.reduce_partials(self.fold_part, flat.gv(cl.fc_coord.bias), G, Hd0, per, True, Hd0 * nd)
        if LC > 0:
            ops.reduce_partials(self.fold_part, flat.gv(cl.fc_latent.weight), G, Hd0 * LC, per,
                                True, Hd0 * (nd + 1))

    def backward_fused(self, head, cond, w, beta, z=None):
        """Fused-decoder variant of backward() that also runs the latent backward:
        weight-gradient partial reduction, then ONE launch for dUv gather + fold backward +
        latent backward (-> head.gmu / head.gs_pre), then the fold-weight partial reduction.
        head=None (latent shared by several instances, jiVAE): only dz (self.gz) is produced
        from the instance codes `z`."""
        assert self.spatial and self.use_tc
        eng = self.engine
        m, flat = eng.model, eng.flat
        dec = m.decoder
        cl = dec.coord_latent
        L = linear_layers(dec.fc_layers)
        n_w = TC_WGRAD_FLOATS
        base = flat.offset(L[0].weight)
        # the two partial-sum reductions only feed the optimizer: side stream, concurrent with the
        # latent / encoder backward chain
        with eng.fork_side():
            ops.reduce_partials(self.wgrad_part, flat.g[base:base + n_w], self.tc_sizes.ctas,
                                n_w, TC_WGRAD_STRIDE, True)
        if head is None:
            ops.latent_side_bwd(self.fold_cfg, z, cond, cl.fc_coord.weight.data,
                                cl.fc_latent.weight.data, None, self.gUv_part, self.N, self.gz,
                                self.gcond, self.fold_part, None, None, None, None, 0.0, None, None)
        else:
            ops.latent_side_bwd(self.fold_cfg, head.z, cond, cl.fc_coord.weight.data,
                                cl.fc_latent.weight.data, None, self.gUv_part, self.N, self.gz,
                                self.gcond, self.fold_part, head.eps, head.sigma, head.s_pre, w,
                                beta, head.gmu, head.gs_pre)
        with eng.fork_side():
            self._reduce_fold_partials(self.G_side)
        return self.gz

    def backward(self, z, cond):
        """Accumulates decoder weight gradients; returns dloss/dz [I,Zf]."""
        eng = self.engine
        m, flat = eng.model, eng.flat
        dec = m.decoder
        if self.spatial:
            cl = dec.coord_latent
            if self.use_tc:
                L = linear_layers(dec.fc_layers)
                n_w = TC_WGRAD_FLOATS
                # partial layout == flat layout of (fc0.w, fc0.b, fc2.w, fc2.b, out.w, out.b)
                base = flat.offset(L[0].weight)
                ops.reduce_partials(self.wgrad_part, flat.g[base:base + n_w], self.tc_sizes.ctas,
                                    n_w, TC_WGRAD_STRIDE, True)
                ops.sdec_tc_gather_gUv(self.gUv_part, self.gUv, self.I, self.N)
            else:
                hl = self.dmlp.h[-1] if self.dmlp.layers else self.h0
                d_hl = self.dec_scratch[2][:hl.numel()].view_as(hl)
                ops.linear_bwd(hl, dec.out.weight.data, None, None, self.dlogit, self.dlogit, d_hl,
                               False, flat.gv(dec.out.weight), flat.gv(dec.out.bias), None)
                dh0 = self.dmlp.backward(d_hl, self.dec_scratch, True)
                ops.sdec_h0_bwd(dh0, self.h0, self.gUv, m._H, m._W, m.ndim)
            ops.fold_bwd(self.fold_cfg, z, cond, cl.fc_coord.weight.data,
                         cl.fc_latent.weight.data, self.gUv, self.gz, self.gcond, self.fold_part)
            self._reduce_fold_partials()
        else:
            hl = self.dmlp.h[-1]
            d_hl = self.dec_scratch[2][:hl.numel()].view_as(hl)
            ops.linear_bwd(hl, dec.out.weight.data, None, None, self.dlogit, self.dlogit, d_hl,
                           False, flat.gv(dec.out.weight), flat.gv(dec.out.bias), None)
            d_in = self.dmlp.backward(d_hl, self.dec_scratch, True)
            self.gz.copy_(d_in[:, :self.Zf])
            if self.gcond is not None:
                self.gcond.copy_(d_in[:, self.Zf:])
        return self.gz


class GaussHead:
    """fc11 / fc12 heads + reparameterised sample + their backward, for M rows."""

    def __init__(self, engine, enc, M, Z):
        f32 = dict(device=engine.device, dtype=torch.float32)
        self.engine, self.enc, self.M, self.Z = engine, enc, M, Z
        self.eps = torch.zeros(M, Z, **f32)
        self.mu = torch.empty(M, Z, **f32)
        self.s_pre = torch.empty(M, Z, **f32)
        self.sigma = torch.empty(M, Z, **f32)
        self.z = torch.empty(M, Z, **f32)
        self.kl = torch.empty(M, **f32)
        self.gmu = torch.empty(M, Z, **f32)
        self.gs_pre = torch.empty(M, Z, **f32)
        self.dh = torch.empty(M, enc.fc11.in_features, **f32)

    def forward(self, h, gen_eps, n_rank_elems=None):
        eng, enc = self.engine, self.enc
        if gen_eps:
            ops.randn(self.eps, eng.seed, eng.step_counter, eng.eps_first_index(self.eps.numel()))
        ops.linear_fwd(h, enc.fc11.weight.data, enc.fc11.bias.data, None, out=self.mu)
        ops.linear_fwd(h, enc.fc12.weight.data, enc.fc12.bias.data, None, out=self.s_pre)
        ops.latent_fwd(self.mu, self.s_pre, self.eps, self.sigma, self.z, self.kl)

    def backward(self, h, gz, w, beta):
        """-> self.dh = dloss/dh through both heads."""
        flat, enc = self.engine.flat, self.enc
        ops.latent_bwd(gz, self.eps, self.sigma, self.s_pre, self.z, w, beta, self.gmu, self.gs_pre)
        ops.linear_bwd(h, enc.fc11.weight.data, None, None, self.gmu, self.gmu, self.dh, False,
                       flat.gv(enc.fc11.weight), flat.gv(enc.fc11.bias), None)
        ops.linear_bwd(h, enc.fc12.weight.data, None, None, self.gs_pre, self.gs_pre, self.dh,
                       True, flat.gv(enc.fc12.weight), flat.gv(enc.fc12.bias), None)
        return self.dh


def _mlp_scratch(layers, M, device, with_input=False):
    """ping-pong buffers for MLP.backward (with_input: room for the input gradient too)"""
    first = 0 if with_input else 1
    wmax = max([l.in_features for l in layers[first:]] + [layers[-1].out_features, 1])
    return [torch.empty(M * wmax, device=device, dtype=torch.float32) for _ in range(2)]


class SpatialVAEProgram(StepProgram):
    """iVAE (Trace_ELBO) step: reference models/ivae.py:165-221.  Also the
    supervised ssiVAE step (ys observed): same trace plus the constant
    log p(y) = log(1/K) per sample (models/ssivae.py:153-215)."""

    def __init__(self, engine, B, has_y, cond_dim=None, loss_const=0.0):
        super().__init__(engine, B, has_y)
        m = engine.model
        dev, flat = engine.device, engine.flat
        self.N = m._n_pix
        self.Z = m.z_dim
        c_model = m.c_dim if cond_dim is None else cond_dim
        self.C = c_model if has_y else 0
        if c_model > 0 and not has_y:
            raise ValueError("model was built with c_dim={} but no y was passed".format(c_model))
        self.loss_const = loss_const
        # the ELBO reduction may run on the side stream only if nothing else in the step
        # accumulates into the loss slot (SsRegProgram adds log-prob terms of y: it may not)
        self.side_loss = True
        N, Z, C = self.N, self.Z, self.C
        f32 = dict(device=dev, dtype=torch.float32)
        self.enc_in = torch.zeros(B, N + C, **f32)
        self.x = torch.zeros(B, N, **f32) if C > 0 else self.enc_in
        self.y = torch.zeros(B, C, **f32) if C > 0 else None
        enc = m.encoder_z
        self.dec = DecoderOps(engine, B, B, C)
        # a convolutional encoder installed with set_encoder (reference models/base.py:173-176;
        # BASELINE configs[1] words the iVAE as "conv encoder / fc decoder")
        self.conv_enc = hasattr(enc, "feature_extractor")
        if self.conv_enc:
            from .conv_engine import ConvGaussEncoder
            if C > 0:
                raise NotImplementedError("pyroved_b200: a convolutional encoder_z cannot take the "
                                          "conditioning vector y (the reference's convEncoderNet "
                                          "has no such input either)")
            self.head = ConvGaussEncoder(engine, enc, B, Z)
            if _numel_shape(self.head.in_shape) != N:
                raise ValueError("encoder input {} does not match data_dim {}".format(
                    self.head.in_shape, m._data_dim))
            self.fused = False
            return
        self.head = GaussHead(engine, enc, B, Z)
        enc_layers = linear_layers(enc.fc_layers)
        self.fused = (not engine.force_generic and
                      FusedStack.eligible(enc_layers, [enc.fc11, enc.fc12], B))
        if self.fused:
            fold = self.dec.fold_ctx(self.y) if self.dec.spatial else None
            self.enc = FusedStack(engine, enc_layers, enc.activation, [enc.fc11, enc.fc12], B,
                                  gauss_head=self.head, fold=fold)
        else:
            self.enc = MLP(enc_layers, enc.activation, B, dev, flat)
            self.enc_scratch = _mlp_scratch(self.enc.layers, B, dev)

    # convenient aliases (tests / inference read these)
    eps = property(lambda s: s.head.eps)
    mu = property(lambda s: s.head.mu)
    sigma = property(lambda s: s.head.sigma)
    z = property(lambda s: s.head.z)
    loc = property(lambda s: s.dec.loc)
    ll = property(lambda s: s.dec.ll)
    use_tc = property(lambda s: s.dec.use_tc)

    def load(self, x, y):
        B, N = self.B, self.N
        x = x.reshape(B, -1)
        if x.shape[1] != N:
            raise ValueError("expected {} features per sample, got {}".format(N, x.shape[1]))
        if self.C > 0:
            self.x.copy_(x, non_blocking=True)
            self.y.copy_(y.reshape(B, -1), non_blocking=True)
            self.enc_in[:, :N].copy_(self.x)
            self.enc_in[:, N:].copy_(self.y)
        else:
            self.enc_in.copy_(x, non_blocking=True)

    def forward(self, beta, want_grad, gen_eps):
        flat = self.engine.flat
        self.dec.prepare()
        if self.fused:
            self.enc.forward(self.enc_in, gen_eps)
            self.dec.forward(self.head.z, self.y, self.x, None, want_grad, kl=self.head.kl,
                             beta=float(beta), loss_out=flat.loss, uv_ready=self.dec.spatial,
                             side_loss=self.side_loss)
            return
        if self.conv_enc:
            self.head.forward(self.enc_in, gen_eps)
        else:
            h = self.enc.forward(self.enc_in)
            self.head.forward(h, gen_eps)
        self.dec.forward(self.head.z, self.y, self.x, None, want_grad, kl=self.head.kl,
                         beta=float(beta), loss_out=flat.loss)

    def backward(self, beta):
        if self.conv_enc:
            gz = self.dec.backward(self.head.z, self.y)
            self.head.backward(gz, None, beta)
            return
        if self.fused and self.dec.spatial and self.dec.use_tc:
            self.dec.backward_fused(self.head, self.y, None, beta)
            self.enc.backward([self.head.gmu, self.head.gs_pre])
            return
        gz = self.dec.backward(self.head.z, self.y)
        if self.fused:
            ops.latent_bwd(gz, self.head.eps, self.head.sigma, self.head.s_pre, self.head.z, None,
                           beta, self.head.gmu, self.head.gs_pre)
            self.enc.backward([self.head.gmu, self.head.gs_pre])
            return
        dh = self.head.backward(self.enc.h[-1], gz, None, beta)
        self.enc_dx = self.enc.backward(dh, self.enc_scratch, getattr(self, "need_enc_dx", False))


class SsRegProgram(SpatialVAEProgram):
    """ss_reg_iVAE under Trace_ELBO (reference models/ss_reg_ivae.py:152-219).
    Supervised (ys observed): the conditional iVAE trace plus log N(ys; 0, s).
    Unsupervised: y = c(x) + s eps_y is a reparameterised sample of q(y|x) = N(c(x), s); it
    conditions encoder_z and the decoder, is scored under the prior N(0, s), and its gradient
    (decoder + encoder_z + prior) flows back into the regressor."""

    def __init__(self, engine, B, has_y):
        m = engine.model
        super().__init__(engine, B, True, cond_dim=m.reg_dim)
        self.side_loss = False      # normal_logprob below adds to the loss slot on the main stream
        self.sup = has_y
        self.sig = float(m.reg_sig)
        self.has_y = has_y
        if has_y:
            return
        dev, flat = engine.device, engine.flat
        f32 = dict(device=dev, dtype=torch.float32)
        C = m.reg_dim
        reg = m.encoder_y
        layers = linear_layers(reg.fc_layers)
        self.eps_y = torch.zeros(B, C, **f32)
        self.gy = torch.empty(B, C, **f32)
        self.dec.gcond = torch.empty(B, C, **f32)
        self.need_enc_dx = True
        if not self.fused:
            self.enc_scratch = _mlp_scratch(self.enc.layers, B, dev, with_input=True)
        self.reg_fused = (not engine.force_generic) and FusedStack.eligible(layers, [reg.out], B)
        if self.reg_fused:
            self.reg = FusedStack(engine, layers, reg.activation, [reg.out], B)
            self.c = self.reg.hout[0]
        else:
            self.reg = MLP(layers, reg.activation, B, dev, flat)
            self.reg_scratch = _mlp_scratch(layers, B, dev)
            self.c = torch.empty(B, C, **f32)
            self.dh_r = torch.empty(B, reg.out.in_features, **f32)

    def grad_modules(self):
        m = self.engine.model
        return [m.encoder_z, m.decoder] + ([] if self.sup else [m.encoder_y])

    def set_eps(self, eps):
        if isinstance(eps, dict):
            self.eps.copy_(eps["z"].reshape(self.eps.shape), non_blocking=True)
            if not self.sup:
                self.eps_y.copy_(eps["y"].reshape(self.eps_y.shape), non_blocking=True)
        else:
            super().set_eps(eps)

    def load(self, x, y):
        if self.sup:
            return super().load(x, y)
        B, N = self.B, self.N
        self.x.copy_(x.reshape(B, -1), non_blocking=True)
        self.enc_in[:, :N].copy_(self.x)

    def forward(self, beta, want_grad, gen_eps):
        eng = self.engine
        flat = eng.flat
        if not self.sup:
            reg = eng.model.encoder_y
            if self.reg_fused:
                self.reg.forward(self.x)
            else:
                h = self.reg.forward(self.x)
                ops.linear_fwd(h, reg.out.weight.data, reg.out.bias.data, None, out=self.c)
            if gen_eps:
                ops.randn(self.eps_y, eng.seed + 0x5bd1e995, eng.step_counter,
                          eng.eps_first_index(self.eps_y.numel()))
            ops.axpy_out(self.c, self.eps_y, self.sig, self.y)        # y = c + s eps_y
            self.enc_in[:, self.N:].copy_(self.y)
        super().forward(beta, want_grad, gen_eps)
        ops.normal_logprob(self.y, None, self.sig, -1.0, flat.loss)   # - log p(y)
        if not self.sup:
            ops.normal_logprob(self.y, self.c, self.sig, 1.0, flat.loss)   # + log q(y|x)

    def backward(self, beta):
        super().backward(beta)
        if self.sup:
            return
        eng = self.engine
        flat = eng.flat
        reg = eng.model.encoder_y
        # dloss/dy = decoder path + encoder_z path + prior term y / s^2
        if self.fused:
            W0 = self.enc.layers[0].weight.data
            ops.linear_dx_cols(self.enc.dpre[0], W0, self.gy, self.N)
        else:
            self.gy.copy_(self.enc_dx[:, self.N:])
        ops.axpy_out(self.gy, self.dec.gcond, 1.0, self.gy)
        ops.axpy_out(self.gy, self.y, 1.0 / (self.sig * self.sig), self.gy)
        if self.reg_fused:
            self.reg.backward([self.gy])
        else:
            hr = self.reg.h[-1]
            ops.linear_bwd(hr, reg.out.weight.data, None, None, self.gy, self.gy, self.dh_r, False,
                           flat.gv(reg.out.weight), flat.gv(reg.out.bias), None)
            self.reg.backward(self.dh_r, self.reg_scratch, False)


class RegressorAuxProgram(StepProgram):
    """ss_reg_iVAE auxiliary step (reference models/ss_reg_ivae.py:229-242):
    loss = -mult * sum log N(ys; c(x), s); no sites when ys is None."""

    def __init__(self, engine, B, has_y):
        super().__init__(engine, B, has_y)
        m = engine.model
        dev, flat = engine.device, engine.flat
        f32 = dict(device=dev, dtype=torch.float32)
        self.N, self.C = m._n_pix, m.reg_dim
        self.sig = float(m.reg_sig)
        self.x = torch.zeros(B, self.N, **f32)
        self.y = torch.zeros(B, self.C, **f32)
        self.gc = torch.empty(B, self.C, **f32)
        self.eps = torch.zeros(1, **f32)
        reg = m.encoder_y
        layers = linear_layers(reg.fc_layers)
        self.reg_fused = (not engine.force_generic) and FusedStack.eligible(layers, [reg.out], B)
        if self.reg_fused:
            self.reg = FusedStack(engine, layers, reg.activation, [reg.out], B)
            self.c = self.reg.hout[0]
        else:
            self.reg = MLP(layers, reg.activation, B, dev, flat)
            self.reg_scratch = _mlp_scratch(layers, B, dev)
            self.c = torch.empty(B, self.C, **f32)
            self.dh_r = torch.empty(B, reg.out.in_features, **f32)

    def grad_modules(self):
        return [self.engine.model.encoder_y] if self.has_y else []

    def load(self, x, y):
        self.x.copy_(x.reshape(self.B, -1), non_blocking=True)
        if y is not None:
            self.y.copy_(y.reshape(self.B, -1), non_blocking=True)

    def forward(self, mult, want_grad, gen_eps):
        if not self.has_y:
            return
        reg = self.engine.model.encoder_y
        if self.reg_fused:
            self.reg.forward(self.x)
        else:
            h = self.reg.forward(self.x)
            ops.linear_fwd(h, reg.out.weight.data, reg.out.bias.data, None, out=self.c)
        ops.normal_logprob(self.y, self.c, self.sig, -float(mult), self.engine.flat.loss, self.gc)

    def backward(self, mult):
        if not self.has_y:
            return
        flat = self.engine.flat
        reg = self.engine.model.encoder_y
        if self.reg_fused:
            self.reg.backward([self.gc])
        else:
            hr = self.reg.h[-1]
            ops.linear_bwd(hr, reg.out.weight.data, None, None, self.gc, self.gc, self.dh_r, False,
                           flat.gv(reg.out.weight), flat.gv(reg.out.bias), None)
            self.reg.backward(self.dh_r, self.reg_scratch, False)


class EnumVAEProgram(StepProgram):
    """TraceEnum_ELBO step with one enumerated discrete latent.

    kind == "jivae"  (reference models/jivae.py:152-220): encoder(x) -> mu, sigma,
        alpha; the continuous code is shared by the K enumerated classes.
    kind == "ssivae" (unsupervised; reference models/ssivae.py:153-215):
        alpha = classifier(x); encoder_z([x, onehot_k]) per class.
    Instance i = k*B + b; alpha_bk weights every downstream cost.
    """

    def __init__(self, engine, B, kind):
        super().__init__(engine, B, False)
        m = engine.model
        dev, flat = engine.device, engine.flat
        self.kind = kind
        self.N, self.Z = m._n_pix, m.z_dim
        K = m.discrete_dim if kind == "jivae" else m.num_classes
        self.K = K
        I = K * B
        self.I = I
        N, Z = self.N, self.Z
        f32 = dict(device=dev, dtype=torch.float32)
        self.x = torch.zeros(B, N, **f32)
        self.onehot = torch.zeros(I, K, **f32)
        self.onehot.view(K, B, K)[torch.arange(K), :, torch.arange(K)] = 1.0
        self.logits = torch.empty(B, K, **f32)
        self.alpha = torch.empty(B, K, **f32)
        self.w = torch.empty(I, **f32)
        self.glogits = torch.zeros(B * K + B, **f32)
        self.cost = torch.empty(I, **f32)
        enc = m.encoder_z
        enc_layers = linear_layers(enc.fc_layers)
        self.dec = DecoderOps(engine, I, B, K)
        ok = not engine.force_generic
        self.glog_view = self.glogits[:B * K].view(B, K)
        if kind == "jivae":
            self.head = GaussHead(engine, enc, B, Z)
            self.z_rep = torch.empty(I, Z, **f32)
            self.gz_b = torch.empty(B, Z, **f32)
            heads = [enc.fc11, enc.fc12, enc.fc13]
            self.fused = ok and FusedStack.eligible(enc_layers, heads, B)
            if self.fused:
                self.enc = FusedStack(engine, enc_layers, enc.activation, heads, B,
                                      gauss_head=self.head)
                self.logits = self.enc.hout[2]
            else:
                self.enc = MLP(enc_layers, enc.activation, B, dev, flat)
                self.enc_scratch = _mlp_scratch(self.enc.layers, B, dev)
        else:
            self.enc_in = torch.zeros(I, N + K, **f32)
            self.enc_in[:, N:].copy_(self.onehot)
            self.head = GaussHead(engine, enc, I, Z)
            cls = m.encoder_y
            cls_layers = linear_layers(cls.fc_layers)
            self.fused = (ok and FusedStack.eligible(enc_layers, [enc.fc11, enc.fc12], I)
                          and FusedStack.eligible(cls_layers, [cls.out], B))
            if self.fused:
                fold = self.dec.fold_ctx(self.onehot) if self.dec.spatial else None
                self.enc = FusedStack(engine, enc_layers, enc.activation, [enc.fc11, enc.fc12], I,
                                      gauss_head=self.head, fold=fold)
                self.cls = FusedStack(engine, cls_layers, cls.activation, [cls.out], B)
                self.logits = self.cls.hout[0]
            else:
                self.enc = MLP(enc_layers, enc.activation, I, dev, flat)
                self.enc_scratch = _mlp_scratch(self.enc.layers, I, dev)
                self.cls = MLP(cls_layers, cls.activation, B, dev, flat)
                self.cls_scratch = _mlp_scratch(self.cls.layers, B, dev)
                self.dh_c = torch.empty(B, cls.out.in_features, **f32)

    eps = property(lambda s: s.head.eps)
    mu = property(lambda s: s.head.mu)
    sigma = property(lambda s: s.head.sigma)
    loc = property(lambda s: s.dec.loc)
    ll = property(lambda s: s.dec.ll)
    use_tc = property(lambda s: s.dec.use_tc)

    def grad_modules(self):
        m = self.engine.model
        mods = [m.encoder_z, m.decoder]
        if self.kind == "ssivae":
            mods.append(m.encoder_y)
        return mods

    def load(self, x, y):
        B, N, K = self.B, self.N, self.K
        x = x.reshape(B, -1)
        if x.shape[1] != N:
            raise ValueError("expected {} features per sample, got {}".format(N, x.shape[1]))
        self.x.copy_(x, non_blocking=True)
        if self.kind == "ssivae":
            self.enc_in.view(K, B, N + K)[:, :, :N].copy_(self.x.unsqueeze(0).expand(K, B, N))

    def forward(self, beta, want_grad, gen_eps):
        m = self.engine.model
        flat = self.engine.flat
        K, B, Z = self.K, self.B, self.Z
        self.dec.prepare()
        if self.kind == "jivae":
            b0, b1 = beta
            enc = m.encoder_z
            if self.fused:
                self.enc.forward(self.x, gen_eps)      # hidden stack + 3 heads + sample: 2 launches
            else:
                h = self.enc.forward(self.x)
                self.head.forward(h, gen_eps)
                ops.linear_fwd(h, enc.fc13.weight.data, enc.fc13.bias.data, None, out=self.logits)
            ops.enum_head_fwd(self.logits, self.alpha, self.w)
            self.z_rep.view(K, B, Z).copy_(self.head.z.unsqueeze(0).expand(K, B, Z))
            self.dec.forward(self.z_rep, self.onehot, self.x, self.w, want_grad)
            ops.weighted_sum(self.head.kl, None, -float(b0), flat.loss)
            ops.enum_head_bwd(self.alpha, self.dec.ll, b1, self.glogits, flat.loss)
        else:
            cls = m.encoder_y
            if self.fused:
                self.cls.forward(self.x)
                ops.enum_head_fwd(self.logits, self.alpha, self.w)
                self.enc.forward(self.enc_in, gen_eps)
                self.dec.forward(self.head.z, self.onehot, self.x, self.w, want_grad,
                                 uv_ready=self.dec.spatial)
            else:
                hc = self.cls.forward(self.x)
                ops.linear_fwd(hc, cls.out.weight.data, cls.out.bias.data, None, out=self.logits)
                ops.enum_head_fwd(self.logits, self.alpha, self.w)
                h = self.enc.forward(self.enc_in)
                self.head.forward(h, gen_eps)
                self.dec.forward(self.head.z, self.onehot, self.x, self.w, want_grad)
            ops.axpy_out(self.dec.ll, self.head.kl, float(beta), self.cost)
            ops.enum_head_bwd(self.alpha, self.cost, 1.0, self.glogits, flat.loss)

    def backward(self, beta):
        m = self.engine.model
        flat = self.engine.flat
        K, B, Z = self.K, self.B, self.Z
        if self.kind == "jivae":
            b0, b1 = beta
            enc = m.encoder_z
            tc = self.dec.spatial and self.dec.use_tc
            if self.fused and tc:
                gz = self.dec.backward_fused(None, self.onehot, None, 0.0, z=self.z_rep)
            else:
                gz = self.dec.backward(self.z_rep, self.onehot)
            ops.reduce_partials(gz, self.gz_b, K, B * Z, B * Z, False)
            if self.fused:
                hd = self.head
                ops.latent_bwd(self.gz_b, hd.eps, hd.sigma, hd.s_pre, hd.z, None, b0, hd.gmu,
                               hd.gs_pre)
                self.enc.backward([hd.gmu, hd.gs_pre, self.glog_view])
                return
            h = self.enc.h[-1]
            dh = self.head.backward(h, self.gz_b, None, b0)
            ops.linear_bwd(h, enc.fc13.weight.data, None, None, self.glogits, self.glogits, dh,
                           True, flat.gv(enc.fc13.weight), flat.gv(enc.fc13.bias), None)
            self.enc.backward(dh, self.enc_scratch, False)
        else:
            cls = m.encoder_y
            if self.fused:
                hd = self.head
                if self.dec.spatial and self.dec.use_tc:
                    self.dec.backward_fused(hd, self.onehot, self.w, beta)
                else:
                    gz = self.dec.backward(hd.z, self.onehot)
                    ops.latent_bwd(gz, hd.eps, hd.sigma, hd.s_pre, hd.z, self.w, beta, hd.gmu,
                                   hd.gs_pre)
                self.enc.backward([hd.gmu, hd.gs_pre])
                self.cls.backward([self.glog_view])
                return
            gz = self.dec.backward(self.head.z, self.onehot)
            dh = self.head.backward(self.enc.h[-1], gz, self.w, beta)
            self.enc.backward(dh, self.enc_scratch, False)
            hc = self.cls.h[-1]
            ops.linear_bwd(hc, cls.out.weight.data, None, None, self.glogits, self.glogits,
                           self.dh_c, False, flat.gv(cls.out.weight), flat.gv(cls.out.bias), None)
            self.cls.backward(self.dh_c, self.cls_scratch, False)


class ClassifierAuxProgram(StepProgram):
    """ssiVAE auxiliary step (reference models/ssivae.py:229-248):
    loss = -mult * sum_b log Cat(y_b | classifier(x_b)); no sites when ys is None."""

    def __init__(self, engine, B, has_y):
        super().__init__(engine, B, has_y)
        m = engine.model
        dev, flat = engine.device, engine.flat
        f32 = dict(device=dev, dtype=torch.float32)
        self.N, self.K = m._n_pix, m.num_classes
        self.x = torch.zeros(B, self.N, **f32)
        self.y = torch.zeros(B, self.K, **f32)
        cls = m.encoder_y
        layers = linear_layers(cls.fc_layers)
        self.glogits = torch.zeros(B * self.K + B, **f32)
        self.glog_view = self.glogits[:B * self.K].view(B, self.K)
        self.eps = torch.zeros(1, **f32)
        self.fused = (not engine.force_generic) and FusedStack.eligible(layers, [cls.out], B)
        if self.fused:
            self.cls = FusedStack(engine, layers, cls.activation, [cls.out], B)
            self.logits = self.cls.hout[0]
        else:
            self.cls = MLP(layers, cls.activation, B, dev, flat)
            self.cls_scratch = _mlp_scratch(self.cls.layers, B, dev)
            self.logits = torch.empty(B, self.K, **f32)
            self.dh_c = torch.empty(B, cls.out.in_features, **f32)

    def grad_modules(self):
        return [self.engine.model.encoder_y] if self.has_y else []

    def load(self, x, y):
        self.x.copy_(x.reshape(self.B, -1), non_blocking=True)
        if y is not None:
            self.y.copy_(y.reshape(self.B, -1), non_blocking=True)

    def forward(self, mult, want_grad, gen_eps):
        if not self.has_y:
            return
        cls = self.engine.model.encoder_y
        if self.fused:
            self.cls.forward(self.x)
        else:
            hc = self.cls.forward(self.x)
            ops.linear_fwd(hc, cls.out.weight.data, cls.out.bias.data, None, out=self.logits)
        ops.class_nll(self.logits, self.y, mult, self.glogits, self.engine.flat.loss)

    def backward(self, mult):
        if not self.has_y:
            return
        flat = self.engine.flat
        cls = self.engine.model.encoder_y
        if self.fused:
            self.cls.backward([self.glog_view])
            return
        hc = self.cls.h[-1]
        ops.linear_bwd(hc, cls.out.weight.data, None, None, self.glogits, self.glogits, self.dh_c,
                       False, flat.gv(cls.out.weight), flat.gv(cls.out.bias), None)
        self.cls.backward(self.dh_c, self.cls_scratch, False)


class SVIEngine:
    """Drop-in for the object the reference keeps in `SVItrainer.svi`
    (`pyro.infer.SVI`): `.step(x[, y], **kwargs) -> float` runs one optimisation
    step on a mini-batch and returns the (batch-sum) loss."""

    def __init__(self, model, lr=1e-3, enumerate_parallel=False, seed=1, device=None,
                 use_graphs=None, force_generic=None, data_parallel=None):
        if device is None:
            device = getattr(model, "device", "cuda")
        self.device = torch.device(device if str(device) != "cuda" else "cuda:{}".format(
            torch.cuda.current_device() if torch.cuda.is_available() else 0))
        if self.device.type != "cuda":
            raise RuntimeError(
                "pyroved_b200 runs on CUDA devices only (no CPU fallback); got device={!r}".format(
                    str(device)))
        self.model = model
        self.lr = float(lr)
        self.enumerate_parallel = enumerate_parallel
        self.seed = int(seed)
        # side stream for work that only feeds the optimizer / the loss read-back (forked and
        # joined inside the step, so it becomes a parallel branch of the step's CUDA graph)
        self.overlap = os.environ.get("PVB_SIDE_STREAM", "1") != "0"
        self.side = torch.cuda.Stream(self.device)
        self._side_used = False
        self.peer = None          # parallel.PeerExchange (staging buffers of the NVLink exchange)
        # data_parallel=False: a purely local engine even inside an initialised process group
        # (e.g. a single-GPU cross-check next to a data-parallel run); None: follow the group
        self.data_parallel = data_parallel is not False
        self.flat = FlatParams(model, self.device)
        # the optimizer kernels leave the gradient buffer zeroed; anything else (loss_and_grads,
        # evaluate_loss) leaves it dirty and the next step clears it first
        self._g_dirty = False
        self._make_peer()
        self.step_counter = torch.zeros(1, device=self.device, dtype=torch.int32)
        self.adam_ticket = torch.zeros(1, device=self.device, dtype=torch.int32)
        # pinned host ring the optimizer kernel writes each step's loss into (slot = step count &
        # (LOSS_RING - 1)): the step's result reaches the host without a copy call
        self.loss_ring = torch.zeros(ops.LOSS_RING, dtype=torch.float32).pin_memory()
        self.updates_done = 0     # host mirror of step_counter
        self.programs = {}
        self.graphs = {}
        if use_graphs is None:
            use_graphs = os.environ.get("PVB_CUDA_GRAPHS", "1") != "0"
        self.use_graphs = use_graphs
        # force_generic=True: the exact fp32 kernels everywhere (no tcgen05 / fused small-batch
        # stacks); None: taken from PVB_FORCE_GENERIC once, here
        if force_generic is None:
            force_generic = os.environ.get("PVB_FORCE_GENERIC", "0") == "1"
        self.force_generic = bool(force_generic)
        self.launches_per_step = 0
        self.last_loss_const = 0.0
        # data-parallel state (pyroved_b200.parallel)
        self.world_size = 1
        self.rank = 0
        self.process_group = None
        self._attach_distributed()

    # ---- distributed -----------------------------------------------------
    def _attach_distributed(self):
        if self.data_parallel:
            self.rank, self.world_size = parallel.rank_world()

    def _make_peer(self):
        """Staging buffers of the fused NVLink exchange, sized to the flat gradient buffer
        (collective: every rank builds / rebuilds its engine at the same point).  Falls back to
        NCCL's all-reduce, with a warning, where symmetric memory is unavailable."""
        self.peer = None
        if not (self.data_parallel and parallel.peer_exchange_enabled()):
            return
        try:
            self.peer = parallel.PeerExchange(self.flat.total + 4, self.device)
        except Exception as err:          # noqa: BLE001 -- any failure means "use NCCL"
            warnings.warn("pyroved_b200: symmetric-memory exchange unavailable ({}); using the "
                          "NCCL all-reduce".format(err))
            self.peer = None

    def _update_exchange(self):
        """gradient all-reduce + Adam in one kernel over NVLink peer memory"""
        flat, pe = self.flat, self.peer
        ops.peer_allreduce_adam(flat.p, flat.m, flat.v, flat.g, flat.total, pe.stage_ptrs,
                                pe.peer_flags, pe.state, pe.rank, pe.world, self.lr,
                                self.step_counter, flat.first_step, loss_ring=self.loss_ring,
                                two_shot=pe.two_shot)

    def eps_first_index(self, n_local):
        """Global index of this rank's first noise element: the noise of a
        sample depends on its GLOBAL position, so the ELBO does not depend on
        the number of GPUs."""
        return parallel.noise_first_index(self.rank, n_local)

    # ---- program selection -------------------------------------------------
    def tc_eligible(self, dec, N):
        if self.force_generic or not ops.has_tcgen05():
            return False
        layers = linear_layers(dec.fc_layers)
        return (len(layers) == 2 and all(l.in_features == 128 and l.out_features == 128
                                         for l in layers)
                and dec.coord_latent.fc_coord.out_features == 128
                and dec.activation == "tanh" and N >= 32
                and self.model.sampler_d.name in ("bernoulli", "gaussian", "continuous_bernoulli"))

    def _program(self, B, has_y, mode="main"):
        key = (B, has_y, mode)
        prog = self.programs.get(key)
        if prog is None:
            prog = self.model._make_program(self, B, has_y, mode)
            self.programs[key] = prog
        return prog

    # ---- one step -------------------------------------------------------------
    def _run(self, prog, beta, train, gen_eps, update):
        # (the gradient buffer is clean here: the optimizer kernels zero it as they consume it,
        # _step_on_device clears it after a call that left gradients behind)
        prog.forward(beta, train, gen_eps)
        if train:
            prog.backward(beta)
        self.join_side()
        if update:
            self._update()

    def fork_side(self):
        """Context: launches go to the side stream, ordered after everything queued so far on
        the current stream (PVB_SIDE_STREAM=0: they stay on the current stream)."""
        cur = torch.cuda.current_stream(self.device)
        if not self.overlap:
            return torch.cuda.stream(cur)
        self.side.wait_stream(cur)
        self._side_used = True
        return torch.cuda.stream(self.side)

    def join_side(self):
        if self._side_used:
            torch.cuda.current_stream(self.device).wait_stream(self.side)
            self._side_used = False

    def _update(self):
        flat = self.flat
        ops.adam_flat_step(flat.p, flat.g, flat.m, flat.v, flat.total, self.lr, self.step_counter,
                           self.adam_ticket, flat.first_step, loss_src=flat.loss_pair,
                           loss_ring=self.loss_ring)

    def _allreduce(self):
        parallel.allreduce_sum_(self.flat.g, self.process_group)

    def _execute(self, key, fn):
        """Eager on first sight of `key`, captured on the second, replayed after."""
        if not self.use_graphs:
            fn()
            return
        entry = self.graphs.pop(key, None)
        if entry is None:
            fn()
            self.graphs[key] = "warm"
            self._trim_graphs()
            return
        if entry == "warm":
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.device)
            # No CUDA graph may be DESTROYED while a capture is under way (cudaGraphExecDestroy is not
            # permitted then and invalidates the capture).  Graphs of engines that went out of scope sit in
            # reference cycles until the cyclic collector runs -- which torch.cuda.graph no longer forces --
            # so collect them now and keep the collector off for the duration of the capture.
            guard = os.environ.get("PVB_GC_GUARD", "1") != "0"      # "0": debugging only
            if guard:
                gc.collect()
            gc_was_on = gc.isenabled()
            if guard:
                gc.disable()
            try:
                with torch.cuda.graph(g):
                    fn()
            finally:
                if gc_was_on:
                    gc.enable()
            entry = g
        self.graphs[key] = entry          # most recently used last
        entry.replay()

    MAX_GRAPHS = 48

    def _trim_graphs(self):
        """The graph key carries the scale factors: under KL annealing (a new scale_factor every
        epoch) captured graphs of stale values would pile up; keep the most recently used ones."""
        while len(self.graphs) > self.MAX_GRAPHS:
            self.graphs.pop(next(iter(self.graphs)))

    def step(self, *args, **kwargs):
        return self._step(args, kwargs, train=True)

    def step_aux(self, *args, **kwargs):
        """Auxiliary (supervised) loss step of semi-supervised models
        (reference trainers/auxsvi.py:79-81,99)."""
        return self._step(args, kwargs, train=True, mode="aux")

    def evaluate_loss(self, *args, **kwargs):
        return self._step(args, kwargs, train=False, update=False)

    def loss_and_grads(self, *args, **kwargs):
        """Loss + gradients (left in every parameter's .grad) without an
        optimizer update -- used by the parity tests."""
        return self._step(args, kwargs, train=True, update=False)

    def _step(self, args, kwargs, train=True, update=True, mode="main"):
        # every launch of the step goes to this engine's device, whatever the caller's current one
        with torch.cuda.device(self.device):
            return self._step_on_device(args, kwargs, train, update, mode)

    def _step_on_device(self, args, kwargs, train, update, mode):
        self.flat.ensure() and self._invalidate()
        if self._g_dirty:
            self.flat.g[:self.flat.total + 1].zero_()
            self._g_dirty = False
        kwargs = dict(kwargs)
        x = args[0]
        y = args[1] if len(args) > 1 else None
        eps = kwargs.pop("_eps", None)
        sync = kwargs.pop("_sync", True)
        # _static: x / y live in buffers whose addresses recur (the trainer's staging slots), so the
        # copy into the program's input buffers is captured in the step's graph (keyed by address)
        static = kwargs.pop("_static", False) and self.use_graphs
        beta = self.model._beta(kwargs) if mode == "main" else self.model._aux_scale(kwargs)
        B = x.shape[0]
        prog = self._program(B, y is not None, mode)
        if not static:
            prog.load(x, y)
        if eps is not None and mode == "main":
            prog.set_eps(eps)
        gen_eps = eps is None
        if train:
            self.flat.activate(prog.grad_modules(), self.updates_done)
        if update:
            self.updates_done += 1
        bkey = tuple(beta) if isinstance(beta, (list, tuple)) else float(beta)
        # (model.training: BatchNorm layers follow the module's mode, conv_engine.ConvStack)
        key = (B, y is not None, mode, bkey, train, gen_eps, update, bool(self.model.training))
        if static:
            key += (x.data_ptr(), y.data_ptr() if y is not None else 0)
        if self.world_size > 1 and train and update and self.peer is not None:
            # gradients -> fused NVLink all-reduce + Adam: ONE captured graph, no NCCL call
            def whole():
                if static:
                    prog.load(x, y)
                self._run(prog, beta, True, gen_eps, False)
                self._update_exchange()
            self._execute(key + ("peer",), whole)
        elif self.world_size > 1 and train:
            if static:
                prog.load(x, y)
            self._execute(key + ("grads",), lambda: self._run(prog, beta, True, gen_eps, False))
            self._allreduce()
            if update:
                self._execute(("update",), self._update)
        else:
            def single():
                if static:
                    prog.load(x, y)
                self._run(prog, beta, train, gen_eps, update)
            self._execute(key, single)
        self.last_loss_const = prog.loss_const * self.world_size
        # the optimizer kernels move the step's loss to the "last loss" slot and clear the buffer
        self._g_dirty = not update
        out = self.flat.last_loss if update else self.flat.loss
        if not sync:
            return out                # device scalar, no host synchronisation
        return float(out.item()) + prog.loss_const * self.world_size

    def _invalidate(self):
        self.programs.clear()
        self.graphs.clear()
        self._g_dirty = False        # a rebuilt FlatParams starts from zeroed buffers
        self._make_peer()
        return True
