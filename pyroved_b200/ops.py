"""Tensor-level wrappers over the C ABI (raw device pointers + current stream).

PyTorch is used for device memory and streams only; every arithmetic op on the
hot path is one of the hand-written kernels in csrc/.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import FoldCfg, MlpChainArgs, MlpTailArgs, TcSizes, WgradProblem, check

ACT = {None: 0, "none": 0, "tanh": 1, "relu": 2, "lrelu": 3, "softplus": 4, "gelu": 5,
       "sigmoid": 6}
SAMPLER = {"bernoulli": 0, "gaussian": 1, "continuous_bernoulli": 2}
INV = {"r": 1, "t": 2, "s": 4}


def _p(t):
    """device pointer of a contiguous fp32 CUDA tensor (None -> NULL)"""
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.PvbError("pyroved_b200 kernels need CUDA tensors (no CPU fallback); got a "
                            "{} tensor".format(t.device))
    if t.device.index != torch.cuda.current_device():
        raise _lib.PvbError("tensor lives on {} but the current CUDA device is {}: run the call "
                            "under torch.cuda.device(...) (the trainers and models do this for "
                            "their own device)".format(t.device, torch.cuda.current_device()))
    if t.dtype not in (torch.float32, torch.int32) or not t.is_contiguous():
        raise _lib.PvbError("expected a contiguous float32 tensor, got {} {}".format(
            t.dtype, "non-contiguous" if not t.is_contiguous() else ""))
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def linear_fwd(x, W, b, act, out=None, pre=None):
    M = x.numel() // x.shape[-1]
    N, K = W.shape
    assert x.shape[-1] == K, (x.shape, W.shape)
    if out is None:
        out = torch.empty(x.shape[:-1] + (N,), device=x.device, dtype=torch.float32)
    check(_lib.lib().pvb_linear_fwd(_p(x), _p(W), _p(b), _p(out), _p(pre), M, N, K, ACT[act],
                                    _stream()), "pvb_linear_fwd")
    return out


def linear_bwd(x, W, y, pre, dy, dpre_ws, dx, dx_accumulate, dW, db, act):
    M = x.numel() // x.shape[-1]
    N, K = W.shape
    check(_lib.lib().pvb_linear_bwd(_p(x), _p(W), _p(y), _p(pre), _p(dy), _p(dpre_ws), _p(dx),
                                    int(dx_accumulate), _p(dW), _p(db), M, N, K, ACT[act],
                                    _stream()), "pvb_linear_bwd")


def randn(out, seed, step_counter, first_index=0):
    check(_lib.lib().pvb_randn(_p(out), out.numel(), seed & 0xFFFFFFFFFFFFFFFF,
                               _p(step_counter), first_index, _stream()), "pvb_randn")
    return out


def latent_fwd(mu, s_pre, eps, sigma, z, kl):
    Z = mu.shape[-1]
    check(_lib.lib().pvb_latent_fwd(_p(mu), _p(s_pre), _p(eps), _p(sigma), _p(z), _p(kl),
                                    mu.numel() // Z, Z, _stream()), "pvb_latent_fwd")


def latent_bwd(gz, eps, sigma, s_pre, z, w, beta, gmu, gs_pre):
    Z = z.shape[-1]
    check(_lib.lib().pvb_latent_bwd(_p(gz), _p(eps), _p(sigma), _p(s_pre), _p(z), _p(w),
                                    float(beta), _p(gmu), _p(gs_pre), z.numel() // Z, Z,
                                    _stream()), "pvb_latent_bwd")


def make_fold_cfg(ndim, invariances, latent_dim, cond_dim, hidden, dx_prior, dy_prior, sc_prior):
    inv = 0
    for k in (invariances or []):
        inv |= INV.get(k, 0)
    return FoldCfg(ndim, inv, latent_dim, cond_dim, hidden, float(dx_prior), float(dy_prior),
                   float(sc_prior))


def fold_fwd(cfg, z, cond, Wc, bc, Wz, Uv):
    I = Uv.shape[0]
    check(_lib.lib().pvb_fold_fwd(C.byref(cfg), _p(z), _p(cond), _p(Wc), _p(bc), _p(Wz), _p(Uv), I,
                                  _stream()), "pvb_fold_fwd")


def fold_bwd_num_partials():
    return _lib.lib().pvb_fold_bwd_num_partials()


def fold_bwd(cfg, z, cond, Wc, Wz, gUv, gz, gcond, part):
    I = gUv.shape[0]
    check(_lib.lib().pvb_fold_bwd(C.byref(cfg), _p(z), _p(cond), _p(Wc), _p(Wz), _p(gUv), _p(gz),
                                  _p(gcond), _p(part), I, _stream()), "pvb_fold_bwd")


def sdec_h0_fwd(Uv, h0, H, W, ndim):
    I, _, Hd = Uv.shape
    check(_lib.lib().pvb_sdec_h0_fwd(_p(Uv), _p(h0), I, H, W, ndim, Hd, _stream()),
          "pvb_sdec_h0_fwd")


def sdec_h0_bwd(dh0, h0, gUv, H, W, ndim):
    I, _, Hd = gUv.shape
    check(_lib.lib().pvb_sdec_h0_bwd(_p(dh0), _p(h0), _p(gUv), I, H, W, ndim, Hd, _stream()),
          "pvb_sdec_h0_bwd")


def obs_loglik(logit, x, w, rowll, dlogit, loc, I, B, N, sampler, sigmoid_d, decoder_sig):
    check(_lib.lib().pvb_obs_loglik(_p(logit), _p(x), _p(w), _p(rowll), _p(dlogit), _p(loc), I, B,
                                    N, SAMPLER[sampler], int(bool(sigmoid_d)), float(decoder_sig),
                                    _stream()), "pvb_obs_loglik")


def elbo_reduce(rowll, kl, w, beta, ll, loss_out, accumulate, I, N):
    check(_lib.lib().pvb_elbo_reduce(_p(rowll), _p(kl), _p(w), float(beta), _p(ll), _p(loss_out),
                                     int(accumulate), I, N, _stream()), "pvb_elbo_reduce")


def weighted_sum(v, w, scale, loss_out):
    check(_lib.lib().pvb_weighted_sum(_p(v), _p(w), float(scale), _p(loss_out), v.numel(),
                                      _stream()), "pvb_weighted_sum")


def axpy_out(a, b, beta, out):
    check(_lib.lib().pvb_axpy_out(_p(a), _p(b), float(beta), _p(out), a.numel(), _stream()),
          "pvb_axpy_out")


def enum_head_fwd(logits, alpha, w):
    B, K = logits.shape
    check(_lib.lib().pvb_enum_head_fwd(_p(logits), _p(alpha), _p(w), B, K, _stream()),
          "pvb_enum_head_fwd")


def enum_head_bwd(alpha, cost, beta_d, glogits, loss_out):
    B, K = alpha.shape
    assert glogits.numel() >= B * K + B
    check(_lib.lib().pvb_enum_head_bwd(_p(alpha), _p(cost), float(beta_d), _p(glogits),
                                       _p(loss_out), B, K, _stream()), "pvb_enum_head_bwd")


def class_nll(logits, y_onehot, mult, glogits, loss_out):
    B, K = logits.shape
    assert glogits.numel() >= B * K + B
    check(_lib.lib().pvb_class_nll(_p(logits), _p(y_onehot), float(mult), _p(glogits),
                                   _p(loss_out), B, K, _stream()), "pvb_class_nll")


def reduce_partials(part, out, G, n, stride, accumulate, part_offset=0):
    pp = part.data_ptr() + 4 * part_offset
    check(_lib.lib().pvb_reduce_partials(pp, _p(out), G, n, stride, int(accumulate), _stream()),
          "pvb_reduce_partials")


def counter_add(counter, v):
    check(_lib.lib().pvb_counter_add(_p(counter), v, _stream()), "pvb_counter_add")


def adam_flat(p, g, m, v, n, lr, step_counter, first_step=None, beta1=0.9, beta2=0.999, eps=1e-8):
    check(_lib.lib().pvb_adam_flat(_p(p), _p(g), _p(m), _p(v), n, float(lr), beta1, beta2, eps,
                                   _p(step_counter), _p(first_step), _stream()), "pvb_adam_flat")


def has_tcgen05():
    return bool(_lib.lib().pvb_has_tcgen05())


def sdec_tc_sizes(I, N):
    s = TcSizes()
    check(_lib.lib().pvb_sdec_tc_sizes(I, N, C.byref(s)), "pvb_sdec_tc_sizes")
    return s


def sdec_tc_step(Uv, x, w, W1, b1, W2, b2, wo, bo, rowll, loc, gUv_part, wgrad_part, I, B, H, W,
                 ndim, sampler, sigmoid_d, decoder_sig, backward, packed_w=None):
    """packed_w: output of sdec_tc_pack_weights for (W1, W2): the kernel then fetches its weight
    tiles with two TMA bulk copies instead of converting the fp32 weights in every CTA."""
    check(_lib.lib().pvb_sdec_tc_step(_p(Uv), _p(x), _p(w), _p(W1), _p(b1), _p(W2), _p(b2), _p(wo),
                                      _p(bo), _p(rowll), _p(loc), _p(gUv_part), _p(wgrad_part),
                                      I, B, H, W, ndim, SAMPLER[sampler], int(bool(sigmoid_d)),
                                      float(decoder_sig), int(backward), _p(packed_w), _stream()),
          "pvb_sdec_tc_step")


def sdec_tc_packed_weights(device):
    """buffer for sdec_tc_pack_weights (two fp16 operand tiles)"""
    n = int(_lib.lib().pvb_sdec_tc_packed_weight_bytes())
    return torch.empty(n // 4, device=device, dtype=torch.float32)


def sdec_tc_pack_weights(W1, W2, packed):
    check(_lib.lib().pvb_sdec_tc_pack_weights(_p(W1), _p(W2), _p(packed), _stream()),
          "pvb_sdec_tc_pack_weights")
    return packed


def sdec_tc_gather_gUv(gUv_part, gUv, I, N):
    check(_lib.lib().pvb_sdec_tc_gather_gUv(_p(gUv_part), _p(gUv), I, N, _stream()),
          "pvb_sdec_tc_gather_gUv")


LOSS_RING = 16     # PVB_LOSS_RING of include/pvb.h


def gather_rows(src, idx, dst):
    """dst[r] = src[idx[r]] over the leading dim (idx: int64 device tensor): on-device shuffle."""
    rows = idx.numel()
    row = src[0].numel()
    assert dst.shape[0] >= rows and dst[0].numel() == row and idx.dtype == torch.int64
    assert idx.is_cuda and idx.is_contiguous()
    check(_lib.lib().pvb_gather_rows(_p(src), idx.data_ptr(), _p(dst), rows, row, src.shape[0],
                                     _stream()), "pvb_gather_rows")
    return dst


def _host_ptr(t):
    """device-visible address of a pinned host tensor (unified addressing), None -> NULL"""
    if t is None:
        return None
    if not t.is_pinned():
        raise _lib.PvbError("loss ring must be pinned host memory")
    return t.data_ptr()


def adam_flat_step(p, g, m, v, n, lr, step_counter, ticket, first_step=None, beta1=0.9, beta2=0.999,
                   eps=1e-8, loss_src=None, loss_ring=None):
    """Adam update + step-counter increment in one launch.  g[:n] is zeroed as it is consumed.
    loss_src: two floats {this step's loss accumulator, last loss}: the accumulator moves to the
    second slot and is cleared; with loss_ring (LOSS_RING pinned host floats) the loss is also
    written to ring slot (new step count & (LOSS_RING - 1))."""
    check(_lib.lib().pvb_adam_flat_step(_p(p), _p(g), _p(m), _p(v), n, float(lr), beta1, beta2, eps,
                                        _p(step_counter), _p(first_step), _p(ticket), _p(loss_src),
                                        _host_ptr(loss_ring), _stream()),
          "pvb_adam_flat_step")


def peer_allreduce_adam(p, m, v, g, n, stage_ptrs, peer_flags, state, rank, world, lr, step_counter,
                        first_step=None, beta1=0.9, beta2=0.999, eps=1e-8, loss_ring=None,
                        two_shot=False):
    """All-reduce(SUM) of every rank's [n gradients | loss] buffer over NVLink peer memory fused
    with the Adam update (csrc/pvb_peer.cu).  g: the local gradient buffer (n + 4 floats; zeroed
    as it is consumed, g[n+1] receives the global loss); stage_ptrs: int64 device tensor of
    2 * world pointers (staging buffers by parity and rank); peer_flags: int64 device tensor of
    the ranks' flag blocks."""
    check(_lib.lib().pvb_peer_allreduce_adam(
        _p(p), _p(m), _p(v), _p(g), n, stage_ptrs.data_ptr(), peer_flags.data_ptr(), _p(state),
        int(rank), int(world), int(bool(two_shot)), float(lr), beta1, beta2, eps, _p(step_counter),
        _p(first_step), _host_ptr(loss_ring), _stream()), "pvb_peer_allreduce_adam")


def peer_state_words():
    return int(_lib.lib().pvb_peer_state_words())


def peer_flag_words():
    return int(_lib.lib().pvb_peer_flag_words())


# ---- fused small-batch MLP kernels (argument blocks are built once per program) ---------
def _ptrs(ctype_array, tensors):
    for i, t in enumerate(tensors):
        ctype_array[i] = _p(t)


def make_mlp_tail_args(M, h_in, layers, hs, pres, act, heads, houts, gauss=None, fold=None):
    """layers / heads: lists of nn.Linear; hs / pres / houts: preallocated outputs.
    gauss: dict(eps, sigma, z, kl, gen_eps, seed, step_counter, first_index) or None.
    fold : dict(cfg, cond, Wc, bc, Wz, Uv) or None."""
    a = MlpTailArgs()
    a.M = M
    a.n_layers = len(layers)
    a.w_in = h_in.shape[1]
    a.h_in = _p(h_in)
    for i, l in enumerate(layers):
        a.width[i] = l.out_features
        a.W[i] = _p(l.weight.data)
        a.b[i] = _p(l.bias.data) if l.bias is not None else None
        a.h[i] = _p(hs[i])
        a.pre[i] = _p(pres[i])
    a.act = ACT[act]
    a.n_heads = len(heads)
    for i, l in enumerate(heads):
        a.hdim[i] = l.out_features
        a.hW[i] = _p(l.weight.data)
        a.hb[i] = _p(l.bias.data) if l.bias is not None else None
        a.hout[i] = _p(houts[i])
    a.gauss = 0 if gauss is None else 1
    if gauss is not None:
        a.gen_eps = int(bool(gauss["gen_eps"]))
        a.eps, a.sigma, a.z, a.kl = (_p(gauss[k]) for k in ("eps", "sigma", "z", "kl"))
        a.seed = gauss["seed"] & 0xFFFFFFFFFFFFFFFF
        a.step_counter = _p(gauss["step_counter"])
        a.first_index = gauss["first_index"]
    a.fold = 0 if fold is None else 1
    if fold is not None:
        a.cfg = fold["cfg"]
        a.cond, a.Wc, a.bc, a.Wz, a.Uv = (_p(fold[k]) for k in ("cond", "Wc", "bc", "Wz", "Uv"))
    return a


def mlp_tail_fwd(args):
    check(_lib.lib().pvb_mlp_tail_fwd(C.byref(args), _stream()), "pvb_mlp_tail_fwd")


def make_mlp_chain_args(M, layers, hs, pres, act, dpres, heads, gs):
    a = MlpChainArgs()
    a.M = M
    a.n_layers = len(layers)
    for i, l in enumerate(layers):
        a.width[i] = l.out_features
        a.W[i] = _p(l.weight.data)
        a.h[i] = _p(hs[i])
        a.pre[i] = _p(pres[i])
        a.dpre[i] = _p(dpres[i])
    a.act = ACT[act]
    a.n_heads = len(heads)
    for i, l in enumerate(heads):
        a.hdim[i] = l.out_features
        a.hW[i] = _p(l.weight.data)
        a.g[i] = _p(gs[i])
    return a


def mlp_chain_bwd(args):
    check(_lib.lib().pvb_mlp_chain_bwd(C.byref(args), _stream()), "pvb_mlp_chain_bwd")


def make_wgrad_problems(items):
    """items: list of (d [M,N], x [M,K], dW [N,K], db [N] or None)"""
    arr = (WgradProblem * len(items))()
    for i, (d, x, dW, db) in enumerate(items):
        arr[i].d, arr[i].x, arr[i].dW, arr[i].db = _p(d), _p(x), _p(dW), _p(db)
        arr[i].N, arr[i].K = dW.shape
    return arr


def mlp_wgrad(problems, M):
    check(_lib.lib().pvb_mlp_wgrad(problems, len(problems), M, _stream()), "pvb_mlp_wgrad")


def latent_side_num_partials(I):
    return _lib.lib().pvb_latent_side_num_partials(I)


def latent_side_bwd(cfg, z, cond, Wc, Wz, gUv, gUv_part, N, gz, gcond, part, eps, sigma, s_pre, w,
                    beta, gmu, gs_pre):
    I = z.shape[0]
    check(_lib.lib().pvb_latent_side_bwd(C.byref(cfg), _p(z), _p(cond), _p(Wc), _p(Wz), _p(gUv),
                                         _p(gUv_part), N, _p(gz), _p(gcond), _p(part), _p(eps),
                                         _p(sigma), _p(s_pre), _p(w), float(beta), _p(gmu),
                                         _p(gs_pre), I, _stream()), "pvb_latent_side_bwd")


# ---- convolutional layers (NCHW; 1-D signals as H = 1) ------------------------------------
def _conv_dims(x, W):
    """(B, Cin, Cout, H, Wd, kh, kw) from x [B,Cin,(H,)Wd] and W [Cout,Cin,(kh,)kw]"""
    B, Cin = x.shape[0], x.shape[1]
    if x.dim() == 3:
        H, Wd, kh, kw = 1, x.shape[2], 1, W.shape[2]
    else:
        H, Wd, kh, kw = x.shape[2], x.shape[3], W.shape[2], W.shape[3]
    return B, Cin, W.shape[0], H, Wd, kh, kw


def _conv3_dims(x, W):
    """(B, Cin, Cout, D, H, Wd, k) of a volumetric layer: x [B,Cin,D,H,Wd], W [Cout,Cin,k,k,k]"""
    if not (W.shape[2] == W.shape[3] == W.shape[4]):
        raise NotImplementedError("pyroved_b200: 3-D convolutions need cubic kernels")
    return x.shape[0], x.shape[1], W.shape[0], x.shape[2], x.shape[3], x.shape[4], W.shape[2]


def _vol(x):
    """(BC, D, H, Wd) of an NCDHW tensor"""
    return x.shape[0] * x.shape[1], x.shape[2], x.shape[3], x.shape[4]


def conv_fwd(x, W, b, act, out, pre=None):
    if x.dim() == 5:
        check(_lib.lib().pvb_conv3d_fwd(_p(x), _p(W), _p(b), _p(out), _p(pre), *_conv3_dims(x, W),
                                        ACT[act], _stream()), "pvb_conv3d_fwd")
        return out
    check(_lib.lib().pvb_conv_fwd(_p(x), _p(W), _p(b), _p(out), _p(pre), *_conv_dims(x, W),
                                  ACT[act], _stream()), "pvb_conv_fwd")
    return out


def conv_bwd_data(dpre, W, dx, y_below=None, act_below=None):
    """y_below / act_below: output and activation of the layer below -- dx *= act'(y_below)"""
    if dx.dim() == 5:
        check(_lib.lib().pvb_conv3d_bwd_data(_p(dpre), _p(W), _p(dx), *_conv3_dims(dx, W), _stream()),
              "pvb_conv3d_bwd_data")
        if y_below is not None:
            act_bwd(dx, y_below, None, dx, act_below)
        return
    check(_lib.lib().pvb_conv_bwd_data(_p(dpre), _p(W), _p(dx), *_conv_dims(dx, W), _p(y_below),
                                       ACT[act_below if y_below is not None else None], _stream()),
          "pvb_conv_bwd_data")


def conv_bwd_weight(dpre, x, W, dW, db):
    if x.dim() == 5:
        check(_lib.lib().pvb_conv3d_bwd_weight(_p(dpre), _p(x), _p(dW), _p(db), *_conv3_dims(x, W),
                                               _stream()), "pvb_conv3d_bwd_weight")
        return
    check(_lib.lib().pvb_conv_bwd_weight(_p(dpre), _p(x), _p(dW), _p(db), *_conv_dims(x, W),
                                         _stream()), "pvb_conv_bwd_weight")


def act_bwd(dy, y, pre, dpre, act):
    check(_lib.lib().pvb_act_bwd(_p(dy), _p(y), _p(pre), _p(dpre), y.numel(), ACT[act], _stream()),
          "pvb_act_bwd")


def bn_workspace(C, device):
    n = _lib.lib().pvb_bn_workspace_bytes(int(C))
    return torch.empty((n + 3) // 4, device=device, dtype=torch.float32)


def _bn_dims(x):
    B, C = x.shape[0], x.shape[1]
    return B, C, (x.numel() // (B * C) if B * C else 1)


def bn_fwd(x, bn, y, save_mean, save_invstd, ws, training=None):
    """nn.BatchNorm{1,2}d module `bn` on x [B, C, *]; training=None -> bn.training.  Updates the
    module's running statistics in training mode like torch does."""
    if bn.momentum is None:
        raise NotImplementedError("pyroved_b200: BatchNorm momentum=None (cumulative average)")
    training = bn.training if training is None else training
    track = bn.track_running_stats and bn.running_mean is not None
    if not training and not track:
        training = True          # torch: no running statistics -> batch statistics in eval too
    upd = track and training
    check(_lib.lib().pvb_bn_fwd(
        _p(x), _p(bn.weight.data if bn.affine else None), _p(bn.bias.data if bn.affine else None),
        _p(bn.running_mean if (upd or not training) else None),
        _p(bn.running_var if (upd or not training) else None),
        (bn.num_batches_tracked.data_ptr() if upd and bn.num_batches_tracked.is_cuda else None), _p(y), _p(save_mean), _p(save_invstd), _p(ws),
        *_bn_dims(x), float(bn.eps), float(bn.momentum), int(bool(training)), _stream()), "pvb_bn_fwd")
    return y


def bn_uses_batch_stats(bn, training=None):
    """torch's rule: batch statistics in training mode, and in eval mode when the module tracks
    no running statistics."""
    training = bn.training if training is None else training
    track = bn.track_running_stats and bn.running_mean is not None
    return bool(training or not track)


def bn_bwd(dy, x, bn, save_mean, save_invstd, dx, dgamma, dbeta, ws, training=None):
    """Backward of bn_fwd under the same mode (training=None -> bn.training)."""
    check(_lib.lib().pvb_bn_bwd(_p(dy), _p(x), _p(bn.weight.data if bn.affine else None),
                                _p(save_mean), _p(save_invstd), _p(dx), _p(dgamma), _p(dbeta), _p(ws),
                                *_bn_dims(x), int(bn_uses_batch_stats(bn, training)), _stream()),
          "pvb_bn_bwd")
    return dx


def _plane(x):
    """(BC, H, Wd, two_d) of an NC(H)W tensor"""
    if x.dim() == 3:
        return x.shape[0] * x.shape[1], 1, x.shape[2], 0
    return x.shape[0] * x.shape[1], x.shape[2], x.shape[3], 1


def maxpool2_fwd(x, y):
    if x.dim() == 5:
        check(_lib.lib().pvb_maxpool3d_fwd(_p(x), _p(y), *_vol(x), _stream()), "pvb_maxpool3d_fwd")
        return
    check(_lib.lib().pvb_maxpool2_fwd(_p(x), _p(y), *_plane(x), _stream()), "pvb_maxpool2_fwd")


def maxpool2_bwd(x, dy, dx, act=None):
    """act: activation whose OUTPUT is x -- its derivative is folded into dx (1-D / 2-D only)"""
    if x.dim() == 5:
        check(_lib.lib().pvb_maxpool3d_bwd(_p(x), _p(dy), _p(dx), *_vol(x), _stream()),
              "pvb_maxpool3d_bwd")
        if act is not None:
            act_bwd(dx, x, None, dx, act)
        return
    check(_lib.lib().pvb_maxpool2_bwd(_p(x), _p(dy), _p(dx), *_plane(x), ACT[act], _stream()),
          "pvb_maxpool2_bwd")


def upsample2_fwd(x, y, bilinear):
    if x.dim() == 5:
        if bilinear:
            raise NotImplementedError("pyroved_b200: 3-D up-sampling is 'nearest' only")
        check(_lib.lib().pvb_upsample3d_fwd(_p(x), _p(y), *_vol(x), _stream()), "pvb_upsample3d_fwd")
        return
    check(_lib.lib().pvb_upsample2_fwd(_p(x), _p(y), *_plane(x), int(bool(bilinear)), _stream()),
          "pvb_upsample2_fwd")


def upsample2_bwd(dy, dx, bilinear, y_below=None, act_below=None):
    """y_below / act_below: output and activation of the layer below -- dx *= act'(y_below)"""
    if dx.dim() == 5:
        check(_lib.lib().pvb_upsample3d_bwd(_p(dy), _p(dx), *_vol(dx), _stream()), "pvb_upsample3d_bwd")
        if y_below is not None:
            act_bwd(dx, y_below, None, dx, act_below)
        return
    check(_lib.lib().pvb_upsample2_bwd(_p(dy), _p(dx), *_plane(dx), int(bool(bilinear)), _p(y_below),
                                       ACT[act_below if y_below is not None else None], _stream()),
          "pvb_upsample2_bwd")


def normal_logprob(y, loc, sigma, scale, loss_out, gloc=None):
    check(_lib.lib().pvb_normal_logprob(_p(y), _p(loc), float(sigma), float(scale), _p(loss_out),
                                        _p(gloc), y.numel(), _stream()), "pvb_normal_logprob")


def linear_dx_cols(dpre, W, dx_cols, col0, accumulate=False):
    M, N = dpre.shape
    K = W.shape[1]
    check(_lib.lib().pvb_linear_dx_cols(_p(dpre), _p(W), _p(dx_cols), M, N, K, col0,
                                        dx_cols.shape[1], int(accumulate), _stream()),
          "pvb_linear_dx_cols")


# ---- tensor-core convolutions (same tensors; operands converted on the fly) ----------------
def conv_tc_supported(W):
    if W.dim() == 5:
        return False
    kh, kw = (1, W.shape[2]) if W.dim() == 3 else (W.shape[2], W.shape[3])
    return bool(_lib.lib().pvb_conv_tc_supported(W.shape[1], W.shape[0], kh, kw))


def conv_tc_wgrad_supported(W):
    if W.dim() == 5:
        return False
    kh, kw = (1, W.shape[2]) if W.dim() == 3 else (W.shape[2], W.shape[3])
    return bool(_lib.lib().pvb_conv_tc_wgrad_supported(W.shape[1], W.shape[0], kh, kw))


def conv_tc_workspace(W):
    kh, kw = (1, W.shape[2]) if W.dim() == 3 else (W.shape[2], W.shape[3])
    n = _lib.lib().pvb_conv_tc_workspace_bytes(W.shape[1], W.shape[0], kh, kw)
    return torch.empty((n + 3) // 4, device=W.device, dtype=torch.float32)


def conv_tc_prep(W, ws, mode):
    """Repack W for mode 0 (forward) / 1 (backward data) into ws; the conv calls then take
    prepped=True (weights are constant within an optimizer step)."""
    kh, kw = (1, W.shape[2]) if W.dim() == 3 else (W.shape[2], W.shape[3])
    check(_lib.lib().pvb_conv_tc_prep(_p(W), _p(ws), W.shape[1], W.shape[0], kh, kw, int(mode),
                                      _stream()), "pvb_conv_tc_prep")


def conv_tc_fwd(x, W, b, act, out, ws, pre=None, prepped=False):
    check(_lib.lib().pvb_conv_tc_pix(_p(x), _p(W), _p(b), _p(out), _p(pre), _p(ws),
                                     *_conv_dims(x, W), ACT[act], 0 | (2 if prepped else 0),
                                     _stream()), "pvb_conv_tc_pix")
    return out


def conv_tc_bwd_data(dpre, W, dx, ws, y_below=None, act_below=None, prepped=False):
    """dx = conv_transpose(dpre, W); with y_below (the output of the layer below, same shape as dx)
    and its activation: dx *= act'(y_below), i.e. dx is that layer's dpre."""
    check(_lib.lib().pvb_conv_tc_pix(_p(dpre), _p(W), None, _p(dx), _p(y_below), _p(ws),
                                     *_conv_dims(dx, W), ACT[act_below if y_below is not None else None],
                                     1 | (2 if prepped else 0), _stream()), "pvb_conv_tc_pix")


def _wgrad_scratch_floats(W):
    kh, kw = (1, W.shape[2]) if W.dim() == 3 else (W.shape[2], W.shape[3])
    return (_lib.lib().pvb_conv_tc_wgrad_scratch_bytes(W.shape[1], W.shape[0], kh, kw) + 3) // 4


def conv_tc_wgrad_scratch(weights, device, shared=True):
    """Zeroed scratch for conv_tc_bwd_weight.  shared=True: ONE buffer large enough for every weight in
    `weights` (each folding call leaves it zeroed, so it serves all layers that run on one stream);
    shared=False: a list of per-layer slices of one buffer, for calls with fold=False that are folded
    together by conv_tc_wgrad_fold."""
    sizes = [_wgrad_scratch_floats(W) for W in weights]
    if shared:
        return torch.zeros(max(sizes + [4]), device=device, dtype=torch.float32)
    sizes = [(n + 3) // 4 * 4 for n in sizes]
    buf = torch.zeros(sum(sizes) + 4, device=device, dtype=torch.float32)
    out, o = [], 0
    for n in sizes:
        out.append(buf[o:o + n])
        o += n
    return out


def conv_tc_wgrad_fold(layers):
    """layers: [(scratch, W, dW, db)] of conv_tc_bwd_weight(..., fold=False) calls; adds every scratch copy
    into its dW / db in one launch and leaves the scratches zeroed."""
    if not layers:
        return
    arr = (_lib.WgradFold * len(layers))()
    for i, (sc, W, dW, db) in enumerate(layers):
        taps = W.shape[2] if W.dim() == 3 else W.shape[2] * W.shape[3]
        arr[i] = _lib.WgradFold(_p(sc), _p(dW), _p(db), W.shape[1], W.shape[0], taps)
    check(_lib.lib().pvb_conv_tc_wgrad_fold(arr, len(layers), _stream()), "pvb_conv_tc_wgrad_fold")


def conv_tc_bwd_weight(dpre, x, W, dW, db, scratch=None, fold=True):
    check(_lib.lib().pvb_conv_tc_wgrad(_p(dpre), _p(x), _p(dW), _p(db), *_conv_dims(x, W),
                                       _p(scratch), 1 if fold else 0, _stream()), "pvb_conv_tc_wgrad")
