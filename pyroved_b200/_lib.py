"""ctypes binding of libpvb.so (the C ABI declared in include/pvb.h).

The CUDA library is the ONLY compute path of this package: if it is missing
the import of any compute entry point fails loudly (no CPU / eager fallback).
Build it with `python -c "import __graft_entry__ as g; g.build()"` or
`pyroved_b200/csrc/build.sh`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PVB_LIB: an alternative build of the library (kernel experiments, tools/build_variants.sh)
LIB_PATH = os.environ.get("PVB_LIB") or os.path.join(_HERE, "csrc", "libpvb.so")

_f = C.c_void_p      # device pointer (float*)
_i64 = C.c_int64
_i32 = C.c_int
_fl = C.c_float
_st = C.c_void_p     # cudaStream_t


class FoldCfg(C.Structure):
    """pvb_fold_cfg"""
    _fields_ = [("ndim", C.c_int32), ("inv", C.c_int32), ("latent_dim", C.c_int32),
                ("cond_dim", C.c_int32), ("hidden", C.c_int32),
                ("dx_prior", C.c_float), ("dy_prior", C.c_float), ("sc_prior", C.c_float)]


class TcSizes(C.Structure):
    """pvb_tc_sizes"""
    _fields_ = [("tiles", C.c_int64), ("ctas", C.c_int32),
                ("gUv_part_floats", C.c_int64), ("wgrad_part_floats", C.c_int64)]


class WgradFold(C.Structure):
    """pvb_wgrad_fold"""
    _fields_ = [("scratch", C.c_void_p), ("dW", C.c_void_p), ("db", C.c_void_p),
                ("Cin", C.c_int32), ("Cout", C.c_int32), ("taps", C.c_int32)]


class MlpTailArgs(C.Structure):
    """pvb_mlp_tail_args"""
    _fields_ = [("M", C.c_int64), ("n_layers", C.c_int32), ("w_in", C.c_int32), ("h_in", C.c_void_p),
                ("width", C.c_int32 * 3), ("W", C.c_void_p * 3), ("b", C.c_void_p * 3),
                ("h", C.c_void_p * 3), ("pre", C.c_void_p * 3), ("act", C.c_int32),
                ("n_heads", C.c_int32), ("hdim", C.c_int32 * 3), ("hW", C.c_void_p * 3),
                ("hb", C.c_void_p * 3), ("hout", C.c_void_p * 3),
                ("gauss", C.c_int32), ("gen_eps", C.c_int32), ("eps", C.c_void_p),
                ("sigma", C.c_void_p), ("z", C.c_void_p), ("kl", C.c_void_p),
                ("seed", C.c_uint64), ("step_counter", C.c_void_p), ("first_index", C.c_int64),
                ("fold", C.c_int32), ("cfg", FoldCfg), ("cond", C.c_void_p), ("Wc", C.c_void_p),
                ("bc", C.c_void_p), ("Wz", C.c_void_p), ("Uv", C.c_void_p)]


class MlpChainArgs(C.Structure):
    """pvb_mlp_chain_args"""
    _fields_ = [("M", C.c_int64), ("n_layers", C.c_int32), ("width", C.c_int32 * 4),
                ("W", C.c_void_p * 4), ("h", C.c_void_p * 4), ("pre", C.c_void_p * 4),
                ("act", C.c_int32), ("dpre", C.c_void_p * 4), ("n_heads", C.c_int32),
                ("hdim", C.c_int32 * 3), ("hW", C.c_void_p * 3), ("g", C.c_void_p * 3)]


class WgradProblem(C.Structure):
    """pvb_wgrad_problem"""
    _fields_ = [("d", C.c_void_p), ("x", C.c_void_p), ("dW", C.c_void_p), ("db", C.c_void_p),
                ("N", C.c_int32), ("K", C.c_int32)]


# name -> argtypes ; every entry returns int except where noted.  This table
# must list every symbol of include/pvb.h (tests/test_abi.py checks it).
SIGNATURES = {
    "pvb_version": [],
    "pvb_last_error_string": [],
    "pvb_launch_count": [],
    "pvb_has_tcgen05": [],
    "pvb_linear_fwd": [_f, _f, _f, _f, _f, _i64, _i32, _i32, _i32, _st],
    "pvb_linear_bwd": [_f, _f, _f, _f, _f, _f, _f, _i32, _f, _f, _i64, _i32, _i32, _i32, _st],
    "pvb_randn": [_f, _i64, C.c_uint64, _f, _i64, _st],
    "pvb_latent_fwd": [_f, _f, _f, _f, _f, _f, _i64, _i32, _st],
    "pvb_latent_bwd": [_f, _f, _f, _f, _f, _f, _fl, _f, _f, _i64, _i32, _st],
    "pvb_fold_fwd": [C.POINTER(FoldCfg), _f, _f, _f, _f, _f, _f, _i64, _st],
    "pvb_fold_bwd_num_partials": [],
    "pvb_fold_bwd": [C.POINTER(FoldCfg), _f, _f, _f, _f, _f, _f, _f, _f, _i64, _st],
    "pvb_sdec_h0_fwd": [_f, _f, _i64, _i32, _i32, _i32, _i32, _st],
    "pvb_sdec_h0_bwd": [_f, _f, _f, _i64, _i32, _i32, _i32, _i32, _st],
    "pvb_obs_loglik": [_f, _f, _f, _f, _f, _f, _i64, _i64, _i32, _i32, _i32, _fl, _st],
    "pvb_elbo_reduce": [_f, _f, _f, _fl, _f, _f, _i32, _i64, _i32, _st],
    "pvb_weighted_sum": [_f, _f, _fl, _f, _i64, _st],
    "pvb_axpy_out": [_f, _f, _fl, _f, _i64, _st],
    "pvb_enum_head_fwd": [_f, _f, _f, _i64, _i32, _st],
    "pvb_enum_head_bwd": [_f, _f, _fl, _f, _f, _i64, _i32, _st],
    "pvb_class_nll": [_f, _f, _fl, _f, _f, _i64, _i32, _st],
    "pvb_reduce_partials": [_f, _f, _i32, _i64, _i64, _i32, _st],
    "pvb_counter_add": [_f, _i32, _st],
    "pvb_adam_flat": [_f, _f, _f, _f, _i64, _fl, _fl, _fl, _fl, _f, _f, _st],
    "pvb_adam_flat_step": [_f, _f, _f, _f, _i64, _fl, _fl, _fl, _fl, _f, _f, _f, _f, _f, _st],
    "pvb_gather_rows": [_f, C.c_void_p, _f, _i64, _i64, _i64, _st],
    "pvb_mlp_tail_fwd": [C.POINTER(MlpTailArgs), _st],
    "pvb_mlp_chain_bwd": [C.POINTER(MlpChainArgs), _st],
    "pvb_mlp_wgrad": [C.POINTER(WgradProblem), _i32, _i64, _st],
    "pvb_latent_side_num_partials": [_i64],
    "pvb_latent_side_bwd": [C.POINTER(FoldCfg), _f, _f, _f, _f, _f, _f, _i32, _f, _f, _f, _f, _f, _f,
                            _f, _fl, _f, _f, _i64, _st],
    "pvb_conv_fwd": [_f, _f, _f, _f, _f, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _st],
    "pvb_conv_bwd_data": [_f, _f, _f, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f, _i32, _st],
    "pvb_conv_bwd_weight": [_f, _f, _f, _f, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _st],
    "pvb_act_bwd": [_f, _f, _f, _f, _i64, _i32, _st],
    "pvb_conv3d_fwd": [_f, _f, _f, _f, _f, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _st],
    "pvb_conv3d_bwd_data": [_f, _f, _f, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _st],
    "pvb_conv3d_bwd_weight": [_f, _f, _f, _f, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _st],
    "pvb_maxpool3d_fwd": [_f, _f, _i64, _i32, _i32, _i32, _st],
    "pvb_maxpool3d_bwd": [_f, _f, _f, _i64, _i32, _i32, _i32, _st],
    "pvb_upsample3d_fwd": [_f, _f, _i64, _i32, _i32, _i32, _st],
    "pvb_upsample3d_bwd": [_f, _f, _i64, _i32, _i32, _i32, _st],
    "pvb_peer_flag_words": [],
    "pvb_peer_state_words": [],
    "pvb_peer_allreduce_adam": [_f, _f, _f, _f, _i64, _f, _f, _f, _i32, _i32, _i32, _fl, _fl, _fl, _fl,
                                _f, _f, _f, _st],
    "pvb_bn_workspace_bytes": [_i32],
    "pvb_bn_fwd": [_f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _i32, _i32, _i64, _fl, _fl, _i32, _st],
    "pvb_bn_bwd": [_f, _f, _f, _f, _f, _f, _f, _f, _f, _i32, _i32, _i64, _i32, _st],
    "pvb_maxpool2_fwd": [_f, _f, _i64, _i32, _i32, _i32, _st],
    "pvb_maxpool2_bwd": [_f, _f, _f, _i64, _i32, _i32, _i32, _i32, _st],
    "pvb_upsample2_fwd": [_f, _f, _i64, _i32, _i32, _i32, _i32, _st],
    "pvb_upsample2_bwd": [_f, _f, _i64, _i32, _i32, _i32, _i32, _f, _i32, _st],
    "pvb_normal_logprob": [_f, _f, _fl, _fl, _f, _f, _i64, _st],
    "pvb_linear_dx_cols": [_f, _f, _f, _i64, _i32, _i32, _i32, _i32, _i32, _st],
    "pvb_conv_tc_supported": [_i32, _i32, _i32, _i32],
    "pvb_conv_tc_wgrad_supported": [_i32, _i32, _i32, _i32],
    "pvb_conv_tc_workspace_bytes": [_i32, _i32, _i32, _i32],
    "pvb_conv_tc_prep": [_f, _f, _i32, _i32, _i32, _i32, _i32, _st],
    "pvb_conv_tc_pix": [_f, _f, _f, _f, _f, _f, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _st],
    "pvb_conv_tc_wgrad_scratch_bytes": [_i32, _i32, _i32, _i32],
    "pvb_conv_tc_wgrad": [_f, _f, _f, _f, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f, _i32, _st],
    "pvb_conv_tc_wgrad_fold": [C.POINTER(WgradFold), _i32, _st],
    "pvb_sdec_tc_sizes": [_i64, _i32, C.POINTER(TcSizes)],
    "pvb_sdec_tc_step": [_f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _i64, _i64,
                         _i32, _i32, _i32, _i32, _i32, _fl, _i32, _f, _st],
    "pvb_sdec_tc_packed_weight_bytes": [],
    "pvb_sdec_tc_pack_weights": [_f, _f, _f, _st],
    "pvb_sdec_tc_gather_gUv": [_f, _f, _i64, _i32, _st],
}

# floats per CTA in pvb_sdec_tc_step's weight-gradient partials (PVB_TC_WGRAD_FLOATS)
TC_WGRAD_FLOATS = 2 * 128 * 128 + 2 * 128 + 128 + 1
TC_WGRAD_STRIDE = (TC_WGRAD_FLOATS + 3) // 4 * 4

_LIB = None


class PvbError(RuntimeError):
    pass


def lib():
    """Load libpvb.so once; raise loudly if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise PvbError(
                "pyroved_b200: CUDA extension {} is missing. There is no CPU fallback; "
                "build it with pyroved_b200/csrc/build.sh (needs nvcc, sm_100a).".format(LIB_PATH))
        handle = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = (C.c_char_p if name == "pvb_last_error_string" else
                          C.c_longlong if name in ("pvb_launch_count",
                                                   "pvb_conv_tc_workspace_bytes",
                                                   "pvb_conv_tc_wgrad_scratch_bytes",
                                                   "pvb_bn_workspace_bytes",
                                                   "pvb_sdec_tc_packed_weight_bytes") else C.c_int)
        _LIB = handle
    return _LIB


def check(rc, what=""):
    if rc != 0:
        msg = lib().pvb_last_error_string()
        raise PvbError("{} failed (code {}): {}".format(
            what or "libpvb call", rc, msg.decode() if msg else ""))
