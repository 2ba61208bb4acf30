"""ctypes binding of liblfs2.so (the C ABI declared in include/lfs2.h).

There is deliberately no fallback: if the CUDA library is missing or a call fails, the
product path raises.  Build it with ``python -m lightningfastspeech2_b200.build`` (or
``__graft_entry__.build()``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblfs2.so")

_vp, _i, _f, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong

# name -> argtypes; every function returns int except lfs2_last_error
SIGNATURES = {
    "lfs2_version": [],
    "lfs2_device_arch": [],
    "lfs2_speaker_proj": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "lfs2_embed_pe_spk": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "lfs2_add_pe_spk": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "lfs2_linear": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "lfs2_conv1d_dense": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "lfs2_dwconv1d": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "lfs2_dwconv1d_planes": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "lfs2_merge_planes": [_vp, _vp, _vp, ctypes.c_longlong, _vp],
    "lfs2_attention": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "lfs2_add_layernorm": [_vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp],
    "lfs2_add_layernorm_planes": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp],
    "lfs2_add_layernorm_planes_limited": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _i, _vp],
    "lfs2_rowdot_mask": [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "lfs2_bucket_embed_add": [_vp, _vp, _f, _f, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "lfs2_prior_embed": [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _vp],
    "lfs2_duration_round_guard": [_vp, _vp, _vp, _i, _i, _vp],
    "lfs2_length_regulate_scan": [_vp, _i, _vp, _vp, _vp, _i, _i, _vp],
    "lfs2_length_regulate_scatter": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "lfs2_length_regulate_scatter_ex": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "lfs2_gemm_tc": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i,
                     _vp],
    "lfs2_gemm_tc_limited": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp,
                             _i, _vp, _i, _vp, _vp],
    "lfs2_gemm_tc_ex": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _i,
                        _i, _vp, _i, _vp, _vp, _vp],
    "lfs2_attention_tc_ex": [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "lfs2_attention_tc_wide": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp],
    "lfs2_predictor_layer_tc": [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                _i, _vp, _vp],
    "lfs2_dwconv1d_planes_limited": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp],
    "lfs2_dwconv1d_planes_ex": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp],
    "lfs2_mask_lengths": [_vp, _vp, _i, _i, _vp],
    "lfs2_zero_masked_rows": [_vp, _vp, _ll, _i, _vp],
    "lfs2_ffn_fused_tc": [_vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _i, _vp],
    "lfs2_ffn_fused_tc_limited": [_vp, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _i,
                                  _vp, _i, _vp, _vp],
    "lfs2_ffn_fused_tc_ex": [_vp, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp,
                             _i, _vp, _i, _vp, _vp],
    "lfs2_attention_tc_limited": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "lfs2_attention_tc_workspace_bytes": [_i],
    "lfs2_attention_tc": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "lfs2_split_bf16": [_vp, _vp, _vp, ctypes.c_longlong, _vp],
    "lfs2_split_bf16_ex": [_vp, _vp, _vp, _vp, ctypes.c_longlong, _vp],
    "lfs2_split_f16": [_vp, _vp, _vp, ctypes.c_longlong, _vp],
    "lfs2_gemm_tc2": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _ll, _ll, _i, _i, _i, _i, _i, _i, _i, _vp],
    "lfs2_attn_softmax_planes": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp],
    "lfs2_attn_softmax_planes_drop": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, ctypes.c_ulonglong,
                                      ctypes.c_uint, _vp],
    "lfs2_attn_ds_planes_drop": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, ctypes.c_ulonglong, ctypes.c_uint,
                                 _vp],
    "lfs2_attn_delta": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "lfs2_attn_ds_planes": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp],
    # HiFi-GAN generator glue
    "lfs2_lrelu_planes": [_vp, _vp, _vp, _vp, _ll, _f, _vp],
    "lfs2_mean3_lrelu_planes": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _f, _f, _vp],
    "lfs2_mel_to_planes": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "lfs2_conv_post_tanh": [_vp, _vp, _vp, _vp, _vp, _f, _vp, _i, _i, _i, _i, _vp],
    # FastDiff variance adaptor glue
    "lfs2_diffusion_step_embed": [_vp, _vp, _i, _i, _vp],
    "lfs2_swish": [_vp, _ll, _vp],
    "lfs2_diffusion_input": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "lfs2_diffusion_mix": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _i, _i, _vp],
    # train-step config
    "lfs2_add_layernorm_train": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, ctypes.c_ulonglong, ctypes.c_uint, _vp],
    "lfs2_layernorm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "lfs2_layernorm_bwd_drop": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, ctypes.c_ulonglong, ctypes.c_uint,
                                _vp],
    "lfs2_layernorm_bwd_ex": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, ctypes.c_ulonglong,
                              ctypes.c_uint, _vp],
    "lfs2_gemm_tn": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "lfs2_colsum": [_vp, _vp, _i, _i, _vp],
    "lfs2_relu_bwd": [_vp, _vp, _vp, _ll, _vp],
    "lfs2_relu_bwd_scaled": [_vp, _vp, _vp, _ll, _f, _vp],
    "lfs2_relu_bwd_planes": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp],
    "lfs2_weight_planes_batched": [_vp, _i, _i, _vp],
    "lfs2_add_inplace": [_vp, _vp, _ll, _vp],
    "lfs2_transpose": [_vp, _vp, _i, _i, _vp],
    "lfs2_dwconv1d_bwd_w": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "lfs2_attention_lse": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "lfs2_attention_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "lfs2_length_regulate_bwd": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "lfs2_embedding_bwd": [_vp, _vp, _vp, _i, _i, _i, _ll, _vp],
    "lfs2_decoder_input_planes": [_vp, _vp, _f, _f, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp],
    "lfs2_sdp_dwconv": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "lfs2_sdp_ln_gelu": [_vp, _vp, _vp, _f, _vp, _vp, _i, _i, _vp],
    "lfs2_sdp_flow_pre": [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "lfs2_sdp_spline_inverse": [_vp, _i, _vp, _i, _vp, _i, _f, _i, _vp],
    "lfs2_sdp_affine_reverse": [_vp, _vp, _vp, _vp, _i, _i, _vp],
    "lfs2_sdp_durations": [_vp, _vp, _vp, _i, _i, _vp],
    "lfs2_pack_valid_rows": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "lfs2_bucket_embed_add_oop": [_vp, _vp, _vp, _f, _f, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "lfs2_rowdot_mask_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "lfs2_sum_over_time": [_vp, _vp, _i, _i, _i, _vp],
    "lfs2_fold_pw_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "lfs2_fold_pw_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "lfs2_dropout": [_vp, _vp, _ll, _f, ctypes.c_ulonglong, ctypes.c_uint, _vp],
    "lfs2_dropout_planes": [_vp, _vp, _vp, _vp, _ll, _f, ctypes.c_ulonglong, ctypes.c_uint, _vp],
    "lfs2_masked_loss": [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp],
    "lfs2_sumsq": [_vp, _vp, _ll, _vp],
    "lfs2_scale_by": [_vp, _vp, _ll, _vp],
    "lfs2_adamw_step": [_vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _i, _f, _f, _vp, _i, _vp],
}
# functions whose return type is not int
RESTYPES = {"lfs2_attention_bwd_workspace_bytes": (ctypes.c_longlong, [_i, _i, _i]),
            "lfs2_gemm_tc_limited_workspace_bytes": (ctypes.c_longlong, [_i, _i]),
            "lfs2_predictor_layer_tc_workspace_bytes": (ctypes.c_longlong, [_i, _i]),
            "lfs2_ffn_fused_tc_limited_workspace_bytes": (ctypes.c_longlong, [_i, _i])}

class Operand(ctypes.Structure):
    """lfs2_operand of include/lfs2.h"""
    _fields_ = [("mn_major", _i), ("d0", _i), ("d1", _i), ("d2", _i), ("col0", _i), ("hstride", _i), ("per_z", _i)]


_lib = None
CALLS = 0  # C-ABI launches issued by this process (bench.py reports the per-step count)


class Lfs2Error(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises if liblfs2.so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Lfs2Error(
            f"{LIB_PATH} not found: the CUDA library has not been built "
            "(python -m lightningfastspeech2_b200.build). There is no CPU fallback.")
    h = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(h, name)  # AttributeError if the .so does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    for name, (restype, argtypes) in RESTYPES.items():
        fn = getattr(h, name)
        fn.argtypes = argtypes
        fn.restype = restype
    h.lfs2_last_error.argtypes = []
    h.lfs2_last_error.restype = ctypes.c_char_p
    _lib = h
    return h


def check(rc, name):
    if rc != 0:
        msg = lib().lfs2_last_error().decode(errors="replace")
        codes = {-1: "INVALID_ARG", -2: "UNSUPPORTED", -3: "CUDA"}
        if rc == -2:
            raise NotImplementedError(f"{name}: {msg}")
        raise Lfs2Error(f"{name} failed ({codes.get(rc, rc)}): {msg}")
