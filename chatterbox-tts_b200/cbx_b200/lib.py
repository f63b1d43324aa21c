"""ctypes binding of include/cbx_b200.h.  Fails loudly when the CUDA library is missing:
there is no CPU or PyTorch fallback for the hot path."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcbx_b200.so")


class CbxConfig(C.Structure):
    _fields_ = [("t3_layers", C.c_int), ("enc_blocks", C.c_int), ("up_blocks", C.c_int), ("cfm_blocks", C.c_int),
                ("cfm_mid", C.c_int), ("cfm_steps", C.c_int), ("cfm_cfg_rate", C.c_float), ("max_streams", C.c_int),
                ("max_seq", C.c_int), ("max_text", C.c_int), ("max_s3_tokens", C.c_int), ("max_prompt_tokens", C.c_int),
                ("n_voices", C.c_int), ("n_lanes", C.c_int)]


class S3GenCall(C.Structure):
    """cbx_s3gen_call of include/cbx_b200.h"""
    _fields_ = [("voice", C.c_int), ("tokens_h", C.c_void_p), ("n", C.c_int), ("cache_source_d", C.c_void_p), ("m", C.c_int64),
                ("wav_out_d", C.c_void_p), ("source_out_d", C.c_void_p), ("mel_out_d", C.c_void_p), ("seed", C.c_uint64), ("emit_from", C.c_int64)]


class T3OpenReq(C.Structure):
    """cbx_t3_open_req of include/cbx_b200.h"""
    _fields_ = [("voice", C.c_int), ("text_ids_h", C.c_void_p), ("n_text", C.c_int), ("cfg_weight", C.c_float), ("temperature", C.c_float),
                ("repetition_penalty", C.c_float), ("min_p", C.c_float), ("top_p", C.c_float), ("seed", C.c_uint64), ("max_new_tokens", C.c_int)]


class SgemmArgs(C.Structure):
    """cbx_sgemm_args of include/cbx_b200.h"""
    _fields_ = [("A", C.c_void_p), ("lda", C.c_int64), ("a_bs", C.c_int64), ("kc", C.c_int), ("a_stride", C.c_int), ("a_dil", C.c_int), ("a_pad", C.c_int),
                ("a_rows", C.c_int64), ("a_scale", C.c_void_p), ("a_shift", C.c_void_p), ("a_relu", C.c_int),
                ("W", C.c_void_p), ("ldw", C.c_int64), ("w_bs", C.c_int64), ("w_trans", C.c_int),
                ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("batch", C.c_int), ("alpha", C.c_float),
                ("bias", C.c_void_p), ("o_scale", C.c_void_p), ("o_shift", C.c_void_p), ("act", C.c_int),
                ("mul", C.c_void_p), ("ldm", C.c_int64), ("mul_bs", C.c_int64), ("mul_div", C.c_int),
                ("res", C.c_void_p), ("res2", C.c_void_p), ("ldr", C.c_int64), ("res_bs", C.c_int64),
                ("C", C.c_void_p), ("ldc", C.c_int64), ("c_bs", C.c_int64)]


# every exported symbol of include/cbx_b200.h: name -> (restype, argtypes)
_P, _I, _F, _L, _U64 = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_uint64
SIGNATURES = {
    "cbx_abi_version": (_I, []),
    "cbx_last_error": (C.c_char_p, []),
    "cbx_engine_create": (_I, [C.POINTER(CbxConfig), _I, C.POINTER(_P)]),
    "cbx_engine_destroy": (None, [_P]),
    "cbx_manifest_create": (_I, [C.POINTER(CbxConfig), C.POINTER(_P)]),
    "cbx_tensor_count": (_I, [_P]),
    "cbx_tensor_info": (_I, [_P, _I, C.c_char_p, _I, C.POINTER(_L), C.POINTER(_I)]),
    "cbx_tensor_upload": (_I, [_P, C.c_char_p, _P, _L]),
    "cbx_finalize": (_I, [_P]),
    "cbx_voice_put": (_I, [_P, _I, _P, _P, _I, _F, _P, _I, _P, _I, _P, _P]),
    "cbx_voice_drop": (_I, [_P, _I]),
    "cbx_t3_open": (_I, [_P, _I, _P, _I, _F, _F, _F, _F, _F, _U64, _I, C.POINTER(_I), _P]),
    "cbx_t3_open_batch": (_I, [_P, C.POINTER(T3OpenReq), _I, _P, _P]),
    "cbx_t3_step": (_I, [_P, _P, _I, _I, _P, _P]),
    "cbx_t3_set_persistent": (_I, [_P, _I]),
    "cbx_t3_set_priority": (_I, [_P, _I]),
    "cbx_t3_set_alignment_eos": (_I, [_P, _I, _I]),
    "cbx_t3_alignment_peek": (_I, [_P, _I, _P, _P, _P, _P, _P]),
    "cbx_t3_alignment_poke": (_I, [_P, _I, _P, _P, _P]),
    "cbx_op_alignment_run": (_I, [_P, _I, _I, _I, _P, _P]),
    "cbx_t3_poll": (_I, [_P, _I, C.POINTER(_I), C.POINTER(_I), _P]),
    "cbx_t3_tokens": (_I, [_P, _I, _I, _I, _P, _P]),
    "cbx_t3_logits": (_I, [_P, _I, _P, _P]),
    "cbx_t3_close": (_I, [_P, _I]),
    "cbx_t3_stats": (_I, [_P, C.POINTER(_I), C.POINTER(_I)]),
    "cbx_engine_health": (_I, [_P]),
    "cbx_s3gen_infer": (_I, [_P, _I, _P, _I, _P, _L, _P, _P, _P, _P, _P, _U64, _P]),
    "cbx_s3gen_infer_batch": (_I, [_P, C.POINTER(S3GenCall), _I, _P]),
    "cbx_flow_infer": (_I, [_P, _I, _P, _I, _P, _P]),
    "cbx_hift_infer": (_I, [_P, _P, _I, _P, _L, _P, _P, _P, _P, _U64, _P]),
    "cbx_hift_infer_window": (_I, [_P, _P, _I, _P, _L, _P, _P, _U64, _I, _P]),
    "cbx_hift_f0": (_I, [_P, _P, _I, _P, _P]),
    "cbx_hift_source": (_I, [_P, _P, _I, _P, _P, _U64, _P, _P]),
    "cbx_crossfade_pcm": (_I, [_P, _P, _L, _P, _I, _P, _P, _P, _P]),
    "cbx_gpu_launches": (_L, [_P]),
    "cbx_gemm_tc_launches": (C.c_longlong, []),
    "cbx_gemm_mma_launches": (C.c_longlong, []),
    "cbx_attn_tc_launches": (C.c_longlong, []),
    "cbx_attn_fa_launches": (C.c_longlong, []),
    "cbx_attn_fa_trace": (_I, [_P]),
    "cbx_t3_tc_launches": (C.c_longlong, []),
    "cbx_gemm_tc_trace": (_I, [_P]),
    "cbx_t3_mega_trace": (_I, [_P]),
    "cbx_t3_mega_prof": (_I, [_P]),
    "cbx_profile_begin": (_I, []),
    "cbx_profile_end": (_I, [_P, _P, _P, _I]),
    "cbx_op_gemm": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "cbx_op_gemm_ex": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "cbx_op_cfm_tail": (_I, [_I, _I] + [_P] * 15),
    "cbx_cfm_tail_launches": (C.c_longlong, []),
    "cbx_cfm_tail_trace": (_I, [_P]),
    "cbx_op_attention": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "cbx_op_attention_bias": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "cbx_cond_sgemm": (_I, [C.POINTER(SgemmArgs), _P]),
    "cbx_cond_frames_dft": (_I, [_P, _L, _I, _I, _I, _I, _P, _I, _F, _I, _I, _P, _I, _P]),
    "cbx_cond_resample": (_I, [_P, _L, _P, _L, _I, _I, _P, _I, _I, _P]),
    "cbx_cond_layernorm": (_I, [_P, _L, _P, _L, _I, _I, _P, _P, _F, _P]),
    "cbx_cond_softmax": (_I, [_P, _L, _L, _I, _I, _I, _P]),
    "cbx_cond_rotary": (_I, [_P, _L, _I, _I, _I, _F, _P]),
    "cbx_cond_dwconv_add": (_I, [_P, _L, _P, _I, _P, _L, _I, _I, _P]),
    "cbx_cond_fsq": (_I, [_P, _P, _I, _P]),
    "cbx_cond_mel_log": (_I, [_P, _L, _I, _F, _P, _P]),
    "cbx_cond_col_stats": (_I, [_P, _L, _I, _I, _I, _P, _P, _P, _I, _P, _P]),
    "cbx_cond_l2norm_rows": (_I, [_P, _I, _I, _I, _P]),
    "cbx_cond_mean_rows": (_I, [_P, _I, _I, _P, _P]),
    "cbx_cond_conv2d": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "cbx_cond_lstm_layer": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
}

_lib = None


def load():
    """dlopen libcbx_b200.so and attach signatures.  Raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  The Chatterbox hot path has no CPU/PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class CbxError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise CbxError(load().cbx_last_error().decode())
