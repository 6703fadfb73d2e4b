"""ctypes binding of libb200lic.so (the C ABI declared in include/b200lic.h).

There is deliberately no fallback: if the shared library is missing, or the device is not sm_100, every
op raises.  Build with `python -c "import __graft_entry__ as g; g.build()"` or `make -C rdo_ptq_b200/csrc`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200LIC_LIB: another build of the SAME library (A/B experiments on kernel variants); never a different implementation
LIB_PATH = os.environ.get("B200LIC_LIB") or os.path.join(_HERE, "lib", "libb200lic.so")

ERR_NAMES = {0: "OK", -1: "ERR_ARG", -2: "ERR_ARCH", -3: "ERR_CUDA", -4: "ERR_UNSUPPORTED"}
ACT_NONE, ACT_RELU, ACT_LEAKY_RELU = 0, 1, 2
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC = 0, 1, 2


class B200LicError(RuntimeError):
    def __init__(self, fn, code, msg):
        super().__init__(f"b200lic_{fn} failed: {ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class CalibSched(C.Structure):
    """Mirror of `b200lic_calib_sched` (lives in device memory; this mirror is for size/offset checks and read-back)."""
    _fields_ = [("step", C.c_int), ("lr_over_bc1", C.c_float), ("inv_sqrt_bc2", C.c_float), ("reg_b", C.c_float)]


class ConvDesc(C.Structure):
    """Mirror of `b200lic_conv_desc`."""
    _fields_ = [("N", C.c_int), ("Cin", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("Cout", C.c_int), ("Ho", C.c_int), ("Wo", C.c_int),
                ("KH", C.c_int), ("KW", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
                ("act", C.c_int), ("act_slope", C.c_float), ("engine", C.c_int),
                ("in_square", C.c_int), ("gdn_mode", C.c_int), ("fixed_point", C.c_int), ("k_taps", C.c_int)]


_P, _I, _F, _LL, _SZ, _ULL, _DBL = (C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t, C.c_ulonglong,
                                     C.c_double)
_D = C.POINTER(ConvDesc)

# name -> argtypes (the trailing stream argument is appended automatically)
SIGNATURES = {
    "wq_init_minmax": [_P, _I, _I, _I, _I, _I, _I, _P, _P],
    "wq_init_search": [_P, _I, _I, _I, _I, _I, _I, _DBL, _F, _I, _P, _P],
    "wq_fake_quant": [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P],
    "wq_dequant_u8": [_P, _P, _P, _I, _I, _I, _P],
    "adaround_init_alpha": [_P, _P, _I, _I, _I, _P],
    "adaround_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P],
    "adaround_bwd_adam": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _F, _F, _F, _F, _F, _P, _P],
    "actq_stats_init": [_P, _I],
    "actq_stats": [_P, _I, _I, _I, _P],
    "actq_apply": [_P, _P, _I, _I, _I, _I, _P, _P],
    "actq_fused": [_P, _I, _I, _I, _I, _P, _P],
    "actq_apply_stage": [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P],
    "fixed_point": [_P, _SZ, _I, _I, _P],
    "gaussian_lik_fwd": [_P, _P, _P, _I, _I, _I, _LL, _F, _F, _P, _P, _P],
    "round_latent": [_P, _P, _SZ, _P],
    "factorized_lik_fwd": [_P, _P, _P, _P, _I, _I, _I, _F, _P, _P, _P],
    "factorized_table": [_P, _P, _I, _F, _P],
    "gaussian_lik_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _LL, _LL, _F, _F, _I, _P, _P, _P],
    "factorized_lik_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _I, _P],
    "lsq_delta_grad": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _P, _P, _I, _F, _F, _F, _F],
    "lsq_delta_grad_sched": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _F, _F, _F, _F],
    "lp_loss_fwd_bwd": [_P, _P, _SZ, _F, _F, _F, _P, _P],
    "lp_loss_fwd_bwd_sched": [_P, _P, _P, _I, _SZ, _SZ, _I, _I, _P, _F, _F, _F, _P, _P],
    "sq_err_sum": [_P, _P, _SZ, _P],
    "bits_sum": [_P, _SZ, _P],
    "stage_tokens": [_P, _SZ, _I, _I, _P, _P],
    "layernorm_fwd": [_P, _P, _P, _SZ, _I, _F, _P],
    "gelu_fwd": [_P, _SZ, _P],
    "window_attn_softmax": [_P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "window_attn_apply": [_P, _P, _I, _I, _I, _I, _P],
    "actq_tokens": [_P, _SZ, _I, _I, _P, _P],
    "ssim_level": [_P, _P, _P, _I, _I, _I, _F, _F, _P],
    "avg_pool2": [_P, _I, _I, _I, _I, _I, _P],
    "msssim_combine": [_P, _P, _I, _I, _P, _P],
    "conv_fwd": [_D, _P, _P, _P, _P, _P, _P, _P, _SZ],
    "deconv_fwd": [_D, _P, _P, _P, _P, _P, _SZ],
    "conv_fwd_wq": [_D, _P, _P, _P, _P, _P, _P, _SZ],
    "deconv_fwd_wq": [_D, _P, _P, _P, _P, _P, _P, _SZ],
    "wq_int_weights": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "conv_wgrad": [_D, _P, _P, _P, _P, _SZ],
    "deconv_wgrad": [_D, _P, _P, _P, _P, _SZ],
    "conv_wgrad_staged": [_D, _I, _P, _P, _P, _P, _P, _SZ],
    "conv_dgrad": [_D, _P, _P, _P, _P, _SZ],
    "deconv_dgrad": [_D, _P, _P, _P, _P, _SZ],
    "gdn_reparam_fwd": [_P, _SZ, _F, _F, _P],
    "gdn_reparam_bwd": [_P, _P, _SZ, _F, _P],
    "gdn_bwd_prep": [_P, _P, _P, _SZ, _I, _P, _P],
    "gdn_bwd_finish": [_P, _P, _P, _SZ, _P],
    "add_act": [_P, _P, _SZ, _I, _F, _P],
    "act_bwd": [_P, _P, _SZ, _I, _F, _P],
    "gather_mix": [_P, _P, _P, _SZ, _SZ, _F, _ULL, _P, _P],
    "calib_sched_tick": [_P, _I, _DBL, _DBL, _DBL, _F, _F, _F],
    "adaround_bwd_adam_sched": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _F, _F, _F, _F, _F, _P],
    "gather_mix_sched": [_P, _P, _P, _I, _SZ, _SZ, _F, _ULL, _I, _I, _P, _P],
    "conv_pack_weights": [_D, _I, _P, _P, _SZ],
    "quant_pack_weights": [_D, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _SZ, _P],
    "conv_fwd_packed": [_D, _I, _P, _P, _P, _P, _P, _P, _P, _P, _SZ],
    "stage_mix_sched": [_P, _P, _P, _I, _I, _I, _I, _F, _ULL, _I, _I, _P, _I, _P, _P, _I, _P],
    "lp_loss_stage_sched": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _F, _F, _F, _I, _F, _P, _P, _I, _P, _P, _P, _I, _P],
    "conv_wgrad_prepared": [_D, _I, _P, _P, _I, _P, _P, _P, _SZ],
    "conv_wgrad_adam_sched": [_D, _I, _P, _P, _I, _P, _P, _SZ, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _F, _F, _F,
                              _F, _F, _P, _P],
    "xgpu_reduce_adam_sched": [_P, _P, _P, _P, _I, _I, _SZ, _SZ, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _F, _F, _F, _F, _F,
                               _P, _I],
    "gdn_fwd_fused": [_P, _P, _I, _P, _P, _I, _I, _I, _I, _P],
    "selftest_fast_div": [_ULL, _ULL, _P],
    "im2col_stage": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I],
    "col2im": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "rans_symbols": [_P, _P, _I, _P, _P, _I, _F, _I, _I, _SZ, _P, _P],
    "rans_encode_sizes": [_P, _P, C.c_uint, C.c_uint, _P, _P, _P, _I, _P],
    "rans_encode_write": [_P, _P, C.c_uint, C.c_uint, _P, _P, _P, _I, _P, _P],
    "rans_decode": [_P, _P, C.c_uint, C.c_uint, _P, _P, _P, _P, _I, _P],
    "attn_gate": [_P, _P, _P, _SZ, _P],
    "abs": [_P, _SZ, _P],
    "pixel_shuffle": [_P, _I, _I, _I, _I, _I, _I, _F, _P],
    "pixel_unshuffle": [_P, _I, _I, _I, _I, _I, _P],
}
PLAIN = {"version": (C.c_int, []), "last_error_string": (C.c_char_p, []), "device_check": (C.c_int, []),
         "launch_count": (C.c_ulonglong, []), "simt_fallback_count": (C.c_ulonglong, []), "factorized_table_floats": (C.c_int, []), "conv_workspace_bytes": (C.c_size_t, [_D, C.c_int]),
         "debug_timeline": (C.c_int, [C.c_void_p, C.c_int]),
         "set_option": (C.c_int, [C.c_char_p, C.c_int]),
         "gdn_fused_ok": (C.c_int, [C.c_int, C.c_int]),
         "pmf_to_quantized_cdf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
         "conv_staged_view": (C.c_int, [_D, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_void_p)]),
         "conv_packed_weight_bytes": (C.c_size_t, [_D, C.c_int]),
         "conv_stats_once": (C.c_int, [C.c_void_p]), "conv_stats_pending": (C.c_int, []),
         "conv_plan_info": (C.c_int, [_D, C.c_int, C.POINTER(C.c_int)]),
         "conv_x_slot": (C.c_int, [_D, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_int)]),
         "conv_dy_slot": (C.c_int, [_D, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_int)])}
OP_CONV_FWD, OP_DECONV_FWD, OP_CONV_DGRAD, OP_DECONV_DGRAD, OP_CONV_WGRAD, OP_DECONV_WGRAD = range(6)

_lib = None


def exported_names():
    """Every symbol include/b200lic.h declares (checked by the CPU test-suite)."""
    return ["b200lic_" + n for n in list(SIGNATURES) + list(PLAIN)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built: the sm_100a CUDA library is mandatory (no CPU fallback); "
                              "run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, "b200lic_" + name)
            fn.restype, fn.argtypes = C.c_int, list(args) + [C.c_void_p]
        for name, (res, args) in PLAIN.items():
            fn = getattr(L, "b200lic_" + name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def call(name, *args, stream=None):
    """Invoke b200lic_<name>(*args, stream); raise B200LicError on a non-zero return."""
    L = lib()
    rc = getattr(L, "b200lic_" + name)(*args, stream)
    if rc != 0:
        raise B200LicError(name, rc, L.b200lic_last_error_string().decode())


def launch_count():
    return int(lib().b200lic_launch_count())


def simt_fallback_count():
    """Calls that ran on the exact-fp32 SIMT engine because the tensor-core engine rejected the shape (ENGINE_AUTO)."""
    return int(lib().b200lic_simt_fallback_count())
