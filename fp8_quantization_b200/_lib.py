"""ctypes binding of libfp8fq.so -- the only way the package reaches the GPU.

There is deliberately NO fallback: if the library is missing, or a tensor is not a contiguous
fp32 CUDA tensor, the call raises.  (north_star: "no Triton, no multi-backend dispatch, no CPU
fallback".)
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FP8FQ_LIB lets the tuning scripts under tools/ load an alternative build of the SAME library
LIB_PATH = os.environ.get("FP8FQ_LIB") or os.path.join(_HERE, "libfp8fq.so")

_c_f = ctypes.c_float
_c_d = ctypes.c_double
_c_i = ctypes.c_int
_c_l = ctypes.c_int64
_c_p = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/fp8fq.h one to one
SIGNATURES = {
    "fp8fq_version": (_c_i, []),
    "fp8fq_build_info": (ctypes.c_char_p, []),
    "fp8fq_launch_count": (_c_l, []),
    "fp8fq_format_split": (_c_i, [_c_f, _c_i, _c_i, ctypes.POINTER(_c_i), ctypes.POINTER(_c_i), ctypes.POINTER(_c_i)]),
    "fp8fq_table_stride": (_c_l, [_c_f, _c_i, _c_i]),
    "fp8fq_table_floats": (_c_l, [_c_f, _c_i, _c_i, _c_l]),
    "fp8fq_prepare_f32": (_c_i, [_c_p, _c_l, _c_f, _c_i, _c_i, _c_p, _c_p]),
    "fp8fq_set_range_prepare_f32": (_c_i, [_c_p, _c_p, _c_l, _c_p, _c_f, _c_i, _c_i, _c_p, _c_p]),
    "fp8fq_fake_quant_f32": (_c_i, [_c_p, _c_p, _c_p, _c_l, _c_l, _c_l, _c_f, _c_i, _c_i, _c_p]),
    "fp8fq_fake_quant_codes_f32": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_l, _c_l, _c_l, _c_f, _c_i, _c_i, _c_p]),
    "fp8fq_bn_act_quant_f32": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_l, _c_l, _c_l, _c_i, _c_i, _c_p, _c_f, _c_i, _c_i, _c_p]),
    "fp8fq_bn_fold_f32": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_f, _c_l, _c_p, _c_p, _c_p]),
    "fp8fq_bn_pack_f32": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_f, _c_l, _c_p, _c_p]),
    "fp8fq_fake_quant_backward_f32": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_l, _c_l, _c_l, _c_f, _c_i, _c_i, _c_p, _c_p]),
    "fp8fq_uniform_table_floats": (_c_l, [_c_l]),
    "fp8fq_uniform_prepare_f32": (_c_i, [_c_p, _c_p, _c_l, _c_i, _c_i, _c_i, _c_f, _c_p, _c_p, _c_p, _c_p, _c_p]),
    "fp8fq_uniform_quant_f32": (_c_i, [_c_p, _c_p, _c_p, _c_l, _c_l, _c_l, _c_p]),
    "fp8fq_add_act_quant_f32": (_c_i, [_c_p, _c_p, _c_p, _c_l, _c_i, _c_p, _c_f, _c_i, _c_i, _c_p]),
    "fp8fq_minmax_workspace_bytes": (_c_l, []),
    "fp8fq_minmax_f32": (_c_i, [_c_p, _c_l, _c_l, _c_l, _c_p, _c_p, _c_i, _c_i, _c_d, _c_p, _c_p]),
    "fp8fq_estimate_prepare_f32": (_c_i, [_c_p, _c_l, _c_l, _c_l, _c_p, _c_p, _c_i, _c_i, _c_d, _c_p, _c_f, _c_i, _c_i, _c_p, _c_p, _c_p]),
    "fp8fq_mse_table_floats": (_c_l, [ctypes.POINTER(_c_f), _c_i, _c_i, _c_i, _c_l, _c_l]),
    "fp8fq_mse_grid_f32": (_c_i, [_c_p, _c_l, _c_l, _c_l, _c_p, _c_l, ctypes.POINTER(_c_f), _c_i, _c_i, _c_i, _c_p, _c_p, _c_p]),
    "fp8fq_fake_quant_host_f32": (_c_i, [_c_p, _c_p, _c_p, _c_l, _c_l, _c_l, _c_f, _c_i, _c_i, _c_i]),
}



class TensorDesc(ctypes.Structure):
    """fp8fq_tensor_desc (include/fp8fq.h)."""

    _fields_ = [("x", ctypes.c_void_p), ("y", ctypes.c_void_p), ("table", ctypes.c_void_p), ("C", ctypes.c_int64),
                ("inner", ctypes.c_int64)]


SIGNATURES["fp8fq_fake_quant_multi_f32"] = (_c_i, [ctypes.POINTER(TensorDesc), _c_i, _c_f, _c_i, _c_i, _c_p])
SIGNATURES["fp8fq_bn_quant_add_act_quant_f32"] = (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_l, _c_l, _c_l, _c_i, _c_i, _c_p, _c_f,
                                                         _c_i, _c_i, _c_p, _c_f, _c_i, _c_i, _c_p])

SIGNATURES["fp8fq_bn_act_quant_nhwc_f32"] = (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_l, _c_l, _c_i, _c_i, _c_p, _c_f, _c_i, _c_i,
                                                    _c_p])
SIGNATURES["fp8fq_bn_quant_add_act_quant_nhwc_f32"] = (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_l, _c_l, _c_i, _c_i, _c_p,
                                                              _c_f, _c_i, _c_i, _c_p, _c_f, _c_i, _c_i, _c_p])

SIGNATURES["fp8fq_bn_act_estimate_prepare_f32"] = (_c_i, [_c_p, _c_l, _c_l, _c_l, _c_i, _c_p, _c_p, _c_i, _c_i, _c_p, _c_p,
                                                          _c_i, _c_i, _c_d, _c_p, _c_f, _c_i, _c_i, _c_p, _c_p, _c_p])

SIGNATURES["fp8fq_space_to_depth2_nhwc_f32"] = (_c_i, [_c_p, _c_p, _c_l, _c_l, _c_l, _c_l, _c_l, _c_l, _c_l, _c_p])

SIGNATURES["fp8fq_max_pool2d_nhwc_f32"] = (_c_i, [_c_p, _c_p, _c_l, _c_l, _c_l, _c_l, _c_i, _c_i, _c_i, _c_i, _c_i, _c_i, _c_p])

SIGNATURES["fp8fq_u8_normalize_nchw_f32"] = (_c_i, [_c_p, _c_p, _c_p, _c_l, _c_l, _c_l, _c_p])

SIGNATURES["fp8fq_dp_finish_prepare_f32"] = (_c_i, [_c_p, _c_l, _c_p, _c_p, _c_i, _c_i, _c_d, _c_p, _c_f, _c_i, _c_i, _c_p, _c_p])

_c_u = ctypes.c_uint
SIGNATURES["fp8fq_dp_exchange_words"] = (_c_l, [_c_i])
SIGNATURES["fp8fq_estimate_prepare_p2p_f32"] = (_c_i, [_c_p, _c_l, _c_p, _c_p, _c_i, _c_i, _c_d, _c_p, _c_f, _c_i, _c_i, _c_p, _c_p,
                                                      _c_p, _c_i, _c_i, _c_u, _c_p])
SIGNATURES["fp8fq_bn_act_estimate_prepare_p2p_f32"] = (_c_i, [_c_p, _c_l, _c_l, _c_l, _c_i, _c_p, _c_p, _c_i, _c_i, _c_p, _c_p,
                                                             _c_i, _c_i, _c_d, _c_p, _c_f, _c_i, _c_i, _c_p, _c_p, _c_p, _c_i,
                                                             _c_i, _c_u, _c_p])

_lib = None


class Fp8fqError(RuntimeError):
    """Non-zero return code from libfp8fq.so (mapped from the C ABI's int codes)."""


_ERR_NAMES = {-1: "FP8FQ_ERR_BAD_ARG", -2: "FP8FQ_ERR_UNSUPPORTED", -3: "FP8FQ_ERR_ALIGNMENT", -4: "FP8FQ_ERR_WORKSPACE"}


def lib():
    """Loads libfp8fq.so (once).  Raises if it has not been built -- there is no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Fp8fqError(
                f"{LIB_PATH} not found: build it with `python -m fp8_quantization_b200.build` "
                "(or __graft_entry__.build()); this package has no non-CUDA fallback"
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code, what):
    if code == 0:
        return
    if code < 0:
        raise Fp8fqError(f"{what}: {_ERR_NAMES.get(code, code)}")
    raise Fp8fqError(f"{what}: CUDA error {code}")
