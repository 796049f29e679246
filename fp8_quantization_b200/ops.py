"""Tensor-level wrappers over the C ABI (include/fp8fq.h).  PyTorch is plumbing here: it owns the
device memory and the stream; every function below is one (or a fixed small number of) kernel
launch(es) from libfp8fq.so on the current CUDA stream, with no host synchronisation.
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import Fp8fqError, TensorDesc, check, lib

ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 2
EST_CURRENT, EST_ALL, EST_RUNNING = 0, 1, 2
EST_DP_STATS = 3   # statistics only, packed as [-min | max | NaN flag] for one MAX all-reduce (fp8fq.h)


try:  # raw handle of the current stream without constructing a torch.cuda.Stream (~0.2 us instead of ~2 us)
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:  # pragma: no cover
    _raw_stream = None


def _stream():
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


# ---- host binding of the hot path ---------------------------------------------------------------------------------------
# Two bindings of the SAME C ABI: the torch-registered operators of libfp8fq_torch.so (csrc/fp8fq_torch.cpp:
# TORCH_LIBRARY(fp8fq, ...) -- argument checks, output allocation and the stream look-up in C++, one dispatcher call per
# op) for the ops of the validate and calibration forwards, and ctypes (_lib.py) for the whole ABI.  The torch binding
# is used for CUDA tensors whenever the library has been built; FP8FQ_BINDING=ctypes forces ctypes (so does FP8FQ_LIB,
# an alternative build of libfp8fq.so the operator library is not linked against).
_torch_ops = None
_torch_ops_tried = False


def torch_binding():
    """``torch.ops.fp8fq`` if libfp8fq_torch.so is available and selected, else None."""
    global _torch_ops, _torch_ops_tried
    if not _torch_ops_tried:
        _torch_ops_tried = True
        import os

        from ._lib import _HERE

        path = os.path.join(_HERE, "libfp8fq_torch.so")
        if (os.environ.get("FP8FQ_BINDING", "torch") == "torch" and not os.environ.get("FP8FQ_LIB")
                and os.path.exists(path)):
            try:
                torch.ops.load_library(path)
            except (OSError, RuntimeError) as exc:   # e.g. built against another torch: the ctypes binding of the SAME
                import warnings                       # kernels takes over (this is a choice of binding, not of path)

                warnings.warn(f"libfp8fq_torch.so could not be loaded ({exc}); using the ctypes binding of libfp8fq.so")
                return None
            if torch.ops.fp8fq.abi_version() != lib().fp8fq_version():
                raise Fp8fqError("libfp8fq_torch.so and libfp8fq.so disagree about the ABI version: rebuild both")
            _torch_ops = torch.ops.fp8fq
    return _torch_ops


def _via_torch(t) -> bool:
    return isinstance(t, torch.Tensor) and t.is_cuda and torch_binding() is not None


def _torch_call(fn, *args):
    try:
        return fn(*args)
    except torch.OutOfMemoryError:    # (a RuntimeError subclass: not an argument error, keep its type)
        raise
    except RuntimeError as exc:       # TORCH_CHECK -> the package's error type
        raise Fp8fqError(str(exc).split("\n")[0]) from None


class nvtx_range:
    """NVTX range around a phase of the flow (calibration, BN re-estimation, validate, one graphed forward) when
    FP8FQ_NVTX=1: shows up in Nsight Systems / ncu --nvtx; free otherwise.  (The reference has no tracing hooks.)"""

    _on = None

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if nvtx_range._on is None:
            import os

            nvtx_range._on = os.environ.get("FP8FQ_NVTX", "0") == "1" and torch.cuda.is_available()
        if nvtx_range._on:
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if nvtx_range._on:
            torch.cuda.nvtx.range_pop()
        return False


def on_device(t) -> bool:
    """The one gate that decides whether a tensor may go to the kernels: it must live on a CUDA device.  Every fused
    path of the module layer asks this (and ``_require`` enforces it), so the package has no CPU path.  (Being the
    single gate is also what lets the test-suite put a functional simulation of the kernels behind the same Python
    layer; nothing in the package itself ever does.)"""
    return isinstance(t, torch.Tensor) and t.is_cuda


def default_device() -> torch.device:
    """Where range scalars given as Python floats are materialised: the current CUDA device."""
    return torch.device("cuda", torch.cuda.current_device())


def is_channels_last(t: torch.Tensor) -> bool:
    """Dense channel-innermost memory ([N, H, W, C] / [N, D, H, W, C]) that is not also plain contiguous."""
    if t.is_contiguous():
        return False
    return ((t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))
            or (t.dim() == 5 and t.is_contiguous(memory_format=torch.channels_last_3d)))


def dense(t: torch.Tensor) -> torch.Tensor:
    """``t`` itself when its memory is dense in one of the two layouts the kernels address (contiguous, or
    channels_last -- the layout cuDNN's tensor-core convolutions produce natively); a contiguous copy otherwise."""
    return t if (t.is_contiguous() or is_channels_last(t)) else t.contiguous()


def _require(t: torch.Tensor, name: str):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise Fp8fqError(f"{name} must be a CUDA tensor (this engine has no CPU path); got device {t.device}")
    if t.dtype != torch.float32:
        raise Fp8fqError(f"{name} must be float32; got {t.dtype}")
    if not t.is_contiguous() and not is_channels_last(t):
        raise Fp8fqError(f"{name} must be contiguous (or dense channels_last)")
    if t.device.index != torch.cuda.current_device():
        raise Fp8fqError(f"{name} lives on {t.device} but the current CUDA device is {torch.cuda.current_device()}; "
                         "kernels are launched on the current device's current stream (use torch.cuda.device(...))")


def _require_same_layout(a: torch.Tensor, b: torch.Tensor, what: str):
    if a.shape != b.shape:
        raise Fp8fqError(f"{what}: shape mismatch")
    if a.stride() != b.stride():
        raise Fp8fqError(f"{what}: both tensors must use the same memory layout (strides {a.stride()} vs {b.stride()})")


def _out_like(x: torch.Tensor, out, what: str) -> torch.Tensor:
    """Output buffer of an elementwise op: a fresh tensor with x's layout, or the caller's -- which must be a dense
    fp32 CUDA tensor with exactly x's shape and strides (the kernels address input and output identically)."""
    if out is None:
        return torch.empty_like(x)
    _require(out, "out")
    _require_same_layout(x, out, what)
    return out


def _opt_ptr(t):
    return t.data_ptr() if t is not None else None


_stride_cache = {}


def _require_table(table: torch.Tensor, C: int, mantissa_bits: float, n_bits: int, sign_bits: int, what: str):
    """``table`` must hold the C channel tables of this format (include/fp8fq.h): the kernels index it by the format's
    stride, so a table built for another format / channel count would be read out of bounds."""
    key = (float(mantissa_bits), int(n_bits), int(sign_bits))
    stride = _stride_cache.get(key)
    if stride is None:
        stride = _stride_cache[key] = table_stride(*key)
    _require(table, what)
    if table.numel() < stride * max(int(C), 1):
        raise Fp8fqError(f"{what}: {table.numel()} floats, but {C} channel table(s) of this format need {stride * max(int(C), 1)}"
                         " (was it prepared for another format or channel count?)")


def _require_state(t: torch.Tensor, C: int, what: str):
    """Per-channel estimator / range buffers written by the kernels: fp32, dense, at least C entries."""
    _require(t, what)
    if t.numel() < C:
        raise Fp8fqError(f"{what}: {t.numel()} entries for {C} channel(s)")


def format_split(mantissa_bits: float, n_bits: int, sign_bits: int):
    """(M, E, K) of fp8_quantizer.py:105-106; K = number of exponent codes."""
    M, E, K = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(lib().fp8fq_format_split(float(mantissa_bits), int(n_bits), int(sign_bits), M, E, K), "format_split")
    return M.value, E.value, K.value


def table_stride(mantissa_bits: float, n_bits: int, sign_bits: int) -> int:
    s = lib().fp8fq_table_stride(float(mantissa_bits), int(n_bits), int(sign_bits))
    if s < 0:
        check(int(s), "table_stride")
    return int(s)


def new_table(C: int, mantissa_bits: float, n_bits: int, sign_bits: int, device) -> torch.Tensor:
    return torch.empty(table_stride(mantissa_bits, n_bits, sign_bits) * C, dtype=torch.float32, device=device)


def prepare(maxval: torch.Tensor, mantissa_bits: float, n_bits: int, sign_bits: int, out: torch.Tensor = None):
    """Table for maxval [C] (fp8_quantizer.py:105-113,128,130)."""
    _require(maxval, "maxval")
    C = maxval.numel()
    if out is None:
        out = new_table(C, mantissa_bits, n_bits, sign_bits, maxval.device)
    else:
        _require_table(out, C, mantissa_bits, n_bits, sign_bits, "table")
    check(lib().fp8fq_prepare_f32(maxval.data_ptr(), C, float(mantissa_bits), int(n_bits), int(sign_bits),
                                  out.data_ptr(), _stream()), "fp8fq_prepare_f32")
    return out


def set_range_prepare(xmin: torch.Tensor, xmax: torch.Tensor, mantissa_bits: float, n_bits: int, sign_bits: int,
                      maxval_out: torch.Tensor = None, table_out: torch.Tensor = None):
    """set_quant_range (fp8_quantizer.py:236-237) + table; returns (maxval [C], table)."""
    _require(xmin, "x_min")
    _require(xmax, "x_max")
    C = xmax.numel()
    if xmin.numel() != C:
        raise Fp8fqError("x_min and x_max must have the same number of elements")
    if maxval_out is None:
        maxval_out = torch.empty(C, dtype=torch.float32, device=xmax.device)
    else:
        _require_state(maxval_out, C, "maxval_out")
    if table_out is None:
        table_out = new_table(C, mantissa_bits, n_bits, sign_bits, xmax.device)
    else:
        _require_table(table_out, C, mantissa_bits, n_bits, sign_bits, "table_out")
    check(lib().fp8fq_set_range_prepare_f32(xmin.data_ptr(), xmax.data_ptr(), C, maxval_out.data_ptr(),
                                            float(mantissa_bits), int(n_bits), int(sign_bits),
                                            table_out.data_ptr(), _stream()), "fp8fq_set_range_prepare_f32")
    return maxval_out, table_out


def fake_quant(x: torch.Tensor, table: torch.Tensor, C: int, mantissa_bits: float, n_bits: int, sign_bits: int,
               out: torch.Tensor = None):
    """FPQuantizer.forward (fp8_quantizer.py:91-133).  C == 1: per tensor; else channel = dim 0."""
    if _via_torch(x):
        return _torch_call(_torch_ops.fake_quant, x, table, int(C), float(mantissa_bits), int(n_bits), int(sign_bits), out)
    _require(x, "x")   # channels_last x: elementwise over the same memory; channel = dim 0 stays the outermost stride
    _require_table(table, C, mantissa_bits, n_bits, sign_bits, "table")
    n = x.numel()
    out = _out_like(x, out, "fake_quant")
    inner = n // C if C > 0 else 0
    check(lib().fp8fq_fake_quant_f32(x.data_ptr(), out.data_ptr(), table.data_ptr(), n, C, inner,
                                     float(mantissa_bits), int(n_bits), int(sign_bits), _stream()),
          "fp8fq_fake_quant_f32")
    return out


def fake_quant_codes(x: torch.Tensor, table: torch.Tensor, C: int, mantissa_bits: float, n_bits: int, sign_bits: int):
    """Returns (y, codes int32) -- codes = sign<<31 | e<<16 | q (parity tests)."""
    _require(x, "x")
    _require_table(table, C, mantissa_bits, n_bits, sign_bits, "table")
    n = x.numel()
    y = torch.empty_like(x)
    codes = torch.empty_like(x, dtype=torch.int32)
    check(lib().fp8fq_fake_quant_codes_f32(x.data_ptr(), y.data_ptr(), codes.data_ptr(), table.data_ptr(), n, C,
                                           n // C, float(mantissa_bits), int(n_bits), int(sign_bits), _stream()),
          "fp8fq_fake_quant_codes_f32")
    return y, codes


def bn_fold(mean, var, gamma, beta, eps: float):
    """Eval-mode batch norm as a per-channel affine map (scale, shift)."""
    _require(mean, "running_mean")
    _require(var, "running_var")
    C = mean.numel()
    scale = torch.empty_like(mean)
    shift = torch.empty_like(mean)
    check(lib().fp8fq_bn_fold_f32(mean.data_ptr(), var.data_ptr(),
                                  gamma.data_ptr() if gamma is not None else None,
                                  beta.data_ptr() if beta is not None else None, float(eps), C,
                                  scale.data_ptr(), shift.data_ptr(), _stream()), "fp8fq_bn_fold_f32")
    return scale, shift


def bn_pack(mean, var, gamma, beta, eps: float):
    """Packed parameters [mean | gamma | rsqrt(var + eps) | beta] for the bit-exact batch-norm mode (bn_mode 1)."""
    _require(mean, "running_mean")
    _require(var, "running_var")
    C = mean.numel()
    packed = torch.empty(4 * C, dtype=torch.float32, device=mean.device)
    check(lib().fp8fq_bn_pack_f32(mean.data_ptr(), var.data_ptr(), _opt_ptr(gamma), _opt_ptr(beta), float(eps), C,
                                  packed.data_ptr(), _stream()), "fp8fq_bn_pack_f32")
    return packed


# The fused batch-norm epilogues index with 32 bits (their row / channel arithmetic is built on 32-bit multiply-high
# divisions); the C ABI answers FP8FQ_ERR_UNSUPPORTED from this size on and the callers compose the unfused steps
# (F.batch_norm, activation, the 64-bit-indexed plain quantiser) instead.  (2^31: the channel-innermost variants keep
# tile offsets beyond the last element in 32 bits as well.)
MAX_FUSED_ELEMS = 1 << 31


def bn_act_quant(x, bn_scale, bn_shift, act: int, table, mantissa_bits: float, n_bits: int, sign_bits: int,
                 bn_mode: int = 0, out=None):
    """quantized_folded_bn.py:39-55 in one pass: Q(act(bn(x))), x is [N, C, *spatial] contiguous.
    bn_mode 0: (bn_scale, bn_shift) from bn_fold; bn_mode 1: bn_scale = bn_pack(...) (bit-exact ATen arithmetic).
    Returns None for tensors of MAX_FUSED_ELEMS elements or more (the caller composes the unfused ops)."""
    if _via_torch(x):
        return _torch_call(_torch_ops.bn_act_quant, x, bn_scale, bn_shift, int(act), table, float(mantissa_bits), int(n_bits),
                           int(sign_bits), int(bn_mode), out)
    _require(x, "x")
    if x.numel() >= MAX_FUSED_ELEMS:
        return None
    Cbn = bn_scale.numel() // (4 if bn_mode == 1 else 1)
    rows, hw = _rows_hw(x, Cbn)
    _require_table(table, 1, mantissa_bits, n_bits, sign_bits, "table")
    out = _out_like(x, out, "bn_act_quant")
    if hw == 1 or is_channels_last(x):   # channel-innermost memory: [N, C] Linear outputs, channels_last activations
        check(lib().fp8fq_bn_act_quant_nhwc_f32(x.data_ptr(), out.data_ptr(), bn_scale.data_ptr(), _opt_ptr(bn_shift),
                                                x.numel() // Cbn, Cbn, int(act), int(bn_mode), table.data_ptr(),
                                                float(mantissa_bits), int(n_bits), int(sign_bits), _stream()),
              "fp8fq_bn_act_quant_nhwc_f32")
        return out
    check(lib().fp8fq_bn_act_quant_f32(x.data_ptr(), out.data_ptr(), bn_scale.data_ptr(), _opt_ptr(bn_shift), rows,
                                       hw, Cbn, int(act), int(bn_mode), table.data_ptr(), float(mantissa_bits),
                                       int(n_bits), int(sign_bits), _stream()), "fp8fq_bn_act_quant_f32")
    return out


def _rows_hw(x, Cbn):
    if x.dim() < 2 or x.shape[1] != Cbn:
        raise Fp8fqError("x must be [N, C, ...] with C == number of batch-norm channels")
    hw = 1
    for d in x.shape[2:]:
        hw *= d
    return x.shape[0] * Cbn, hw


def bn_quant_add_act_quant(x, residual, bn_scale, bn_shift, act: int, table_inner, fmt_inner, table_outer, fmt_outer,
                           bn_mode: int = 0, out=None):
    """Whole residual-block tail (models/resnet_quantized.py:39-46) in one pass:
    Q_outer(act(Q_inner(bn(x)) + residual)).  fmt_* = (mantissa_bits, n_bits, sign_bits).
    Returns None when the fused variant does not cover the shape (caller composes the two kernels)."""
    if _via_torch(x):
        return _torch_call(_torch_ops.bn_quant_add_act_quant, x, residual, bn_scale, bn_shift, int(act), table_inner,
                           float(fmt_inner[0]), int(fmt_inner[1]), int(fmt_inner[2]), table_outer, float(fmt_outer[0]),
                           int(fmt_outer[1]), int(fmt_outer[2]), int(bn_mode), out)
    _require(x, "x")
    _require(residual, "residual")
    _require_same_layout(x, residual, "bn_quant_add_act_quant")
    if x.numel() >= MAX_FUSED_ELEMS:
        return None
    Cbn = bn_scale.numel() // (4 if bn_mode == 1 else 1)
    rows, hw = _rows_hw(x, Cbn)
    _require_table(table_inner, 1, *fmt_inner, "table_inner")
    _require_table(table_outer, 1, *fmt_outer, "table_outer")
    out = _out_like(x, out, "bn_quant_add_act_quant")
    if hw == 1 or is_channels_last(x):
        check(lib().fp8fq_bn_quant_add_act_quant_nhwc_f32(
            x.data_ptr(), residual.data_ptr(), out.data_ptr(), bn_scale.data_ptr(), _opt_ptr(bn_shift),
            x.numel() // Cbn, Cbn, int(act), int(bn_mode), table_inner.data_ptr(), float(fmt_inner[0]),
            int(fmt_inner[1]), int(fmt_inner[2]), table_outer.data_ptr(), float(fmt_outer[0]), int(fmt_outer[1]),
            int(fmt_outer[2]), _stream()), "fp8fq_bn_quant_add_act_quant_nhwc_f32")
        return out
    code = lib().fp8fq_bn_quant_add_act_quant_f32(
        x.data_ptr(), residual.data_ptr(), out.data_ptr(), bn_scale.data_ptr(), _opt_ptr(bn_shift), rows, hw,
        Cbn, int(act), int(bn_mode), table_inner.data_ptr(), float(fmt_inner[0]), int(fmt_inner[1]),
        int(fmt_inner[2]), table_outer.data_ptr(), float(fmt_outer[0]), int(fmt_outer[1]), int(fmt_outer[2]), _stream())
    if code == -2:
        return None
    check(code, "fp8fq_bn_quant_add_act_quant_f32")
    return out


def fake_quant_multi(xs, tables, Cs, mantissa_bits: float, n_bits: int, sign_bits: int, outs=None):
    """Per-channel fake-quant of several tensors of one format in ONE launch (hijacker.py:88-98 for every
    layer of a forward).  Returns the list of outputs."""
    if outs is None and len(xs) > 0 and _via_torch(xs[0]):
        return list(_torch_call(_torch_ops.fake_quant_multi, list(xs), list(tables), [int(c) for c in Cs],
                                float(mantissa_bits), int(n_bits), int(sign_bits)))
    if outs is None:
        outs = [torch.empty_like(x) for x in xs]
    descs = (TensorDesc * len(xs))()
    for d, x, y, t, C in zip(descs, xs, outs, tables, Cs):
        _require(x, "x")
        _require_table(t, C, mantissa_bits, n_bits, sign_bits, "table")
        _out_like(x, y, "fake_quant_multi")
        d.x, d.y, d.table, d.C, d.inner = x.data_ptr(), y.data_ptr(), t.data_ptr(), C, x.numel() // C
    check(lib().fp8fq_fake_quant_multi_f32(descs, len(xs), float(mantissa_bits), int(n_bits), int(sign_bits),
                                           _stream()), "fp8fq_fake_quant_multi_f32")
    return outs


def add_act_quant(a, b, act: int, table, mantissa_bits: float, n_bits: int, sign_bits: int, out=None):
    """models/resnet_quantized.py:43-46 in one pass: Q(act(a + b))."""
    if _via_torch(a):
        return _torch_call(_torch_ops.add_act_quant, a, b, int(act), table, float(mantissa_bits), int(n_bits), int(sign_bits), out)
    _require(a, "a")
    _require(b, "b")
    _require_same_layout(a, b, "add_act_quant")
    _require_table(table, 1, mantissa_bits, n_bits, sign_bits, "table")
    out = _out_like(a, out, "add_act_quant")
    check(lib().fp8fq_add_act_quant_f32(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), int(act),
                                        table.data_ptr(), float(mantissa_bits), int(n_bits), int(sign_bits),
                                        _stream()), "fp8fq_add_act_quant_f32")
    return out


def fake_quant_backward(grad_y, x, table, C: int, mantissa_bits: float, n_bits: int, sign_bits: int):
    """STE backward (fp8fq_fake_quant_backward_f32): returns (grad_x, acc [C, 2] float64) with acc[:, 0] the clipping
    term and acc[:, 1] the scale term of d/dmaxval."""
    _require(x, "x")
    _require(grad_y, "grad_y")
    _require_same_layout(x, grad_y, "fake_quant_backward")
    _require_table(table, C, mantissa_bits, n_bits, sign_bits, "table")
    n = x.numel()
    gx = torch.empty_like(x)
    acc = torch.empty(C, 2, dtype=torch.float64, device=x.device)
    check(lib().fp8fq_fake_quant_backward_f32(grad_y.data_ptr(), x.data_ptr(), gx.data_ptr(), table.data_ptr(), n, C,
                                              n // C if C else 0, float(mantissa_bits), int(n_bits), int(sign_bits),
                                              acc.data_ptr(), _stream()), "fp8fq_fake_quant_backward_f32")
    return gx, acc


def uniform_prepare(xmin, xmax, n_bits: int, symmetric: bool, eps: float, aten_cuda_scalar_div: bool = True):
    """set_quant_range of the INT uniform quantisers (uniform_quantizers.py:224-246, 303-314) on the device.
    Returns (delta [C], zero_float [C], signed [1] as 0./1., table)."""
    _require(xmin, "x_min")
    _require(xmax, "x_max")
    C = xmax.numel()
    delta = torch.empty(C, dtype=torch.float32, device=xmax.device)
    zero_float = torch.empty(C, dtype=torch.float32, device=xmax.device)
    signed = torch.empty(1, dtype=torch.float32, device=xmax.device)
    table = torch.empty(int(lib().fp8fq_uniform_table_floats(C)), dtype=torch.float32, device=xmax.device)
    check(lib().fp8fq_uniform_prepare_f32(xmin.data_ptr(), xmax.data_ptr(), C, int(n_bits), 1 if symmetric else 0,
                                          1 if aten_cuda_scalar_div else 0, float(eps), delta.data_ptr(), zero_float.data_ptr(), signed.data_ptr(),
                                          table.data_ptr(), _stream()), "fp8fq_uniform_prepare_f32")
    return delta, zero_float, signed, table


def uniform_quant(x, table, C: int, out=None):
    """Asymmetric/SymmetricUniformQuantizer.forward (uniform_quantizers.py:107-164), one launch."""
    _require(x, "x")
    n = x.numel()
    out = _out_like(x, out, "uniform_quant")
    check(lib().fp8fq_uniform_quant_f32(x.data_ptr(), out.data_ptr(), table.data_ptr(), n, C, n // C if C else 0,
                                        _stream()), "fp8fq_uniform_quant_f32")
    return out


def space_to_depth2(x: torch.Tensor, pad: int, hs: int, ws: int) -> torch.Tensor:
    """2x2 space-to-depth of an NCHW image [N, C <= 4, H, W] with ``pad`` zero rows / columns on the top / left (and
    whatever is needed on the bottom / right to fill ``hs x ws``): returns [N, 16, hs, ws] in channels_last memory with
    channel ``c*4 + p*2 + q`` = padded pixel ``(2Y + p, 2X + q)`` of input channel c (fp8fq_space_to_depth2_nhwc_f32)."""
    _require(x, "x")
    if x.dim() != 4 or not x.is_contiguous() or x.shape[1] > 4:
        raise Fp8fqError("space_to_depth2: x must be an NCHW-contiguous [N, C <= 4, H, W] tensor")
    N, C, H, W = x.shape
    y = torch.empty((N, 16, hs, ws), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    check(lib().fp8fq_space_to_depth2_nhwc_f32(x.data_ptr(), y.data_ptr(), N, C, H, W, int(pad), int(hs), int(ws),
                                               _stream()), "fp8fq_space_to_depth2_nhwc_f32")
    return y


def normalize_lut(mean, std, device) -> torch.Tensor:
    """[C, 256] table of torchvision's ToTensor + Normalize for every byte value, evaluated with the same fp32 tensor
    operations (utils/imagenet_dataloaders.py:66-81: ``img.to(float32).div(255)``, then ``sub_(mean).div_(std)``)."""
    v = torch.arange(256, dtype=torch.float32).div(255).repeat(len(mean), 1)
    m = torch.as_tensor(mean, dtype=torch.float32).view(-1, 1)
    sd = torch.as_tensor(std, dtype=torch.float32).view(-1, 1)
    return v.sub_(m).div_(sd).contiguous().to(device)


def normalize_u8(x: torch.Tensor, lut: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """uint8 NCHW image batch -> normalised fp32 NCHW batch through ``lut`` (normalize_lut): one HBM-bound pass,
    bit-identical to ToTensor + Normalize (fp8fq_u8_normalize_nchw_f32)."""
    if not isinstance(x, torch.Tensor) or x.dtype != torch.uint8 or x.dim() != 4 or not x.is_contiguous():
        raise Fp8fqError("normalize_u8: x must be a contiguous uint8 [N, C, H, W] tensor")
    if not on_device(x):
        raise Fp8fqError(f"normalize_u8: x must be a CUDA tensor (this engine has no CPU path); got device {x.device}")
    N, C, H, W = x.shape
    _require(lut, "lut")
    if lut.numel() != C * 256:
        raise Fp8fqError(f"normalize_u8: lut must be [C, 256] = [{C}, 256]")
    if out is None:
        out = torch.empty((N, C, H, W), dtype=torch.float32, device=x.device)
    else:
        _require(out, "out")
        if tuple(out.shape) != (N, C, H, W) or not out.is_contiguous():
            raise Fp8fqError("normalize_u8: out must be a contiguous fp32 tensor of x's shape")
    check(lib().fp8fq_u8_normalize_nchw_f32(x.data_ptr(), lut.data_ptr(), out.data_ptr(), N, C, H * W, _stream()),
          "fp8fq_u8_normalize_nchw_f32")
    return out


def max_pool2d_channels_last(x: torch.Tensor, kernel, stride, padding) -> torch.Tensor:
    """F.max_pool2d (floor mode, dilation 1) of a channels_last [N, C, H, W] tensor, C % 4 == 0, in one HBM-bound
    pass (fp8fq_max_pool2d_nhwc_f32); same bits as ATen, NaN propagation included."""
    _require(x, "x")
    if x.dim() != 4 or not (is_channels_last(x) or (x.shape[2] == 1 and x.shape[3] == 1)):
        raise Fp8fqError("max_pool2d_channels_last: x must be a dense channels_last [N, C, H, W] tensor")
    N, C, H, W = x.shape
    (kh, kw), (sh, sw), (ph, pw) = kernel, stride, padding
    ho, wo = (H + 2 * ph - kh) // sh + 1, (W + 2 * pw - kw) // sw + 1
    y = torch.empty((N, C, ho, wo), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    check(lib().fp8fq_max_pool2d_nhwc_f32(x.data_ptr(), y.data_ptr(), N, H, W, C, kh, kw, sh, sw, ph, pw, _stream()),
          "fp8fq_max_pool2d_nhwc_f32")
    return y


_workspaces = {}


def _workspace(device):
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream())
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(int(lib().fp8fq_minmax_workspace_bytes()) // 4, dtype=torch.int32, device=device)
        _workspaces[key] = ws
    return ws


def minmax(x, per_channel: bool, cur_min, cur_max, est_mode: int, initialized: bool, momentum: float = 0.9):
    """Estimator min/max + state update in place (range_estimators.py:61-125).  cur_min/cur_max: [C]."""
    if _via_torch(x):
        _torch_call(_torch_ops.minmax, x, bool(per_channel), cur_min, cur_max, int(est_mode), bool(initialized),
                    float(momentum), _workspace(x.device))
        return cur_min, cur_max
    _require(x, "x")
    n = x.numel()
    C = x.shape[0] if per_channel else 1
    _require_state(cur_min, C, "cur_min")
    _require_state(cur_max, C, "cur_max")
    check(lib().fp8fq_minmax_f32(x.data_ptr(), n, C, n // C, cur_min.data_ptr(), cur_max.data_ptr(), int(est_mode),
                                 1 if initialized else 0, float(momentum), _workspace(x.device).data_ptr(),
                                 _stream()), "fp8fq_minmax_f32")
    return cur_min, cur_max


def estimate_prepare(x, per_channel: bool, cur_min, cur_max, est_mode: int, initialized: bool, momentum: float,
                     maxval_out, mantissa_bits: float, n_bits: int, sign_bits: int, table_out):
    """estimator + set_quant_range + table in one launch (quantization_manager.py:114-122 minus the quantiser)."""
    if _via_torch(x):
        _torch_call(_torch_ops.estimate_prepare, x, bool(per_channel), cur_min, cur_max, int(est_mode), bool(initialized),
                    float(momentum), maxval_out, float(mantissa_bits), int(n_bits), int(sign_bits), table_out,
                    _workspace(x.device))
        return maxval_out, table_out
    _require(x, "x")
    n = x.numel()
    C = x.shape[0] if per_channel else 1
    _require_state(cur_min, C, "cur_min")
    _require_state(cur_max, C, "cur_max")
    _require_state(maxval_out, C, "maxval_out")
    _require_table(table_out, C, mantissa_bits, n_bits, sign_bits, "table_out")
    check(lib().fp8fq_estimate_prepare_f32(x.data_ptr(), n, C, n // C, cur_min.data_ptr(), cur_max.data_ptr(),
                                           int(est_mode), 1 if initialized else 0, float(momentum),
                                           maxval_out.data_ptr(), float(mantissa_bits), int(n_bits), int(sign_bits),
                                           table_out.data_ptr(), _workspace(x.device).data_ptr(), _stream()),
          "fp8fq_estimate_prepare_f32")
    return maxval_out, table_out


def bn_act_estimate_prepare(x, bn_scale, bn_shift, act: int, bn_mode: int, cur_min, cur_max, est_mode: int,
                            initialized: bool, momentum: float, maxval_out=None, fmt=None, table_out=None) -> bool:
    """Per-tensor estimator statistics of act(bn(x)) without materialising it (+ estimator update, and with
    ``table_out``: set_quant_range + table) -- fp8fq_bn_act_estimate_prepare_f32.  Returns False when the shape is
    not covered by the fused kernel (the caller then composes the unfused ops)."""
    if _via_torch(x):
        mb, nb, sb = fmt if fmt is not None else (0.0, 0, 0)
        return bool(_torch_call(_torch_ops.bn_act_estimate_prepare, x, bn_scale, bn_shift, int(act), int(bn_mode), cur_min,
                                cur_max, int(est_mode), bool(initialized), float(momentum), maxval_out, float(mb), int(nb),
                                int(sb), table_out, _workspace(x.device)))
    _require(x, "x")
    if x.numel() >= MAX_FUSED_ELEMS:
        return False
    Cbn = bn_scale.numel() // (4 if bn_mode == 1 else 1)
    rows, hw = _rows_hw(x, Cbn)
    nhwc = hw == 1 or is_channels_last(x)
    mb, nb, sb = fmt if fmt is not None else (0.0, 0, 0)
    _require_state(cur_min, 1, "cur_min")
    _require_state(cur_max, 1, "cur_max")
    if maxval_out is not None:
        _require_state(maxval_out, 1, "maxval_out")
    if table_out is not None:
        _require_table(table_out, 1, mb, nb, sb, "table_out")
    code = lib().fp8fq_bn_act_estimate_prepare_f32(
        x.data_ptr(), x.numel() // Cbn if nhwc else rows, hw, Cbn, 1 if nhwc else 0, bn_scale.data_ptr(),
        _opt_ptr(bn_shift), int(bn_mode), int(act), cur_min.data_ptr(), cur_max.data_ptr(), int(est_mode),
        1 if initialized else 0, float(momentum), _opt_ptr(maxval_out), float(mb), int(nb), int(sb),
        _opt_ptr(table_out), _workspace(x.device).data_ptr(), _stream())
    if code == -2:
        return False
    check(code, "fp8fq_bn_act_estimate_prepare_f32")
    return True


def estimate_prepare_p2p(x, cur_min, cur_max, est_mode: int, initialized: bool, momentum: float, maxval_out, fmt, table_out,
                         xchg):
    """Per-tensor calibration step of a data-parallel run in ONE launch: local statistics, MAX over the ranks through
    NVLink peer memory (no collective call), estimator rule, set_quant_range, table (fp8fq_estimate_prepare_p2p_f32).
    ``xchg`` = (device tensor of the ranks' exchange-buffer pointers, rank, world, epoch): dist.PeerExchange.next()."""
    _require(x, "x")
    ptrs, rank, world, epoch = xchg
    mb, nb, sb = fmt
    _require_state(cur_min, 1, "cur_min")
    _require_state(cur_max, 1, "cur_max")
    _require_state(maxval_out, 1, "maxval_out")
    _require_table(table_out, 1, mb, nb, sb, "table_out")
    check(lib().fp8fq_estimate_prepare_p2p_f32(x.data_ptr(), x.numel(), cur_min.data_ptr(), cur_max.data_ptr(), int(est_mode),
                                               1 if initialized else 0, float(momentum), maxval_out.data_ptr(), float(mb),
                                               int(nb), int(sb), table_out.data_ptr(), _workspace(x.device).data_ptr(),
                                               ptrs.data_ptr(), int(rank), int(world), int(epoch), _stream()),
          "fp8fq_estimate_prepare_p2p_f32")
    return maxval_out, table_out


def bn_act_estimate_prepare_p2p(x, bn_scale, bn_shift, act: int, bn_mode: int, cur_min, cur_max, est_mode: int,
                                initialized: bool, momentum: float, maxval_out, fmt, table_out, xchg) -> bool:
    """The BN-fused twin (fp8fq_bn_act_estimate_prepare_p2p_f32): statistics of act(bn(x)) without materialising it, the
    exchange over peer memory, range and table -- one launch.  False when the shape is not covered by the fused kernel
    (decided from the shape alone, hence identically on every rank; the exchange epoch is then not consumed)."""
    _require(x, "x")
    if x.numel() >= MAX_FUSED_ELEMS:
        return False
    ptrs, rank, world, epoch = xchg
    Cbn = bn_scale.numel() // (4 if bn_mode == 1 else 1)
    rows, hw = _rows_hw(x, Cbn)
    nhwc = hw == 1 or is_channels_last(x)
    mb, nb, sb = fmt
    _require_state(cur_min, 1, "cur_min")
    _require_state(cur_max, 1, "cur_max")
    _require_state(maxval_out, 1, "maxval_out")
    _require_table(table_out, 1, mb, nb, sb, "table_out")
    code = lib().fp8fq_bn_act_estimate_prepare_p2p_f32(
        x.data_ptr(), x.numel() // Cbn if nhwc else rows, hw, Cbn, 1 if nhwc else 0, bn_scale.data_ptr(), _opt_ptr(bn_shift),
        int(bn_mode), int(act), cur_min.data_ptr(), cur_max.data_ptr(), int(est_mode), 1 if initialized else 0,
        float(momentum), maxval_out.data_ptr(), float(mb), int(nb), int(sb), table_out.data_ptr(),
        _workspace(x.device).data_ptr(), ptrs.data_ptr(), int(rank), int(world), int(epoch), _stream())
    if code == -2:
        return False
    check(code, "fp8fq_bn_act_estimate_prepare_p2p_f32")
    return True


def dp_finish_prepare(packed, cur_min, cur_max, est_mode: int, initialized: bool, momentum: float, maxval_out, fmt,
                      table_out):
    """Second half of a data-parallel calibration step (fp8fq_dp_finish_prepare_f32): ``packed`` = the all-reduced
    [-min (C) | max (C) | NaN flag (C)]; applies the estimator rule to (cur_min, cur_max), set_quant_range and builds the table."""
    _require(packed, "packed")
    C = packed.numel() // 3
    _require_state(cur_min, C, "cur_min")
    _require_state(cur_max, C, "cur_max")
    mb, nb, sb = fmt
    _require_state(maxval_out, C, "maxval_out")
    _require_table(table_out, C, mb, nb, sb, "table_out")
    check(lib().fp8fq_dp_finish_prepare_f32(packed.data_ptr(), C, cur_min.data_ptr(), cur_max.data_ptr(), int(est_mode),
                                            1 if initialized else 0, float(momentum), maxval_out.data_ptr(), float(mb),
                                            int(nb), int(sb), table_out.data_ptr(), _stream()),
          "fp8fq_dp_finish_prepare_f32")
    return maxval_out, table_out


def mse_grid(x, per_channel: bool, grid: torch.Tensor, mbit_list, n_bits: int, sign_bits: int, mses: torch.Tensor):
    """mses[m, g, c] += mean((x - Q(x; grid[g, c], mbit_list[m]))^2)  (range_estimators.py:337-347)."""
    _require(x, "x")
    _require(grid, "grid")
    _require(mses, "mses")
    n = x.numel()
    C = x.shape[0] if per_channel else 1
    G = grid.shape[0]
    Mn = len(mbit_list)
    if grid.numel() != G * C or mses.numel() != Mn * G * C:
        raise Fp8fqError(f"mse_grid: grid must be [G, C] = [{G}, {C}] and mses [Mn, G, C] = [{Mn}, {G}, {C}]")
    arr = (ctypes.c_float * Mn)(*[float(m) for m in mbit_list])
    tf = lib().fp8fq_mse_table_floats(arr, Mn, int(n_bits), int(sign_bits), G, C)
    if tf < 0:
        check(int(tf), "fp8fq_mse_table_floats")
    scratch = torch.empty(int(tf) + 2, dtype=torch.float32, device=x.device)
    check(lib().fp8fq_mse_grid_f32(x.data_ptr(), n, C, n // C, grid.data_ptr(), G, arr, Mn, int(n_bits),
                                   int(sign_bits), mses.data_ptr(), scratch.data_ptr(), _stream()),
          "fp8fq_mse_grid_f32")
    return mses


def fake_quant_host(x_host: torch.Tensor, maxval_host: torch.Tensor, mantissa_bits: float, n_bits: int,
                    sign_bits: int, per_channel: bool = False, out: torch.Tensor = None, device: int = None):
    """End-to-end call with HOST tensors (fp8fq_fake_quant_host_f32): H2D, quantise, D2H inside."""
    if x_host.is_cuda or x_host.dtype != torch.float32 or not x_host.is_contiguous():
        raise Fp8fqError("fake_quant_host: x must be a contiguous float32 CPU tensor")
    if out is None:
        out = torch.empty_like(x_host)
    C = x_host.shape[0] if per_channel else 1
    n = x_host.numel()
    mv = maxval_host.detach().to(torch.float32).contiguous().cpu()
    if mv.numel() != C:
        raise Fp8fqError("fake_quant_host: maxval must have one entry per channel")
    dev = torch.cuda.current_device() if device is None else device
    check(lib().fp8fq_fake_quant_host_f32(x_host.data_ptr(), out.data_ptr(), mv.data_ptr(), n, C, n // C,
                                          float(mantissa_bits), int(n_bits), int(sign_bits), int(dev)),
          "fp8fq_fake_quant_host_f32")
    return out


def launch_count() -> int:
    return int(lib().fp8fq_launch_count())
