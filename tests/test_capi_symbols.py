"""CPU: libfp8fq.so loads without a GPU, exports every symbol include/fp8fq.h declares, and its
host-side helpers / argument validation behave as documented.  No kernel is launched."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fp8fq.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fp8fq_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(built):
    from fp8_quantization_b200 import _lib

    handle = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/fp8fq.h but not exported"
    # the Python binding table mirrors the header one to one
    assert sorted(_lib.SIGNATURES) == syms


def test_format_split_and_table_sizes(built):
    from fp8_quantization_b200 import ops
    from fp8_quantization_b200._lib import Fp8fqError

    assert ops.format_split(5, 8, 1) == (5, 2, 3)
    assert ops.format_split(4, 8, 1) == (4, 3, 7)
    assert ops.format_split(2.5, 8, 1) == (2, 5, 31)   # torch.round is half-to-even
    assert ops.format_split(3.5, 8, 1) == (4, 3, 7)
    assert ops.format_split(0, 8, 1) == (1, 6, 63)     # clamp(…, 1, n_bits - sign_bits)
    assert ops.format_split(9, 8, 1) == (7, 0, 1)
    assert ops.format_split(9, 8, 0) == (8, 0, 1)
    assert ops.format_split(7, 8, 1) == (7, 0, 1)
    assert ops.table_stride(5, 8, 1) == 8 + 4 + 8 + 4
    with pytest.raises(Fp8fqError):
        ops.format_split(1, 16, 1)  # E = 14 > 7 unsupported
    with pytest.raises(Fp8fqError):
        ops.format_split(float("nan"), 8, 1)


def test_argument_errors_without_gpu(built):
    from fp8_quantization_b200._lib import lib

    L = lib()
    assert L.fp8fq_version() >= 100
    assert b"sm_100a" in L.fp8fq_build_info()
    # null pointers / inconsistent sizes are rejected before any CUDA call
    assert L.fp8fq_fake_quant_f32(None, None, None, 16, 1, 16, 5.0, 8, 1, None) == -1
    assert L.fp8fq_fake_quant_f32(None, None, None, 16, 4, 5, 5.0, 8, 1, None) == -1     # n != C * inner
    assert L.fp8fq_fake_quant_f32(None, None, None, 0, 1, 0, 5.0, 8, 1, None) == 0        # empty tensor: no-op
    assert L.fp8fq_fake_quant_f32(None, None, None, 16, 1, 16, 1.0, 16, 1, None) == -2    # unsupported format
    assert L.fp8fq_prepare_f32(None, 1, 5.0, 8, 1, None, None) == -1
    assert L.fp8fq_minmax_f32(None, 16, 1, 16, None, None, 0, 0, 0.9, None, None) == -1
    assert L.fp8fq_bn_act_quant_f32(None, None, None, None, 4, 0, 4, 0, 0, None, 5.0, 8, 1, None) == -1
    assert L.fp8fq_add_act_quant_f32(None, None, None, 8, 7, None, 5.0, 8, 1, None) == -1  # bad activation code
    assert L.fp8fq_minmax_workspace_bytes() >= 8192


def test_cpu_tensors_are_rejected_loudly(built):
    """There is no CPU fallback: a CPU tensor must raise, not silently run somewhere else."""
    import torch

    import fp8_quantization_b200 as fq

    q = fq.FPQuantizer(8, mantissa_bits=5)
    with pytest.raises(fq.Fp8fqError):
        q(torch.randn(16))
    est = fq.CurrentMinMaxEstimator()
    with pytest.raises(fq.Fp8fqError):
        est(torch.randn(16))


def test_torch_operator_library_loads_and_refuses_cpu_tensors(built):
    """libfp8fq_torch.so (csrc/fp8fq_torch.cpp, TORCH_LIBRARY(fp8fq, ...)): builds here without a GPU, registers the
    hot-path operators with the schemas the Python layer calls, agrees with libfp8fq.so about the ABI version, and --
    like every other entry into the library -- refuses CPU tensors instead of computing anything."""
    import pytest
    import torch

    import fp8_quantization_b200 as fq
    from fp8_quantization_b200 import ops

    T = ops.torch_binding()
    assert T is not None and T.abi_version() == fq.lib().fp8fq_version()
    for name in ("fake_quant", "fake_quant_multi", "bn_act_quant", "add_act_quant", "bn_quant_add_act_quant", "minmax",
                 "estimate_prepare", "bn_act_estimate_prepare"):
        assert hasattr(T, name), name
    x, table = torch.zeros(8), torch.zeros(64)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        T.fake_quant(x, table, 1, 5.0, 8, 1)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        T.add_act_quant(x, x, 1, table, 5.0, 8, 1)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        T.bn_act_quant(x.view(1, 8), table, None, 1, table, 5.0, 8, 1, 0)
    # the package's own wrapper takes the ctypes route for CPU tensors and fails there with its own error type
    with pytest.raises(fq.Fp8fqError):
        ops.fake_quant(x, table, 1, 5.0, 8, 1)
