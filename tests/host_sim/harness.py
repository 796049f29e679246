"""TEST ONLY: puts the host simulation of the kernels (libfp8fq_sim.so: the product's own .cu compiled by g++, see
README.md in this directory) behind the product's Python layer by substituting the package's single device gate
(ops.on_device / default_device), the argument check, the stream handle, the min/max workspace and the loaded library.
Used by the ``simdev`` fixture (tests/conftest.py) and by the spawned workers of the gloo tests.  The product itself
has no CPU path: without this a CPU tensor raises Fp8fqError."""
import ctypes
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def install_simulation():
    """Returns a zero-argument ``restore`` callable."""
    from fp8_quantization_b200 import _lib, ops

    handle = ctypes.CDLL(os.environ.get("FP8FQ_SIM_LIB") or os.path.join(ROOT, "oracle", "_build", "libfp8fq_sim.so"))
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype, fn.argtypes = res, args

    def require(t, name):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if t.dtype != torch.float32:
            raise ops.Fp8fqError(f"{name} must be float32; got {t.dtype}")
        if not t.is_contiguous() and not ops.is_channels_last(t):
            raise ops.Fp8fqError(f"{name} must be contiguous (or dense channels_last)")

    ws = torch.zeros(int(handle.fp8fq_minmax_workspace_bytes()) // 4, dtype=torch.int32)
    saved = (_lib._lib, ops.on_device, ops.default_device, ops._require, ops._stream, ops._workspace)
    _lib._lib = handle
    ops.on_device = lambda t: isinstance(t, torch.Tensor)
    ops.default_device = lambda: torch.device("cpu")
    ops._require = require
    ops._stream = lambda: None
    ops._workspace = lambda device: ws

    def restore():
        _lib._lib, ops.on_device, ops.default_device, ops._require, ops._stream, ops._workspace = saved

    return restore
