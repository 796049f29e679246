"""Drop-in ``FPQuantizer`` backed by libfp8fq.so.

Mirrors the reference's quantizer contract:
  * ``QuantizerBase``  -- quantization/quantizers/base_quantizers.py:8-47
  * ``FPQuantizer``    -- quantization/quantizers/fp8_quantizer.py:151-272
Same constructor kwargs, same public attributes (``n_bits, per_channel, mantissa_bits, maxval,
ebits, default_bias, set_maxval, allow_unsigned, sign_bits, mse_include_mantissa_bits,
learning_maxval, learning_mantissa_bits``), same ``set_quant_range`` semantics -- but ``forward`` is
ONE kernel launch reading each element once and writing it once, and nothing on the calibration or
validation path synchronises with the host (the reference does a D2H copy per activation quantiser
per calibration batch, fp8_quantizer.py:239-240).
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from ._lib import Fp8fqError


class QuantizerNotInitializedError(Exception):
    """quantization/quantizers/utils.py:6-12."""

    def __init__(self):
        super().__init__("Quantizer has not been initialized yet")


class QuantizerBase(nn.Module):
    """base_quantizers.py:8-47 (type contract; properties raise until a subclass defines them)."""

    def __init__(self, n_bits, per_channel=False, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.n_bits = n_bits
        self.per_channel = per_channel
        self.state = None
        self.x_min_fp32 = self.x_max_fp32 = None

    @property
    def is_initialized(self):
        raise NotImplementedError()

    @property
    def x_max(self):
        raise NotImplementedError()

    @property
    def symmetric(self):
        raise NotImplementedError()

    @property
    def x_min(self):
        raise NotImplementedError()

    def forward(self, x_float):
        raise NotImplementedError()

    def _adjust_params_per_channel(self, x):
        raise NotImplementedError()

    def set_quant_range(self, x_min, x_max):
        raise NotImplementedError()

    def extra_repr(self):
        return "n_bits={}, per_channel={}, is_initalized={}".format(self.n_bits, self.per_channel, self.is_initialized)

    def reset(self):
        self._delta = None


def default_maxval(n_bits: int, mantissa_bits: int) -> float:
    """fp8_quantizer.py:171-179: largest value of the format with the default integer bias."""
    ebits = n_bits - mantissa_bits - 1
    return (2 - 2 ** (-mantissa_bits)) * 2 ** (2**ebits - 1 - 2 ** (ebits - 1))


class _FakeQuantSTE(torch.autograd.Function):
    """Autograd node of the fake-quantiser with the reference's gradient semantics (round_ste_func, detached
    exponent code, torch.max/min clamp): forward = the streaming kernel, backward = fq_backward_kernel + a few
    [C]-sized tensor ops.  SURVEY section 8f4 (fp8_quantizer.py:242-254, rounding_utils.py:12-19)."""

    @staticmethod
    def forward(ctx, x, maxval, mantissa_bits, table, C, mbits_host, n_bits, sign_bits):
        x = x.contiguous()
        ctx.save_for_backward(x, maxval.detach(), table)
        ctx.fmt = (C, mbits_host, n_bits, sign_bits)
        return ops.fake_quant(x.detach(), table, C, mbits_host, n_bits, sign_bits)

    @staticmethod
    def backward(ctx, g):
        import math

        x, maxval, table = ctx.saved_tensors
        C, mb, nb, sb = ctx.fmt
        gx, acc = ops.fake_quant_backward(g.contiguous(), x, table, C, mb, nb, sb)
        g_maxval = g_mbits = None
        if ctx.needs_input_grad[1]:
            g_maxval = (acc[:, 0] + acc[:, 1] / maxval.reshape(-1).double()).float().reshape(maxval.shape)
        if ctx.needs_input_grad[2]:
            M, E, _ = ops.format_split(mb, nb, sb)
            rb = round(mb)
            inside = 1.0 if 1 <= rb <= nb - sb else 0.0  # torch.clamp passes the gradient inside [min, max]
            dbias_dM = -(2.0**E) * math.log(2.0) + 2.0 ** (-M) / (2.0 - 2.0 ** (-M))
            coef = math.log(2.0) * (-1.0 - dbias_dM) * inside
            g_mbits = (acc[:, 1].sum() * coef).float().reshape(1)
        return (gx if ctx.needs_input_grad[0] else None), g_maxval, g_mbits, None, None, None, None, None


def quantize_to_fp8_ste_MM(x_float, n_bits, maxval, num_mantissa_bits, sign_bits):
    """The reference's functional entry point, fp8_quantizer.py:91-133, with its signature: ``maxval`` a ``[1]`` (per
    tensor) or ``[C]`` / ``[C, 1, ...]`` (per channel, C = x.shape[0], :108-109) tensor, ``num_mantissa_bits`` a ``[1]``
    tensor or a number.  Two launches -- the per-channel prologue (:105-110) and the streaming kernel (:112-132);
    ``FPQuantizer`` caches the prologue's table, this function rebuilds it on every call.  The format split decides the
    table layout on the host, so a device-resident ``num_mantissa_bits`` costs one ``.item()`` here (the module keeps
    a host copy instead).  With grad mode on and a differentiable input it returns the STE graph node, like the
    reference (round_ste_func, rounding_utils.py:12-19)."""
    x = ops.dense(x_float)
    if not isinstance(maxval, torch.Tensor):
        maxval = torch.tensor([float(maxval)], dtype=torch.float32, device=x.device)
    if isinstance(num_mantissa_bits, torch.Tensor):
        mb = float(num_mantissa_bits.detach().reshape(-1)[0].item())
        mbt = num_mantissa_bits
    else:
        mb = float(num_mantissa_bits)
        mbt = torch.tensor([mb], dtype=torch.float32, device=x.device)
    mv = maxval.reshape(-1)
    if mv.device != x.device or mv.dtype != torch.float32:
        mv = mv.to(device=x.device, dtype=torch.float32)
    mv = mv.contiguous()
    C = mv.numel()
    if C != 1 and (x.dim() == 0 or x.shape[0] != C):
        raise Fp8fqError(f"per-channel maxval has {C} entries but x has shape {tuple(x.shape)}")
    table = ops.prepare(mv.detach(), mb, int(n_bits), int(sign_bits))
    if torch.is_grad_enabled() and (x.requires_grad or mv.requires_grad or mbt.requires_grad):
        if mbt.device != x.device:
            mbt = mbt.to(x.device)
        return _FakeQuantSTE.apply(x, mv, mbt, table, C, mb, int(n_bits), int(sign_bits))
    return ops.fake_quant(x, table, C, mb, int(n_bits), int(sign_bits))


def get_max_value(num_exponent_bits: int = 4, bias: int = 8):
    """fp8_quantizer.py:82-88: largest value of an 8-bit format with integer bias (no inf / NaN codes reserved)."""
    num_fraction_bits = 7 - num_exponent_bits
    return 2 ** (2**num_exponent_bits - 1 - bias) * (2 - 2**-num_fraction_bits)


class FPQuantizer(QuantizerBase):
    """8-bit (runtime bit-split) floating-point fake quantiser -- fp8_quantizer.py:151-272."""

    def __init__(self, *args, scale_domain=None, mantissa_bits=4, maxval=3, set_maxval=False, learn_maxval=False,
                 learn_mantissa_bits=False, mse_include_mantissa_bits=True, allow_unsigned=False, **kwargs):
        super().__init__(*args, **kwargs)
        self.ebits = self.n_bits - mantissa_bits - 1
        self.default_bias = 2 ** (self.ebits - 1)
        mv = maxval if maxval is not None else default_maxval(self.n_bits, mantissa_bits)
        self._table = None
        self._table_key = None
        self._mbits_host = float(mantissa_bits)
        self._mantissa_bits = torch.Tensor([float(mantissa_bits)])
        self._maxval = torch.Tensor([mv])
        self.set_maxval = set_maxval
        self.learning_maxval = learn_maxval
        self.learning_mantissa_bits = learn_mantissa_bits
        self.mse_include_mantissa_bits = mse_include_mantissa_bits
        self.allow_unsigned = allow_unsigned
        self.sign_bits = 1

    # -- attributes the reference exposes as plain tensors (fp8_quantizer.py:183-184) -----------
    @property
    def maxval(self):
        return self._maxval

    @maxval.setter
    def maxval(self, value):
        if not isinstance(value, torch.Tensor):
            value = torch.Tensor([float(value)])
        if value.dim() == 0:
            value = value.reshape(1)
        if isinstance(self._maxval, nn.Parameter) and not isinstance(value, nn.Parameter):
            del self._maxval  # un-register the Parameter before storing a plain tensor under the same name
        self._maxval = value
        self._table_key = None

    @property
    def mantissa_bits(self):
        return self._mantissa_bits

    @mantissa_bits.setter
    def mantissa_bits(self, value):
        # The format split decides the table layout, so the host needs the value; a CUDA tensor
        # assigned here costs one .item() -- the library itself never does that.
        if isinstance(self.__dict__.get("_parameters", {}).get("_mantissa_bits"), nn.Parameter) and not isinstance(
                value, nn.Parameter):
            del self._mantissa_bits
        if isinstance(value, torch.Tensor):
            self._mbits_host = float(value.detach().reshape(-1)[0].item())
            self._mantissa_bits = value
        else:
            self._mbits_host = float(value)
            self._mantissa_bits = torch.Tensor([float(value)])
        self._table_key = None

    # -- state dict: the reference registers learnable ranges as ``maxval`` / ``mantissa_bits`` (fp8_quantizer.py:
    # 248-254); here they live behind properties as ``_maxval`` / ``_mantissa_bits``.  Keys are translated both ways so
    # that checkpoints travel between the two implementations.
    _STATE_NAMES = (("_maxval", "maxval"), ("_mantissa_bits", "mantissa_bits"))

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        for ours, ref in self._STATE_NAMES:
            if prefix + ours in destination:
                destination[prefix + ref] = destination.pop(prefix + ours)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        for ours, ref in self._STATE_NAMES:
            if prefix + ref in state_dict and ours in self._parameters:
                state_dict[prefix + ours] = state_dict.pop(prefix + ref)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._table_key = None

    # -- table management -------------------------------------------------------------------------
    def _ensure_table(self, device):
        mv = self._maxval
        if mv.device != device or mv.dtype != torch.float32 or not mv.is_contiguous():
            mv = mv.detach().to(device=device, dtype=torch.float32).contiguous()
            self._maxval = mv  # same lazy move as fp8_quantizer.py:195-196
        # keyed on the maxval tensor OBJECT (kept alive in _table_src: a tensor re-bound behind the setters' back cannot
        # alias a recycled address) and its version counter
        key = (mv.data_ptr(), mv._version, mv.numel(), self._mbits_host, self.n_bits, self.sign_bits)
        if key != self._table_key or self.__dict__.get("_table_src") is not mv:
            self._table = ops.prepare(mv.detach(), self._mbits_host, self.n_bits, self.sign_bits)
            self._table_key = key
            self.__dict__["_table_src"] = mv
        return self._table

    def adopt_range(self, maxval: torch.Tensor, table: torch.Tensor):
        """Install a (maxval, table) pair produced on device by the fused estimate/set-range kernels."""
        if isinstance(self._maxval, nn.Parameter):
            del self._maxval
        self._maxval = maxval
        self._table = table
        self._table_key = (maxval.data_ptr(), maxval._version, maxval.numel(), self._mbits_host, self.n_bits,
                           self.sign_bits)
        self.__dict__["_table_src"] = maxval

    @property
    def table(self):
        return self._table

    def table_for(self, x: torch.Tensor):
        """(table, C) for quantising ``x`` -- validates the channel layout like fp8_quantizer.py:108-109."""
        table = self._ensure_table(x.device)
        C = self._maxval.numel()
        if C != 1:
            if x.dim() == 0 or x.shape[0] != C:
                raise Fp8fqError(f"per-channel maxval has {C} entries but x has shape {tuple(x.shape)}")
        return table, C

    # -- the hot path ---------------------------------------------------------------------------------
    def forward(self, x_float):
        x = ops.dense(x_float)
        if isinstance(self._mantissa_bits, nn.Parameter):  # learnable mantissa width: re-read the value (one .item())
            mb = float(self._mantissa_bits.detach().reshape(-1)[0].item())
            if mb != self._mbits_host:
                self._mbits_host = mb
                self._table_key = None
        table, C = self.table_for(x)
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or self._maxval.requires_grad
                                                  or self._mantissa_bits.requires_grad)
        if needs_grad:  # straight-through estimator, as the reference's autograd graph
            mbt = self._mantissa_bits
            if mbt.device != x.device and not isinstance(mbt, nn.Parameter):
                mbt = mbt.to(x.device)
            return _FakeQuantSTE.apply(x, self._maxval, mbt, table, C, self._mbits_host, self.n_bits, self.sign_bits)
        return ops.fake_quant(x, table, C, self._mbits_host, self.n_bits, self.sign_bits)

    def quantize_with_codes(self, x_float):
        """(y, codes) -- test hook exposing the reference's intermediate integers (:128, :132)."""
        x = x_float.contiguous()
        table, C = self.table_for(x)
        return ops.fake_quant_codes(x, table, C, self._mbits_host, self.n_bits, self.sign_bits)

    # -- reference API ------------------------------------------------------------------------------
    def is_initialized(self):  # a METHOD in the reference (fp8_quantizer.py:207-208): always truthy
        return True

    def symmetric(self):
        return False

    def effective_bit_width(self):
        return None

    def _make_unsigned(self, x_min):
        if isinstance(x_min, torch.Tensor):
            return self.allow_unsigned and bool(torch.all(x_min >= 0))
        return self.allow_unsigned and x_min >= 0

    def set_quant_range(self, x_min, x_max):
        """fp8_quantizer.py:222-240.  ``maxval`` stays on the device (shape [C] or [1])."""
        if self._make_unsigned(x_min):
            self.sign_bits = 0  # sticky, as in the reference
            self._table_key = None
        if not self.set_maxval:
            return
        if not isinstance(x_max, torch.Tensor):
            dev = self._maxval.device if ops.on_device(self._maxval) else ops.default_device()
            x_max = torch.tensor([float(x_max)], dtype=torch.float32, device=dev)
            x_min = torch.tensor([float(x_min)], dtype=torch.float32, device=dev)
        if not isinstance(x_min, torch.Tensor):
            x_min = torch.full_like(x_max, float(x_min))
        if not ops.on_device(x_max):
            raise Fp8fqError("set_quant_range: range tensors must live on the GPU (no CPU path)")
        x_min = x_min.detach().to(torch.float32).reshape(-1).contiguous()
        x_max = x_max.detach().to(torch.float32).reshape(-1).contiguous()
        if x_min.numel() != x_max.numel():
            x_min = x_min.expand_as(x_max).contiguous()
        maxval, table = ops.set_range_prepare(x_min, x_max, self._mbits_host, self.n_bits, self.sign_bits)
        self.adopt_range(maxval, table)

    def make_range_trainable(self):  # fp8_quantizer.py:242-246
        if self.learning_maxval:
            self.learn_maxval()
        if self.learning_mantissa_bits:
            self.learn_mantissa_bits()

    def _param_device(self):
        return self._maxval.device if ops.on_device(self._maxval) else ops.default_device()

    def learn_maxval(self):  # :248-250 (registered as ``_maxval``; ``.maxval`` returns the Parameter)
        self.learning_maxval = True
        if not isinstance(self._maxval, nn.Parameter):
            self._maxval = nn.Parameter(self._maxval.detach().to(self._param_device(), torch.float32).contiguous())
            self._table_key = None

    def learn_mantissa_bits(self):  # :252-254
        self.learning_mantissa_bits = True
        if not isinstance(self._mantissa_bits, nn.Parameter):
            self._mantissa_bits = nn.Parameter(self._mantissa_bits.detach().to(self._param_device(), torch.float32)
                                               .reshape(-1).contiguous())

    def fix_ranges(self):  # :256-260 (the reference calls an undefined helper here; this is what it intends)
        if isinstance(self._maxval, nn.Parameter):
            mv = self._maxval.detach().clone()
            del self._maxval  # un-register before storing a plain tensor under the same name
            self._maxval = mv
            self._table_key = None
        if isinstance(self._mantissa_bits, nn.Parameter):
            mb = self._mantissa_bits.detach().clone()
            del self._mantissa_bits
            self._mantissa_bits = mb
            self._mbits_host = float(mb.reshape(-1)[0].item())
            self._table_key = None

    def extra_repr(self):
        M, E, _ = ops.format_split(self._mbits_host, self.n_bits, self.sign_bits)
        tag = "[per_channel]" if self._maxval.numel() > 1 else "per_tensor"
        return f"Exponent: {E} bits; mantissa: {M} bits; sign: {self.sign_bits}; maxval: {tag}"

    def __deepcopy__(self, memo):
        # LineSearchEstimator deep-copies its quantiser (range_estimators.py:201); tables are plain tensors
        import copy

        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            setattr(new, k, copy.deepcopy(v, memo))
        new._table_key = None
        return new


class AsymmetricUniformQuantizer(QuantizerBase):
    """INT asymmetric uniform fake quantiser -- quantization/quantizers/uniform_quantizers.py:13-256, the reference's
    default ``method`` and comparison baseline.  Forward path, ``scale_domain="linear"``,
    round-to-nearest-even discretizer; ``delta`` / ``zero_float`` live on the device as in the reference."""

    _SYMMETRIC = False
    # ``delta = range / int_max`` divides by a Python scalar: ATen does a true division on the CPU and multiplies by the
    # fp32 reciprocal on CUDA.  True (default) reproduces the reference as run on the GPU; False the CPU run.
    aten_cuda_scalar_div = True

    def __init__(self, n_bits, scale_domain="linear", discretizer=None, discretizer_args=tuple(), grad_scaling=False,
                 eps=1e-8, **kwargs):
        kwargs.pop("mantissa_bits", None)  # fp8_kwargs are forwarded to every quantiser by QuantizedModule
        for k in ("maxval", "set_maxval", "learn_maxval", "learn_mantissa_bits", "mse_include_mantissa_bits",
                  "allow_unsigned"):
            kwargs.pop(k, None)
        super().__init__(n_bits=n_bits, **kwargs)
        if scale_domain != "linear":
            raise NotImplementedError("only scale_domain='linear' is implemented")
        if grad_scaling:
            raise NotImplementedError("forward-only engine: grad_scaling needs the STE backward")
        self.register_buffer("_delta", None)
        self.register_buffer("_zero_float", None)
        self.register_buffer("_signed", None)
        self.scale_domain = scale_domain
        self.grad_scaling = grad_scaling
        self.eps = eps
        self._table = None

    @property
    def delta(self):
        if self._delta is None:
            raise QuantizerNotInitializedError()
        return self._delta

    @property
    def zero_float(self):
        if self._zero_float is None:
            raise QuantizerNotInitializedError()
        return self._zero_float

    @property
    def is_initialized(self):
        return self._delta is not None

    @property
    def symmetric(self):
        return self._SYMMETRIC

    @property
    def int_min(self):
        return 0.0

    @property
    def int_max(self):
        return 2.0**self.n_bits - 1

    @property
    def scale(self):
        return torch.clamp(self.delta, min=self.eps)

    @property
    def zero_point(self):
        return torch.clamp(torch.round(self.zero_float), self.int_min, self.int_max)

    @property
    def x_max(self):
        return self.scale * (self.int_max - self.zero_point)

    @property
    def x_min(self):
        return self.scale * (self.int_min - self.zero_point)

    def _as_device_ranges(self, x_min, x_max):
        if not torch.is_tensor(x_max):
            dev = ops.default_device()
            x_min = torch.tensor([float(x_min)], dtype=torch.float32, device=dev)
            x_max = torch.tensor([float(x_max)], dtype=torch.float32, device=dev)
        if x_min.dim() > 0 and x_min.numel() > 1 and not self.per_channel:
            raise ValueError("x_min and x_max must be a float or 1-D Tensor for per-tensor quantization "
                             "(per_channel=False)")
        if not ops.on_device(x_max):
            raise Fp8fqError("set_quant_range: range tensors must live on the GPU (no CPU path)")
        return (x_min.detach().to(torch.float32).reshape(-1).contiguous(),
                x_max.detach().to(torch.float32).reshape(-1).contiguous())

    def set_quant_range(self, x_min, x_max):
        self.x_min_fp32, self.x_max_fp32 = x_min, x_max
        mn, mx = self._as_device_ranges(x_min, x_max)
        delta, zero_float, signed, table = ops.uniform_prepare(mn, mx, self.n_bits, self._SYMMETRIC, self.eps,
                                                               self.aten_cuda_scalar_div)
        self._delta = delta if delta.numel() > 1 else delta.reshape(())
        if not self._SYMMETRIC:
            self._zero_float = zero_float if zero_float.numel() > 1 else zero_float.reshape(())
        self._signed = signed
        self._table = table

    def forward(self, x_float, *args, **kwargs):
        if self._table is None:
            raise QuantizerNotInitializedError()
        if torch.is_grad_enabled() and x_float.requires_grad:
            raise Fp8fqError("uniform quantiser: forward-only engine; call under torch.no_grad()")
        x = ops.dense(x_float)
        C = self._table.numel() // 8
        if C != 1 and (x.dim() == 0 or x.shape[0] != C):
            raise Fp8fqError(f"per-channel range has {C} entries but x has shape {tuple(x.shape)}")
        return ops.uniform_quant(x, self._table, C)

    def make_range_trainable(self):
        raise NotImplementedError("learnable ranges need the STE backward (SURVEY section 8 row f4)")

    def fix_ranges(self):
        pass

    def reset(self):
        self._delta = None
        self._zero_float = None
        self._table = None


class SymmetricUniformQuantizer(AsymmetricUniformQuantizer):
    """uniform_quantizers.py:259-331."""

    _SYMMETRIC = True

    @property
    def signed(self):
        if self._signed is None:
            raise QuantizerNotInitializedError()
        return bool(self._signed.item())

    @property
    def int_min(self):
        return -(2.0 ** (self.n_bits - 1)) if self.signed else 0

    @property
    def int_max(self):
        return 2.0 ** (self.n_bits - int(self.signed)) - 1

    @property
    def zero_point(self):
        return 0.0

    def generate_grid(self):  # uniform_quantizers.py:328-331
        x_int_rng = torch.arange(self.int_min, self.int_max + 1, device=self.delta.device)
        return self.scale * (x_int_rng - self.zero_point)
