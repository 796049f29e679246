"""CPU restatement of the reference FP8 fake-quant hot path.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it restates (paths relative to
/root/reference).  The arithmetic of the reference lives in PyTorch ATen ops, so
the restatement uses the *same ATen ops in the same order* (anything else would
not be a restatement of its fp32 behaviour: ``bias`` and ``scales`` are
non-integer powers of two computed with fp32 ``log2``/``pow``).  It is device
agnostic: run with CPU tensors it is the reference's CPU path; run with CUDA
tensors (inside ``-m gpu`` tests only) it is "the reference as shipped, run with
--cuda on the same B200".

Parity pin: ``tests/golden/make_golden.py`` runs the real reference and this file
on the same seeded inputs and asserts bit equality before writing the golden
fixtures; ``tests/test_oracle_golden.py`` re-checks this file against the
committed fixtures.
"""
from __future__ import annotations

import torch

# --------------------------------------------------------------------------------------
# quantize_to_fp8_ste_MM  (quantization/quantizers/fp8_quantizer.py:91-133)
# --------------------------------------------------------------------------------------


def mantissa_exponent_split(mantissa_bits: torch.Tensor, n_bits: int, sign_bits: int):
    """fp8_quantizer.py:105-106.  ``round_ste_func`` forward is torch.round
    (rounding_utils.py:12-19), i.e. half-to-even."""
    M = torch.clamp(torch.round(mantissa_bits), 1, n_bits - sign_bits)
    E = n_bits - sign_bits - M
    return M, E


def fake_quant(x, n_bits: int, maxval, mantissa_bits, sign_bits: int, return_codes: bool = False):
    """fp8_quantizer.py:105-133, op for op.

    Returns ``y`` or ``(y, log_scales, q)`` where ``log_scales`` is the exponent
    code e (fp8_quantizer.py:128) and ``q = round(xc / scales)`` the mantissa
    integer (fp8_quantizer.py:132), both as float tensors.
    """
    M, E = mantissa_exponent_split(mantissa_bits, n_bits, sign_bits)
    if maxval.shape[0] != 1 and len(maxval.shape) != len(x.shape):  # :108-109
        maxval = maxval.view([-1] + [1] * (len(x.shape) - 1))
    bias = 2**E - torch.log2(maxval) + torch.log2(2 - 2 ** (-M)) - 1  # :110
    minval = -maxval if sign_bits == 1 else torch.zeros_like(maxval)  # :112
    xc = torch.min(torch.max(x, minval), maxval)  # :113
    log_scales = torch.clamp(torch.floor(torch.log2(torch.abs(xc)) + bias), 1.0)  # :128
    scales = 2.0 ** (log_scales - M - bias)  # :130
    q = torch.round(xc / scales)  # :132
    y = q * scales
    if return_codes:
        return y, log_scales, q
    return y


class _RoundSTE(torch.autograd.Function):
    """rounding_utils.py:12-19: forward torch.round, backward identity."""

    @staticmethod
    def forward(ctx, x):
        return torch.round(x)

    @staticmethod
    def backward(ctx, g):
        return g


def fake_quant_ste(x, n_bits: int, maxval, mantissa_bits, sign_bits: int):
    """fp8_quantizer.py:105-133 with the reference's autograd semantics (round_ste_func, detached log_scales):
    differentiable w.r.t. x, maxval and mantissa_bits -- the checker of the STE backward kernel."""
    M = torch.clamp(_RoundSTE.apply(mantissa_bits), 1, n_bits - sign_bits)
    E = n_bits - sign_bits - M
    if maxval.shape[0] != 1 and len(maxval.shape) != len(x.shape):
        maxval = maxval.view([-1] + [1] * (len(x.shape) - 1))
    bias = 2**E - torch.log2(maxval) + torch.log2(2 - 2 ** (-M)) - 1
    minval = -maxval if sign_bits == 1 else torch.zeros_like(maxval)
    xc = torch.min(torch.max(x, minval), maxval)
    log_scales = torch.clamp((torch.floor(torch.log2(torch.abs(xc)) + bias)).detach(), 1.0)
    scales = 2.0 ** (log_scales - M - bias)
    return _RoundSTE.apply(xc / scales) * scales


def quant_tables(n_bits: int, maxval, mantissa_bits, sign_bits: int):
    """The per-channel quantities the reference derives from (maxval, M): ``bias``
    (:110) and, for every exponent code e in [1, max(1, 2^E-1)], the scale
    ``2 ** (e - M - bias)`` (:130) evaluated with the same ATen ops.  Returns
    ``(bias [C], scales [C, K])``.  Used by tests to compare our device-side
    tables against the reference's own libm."""
    M, E = mantissa_exponent_split(mantissa_bits, n_bits, sign_bits)
    mv = maxval.reshape(-1, 1)
    bias = 2**E - torch.log2(mv) + torch.log2(2 - 2 ** (-M)) - 1
    K = max(1, int(2 ** int(E.item())) - 1)
    e = torch.arange(1, K + 1, dtype=torch.float32, device=maxval.device).view(1, -1)
    scales = 2.0 ** (e - M - bias)
    return bias.reshape(-1), scales


def canonical_codes(y, e, q, M: int):
    """Canonical (sign, exponent-code, mantissa-int) triple.

    The reference's ``q`` reaches ``2^(M+1)`` routinely; ``(e, 2^(M+1))`` is the
    same grid point as ``(e+1, 2^M)``.  Canonical form:
    q == 2^(M+1) -> (e+1, 2^M); q == 0 -> e = 0 (zero has no exponent).
    NaN -> (0, -1, -1).
    """
    top = float(2 ** (M + 1))
    half = float(2**M)
    qa = q.abs()
    ec = torch.where(qa == top, e + 1, e)
    qc = torch.where(qa == top, torch.full_like(qa, half), qa)
    ec = torch.where(qc == 0, torch.zeros_like(ec), ec)
    sign = torch.signbit(y) & (qc != 0)
    nan = torch.isnan(y)
    ec = torch.where(nan, torch.full_like(ec, -1), ec)
    qc = torch.where(nan, torch.full_like(qc, -1), qc)
    sign = sign & ~nan
    return sign.to(torch.int32), ec.to(torch.int32), qc.to(torch.int32)


def default_maxval(n_bits: int, mantissa_bits: int) -> float:
    """fp8_quantizer.py:171-179."""
    ebits = n_bits - mantissa_bits - 1
    default_bias = 2 ** (ebits - 1)
    return (2 - 2 ** (-mantissa_bits)) * 2 ** (2**ebits - 1 - default_bias)


# --------------------------------------------------------------------------------------
# FPQuantizer state (fp8_quantizer.py:151-260) -- only what the hot path touches
# --------------------------------------------------------------------------------------


class OracleFPQuantizer(torch.nn.Module):
    """Restates FPQuantizer's ctor/forward/set_quant_range (fp8_quantizer.py:156-240)."""

    def __init__(self, n_bits, per_channel=False, *, scale_domain=None, mantissa_bits=4, maxval=3,
                 set_maxval=False, learn_maxval=False, learn_mantissa_bits=False,
                 mse_include_mantissa_bits=True, allow_unsigned=False):
        super().__init__()
        self.n_bits = n_bits
        self.per_channel = per_channel
        self.state = None
        self.ebits = n_bits - mantissa_bits - 1
        self.default_bias = 2 ** (self.ebits - 1)
        mv = maxval if maxval is not None else default_maxval(n_bits, mantissa_bits)
        self.maxval = torch.Tensor([mv])
        self.mantissa_bits = torch.Tensor([float(mantissa_bits)])
        self.set_maxval = set_maxval
        self.learning_maxval = learn_maxval
        self.learning_mantissa_bits = learn_mantissa_bits
        self.mse_include_mantissa_bits = mse_include_mantissa_bits
        self.allow_unsigned = allow_unsigned
        self.sign_bits = 1

    def forward(self, x):  # :194-205
        if self.maxval.device != x.device:
            self.maxval = self.maxval.to(x.device)
        if self.mantissa_bits.device != x.device:
            self.mantissa_bits = self.mantissa_bits.to(x.device)
        return fake_quant(x, self.n_bits, self.maxval, self.mantissa_bits, self.sign_bits)

    def is_initialized(self):  # :207-208 (a method, always truthy where the manager tests it)
        return True

    def make_range_trainable(self):
        raise NotImplementedError("oracle covers the forward path only")

    def reset(self):  # base_quantizers.py:46-47
        self._delta = None

    def set_quant_range(self, x_min, x_max):  # :222-240
        if isinstance(x_min, torch.Tensor):
            unsigned = self.allow_unsigned and bool(torch.all(x_min >= 0))
        else:
            unsigned = self.allow_unsigned and x_min >= 0
        if unsigned:
            self.sign_bits = 0  # sticky, never reset (:216-225)
        if self.set_maxval:
            if not isinstance(x_max, torch.Tensor):
                x_max = torch.Tensor([x_max]).to(self.maxval.device)
                x_min = torch.Tensor([x_min]).to(self.maxval.device)
            if self.maxval.device != x_max.device:
                self.maxval = self.maxval.to(x_max.device)
            if self.mantissa_bits.device != x_max.device:
                self.mantissa_bits = self.mantissa_bits.to(x_max.device)
            self.maxval = torch.abs(torch.max(torch.abs(x_min), x_max))
            if len(self.maxval.shape) == 0:
                self.maxval = torch.Tensor([self.maxval])


# --------------------------------------------------------------------------------------
# INT uniform quantisers (quantization/quantizers/uniform_quantizers.py) -- SURVEY section 8f3, the reference's
# comparison baseline.  Forward path only, scale_domain="linear", round-to-nearest-even STE discretizer.
# --------------------------------------------------------------------------------------


class OracleAsymmetricUniform(torch.nn.Module):
    """AsymmetricUniformQuantizer, uniform_quantizers.py:13-256."""

    symmetric = False

    def __init__(self, n_bits, per_channel=False, eps=1e-8):
        super().__init__()
        self.n_bits = n_bits
        self.per_channel = per_channel
        self.eps = eps
        self._delta = None
        self._zero_float = None
        self.state = None

    @property
    def is_initialized(self):  # :68-70
        return self._delta is not None

    @property
    def int_min(self):  # :76-79
        return 0.0

    @property
    def int_max(self):  # :81-84
        return 2.0**self.n_bits - 1

    @property
    def scale(self):  # :86-91
        return torch.clamp(self._delta, min=self.eps)

    @property
    def zero_point(self):  # :93-97
        return torch.clamp(torch.round(self._zero_float), self.int_min, self.int_max)

    def _tensorize_min_max(self, x_min, x_max):  # :193-222
        if not torch.is_tensor(x_min):
            x_min = torch.tensor(x_min).float()
            x_max = torch.tensor(x_max).float()
        x_min = torch.min(x_min, torch.zeros_like(x_min))
        x_max = torch.max(x_max, torch.ones_like(x_max) * self.eps)
        return x_min, x_max

    def set_quant_range(self, x_min, x_max):  # :224-246
        x_min, x_max = self._tensorize_min_max(x_min, x_max)
        self._delta = (x_max - x_min) / self.int_max
        self._zero_float = -x_min / self._delta

    def _params_for(self, x):  # :176-191
        d, z = self._delta, self._zero_float
        if self.per_channel and x.ndim != d.ndim:
            shape = [-1] + [1] * (x.dim() - 1)
            d = d.view(shape)
            z = z.view(shape) if z is not None else None
        return d, z

    def forward(self, x):  # :107-164
        d, z = self._params_for(x)
        scale = torch.clamp(d, min=self.eps)
        zero_point = torch.clamp(torch.round(z), self.int_min, self.int_max)
        x_int = torch.round(x / scale) + zero_point
        x_int = torch.clamp(x_int, self.int_min, self.int_max)
        return scale * (x_int - zero_point)

    def reset(self):
        self._delta = None


class OracleSymmetricUniform(OracleAsymmetricUniform):
    """SymmetricUniformQuantizer, uniform_quantizers.py:259-331."""

    symmetric = True

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._signed = None

    @property
    def int_min(self):  # :290-292
        return -(2.0 ** (self.n_bits - 1)) if bool(self._signed) else 0

    @property
    def int_max(self):  # :294-297
        return 2.0 ** (self.n_bits - int(bool(self._signed))) - 1

    def set_quant_range(self, x_min, x_max):  # :303-314
        x_min, x_max = self._tensorize_min_max(x_min, x_max)
        self._signed = x_min.min() < 0
        self._delta = torch.max(x_min.abs(), x_max) / self.int_max

    def forward(self, x):
        d, _ = self._params_for(x)
        scale = torch.clamp(d, min=self.eps)
        x_int = torch.round(x / scale) + 0.0
        x_int = torch.clamp(x_int, self.int_min, self.int_max)
        return scale * (x_int - 0.0)


# --------------------------------------------------------------------------------------
# Range estimators (quantization/range_estimators.py)
# --------------------------------------------------------------------------------------


def minmax(x, per_channel: bool):
    """range_estimators.py:73-74 / :85-91 / :110-116: per-channel = over dim 0 rows."""
    if per_channel:
        xf = x.reshape(x.shape[0], -1)
        return xf.min(-1)[0], xf.max(-1)[0]
    return x.min(), x.max()


class OracleEstimatorBase(torch.nn.Module):
    """range_estimators.py:15-53."""

    def __init__(self, per_channel=False, quantizer=None):
        super().__init__()
        self.register_buffer("current_xmin", None)
        self.register_buffer("current_xmax", None)
        self.per_channel = per_channel
        self.quantizer = quantizer

    def reset(self):
        self.current_xmin = None
        self.current_xmax = None


class OracleCurrentMinMax(OracleEstimatorBase):
    """range_estimators.py:56-76 (percentile branch unreachable from the CLI)."""

    def __init__(self, percentile=None, **kw):
        super().__init__(**kw)
        if percentile:
            raise NotImplementedError("percentile path is not on the hot path")

    def forward(self, x):
        self.current_xmin, self.current_xmax = minmax(x, self.per_channel)
        return self.current_xmin, self.current_xmax


class OracleAllMinMax(OracleEstimatorBase):
    """range_estimators.py:79-100."""

    def forward(self, x):
        mn, mx = minmax(x, self.per_channel)
        if self.current_xmin is None:
            self.current_xmin, self.current_xmax = mn, mx
        else:
            self.current_xmin = torch.min(self.current_xmin, mn)
            self.current_xmax = torch.max(self.current_xmax, mx)
        return self.current_xmin, self.current_xmax


class OracleRunningMinMax(OracleEstimatorBase):
    """range_estimators.py:103-125."""

    def __init__(self, momentum=0.9, **kw):
        super().__init__(**kw)
        self.momentum = momentum

    def forward(self, x):
        mn, mx = minmax(x, self.per_channel)
        if self.current_xmin is None:
            self.current_xmin, self.current_xmax = mn, mx
        else:
            self.current_xmin = (1 - self.momentum) * mn + self.momentum * self.current_xmin
            self.current_xmax = (1 - self.momentum) * mx + self.momentum * self.current_xmax
        return self.current_xmin, self.current_xmax


class OracleFPMSE(OracleEstimatorBase):
    """FP_MSE_Estimator, range_estimators.py:285-369."""

    NUM_GRID = 111  # :306 (num_candidates / range_margin ctor args are ignored, :286-293)

    def __init__(self, num_candidates=100, opt_method=None, range_margin=0.5, **kw):
        super().__init__(**kw)
        self.num_candidates = num_candidates
        self.mses = self.search_grid = None

    def _define_search_range(self, x, mbit_list):  # :295-316
        x2 = x.reshape(x.shape[0], -1) if self.per_channel else x.reshape(1, -1)
        mxs = [torch.max(torch.abs(row.min()), torch.abs(row.max())) for row in x2]
        if self.search_grid is None:
            lsp = [torch.linspace(0.1 * mx.item(), 1.2 * mx.item(), self.NUM_GRID) for mx in mxs]
            self.search_grid = torch.stack(lsp).to(x.device).transpose(0, 1)  # [111, C]
            self.mses = torch.stack([torch.zeros_like(self.search_grid) for _ in mbit_list])
        return self.search_grid, self.mses

    def forward(self, x):  # :318-369
        qz = self.quantizer
        mbit_list = [float(qz.mantissa_bits)]
        if qz.mse_include_mantissa_bits:
            mbit_list = [float(m) for m in range(1, qz.n_bits - qz.sign_bits)]
        grid, mses = self._define_search_range(x, mbit_list)
        assert mses.shape[1:] == grid.shape
        sign_bits = int(torch.any(x < 0)) if qz.allow_unsigned else 1
        meandims = list(range(x.dim()))
        if self.per_channel:
            meandims = meandims[1:]
        for m, mbits in enumerate(mbit_list):
            qz.mantissa_bits = torch.Tensor([mbits]).to(x.device)
            for i, maxval in enumerate(grid):
                qz.set_quant_range(sign_bits * -1.0 * maxval, maxval)
                xfp = qz(x)
                mses[m, i, :] += ((x - xfp) ** 2).mean(meandims)
        best_mbits_per_channel = mses.min(1)[0].argmin(0)
        best_idx = torch.mode(best_mbits_per_channel).values.item()
        best_mbits = float(mbit_list[best_idx])
        arg = mses[best_idx].argmin(0)
        maxval = torch.tensor([grid[arg[i], i] for i in range(grid.shape[-1])]).to(x.device)
        qz.mantissa_bits = torch.tensor(best_mbits).to(qz.mantissa_bits.device)
        maxval = maxval.to(qz.maxval.device)
        return sign_bits * -1.0 * maxval, maxval


class OracleLineSearch(OracleEstimatorBase):
    """LineSearchEstimator, 1-D grid search (range_estimators.py:133-282; the only branch that exists for a
    quantiser whose ``symmetric`` is truthy, as FPQuantizer's bound method is) -- SURVEY section 8f2."""

    def __init__(self, num_candidates=1000, range_margin=0.5, expand_range=10.0, **kw):
        super().__init__(**kw)
        self.num_candidates = num_candidates
        self.range_margin = range_margin
        self.expand_range = expand_range
        self.loss_array = None
        self.one_sided_dist = None

    def forward(self, data):  # :258-273
        import copy

        import numpy as np

        if self.loss_array is None:
            if self.one_sided_dist is None:
                self.one_sided_dist = bool((data.min() >= 0).item())
            self.channel_groups = len(data) if self.per_channel else 1  # :204-214
            self.loss_array = np.zeros((self.channel_groups, self.num_candidates + 1))
            self.loss_array[:, 0] = np.inf
            self.max_pos_thr = max(abs(float(data.min())), float(data.max())) + self.range_margin
            self.max_search_range = self.max_pos_thr * self.expand_range
        step = self.max_search_range / self.num_candidates  # :169-174
        for i in range(1, self.num_candidates + 1):  # :236-256
            neg_thr = 0 if self.one_sided_dist else -step * i
            pos_thr = step * i
            q = copy.deepcopy(self.quantizer)  # :199-206
            q.per_channel = False
            if neg_thr or pos_thr:
                q.set_quant_range(neg_thr, pos_thr)
            y = q(data)
            per_row = torch.sum(((data - y) ** 2).view(len(data), -1), dim=1)  # :154-162
            self.loss_array[:, i] += per_row.cpu().numpy() if self.per_channel else float(torch.sum(per_row))
        min_cand = self.loss_array.argmin(axis=1)
        xmin = (np.zeros(self.channel_groups) if self.one_sided_dist else -step * min_cand).astype(np.single)
        xmax = (step * min_cand).astype(np.single)
        self.current_xmax = torch.tensor(xmax).to(device=data.device)
        self.current_xmin = torch.tensor(xmin).to(device=data.device)
        return self.current_xmin, self.current_xmax


# --------------------------------------------------------------------------------------
# QuantizationManager.forward (quantization/quantization_manager.py:114-122)
# --------------------------------------------------------------------------------------


def manager_forward(estimator, quantizer, x, estimate: bool):
    if estimate:
        mn, mx = estimator(x)
        quantizer.set_quant_range(mn, mx)
    return quantizer(x)


# --------------------------------------------------------------------------------------
# Fused epilogue restatements used as the checker for the fused kernels
# (quantized_folded_bn.py:39-55; models/resnet_quantized.py:39-46)
# --------------------------------------------------------------------------------------


def bn_act(x, mean, var, gamma, beta, eps, act: str):
    """F.batch_norm in eval mode + activation, as BNFusedHijacker.forward applies them."""
    y = torch.nn.functional.batch_norm(x, mean, var, gamma, beta, False, 0.0, eps)
    if act == "relu":
        y = torch.relu(y)
    elif act == "relu6":
        y = torch.nn.functional.relu6(y)
    return y
