"""Drop-in range estimators backed by libfp8fq.so.

Mirrors quantization/range_estimators.py of the reference:
  RangeEstimatorBase :15-53, CurrentMinMaxEstimator :56-76, AllMinMaxEstimator :79-100,
  RunningMinMaxEstimator :103-125, FP_MSE_Estimator :285-369, enum RangeEstimators :389-393.

Each min/max estimator is ONE pass over x (the reference's x.min() + x.max() read it twice) with the
state update done inside the kernel.  Under data parallelism (fp8_quantization_b200.dist) the
batch statistic is all-reduced across ranks *before* the update rule is applied, so every rank
ends up with the range a single process would have computed on the concatenated batch.
"""
from __future__ import annotations

from enum import Flag, auto
from collections import namedtuple
from functools import partial

import torch
from torch import nn

from . import dist as fq_dist
from . import ops


class BaseEnumOptions(Flag):  # utils/utils.py:297-304
    def __str__(self):
        return self.name

    @classmethod
    def list_names(cls):
        return list(cls.__members__)  # (Flag iteration skips non-integer values on Python >= 3.11)


class ClassEnumOptions(BaseEnumOptions):  # utils/utils.py:307-313
    @property
    def cls(self):
        return self.value.cls

    def __call__(self, *args, **kwargs):
        return self.value.cls(*args, **kwargs)


MethodMap = partial(namedtuple("MethodMap", ["value", "cls"]), auto())  # utils/utils.py:315


class NoDataPassedError(Exception):  # range_estimators.py:382-386
    def __init__(self):
        super().__init__("Data must be pass through the range estimator to be initialized")


class RangeEstimatorBase(nn.Module):
    """range_estimators.py:15-53."""

    def __init__(self, per_channel=False, quantizer=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.register_buffer("current_xmin", None)
        self.register_buffer("current_xmax", None)
        self.per_channel = per_channel
        self.quantizer = quantizer
        # True for estimators whose input is identical on every rank (weights of a replicated model): their statistics
        # need no collective under data parallelism.  Set by QuantizationHijacker.
        self.replicated_input = False

    def _dp(self) -> bool:
        """Exchange this estimator's batch statistics across ranks?"""
        return fq_dist.active() and not self.replicated_input

    def forward(self, x):
        raise NotImplementedError()

    def reset(self):
        self.current_xmin = None
        self.current_xmax = None

    # The reference keeps per-tensor statistics as 0-dim tensors (x.min() / x.max(), range_estimators.py:73-74); the
    # kernels here write [C]-shaped state with C == 1.  Checkpoints use the reference's shapes, both ways.
    _STATE_BUFFERS = ("current_xmin", "current_xmax")

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        if not self.per_channel:
            for k in self._STATE_BUFFERS:
                t = destination.get(prefix + k)
                if t is not None and t.dim() == 1 and t.numel() == 1:
                    destination[prefix + k] = t.reshape(())

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        for k in self._STATE_BUFFERS:
            t, cur = state_dict.get(prefix + k), getattr(self, k)
            if t is not None and cur is not None and t.dim() == 0 and cur.dim() == 1 and cur.numel() == 1:
                state_dict[prefix + k] = t.reshape(1)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def __repr__(self):
        lines = self.extra_repr().split("\n")
        extra_str = lines[0] if len(lines) == 1 else "\n  " + "\n  ".join(lines) + "\n"
        return self._get_name() + "(" + extra_str + ")"


class _MinMaxEstimator(RangeEstimatorBase):
    """Shared driver of the three min/max estimators: one fused kernel, optional DP all-reduce."""

    EST_MODE = ops.EST_CURRENT
    momentum = 0.9

    def _state(self, x):
        """(cur_min, cur_max, initialized) with state buffers of the right shape on x's device."""
        C = x.shape[0] if self.per_channel else 1
        init = self.current_xmin is not None
        if init and (self.current_xmin.numel() != C or self.current_xmin.device != x.device):
            if self.current_xmin.numel() != C:
                raise ops.Fp8fqError("range estimator was initialised with a different channel count")
            self.current_xmin = self.current_xmin.to(x.device)
            self.current_xmax = self.current_xmax.to(x.device)
        if not init:
            self.current_xmin = torch.empty(C, dtype=torch.float32, device=x.device)
            self.current_xmax = torch.empty(C, dtype=torch.float32, device=x.device)
        return self.current_xmin, self.current_xmax, init

    def forward(self, x):
        x = x.detach()
        x = ops.dense(x)
        if self._dp():
            return self._forward_dp(x)
        cmin, cmax, init = self._state(x)
        ops.minmax(x, self.per_channel, cmin, cmax, self.EST_MODE, init, self.momentum)
        return self.current_xmin, self.current_xmax

    def _forward_dp(self, x):
        C = x.shape[0] if self.per_channel else 1
        packed = torch.empty(2 * C, dtype=torch.float32, device=x.device)
        ops.minmax(x, self.per_channel, packed[:C], packed[C:], ops.EST_CURRENT, False)
        return self.dp_merge(packed)

    def dp_merge(self, packed):
        """All-reduce this rank's batch statistic ``packed = [min (C), max (C)]`` and apply the update rule.
        Device agnostic (the CPU/gloo tests drive it directly): after it every rank holds the range a
        single process would have computed on the concatenated batch (min/max are order independent)."""
        C = packed.numel() // 2
        init = self.current_xmin is not None
        bmin, bmax = packed[:C], packed[C:]
        # a MAX all-reduce does not promise to propagate NaN (torch.min / torch.max do): NaN statistics travel as a flag
        nan = (torch.isnan(bmin) | torch.isnan(bmax)).to(packed.dtype)
        ninf = torch.full_like(bmin, float("-inf"))
        wire = torch.cat([torch.where(nan > 0, ninf, -bmin), torch.where(nan > 0, ninf, bmax), nan])
        fq_dist.all_reduce_max(wire)  # ONE collective for [-min, max, NaN flag]
        bad = wire[2 * C:] > 0
        bmin = torch.where(bad, torch.full_like(bmin, float("nan")), -wire[:C])
        bmax = torch.where(bad, torch.full_like(bmax, float("nan")), wire[C:2 * C])
        if not init or self.EST_MODE == ops.EST_CURRENT:
            self.current_xmin, self.current_xmax = bmin.clone(), bmax.clone()
        elif self.EST_MODE == ops.EST_ALL:
            self.current_xmin = torch.min(self.current_xmin, bmin)
            self.current_xmax = torch.max(self.current_xmax, bmax)
        else:
            m = self.momentum
            self.current_xmin = (1 - m) * bmin + m * self.current_xmin
            self.current_xmax = (1 - m) * bmax + m * self.current_xmax
        return self.current_xmin, self.current_xmax

    def fused_supported(self) -> bool:
        return True

    def _dp_packed(self, C, device):
        """Persistent [-min (C) | max (C) | NaN flag (C)] exchange buffer of this estimator (one per estimator: no allocation and no
        packing kernels on the calibration path)."""
        buf = self.__dict__.get("_dp_buf")
        if buf is None or buf.numel() != 3 * C or buf.device != device:
            buf = self.__dict__["_dp_buf"] = torch.empty(3 * C, dtype=torch.float32, device=device)
        return buf

    def _dp_finish(self, packed, x, quantizer):
        """All-reduce + ONE launch: estimator rule, set_quant_range, table (fp8fq_dp_finish_prepare_f32)."""
        C = packed.numel() // 3
        fq_dist.all_reduce_max(packed)            # the one collective: MAX over [-min, max, NaN flag] of every shard
        cmin, cmax, init = self._state(x)
        mb, nb, sb = quantizer._mbits_host, quantizer.n_bits, quantizer.sign_bits
        maxval = torch.empty(C, dtype=torch.float32, device=x.device)
        table = ops.new_table(C, mb, nb, sb, x.device)
        ops.dp_finish_prepare(packed, cmin, cmax, self.EST_MODE, init, self.momentum, maxval, (mb, nb, sb), table)
        quantizer.adopt_range(maxval, table)

    def fused_estimate_prepare(self, x, quantizer):
        """estimator update + set_quant_range + table in ONE launch; installs the result in ``quantizer``.  Under data
        parallelism: statistics launch (writes [-min | max] straight into the exchange buffer), one MAX all-reduce, one
        finishing launch -- every rank ends with the range of the concatenated batch."""
        x = x.detach()
        x = ops.dense(x)
        if self._dp():
            C = x.shape[0] if self.per_channel else 1
            px = fq_dist.peer_exchange(x.device) if C == 1 else None
            if px is not None:      # ONE launch: statistics, exchange over NVLink peer memory, rule, range, table
                cmin, cmax, init = self._state(x)
                mb, nb, sb = quantizer._mbits_host, quantizer.n_bits, quantizer.sign_bits
                maxval = torch.empty(1, dtype=torch.float32, device=x.device)
                table = ops.new_table(1, mb, nb, sb, x.device)
                ops.estimate_prepare_p2p(x, cmin, cmax, self.EST_MODE, init, self.momentum, maxval, (mb, nb, sb), table,
                                         px.next())
                quantizer.adopt_range(maxval, table)
                return x
            packed = self._dp_packed(C, x.device)
            ops.minmax(x, self.per_channel, packed[:C], packed[C:], ops.EST_DP_STATS, False)
            self._dp_finish(packed, x, quantizer)
            return x
        cmin, cmax, init = self._state(x)
        C = cmin.numel()
        mb, nb, sb = quantizer._mbits_host, quantizer.n_bits, quantizer.sign_bits
        maxval = torch.empty(C, dtype=torch.float32, device=x.device)
        table = ops.new_table(C, mb, nb, sb, x.device)
        ops.estimate_prepare(x, self.per_channel, cmin, cmax, self.EST_MODE, init, self.momentum, maxval, mb, nb, sb,
                             table)
        quantizer.adopt_range(maxval, table)
        return x


    def fused_bn_estimate_prepare(self, x, quantizer, bn_scale, bn_shift, bn_mode, act_code) -> bool:
        """Calibration epilogue of a BN-fused layer: the statistics of ``act(bn(x))`` straight from the convolution
        output ``x`` (one read, nothing written), the estimator update, set_quant_range and the quantiser table in ONE
        launch -- or, under data parallelism, the local statistics, the all-reduce and ``set_quant_range``.  Returns
        False (state untouched) when the fused kernel does not cover the shape."""
        assert not self.per_channel
        x = ops.dense(x.detach())
        mb, nb, sb = quantizer._mbits_host, quantizer.n_bits, quantizer.sign_bits
        if self._dp():
            px = fq_dist.peer_exchange(x.device)
            if px is not None:      # ONE launch, the exchange inside it
                was_init = self.current_xmin is not None
                cmin, cmax, init = self._state(x)
                maxval = torch.empty(1, dtype=torch.float32, device=x.device)
                table = ops.new_table(1, mb, nb, sb, x.device)
                if not ops.bn_act_estimate_prepare_p2p(x, bn_scale, bn_shift, act_code, bn_mode, cmin, cmax, self.EST_MODE,
                                                       init, self.momentum, maxval, (mb, nb, sb), table, px.next()):
                    px.unused()     # shape not covered (same on every rank): the caller composes the unfused ops
                    if not was_init:
                        self.current_xmin = self.current_xmax = None
                    return False
                quantizer.adopt_range(maxval, table)
                return True
            packed = self._dp_packed(1, x.device)
            if not ops.bn_act_estimate_prepare(x, bn_scale, bn_shift, act_code, bn_mode, packed[:1], packed[1:],
                                               ops.EST_DP_STATS, False, self.momentum):
                return False
            self._dp_finish(packed, x, quantizer)
            return True
        was_init = self.current_xmin is not None
        cmin, cmax, init = self._state(x)
        maxval = torch.empty(1, dtype=torch.float32, device=x.device)
        table = ops.new_table(1, mb, nb, sb, x.device)
        if not ops.bn_act_estimate_prepare(x, bn_scale, bn_shift, act_code, bn_mode, cmin, cmax, self.EST_MODE, init,
                                           self.momentum, maxval, (mb, nb, sb), table):
            if not was_init:
                self.current_xmin = self.current_xmax = None
            return False
        quantizer.adopt_range(maxval, table)
        return True


class CurrentMinMaxEstimator(_MinMaxEstimator):
    """range_estimators.py:56-76 (the percentile branch is unreachable from the reference's CLI:
    hijacker.py:57 compares a class with an enum member)."""

    EST_MODE = ops.EST_CURRENT

    def __init__(self, percentile=None, *args, **kwargs):
        self.percentile = percentile
        super().__init__(*args, **kwargs)
        if percentile:
            raise NotImplementedError("percentile ranges are a host/numpy path in the reference, out of scope")


class AllMinMaxEstimator(_MinMaxEstimator):
    """range_estimators.py:79-100."""

    EST_MODE = ops.EST_ALL


class RunningMinMaxEstimator(_MinMaxEstimator):
    """range_estimators.py:103-125."""

    EST_MODE = ops.EST_RUNNING

    def __init__(self, momentum=0.9, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.momentum = momentum


class OptMethod(BaseEnumOptions):  # range_estimators.py:128-130
    grid = auto()
    golden_section = auto()


class FP_MSE_Estimator(RangeEstimatorBase):
    """range_estimators.py:285-369 with the 111..666-iteration Python loop replaced by one kernel
    launch per mantissa width that reads x once and sweeps all 111 candidate ranges in registers."""

    NUM_GRID = 111  # range_estimators.py:306; num_candidates / range_margin are ignored there too

    def __init__(self, num_candidates=100, opt_method=OptMethod.grid, range_margin=0.5, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert opt_method == OptMethod.grid
        self.num_candidates = num_candidates
        self.mses = self.search_grid = None

    def _define_search_range(self, x, mbit_list):  # :295-316
        if self.search_grid is None:
            assert self.mses is None
            C = x.shape[0] if self.per_channel else 1
            packed = torch.empty(2 * C, dtype=torch.float32, device=x.device)
            ops.minmax(x, self.per_channel, packed[:C], packed[C:], ops.EST_CURRENT, False)
            absmax = torch.max(packed[:C].abs(), packed[C:].abs())
            if self._dp():
                fq_dist.all_reduce_max(absmax)
            # The reference builds the grid on the host from .item() values with torch.linspace; do the
            # same (one D2H copy of C floats, once per estimator) so the grid is bit-identical.
            vals = absmax.cpu().tolist()
            lsp = [torch.linspace(0.1 * v, 1.2 * v, self.NUM_GRID) for v in vals]
            self.search_grid = torch.stack(lsp).to(x.device).transpose(0, 1).contiguous()  # [111, C]
            self.mses = torch.zeros(len(mbit_list), self.NUM_GRID, C, dtype=torch.float32, device=x.device)
        return self.search_grid, self.mses

    def forward(self, x):  # :318-369
        qz = self.quantizer
        x = x.detach()
        x = ops.dense(x)
        mbit_list = [float(qz._mbits_host)]
        if qz.mse_include_mantissa_bits:
            mbit_list = [float(m) for m in range(1, qz.n_bits - qz.sign_bits)]
        grid, mses = self._define_search_range(x, mbit_list)
        assert mses.shape[1:] == grid.shape, f"{mses.shape}, {grid.shape}"
        sign_bits = 1
        if qz.allow_unsigned:
            # one decision for the GLOBAL batch: ranks whose shards disagree on "any negative value" would otherwise
            # average MSE tables of different formats
            neg = torch.any(x < 0).to(torch.float32).reshape(1)
            if self._dp():
                fq_dist.all_reduce_max(neg)
            sign_bits = int(neg.item())
        if qz.allow_unsigned and sign_bits == 0:
            qz.sign_bits = 0  # what set_quant_range(-0.0 * maxval, maxval) does in the reference loop (:341-342)
        if self._dp():
            inc = torch.zeros_like(mses)
            ops.mse_grid(x, self.per_channel, grid, mbit_list, qz.n_bits, qz.sign_bits, inc)
            fq_dist.all_reduce_mean(inc)
            mses += inc
        else:
            ops.mse_grid(x, self.per_channel, grid, mbit_list, qz.n_bits, qz.sign_bits, mses)
        best_mbits_per_channel = mses.min(1)[0].argmin(0)
        best_idx = int(torch.mode(best_mbits_per_channel).values.item())
        best_mbits = float(mbit_list[best_idx])
        arg = mses[best_idx].argmin(0)  # [C]
        maxval = grid.gather(0, arg.view(1, -1)).reshape(-1)
        qz.mantissa_bits = best_mbits
        return sign_bits * -1.0 * maxval, maxval


class LineSearchEstimator(RangeEstimatorBase):
    """range_estimators.py:133-282 -- the 1-D grid search (the only branch that exists for a quantiser whose
    ``symmetric`` is truthy, as FPQuantizer's is): ``num_candidates`` clipping thresholds ``step * i``, loss = sum of
    squared quantisation error (per row when per_channel), accumulated over calls, argmin.  SURVEY section 8f2: the
    reference deep-copies the quantiser and launches 13 kernels + a D2H copy per candidate (1000 x per call in
    compute_quant_error.py); here all candidates are swept by ONE launch of the MSE-grid kernel that reads the data
    once."""

    def __init__(self, num_candidates=1000, opt_method=OptMethod.grid, range_margin=0.5, expand_range=10.0, *args,
                 **kwargs):
        super().__init__(*args, **kwargs)
        assert opt_method in OptMethod
        if opt_method != OptMethod.grid:
            raise NotImplementedError("only the grid search exists in the reference (golden-section methods are "
                                      "referenced but not defined there)")
        self.opt_method = opt_method
        self.num_candidates = num_candidates
        self.expand_range = expand_range
        self.range_margin = range_margin
        self.loss_array = None
        self.max_pos_thr = self.max_neg_thr = self.max_search_range = None
        self.one_sided_dist = None
        if self.quantizer is None:
            raise NotImplementedError("A Quantizer must be given as an argument to the MSE Range" "Estimator")
        self.max_int_skew = (2**self.quantizer.n_bits) // 4

    @property
    def step_size(self):
        if self.one_sided_dist is None:
            raise NoDataPassedError()
        return self.max_search_range / self.num_candidates

    @property
    def optimization_method(self):  # :180-197 -- only the 1-D grid search is defined in the reference
        if self.one_sided_dist is None:
            raise NoDataPassedError()
        if not (self.one_sided_dist or self.quantizer.symmetric):
            raise NotImplementedError("2-D grid search (asymmetric quantiser, two-sided data) does not exist in the "
                                      "reference either")
        return self.forward

    def quantize(self, x_float, x_min=None, x_max=None):  # :199-206
        import copy

        temp_q = copy.deepcopy(self.quantizer)
        temp_q.per_channel = False
        if x_min or x_max:
            temp_q.set_quant_range(x_min, x_max)
        return temp_q(x_float)

    def loss_fx(self, data, neg_thr, pos_thr, per_channel_loss=False):
        """:161-169 -- the loss of ONE candidate (sum of squared error, per row or in total), as a numpy value; the
        search itself evaluates all candidates in one launch (``forward``)."""
        y = self.quantize(data, x_min=neg_thr, x_max=pos_thr)
        temp_sum = torch.sum(((data - y) ** 2).reshape(len(data), -1), dim=1)
        return (temp_sum if per_channel_loss else torch.sum(temp_sum)).cpu().numpy()

    def _define_search_range(self, data, dmin, dmax):  # :203-234, 1-D branch
        import numpy as np

        self.channel_groups = len(data) if self.per_channel else 1
        self.loss_array = np.zeros((self.channel_groups, self.num_candidates + 1))
        self.loss_array[:, 0] = np.inf
        self.max_pos_thr = max(abs(dmin), dmax) + self.range_margin
        self.max_neg_thr = -self.max_pos_thr * self.expand_range
        self.max_search_range = self.max_pos_thr * self.expand_range

    def forward(self, data):
        import numpy as np

        from .quantizers import FPQuantizer

        qz = self.quantizer
        data = data.detach()
        data = data if data.is_contiguous() else data.contiguous()
        if self.loss_array is None:
            mm = torch.empty(2, dtype=torch.float32, device=data.device)
            ops.minmax(data, False, mm[:1], mm[1:], ops.EST_CURRENT, False)
            if self._dp():            # global-batch extremes, so that every rank searches the same candidates
                mm[:1].neg_()
                fq_dist.all_reduce_max(mm)
                mm[:1].neg_()
            dmin, dmax = mm.tolist()  # the reference reads float(data.min()) / float(data.max()) here too
            if self.one_sided_dist is None:
                self.one_sided_dist = bool(dmin >= 0)
            if not isinstance(qz, FPQuantizer) and not (self.one_sided_dist or qz.symmetric):
                # range_estimators.py:190: the 2-D search of asymmetric quantisers is referenced there but not defined
                raise NotImplementedError("2-D grid search (asymmetric quantiser, two-sided data) does not exist in the "
                                          "reference either")
            self._define_search_range(data, dmin, dmax)
        if not isinstance(qz, FPQuantizer):
            return self._search_per_candidate(data)
        step = self.step_size
        C = self.channel_groups
        if qz.set_maxval:
            thr = torch.tensor([step * i for i in range(1, self.num_candidates + 1)], dtype=torch.float32)
        else:  # set_quant_range ignores the candidates (fp8_quantizer.py:227): every candidate is the current range
            thr = qz.maxval.detach().reshape(-1)[:1].cpu().expand(self.num_candidates).clone()
        grid = thr.to(data.device).view(-1, 1).expand(-1, C).contiguous()
        # the reference evaluates a deep copy whose set_quant_range(0, thr) may switch it to unsigned (:199-206)
        sign_bits = 0 if (qz.allow_unsigned and self.one_sided_dist) else qz.sign_bits
        mses = torch.zeros(1, self.num_candidates, C, dtype=torch.float32, device=data.device)
        ops.mse_grid(data, self.per_channel, grid, [qz._mbits_host], qz.n_bits, sign_bits, mses)
        inner = data.numel() // C
        loss = (mses[0].double() * inner).t().cpu().numpy()  # [C, G] sums of squared error
        self.loss_array[:, 1:] += loss
        min_cand = self.loss_array.argmin(axis=1)
        xmin = (np.zeros(C) if self.one_sided_dist else -step * min_cand).astype(np.single)
        xmax = (step * min_cand).astype(np.single)
        self.current_xmax = torch.tensor(xmax).to(device=data.device)
        self.current_xmin = torch.tensor(xmin).to(device=data.device)
        return self.current_xmin, self.current_xmax

    def _search_per_candidate(self, data):
        """range_estimators.py:236-256 as written, for quantisers the MSE-grid kernel does not cover (the INT uniform
        ones -- the E = 0 row of compute_quant_error.py:24-27): per candidate a deep copy of the quantiser gets the range
        (two launches of this library: range -> table, quantise) and the squared error is summed on the device; the
        losses come back to the host in ONE copy at the end instead of one per candidate."""
        import copy

        import numpy as np

        step = self.step_size
        C = self.channel_groups
        rows = data.reshape(len(data), -1) if self.per_channel else data.reshape(1, -1)
        losses = torch.empty(self.num_candidates, C, dtype=torch.float64, device=data.device)
        for i in range(1, self.num_candidates + 1):
            temp_q = copy.deepcopy(self.quantizer)
            temp_q.per_channel = False                      # :198-199
            temp_q.set_quant_range(0.0 if self.one_sided_dist else -step * i, step * i)
            y = temp_q(data)
            losses[i - 1] = ((rows - y.reshape(rows.shape)) ** 2).sum(dim=1, dtype=torch.float64)
        self.loss_array[:, 1:] += losses.t().cpu().numpy()
        min_cand = self.loss_array.argmin(axis=1)
        xmin = (np.zeros(C) if self.one_sided_dist else -step * min_cand).astype(np.single)
        xmax = (step * min_cand).astype(np.single)
        self.current_xmax = torch.tensor(xmax).to(device=data.device)
        self.current_xmin = torch.tensor(xmin).to(device=data.device)
        return self.current_xmin, self.current_xmax

    def reset(self):
        super().reset()
        self.loss_array = None

    def extra_repr(self):
        return "opt_method={} ,num_candidates={}".format(self.opt_method.name, self.num_candidates)


def estimate_range_line_search(W, quant, num_candidates=None):
    """range_estimators.py:372-379."""
    est = LineSearchEstimator(quantizer=quant) if num_candidates is None else \
        LineSearchEstimator(quantizer=quant, num_candidates=num_candidates)
    return est.forward(W)


class RangeEstimators(ClassEnumOptions):  # range_estimators.py:389-393
    current_minmax = MethodMap(CurrentMinMaxEstimator)
    allminmax = MethodMap(AllMinMaxEstimator)
    running_minmax = MethodMap(RunningMinMaxEstimator)
    MSE = MethodMap(FP_MSE_Estimator)
