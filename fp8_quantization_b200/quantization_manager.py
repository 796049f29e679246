"""QuantizationManager -- mirrors quantization/quantization_manager.py:22-136 of the reference.

State machine and API are the reference's; the one functional difference is that, when both the
estimator and the quantiser are this package's device-resident classes, the calibration step
"estimate -> set_quant_range -> prologue" is a single kernel launch with no host round trip.
"""
from __future__ import annotations

from enum import auto

from torch import nn

from .quantizers import (AsymmetricUniformQuantizer, FPQuantizer, QuantizerBase, QuantizerNotInitializedError,
                         SymmetricUniformQuantizer)
from .range_estimators import (BaseEnumOptions, ClassEnumOptions, MethodMap, RangeEstimatorBase, RangeEstimators,
                               _MinMaxEstimator)


class QMethods(ClassEnumOptions):  # quantization_manager.py:22-25
    symmetric_uniform = MethodMap(SymmetricUniformQuantizer)
    asymmetric_uniform = MethodMap(AsymmetricUniformQuantizer)
    fp_quantizer = MethodMap(FPQuantizer)


class Qstates(BaseEnumOptions):  # quantization_manager.py:131-136
    estimate_ranges = auto()
    fix_ranges = auto()
    learn_ranges = auto()
    estimate_ranges_train = auto()


class QuantizationManager(nn.Module):
    """One quantiser + its range estimator + the calibration state machine: the contract of the reference's
    quantization_manager.py:28-128 (constructor arguments :52-61, states :131-136, forward :114-122).  While the state
    is an estimating one, ``forward`` updates the range from the data before quantising; with this package's
    device-resident estimator / quantiser pair that update and the quantiser's prologue are one launch."""

    def __init__(self, qmethod: QuantizerBase = QMethods.symmetric_uniform.cls,
                 init: RangeEstimatorBase = RangeEstimators.current_minmax.cls, per_channel=False, x_min=None,
                 x_max=None, qparams=None, range_estim_params=None):
        super().__init__()
        self.qmethod, self.init, self.per_channel = qmethod, init, per_channel
        self.qparams = dict(qparams or {})
        self.range_estim_params = dict(range_estim_params or {})
        self.quantizer = qmethod(per_channel=per_channel, **self.qparams)
        self.range_estimator = None
        self._enter(Qstates.estimate_ranges)
        if x_min is None or x_max is None:      # ranges come from data: build the estimator (:81-83)
            self.range_estimator = init(per_channel=per_channel, quantizer=self.quantizer, **self.range_estim_params)
        else:                                   # ranges given: install and freeze them (:76-78)
            self.set_quant_range(x_min, x_max)
            self.fix_ranges()

    def _enter(self, state):
        """The manager and its quantiser always carry the same state (:73, 91, 96, 103, 107)."""
        self.state = state
        self.quantizer.state = state

    @property
    def n_bits(self):
        return self.quantizer.n_bits

    # -- state transitions (:89-111) ----------------------------------------------------------------------------------
    def estimate_ranges(self):
        self._enter(Qstates.estimate_ranges)

    def estimate_ranges_train(self):
        self._enter(Qstates.estimate_ranges_train)

    def fix_ranges(self):
        if not self.quantizer.is_initialized:   # a bound method for FPQuantizer, hence always truthy there (:94)
            raise QuantizerNotInitializedError()
        self._enter(Qstates.fix_ranges)

    def learn_ranges(self):
        self.quantizer.make_range_trainable()
        self._enter(Qstates.learn_ranges)

    def reset_ranges(self):
        self.range_estimator.reset()
        self.quantizer.reset()
        self._enter(Qstates.estimate_ranges)

    def set_quant_range(self, x_min, x_max):
        self.quantizer.set_quant_range(x_min, x_max)

    # -- the hot path ---------------------------------------------------------------------------------------------------
    def estimating(self) -> bool:
        return self.state == Qstates.estimate_ranges or (self.state == Qstates.estimate_ranges_train and self.training)

    def device_resident_calibration(self) -> bool:
        """Estimator, set_quant_range and the table build can all stay on the device (no host-visible decision such
        as the sticky unsigned switch of fp8_quantizer.py:224-225)."""
        q, r = self.quantizer, self.range_estimator
        return isinstance(q, FPQuantizer) and isinstance(r, _MinMaxEstimator) and q.set_maxval and not q.allow_unsigned

    def _fusable(self) -> bool:
        return self.device_resident_calibration() and self.range_estimator.fused_supported()

    def forward(self, x):
        if self.estimating():
            if self._fusable():     # statistics + estimator update + set_quant_range + prologue: one launch
                # (only the statistics see a detached x -- range_estimators.py:73-74 -- the quantiser below gets the
                # graph-attached tensor, so estimate_ranges_train keeps the upstream layers' gradients)
                self.range_estimator.fused_estimate_prepare(x, self.quantizer)
            else:                   # :116-118
                self.set_quant_range(*self.range_estimator(x))
        return self.quantizer(x)

    def extra_repr(self):
        return f"state={self.state.name}"
