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
    """quantization_manager.py:28-128."""

    def __init__(self, qmethod: QuantizerBase = QMethods.symmetric_uniform.cls,  # the reference's defaults (:52-61)
                 init: RangeEstimatorBase = RangeEstimators.current_minmax.cls, per_channel=False, x_min=None,
                 x_max=None, qparams=None, range_estim_params=None):
        super().__init__()
        self.state = Qstates.estimate_ranges
        self.qmethod = qmethod
        self.init = init
        self.per_channel = per_channel
        self.qparams = qparams if qparams else {}
        self.range_estim_params = range_estim_params if range_estim_params else {}
        self.range_estimator = None

        self.quantizer = self.qmethod(per_channel=self.per_channel, **self.qparams)
        self.quantizer.state = self.state

        if x_min is not None and x_max is not None:
            self.set_quant_range(x_min, x_max)
            self.fix_ranges()
        else:
            self.range_estimator = self.init(per_channel=self.per_channel, quantizer=self.quantizer,
                                             **self.range_estim_params)

    @property
    def n_bits(self):
        return self.quantizer.n_bits

    def estimate_ranges(self):
        self.state = Qstates.estimate_ranges
        self.quantizer.state = self.state

    def fix_ranges(self):
        if self.quantizer.is_initialized:
            self.state = Qstates.fix_ranges
            self.quantizer.state = self.state
        else:
            raise QuantizerNotInitializedError()

    def learn_ranges(self):
        self.quantizer.make_range_trainable()
        self.state = Qstates.learn_ranges
        self.quantizer.state = self.state

    def estimate_ranges_train(self):
        self.state = Qstates.estimate_ranges_train
        self.quantizer.state = self.state

    def reset_ranges(self):
        self.range_estimator.reset()
        self.quantizer.reset()
        self.estimate_ranges()

    def estimating(self) -> bool:
        return self.state == Qstates.estimate_ranges or (self.state == Qstates.estimate_ranges_train and self.training)

    def device_resident_calibration(self) -> bool:
        """Estimator, set_quant_range and the table build can all stay on the device (no host-visible decision such
        as the sticky unsigned switch of fp8_quantizer.py:224-225)."""
        q, r = self.quantizer, self.range_estimator
        return isinstance(q, FPQuantizer) and isinstance(r, _MinMaxEstimator) and q.set_maxval and not q.allow_unsigned

    def _fusable(self) -> bool:
        return self.device_resident_calibration() and self.range_estimator.fused_supported()

    def forward(self, x):  # quantization_manager.py:114-122
        if self.estimating():
            if self._fusable():
                x = self.range_estimator.fused_estimate_prepare(x, self.quantizer)
            else:
                cur_xmin, cur_xmax = self.range_estimator(x)
                self.set_quant_range(cur_xmin, cur_xmax)
        return self.quantizer(x)

    def set_quant_range(self, x_min, x_max):
        self.quantizer.set_quant_range(x_min, x_max)

    def extra_repr(self):
        return "state={}".format(self.state.name)
