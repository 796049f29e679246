"""fp8_quantization_b200 -- B200-native FP8 fake-quantisation engine.

Drop-in for the hot path of Qualcomm-AI-research/FP8-quantization: ``FPQuantizer`` and the range
estimators, wired through the reference's ``QuantizationManager`` / hijacker module API.  All
arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI in ``include/fp8fq.h``
(``libfp8fq.so``); there is no CPU or eager-PyTorch fallback.
"""
from ._lib import Fp8fqError, LIB_PATH, lib  # noqa: F401
from . import ops  # noqa: F401
from .quantizers import (AsymmetricUniformQuantizer, FPQuantizer, QuantizerBase,  # noqa: F401
                         QuantizerNotInitializedError, SymmetricUniformQuantizer, get_max_value,
                         quantize_to_fp8_ste_MM)
from .range_estimators import (AllMinMaxEstimator, CurrentMinMaxEstimator, FP_MSE_Estimator,  # noqa: F401
                               LineSearchEstimator, RangeEstimatorBase, RangeEstimators, RunningMinMaxEstimator,
                               estimate_range_line_search)
from .quantization_manager import QMethods, Qstates, QuantizationManager  # noqa: F401
from . import integration  # noqa: F401,E402

__version__ = "0.1.0"
