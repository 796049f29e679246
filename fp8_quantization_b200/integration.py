"""Wiring into an unmodified checkout of the reference (what INTEGRATION.md describes, as code).

The reference has no native binding layer: its injection points are (1) the classes carried in the ``quant_params``
dict (``utils/click_options.py:490-508`` -> ``quantization/base_quantized_classes.py:47-100`` ->
``quantization/quantization_manager.py:72,81-83``) and (2) the module maps of ``quantization/autoquant_utils.py:183-194``.
Nothing here imports the reference; the caller passes its already-imported ``quantization`` package.
"""
import importlib

from . import modules as _m
from . import quantizers as _q
from . import range_estimators as _re
from .quantization_manager import QuantizationManager as _Manager

# reference class name -> our class (same constructor signature, same methods)
CLASS_MAP = {
    "FPQuantizer": _q.FPQuantizer,
    "AsymmetricUniformQuantizer": _q.AsymmetricUniformQuantizer,
    "SymmetricUniformQuantizer": _q.SymmetricUniformQuantizer,
    "CurrentMinMaxEstimator": _re.CurrentMinMaxEstimator,
    "AllMinMaxEstimator": _re.AllMinMaxEstimator,
    "RunningMinMaxEstimator": _re.RunningMinMaxEstimator,
    "FP_MSE_Estimator": _re.FP_MSE_Estimator,
    "LineSearchEstimator": _re.LineSearchEstimator,
}
_CLASS_KEYS = ("method", "act_method", "weight_range_method", "act_range_method")


def patch_quant_params(qparams: dict) -> dict:
    """Route 1 (no edit of the reference): returns a copy of the dict ``quant_params_dict(config)`` built
    (``image_net.py:53``) with every quantiser / range-estimator class replaced by ours of the same name.  Classes
    without a counterpart here (e.g. the percentile or cross-entropy estimators) raise ``KeyError`` -- there is no
    silent fall-back to the reference's eager implementation."""
    out = dict(qparams)
    for key in _CLASS_KEYS:
        cls = out.get(key)
        if cls is None:
            continue
        cls = getattr(cls, "cls", cls)  # enum member (ClassEnumOptions) or plain class
        if cls in CLASS_MAP.values():
            out[key] = cls
            continue
        out[key] = CLASS_MAP[cls.__name__]
    return out


class _Installed:
    """Undo handle of :func:`install_fused_modules`."""

    def __init__(self):
        self._undo = []

    def _set(self, obj, name, value):
        self._undo.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    def _update(self, mapping, new):
        self._undo.append((mapping, None, dict(mapping)))
        mapping.update(new)

    def restore(self):
        for obj, name, old in reversed(self._undo):
            if name is None:
                obj.clear()
                obj.update(old)
            else:
                setattr(obj, name, old)
        self._undo = []


def install_fused_modules(quantization) -> _Installed:
    """Route 2: make the reference's own model builders (``quantize_model`` / ``quantize_sequential``,
    ``models/resnet_quantized.py``, ``models/mobilenet_v2_quantized.py``) emit our fused layers.

    ``quantization`` is the reference's imported top-level package.  Three things are re-pointed, all of them module
    globals the reference only ever uses in ``isinstance`` tests or dict look-ups:

    * ``autoquant_utils.bn_module_map`` / ``non_bn_module_map`` (``autoquant_utils.py:183-194``) -> our hijackers;
    * the name ``QuantizedModule`` in ``autoquant_utils`` (``:298,312,318``) and ``base_quantized_model``
      (``:66-101``) and the name ``QuantizationManager`` in ``autoquant_utils`` (``:143``) -> a tuple of the
      reference's class and ours, so ``QuantizedModel.set_quant_state`` etc. reach both kinds of layer and tied
      activation quantisers are accepted;
    * ``base_quantized_model._set_layer_*`` (``base_quantized_classes.py:17-38``) -> the same rule applied to both
      manager classes, so ``model.fix_ranges()`` / ``estimate_ranges()`` / ``learn_ranges()`` reach our managers.

    Returns a handle whose ``restore()`` undoes everything."""
    from torch import nn

    name = quantization.__name__
    aq = importlib.import_module(name + ".autoquant_utils")
    bqm = importlib.import_module(name + ".base_quantized_model")
    bqc = importlib.import_module(name + ".base_quantized_classes")
    qm = importlib.import_module(name + ".quantization_manager")
    h = _Installed()
    h._update(aq.bn_module_map, {nn.Conv1d: _m.BNQConv1d, nn.Conv2d: _m.BNQConv, nn.Linear: _m.BNQLinear})
    h._update(aq.non_bn_module_map, {nn.Conv1d: _m.QuantConv1d, nn.Conv2d: _m.QuantConv, nn.Linear: _m.QuantLinear,
                                     nn.ConvTranspose1d: _m.QuantConvTranspose1d,
                                     nn.ConvTranspose2d: _m.QuantConvTranspose, nn.LayerNorm: _m.QuantLayerNorm})
    both_modules = (bqc.QuantizedModule, _m.QuantizedModule)
    both_managers = (qm.QuantizationManager, _Manager)
    h._set(aq, "QuantizedModule", both_modules)
    h._set(bqm, "QuantizedModule", both_modules)
    h._set(aq, "QuantizationManager", both_managers)

    def _when_initialized(method):
        def fn(layer):
            if isinstance(layer, both_managers):
                # ``is_initialized`` is a method on FPQuantizer and a property on the uniform quantisers; the
                # reference tests its truthiness without calling it (base_quantized_classes.py:19,25,37)
                if layer.quantizer.is_initialized:
                    getattr(layer, method)()
        return fn

    def _always(method):
        def fn(layer):
            if isinstance(layer, both_managers):
                getattr(layer, method)()
        return fn

    h._set(bqm, "_set_layer_learn_ranges", _when_initialized("learn_ranges"))
    h._set(bqm, "_set_layer_fix_ranges", _when_initialized("fix_ranges"))
    h._set(bqm, "_set_layer_estimate_ranges", _always("estimate_ranges"))
    h._set(bqm, "_set_layer_estimate_ranges_train", _when_initialized("estimate_ranges_train"))
    return h
