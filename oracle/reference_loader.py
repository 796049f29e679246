"""Import the *real* reference (read-only checkout at /root/reference).

TEST INFRASTRUCTURE ONLY.  The reference checkout exists in the build container
but NOT on the GPU box, so nothing that runs under ``-m gpu``, ``smoke()`` or
``bench.py`` may call :func:`load_reference`; it is used by
``tests/golden/make_golden.py`` (golden-vector generation) and by the CPU tests
that are skipped when the checkout is absent.

The reference imports ``timm`` (quantization/hijacker.py:7-8) which is not in
this image; six empty ``nn.Module`` subclasses are enough to satisfy the import
(they are only ever used in an ``isinstance`` tuple, hijacker.py:15-29).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FP8FQ_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "quantization", "quantizers", "fp8_quantizer.py"))


def _install_timm_stub():
    if "timm" in sys.modules:
        return
    from torch import nn

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    timm = mod("timm")
    models = mod("timm.models")
    layers = mod("timm.models.layers")
    acts = mod("timm.models.layers.activations")
    acts_me = mod("timm.models.layers.activations_me")
    timm.models = models
    models.layers = layers
    layers.activations = acts
    layers.activations_me = acts_me
    for name in ("Swish", "HardSwish", "HardSigmoid"):
        setattr(acts, name, type(name, (nn.Module,), {}))
    for name in ("SwishMe", "HardSwishMe", "HardSigmoidMe"):
        setattr(acts_me, name, type(name, (nn.Module,), {}))


_loaded = None


def load_reference():
    """Returns a namespace with the reference's hot-path classes/functions."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the checkout is read-only
    _install_timm_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # The reference has top-level packages called ``quantization``, ``utils`` and ``models``.
    import quantization.quantizers.fp8_quantizer as fp8q
    import quantization.range_estimators as re_
    import quantization.quantization_manager as qm
    import quantization.autoquant_utils as aq
    import quantization.base_quantized_classes as bqc

    ns = types.SimpleNamespace(
        fp8_quantizer=fp8q,
        range_estimators=re_,
        quantization_manager=qm,
        autoquant_utils=aq,
        base_quantized_classes=bqc,
        FPQuantizer=fp8q.FPQuantizer,
        quantize_to_fp8_ste_MM=fp8q.quantize_to_fp8_ste_MM,
    )
    _loaded = ns
    return ns


def load_reference_models():
    ns = load_reference()
    import models.resnet_quantized as rq
    import models.mobilenet_v2_quantized as mq
    import models.mobilenet_v2 as mv2

    ns.resnet_quantized = rq
    ns.mobilenet_v2_quantized = mq
    ns.mobilenet_v2 = mv2
    return ns
