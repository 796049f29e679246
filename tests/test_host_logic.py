"""CPU: host-side mirror of the reference's module API -- construction, configuration contract, state
machine, model rewriting census.  No kernel launches."""
import pytest
import torch
from torch import nn

import fp8_quantization_b200 as fq
from fp8_quantization_b200 import modules, workloads
from fp8_quantization_b200.quantization_manager import Qstates, QuantizationManager


def test_quantizer_constructor_contract():
    """fp8_quantizer.py:156-192: attributes and defaults."""
    q = fq.FPQuantizer(8, mantissa_bits=5, maxval=None)
    assert q.n_bits == 8 and q.per_channel is False and q.sign_bits == 1
    assert q.ebits == 2 and q.default_bias == 2
    assert q.maxval.shape == (1,) and float(q.maxval) == 3.9375        # (2 - 2^-5) * 2^(2^2 - 1 - 2)
    assert q.mantissa_bits.shape == (1,) and float(q.mantissa_bits) == 5.0
    assert float(fq.FPQuantizer(8).maxval) == 3.0 and float(fq.FPQuantizer(8).mantissa_bits) == 4.0
    for M, mv in ((2, 57344.0), (3, 240.0), (4, 15.5), (6, 1.984375)):
        assert float(fq.FPQuantizer(8, mantissa_bits=M, maxval=None).maxval) == mv
    assert q.is_initialized() is True and q.symmetric() is False and q.effective_bit_width() is None
    q.mantissa_bits = torch.tensor(3.0)     # what FP_MSE_Estimator assigns (range_estimators.py:364-366)
    assert q._mbits_host == 3.0
    import copy
    q2 = copy.deepcopy(q)                   # LineSearchEstimator deep-copies its quantiser
    assert q2 is not q and float(q2.maxval) == float(q.maxval)
    assert fq.FPQuantizer(8, learn_maxval=True).learning_maxval


def test_enums_and_manager_state_machine():
    assert fq.RangeEstimators.current_minmax.cls is fq.CurrentMinMaxEstimator
    assert fq.RangeEstimators.MSE.cls is fq.FP_MSE_Estimator
    assert fq.QMethods.fp_quantizer.cls is fq.FPQuantizer
    assert fq.RangeEstimators.list_names() == ["current_minmax", "allminmax", "running_minmax", "MSE"]
    m = QuantizationManager(qmethod=fq.FPQuantizer, init=fq.AllMinMaxEstimator, per_channel=True,
                            qparams=dict(n_bits=8, mantissa_bits=5, set_maxval=True))
    assert m.state == Qstates.estimate_ranges and m.quantizer.state == Qstates.estimate_ranges
    assert m.range_estimator.per_channel and m.range_estimator.quantizer is m.quantizer and m.n_bits == 8
    m.fix_ranges()
    assert m.state == Qstates.fix_ranges and not m.estimating()
    m.estimate_ranges_train()
    m.train()
    assert m.estimating()
    m.eval()
    assert not m.estimating()
    m.reset_ranges()
    assert m.state == Qstates.estimate_ranges and m.range_estimator.current_xmin is None


def test_quant_params_contract_and_module_flags():
    qp = workloads.readme_quant_params(5)
    assert set(qp) == {"method", "n_bits", "n_bits_act", "act_method", "per_channel_weights", "quant_setup",
                       "weight_range_method", "weight_range_options", "act_range_method", "act_range_options",
                       "quantize_input", "fp8_kwargs"}  # utils/click_options.py:490-508
    qp.pop("quant_setup")
    lin = modules.QuantLinear(16, 8, **qp)
    assert lin.weight_quantizer.per_channel and not lin.activation_quantizer.per_channel  # acts: per tensor
    assert isinstance(lin.weight_quantizer.range_estimator, fq.CurrentMinMaxEstimator)
    assert isinstance(lin.activation_quantizer.range_estimator, fq.AllMinMaxEstimator)
    assert lin.weight_quantizer.quantizer.set_maxval and float(lin.weight_quantizer.quantizer.mantissa_bits) == 5
    assert lin.get_quantizer_status() == dict(quant_a=False, quant_w=False)
    lin.quantized()
    assert lin.get_quantizer_status() == dict(quant_a=True, quant_w=True) and bool(lin._quant_a) and bool(lin._quant_w)
    sd = lin.state_dict()
    assert "_quant_a" in sd and "_quant_w" in sd
    lin2 = modules.QuantLinear(16, 8, **qp)
    lin2.load_state_dict(sd)
    assert lin2._qa and lin2._qw
    # full precision mode runs on the CPU (no quantiser involved)
    lin.full_precision()
    y = lin(torch.randn(2, 16))
    assert y.shape == (2, 8)


def test_fold_bn_and_sequential_rewrite():
    qp = workloads.readme_quant_params(5)
    qp.pop("quant_setup")
    seq = nn.Sequential(nn.Conv2d(3, 8, 3, bias=False), nn.BatchNorm2d(8), nn.ReLU(), nn.Conv2d(8, 8, 1), nn.ReLU6(),
                        nn.AdaptiveAvgPool2d(1), nn.Conv2d(8, 4, 1), nn.BatchNorm2d(4))
    seq[1].running_mean.normal_()
    q = modules.quantize_model(seq, tie_activation_quantizers=True, **qp)
    assert [type(m).__name__ for m in q] == ["BNQConv", "QuantConv", "QuantizedActivationWrapper", "BNQConv"]
    assert isinstance(q[0].activation_function, nn.ReLU) and isinstance(q[1].activation_function, nn.ReLU6)
    assert q[3].activation_function is None
    assert torch.equal(q[0].running_mean, seq[1].running_mean) and q[0].bias is None
    assert q[2].activation_quantizer is q[1].activation_quantizer  # tied to the feeding layer
    assert torch.equal(q[1].bias, seq[3].bias)


def test_resnet18_and_mobilenetv2_census():
    """ResNet-18 = 21 per-channel weight + 29 per-tensor activation estimators + 1 tied
    call; MobileNetV2 = 53 weight quantisers, 63 activation estimators in use."""
    model = workloads.resnet18_quantized(**workloads.readme_quant_params(5))
    mgrs = [m for m in model.modules() if isinstance(m, QuantizationManager)]
    assert len(mgrs) == 50
    assert sum(m.per_channel for m in mgrs) == 21
    assert model.avgpool.activation_quantizer is model.features[-1][-1].activation_quantizer
    assert sum(isinstance(m, modules.BNQConv) for m in model.modules()) == 20
    assert isinstance(model.fc, modules.QuantLinear)
    model.set_quant_state(True, True)
    assert all(m._qa and m._qw for m in model.modules() if isinstance(m, modules.QuantizedModule))
    model.fix_ranges()
    assert all(m.state == Qstates.fix_ranges for m in mgrs)
    mb = workloads.mobilenetv2_quantized(**workloads.readme_quant_params(4))
    mgrs = [m for m in mb.modules() if isinstance(m, QuantizationManager)]
    assert sum(m.per_channel for m in mgrs) == 53
    n_params = sum(p.numel() for n, p in mb.named_parameters() if n.endswith("weight"))
    assert n_params == 3469760  # MobileNetV2 weight elements
    for setup in ("LSQ", "FP_logits", "fc4"):
        workloads.resnet18_quantized(**{**workloads.readme_quant_params(5), "quant_setup": setup})
    with pytest.raises(ValueError):
        workloads.resnet18_quantized(**{**workloads.readme_quant_params(5), "quant_setup": "nope"})


def test_learnable_range_parameter_registration():
    """fp8_quantizer.py:242-260: learn_maxval / learn_mantissa_bits register Parameters, fix_ranges and plain
    assignments un-register them again (host bookkeeping only; the device of the Parameter is forced to the CPU
    here because this suite runs without a GPU)."""
    import fp8_quantization_b200 as fq

    q = fq.FPQuantizer(8, mantissa_bits=5, maxval=2.0, learn_maxval=True, learn_mantissa_bits=True)
    q._param_device = lambda: torch.device("cpu")
    assert list(q.parameters()) == []
    q.make_range_trainable()
    assert sorted(n for n, _ in q.named_parameters()) == ["_mantissa_bits", "_maxval"]
    assert isinstance(q.maxval, torch.nn.Parameter) and isinstance(q.mantissa_bits, torch.nn.Parameter)
    q.fix_ranges()
    assert list(q.parameters()) == [] and float(q.maxval) == 2.0 and float(q.mantissa_bits) == 5.0
    q.learn_maxval()
    q.learn_mantissa_bits()
    q.maxval = torch.tensor([3.0])
    q.mantissa_bits = 3
    assert list(q.parameters()) == [] and float(q.maxval) == 3.0 and q._mbits_host == 3.0


def _setup_census(model, quantizer_types, fp32_type):
    """{module path: n_bits | "fp32"} for every quantiser / FP32Acts of a quantised model."""
    out = {}
    for name, m in model.named_modules():
        if isinstance(m, quantizer_types):
            out[name] = m.n_bits
        elif isinstance(m, fp32_type):
            out[name] = "fp32"
    return out


def test_quant_setup_tables():
    """workloads._RESNET_SETUPS / _MOBILENETV2_SETUPS: the per-layer deviations of every quant_setup the reference
    defines (models/resnet_quantized.py:94-124, models/mobilenet_v2_quantized.py:45-84)."""
    from torchvision.models import resnet18

    qp = workloads.readme_quant_params(5)
    qp.pop("quant_setup")
    base = _setup_census(workloads.QuantizedResNet(resnet18(), quant_setup="all", **qp), fq.FPQuantizer, modules.FP32Acts)
    assert set(base.values()) == {8}
    m = workloads.QuantizedResNet(resnet18(), quant_setup="fc4", **dict(qp, n_bits=4))
    c = _setup_census(m, fq.FPQuantizer, modules.FP32Acts)
    assert c["features.0.weight_quantizer.quantizer"] == 8 and c["fc.weight_quantizer.quantizer"] == 4
    assert c["fc.activation_quantizer.quantizer"] == 4
    m = workloads.QuantizedResNet(resnet18(), quant_setup="LSQ_paper", **dict(qp, n_bits=4))
    c = _setup_census(m, fq.FPQuantizer, modules.FP32Acts)
    assert isinstance(m.avgpool, nn.AdaptiveAvgPool2d)
    assert c["features.0.activation_quantizer"] == "fp32" and c["features.2.0.activation_quantizer"] == "fp32"
    assert c["features.2.0.features.0.activation_quantizer.quantizer"] == 4
    assert c["fc.activation_quantizer.quantizer"] == 8 and c["fc.weight_quantizer.quantizer"] == 8
    with pytest.raises(ValueError, match="not supported for Resnet"):
        workloads.QuantizedResNet(resnet18(), quant_setup="nope", **qp)
    m = workloads.QuantizedMobileNetV2(workloads.MobileNetV2(), quant_setup="fc4_dw8", **dict(qp, n_bits=4))
    c = _setup_census(m, fq.FPQuantizer, modules.FP32Acts)
    assert c["features.0.0.weight_quantizer.quantizer"] == 8 and c["classifier.1.weight_quantizer.quantizer"] == 4
    assert c["features.2.conv.1.weight_quantizer.quantizer"] == 8      # depthwise 3x3
    assert c["features.2.conv.0.weight_quantizer.quantizer"] == 4      # 1x1 expansion
    with pytest.raises(ValueError, match="not supported for MobilenetV2"):
        workloads.QuantizedMobileNetV2(workloads.MobileNetV2(), quant_setup="nope", **qp)


def test_quant_setups_match_the_real_reference():
    """Every quant_setup of both model families: same quantiser bit widths and FP32Acts placements, module path by
    module path, as the reference's own constructors produce."""
    from oracle.reference_loader import load_reference_models, reference_available

    if not reference_available():
        pytest.skip("reference checkout not present")
    import contextlib
    import io

    from torchvision.models import resnet18

    R = load_reference_models()
    RE = R.range_estimators
    rqp = dict(method=R.FPQuantizer, act_method=R.FPQuantizer, n_bits=4, n_bits_act=None, per_channel_weights=True,
               weight_range_method=RE.CurrentMinMaxEstimator, weight_range_options={},
               act_range_method=RE.AllMinMaxEstimator, act_range_options={}, quantize_input=False,
               fp8_kwargs=dict(maxval=None, mantissa_bits=2, set_maxval=True, learn_maxval=False,
                               learn_mantissa_bits=False, mse_include_mantissa_bits=False, allow_unsigned=False))
    qp = workloads.readme_quant_params(2)
    qp.pop("quant_setup")
    qp["n_bits"] = 4
    with contextlib.redirect_stdout(io.StringIO()):
        for setup in ("all", "LSQ", "LSQ_paper", "FP_logits", "fc4"):
            ref = R.resnet_quantized.QuantizedResNet(resnet18(), quant_setup=setup, **rqp)
            ours = workloads.QuantizedResNet(resnet18(), quant_setup=setup, **qp)
            assert _setup_census(ours, fq.FPQuantizer, modules.FP32Acts) == \
                _setup_census(ref, R.FPQuantizer, R.base_quantized_classes.FP32Acts), setup
        for setup in ("all", "FP_logits", "fc4", "fc4_dw8", "LSQ", "LSQ_paper"):
            ref = R.mobilenet_v2_quantized.QuantizedMobileNetV2(R.mobilenet_v2.MobileNetV2(), quant_setup=setup, **rqp)
            ours = workloads.QuantizedMobileNetV2(workloads.MobileNetV2(), quant_setup=setup, **qp)
            assert _setup_census(ours, fq.FPQuantizer, modules.FP32Acts) == \
                _setup_census(ref, R.FPQuantizer, R.base_quantized_classes.FP32Acts), setup
