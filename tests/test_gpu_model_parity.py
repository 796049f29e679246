"""GPU: the product's fused forward against the ORACLE-COMPOSED network on the same device, bit for bit.

The checker (oracle/fp8_oracle_models.py) is the reference's quantised ResNet-18 / MobileNetV2 restated with plain
``F.conv2d`` / ``F.batch_norm`` / ``relu`` and the oracle's 13-op quantiser and range estimators -- none of the product's
modules -- and is pinned on the CPU against the real reference's golden ranges and logits
(tests/test_oracle_golden.py::test_oracle_composed_*).  Run with CUDA tensors it is "the reference as shipped, run with
--cuda on this B200".  Requirements, with no tolerance:
  * every quantiser range (50 for ResNet-18 M=5, 123 for MobileNetV2 M=4) bit-equal,
  * calibration-pass logits and fixed-range logits ``torch.equal``,
so a wiring error (wrong tied quantiser, wrong residual operand, a missed range update) cannot hide.
Layout: NCHW, the reference's own (autoquant_utils.py:34-44 forces contiguous operands); same cuDNN settings.
Config 4: FP_MSE_Estimator on the 29 ResNet-18 activation sites for M in {2..7} and with the internal mantissa sweep:
mantissa votes equal at every site, selected ranges equal except on fp32 summation-order ties of the oracle's own
MSE table (counted, stated below).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda:0") if torch.cuda.is_available() else None


def _product_ranges(model):
    from fp8_quantization_b200 import FPQuantizer

    return {n: m.maxval.detach().reshape(-1).clone() for n, m in model.named_modules() if isinstance(m, FPQuantizer)}


def _assert_same_ranges(ours, oracle, expect):
    assert set(ours) == set(oracle) and len(ours) == expect, (len(ours), len(oracle))
    bad = [n for n in ours if not torch.equal(ours[n].view(torch.int32), oracle[n].to(ours[n].device).view(torch.int32))]
    assert not bad, f"{len(bad)} of {expect} ranges differ from the oracle-composed network: {bad[:4]}"


@pytest.mark.parametrize("batch", [16])
def test_config2_resnet18_m5_fused_equals_oracle_composed(batch):
    from torchvision.models import resnet18

    from fp8_quantization_b200 import ops, workloads
    from oracle import fp8_oracle_models as OM

    torch.manual_seed(10)
    net = resnet18().to(DEV).eval()
    model = workloads.QuantizedResNet(net, **workloads.readme_quant_params(5)).to(DEV).eval()
    x = torch.randn(batch, 3, 224, 224, device=DEV, generator=torch.Generator(device=DEV).manual_seed(10))
    model.set_quant_state(True, True)
    with torch.no_grad():
        cal = model(x)                       # state estimate_ranges: fused statistics kernels + fused epilogues
        model.fix_ranges()
        n0 = ops.launch_count()
        logits = model(x)                    # the 23-launch fused validate forward
        assert ops.launch_count() - n0 == 23
    S = OM.OracleSites(5)
    cal_o = OM.resnet_forward(S, net, x)
    S.fix_ranges()
    logits_o = OM.resnet_forward(S, net, x)
    _assert_same_ranges(_product_ranges(model), S.maxvals(), 50)
    assert torch.equal(cal, cal_o)
    assert torch.equal(logits, logits_o)
    assert torch.isfinite(logits).all() and float(logits.std()) > 0


@pytest.mark.parametrize("batch", [8])
def test_config3_mobilenetv2_m4_fused_equals_oracle_composed(batch):
    from fp8_quantization_b200 import workloads
    from oracle import fp8_oracle_models as OM

    torch.manual_seed(10)
    net = workloads.MobileNetV2().to(DEV).eval()
    model = workloads.QuantizedMobileNetV2(net, **workloads.readme_quant_params(4)).to(DEV).eval()
    x = torch.randn(batch, 3, 224, 224, device=DEV, generator=torch.Generator(device=DEV).manual_seed(10))
    model.set_quant_state(True, True)
    with torch.no_grad():
        cal = model(x)
        model.fix_ranges()
        logits = model(x)
    S = OM.OracleSites(4)
    cal_o = OM.mobilenetv2_forward(S, net, x)
    S.fix_ranges()
    logits_o = OM.mobilenetv2_forward(S, net, x)
    _assert_same_ranges(_product_ranges(model), S.maxvals(), 123)   # 116 used + 7 never-called defaults
    assert torch.equal(cal, cal_o)
    assert torch.equal(logits, logits_o)
    assert torch.isfinite(logits).all() and float(logits.std()) > 0


def _capture_estimator_inputs(model, cls):
    """forward-pre hooks on every activation range estimator: (name, input clone) in call order."""
    got, handles = [], []
    for name, m in model.named_modules():
        if isinstance(m, cls) and not m.per_channel:
            handles.append(m.register_forward_pre_hook(lambda mod, a, name=name: got.append((name, mod, a[0].detach().clone()))))
    return got, handles


@pytest.mark.parametrize("M", [2, 3, 4, 5, 6, 7])
def test_config4_mse_estimator_on_resnet18_activation_sites(M):
    """BASELINE config 4 per mantissa width: ResNet-18 calibrated with FP_MSE_Estimator (111-candidate grid,
    range_estimators.py:318-369) on its 29 activation sites; at every site the oracle's estimator is run on the very
    tensor the product's estimator saw.  Grid bit-equal; selected range equal, or -- counted -- a tie of the oracle's
    own fp32 MSE row within 2e-5 relative (its reduction order is not the kernel's); >= 27 of 29 sites exactly equal."""
    import fp8_quantization_b200 as fq
    from fp8_quantization_b200 import workloads
    from oracle import fp8_oracle as O

    torch.manual_seed(10)
    qp = workloads.readme_quant_params(M, act_range_method=fq.FP_MSE_Estimator)
    model = workloads.resnet18_quantized(**qp).to(DEV).eval()
    x = torch.randn(8, 3, 224, 224, device=DEV, generator=torch.Generator(device=DEV).manual_seed(10))
    got, handles = _capture_estimator_inputs(model, fq.FP_MSE_Estimator)
    workloads.pass_data_for_range_estimation([x], model, True, True, 1)
    for h in handles:
        h.remove()
    assert len(got) == 29
    exact = ties = 0
    for name, est, xin in got:
        oq = O.OracleFPQuantizer(8, mantissa_bits=M, set_maxval=True, mse_include_mantissa_bits=False)
        oest = O.OracleFPMSE(quantizer=oq)
        _, omx = oest(xin)
        assert torch.equal(est.search_grid, oest.search_grid), name
        ours = est.quantizer.maxval.reshape(-1)
        torch.testing.assert_close(est.mses, oest.mses, rtol=3e-4, atol=1e-12)
        if torch.equal(ours, omx.reshape(-1).to(ours.device)):
            exact += 1
            continue
        row = oest.mses[0, :, 0]
        gi = int((oest.search_grid[:, 0] - ours).abs().argmin())
        assert float(row[gi]) <= float(row.min()) * (1 + 2e-5), (name, float(row[gi]), float(row.min()))
        ties += 1
    assert exact + ties == 29 and exact >= 27, (exact, ties)


def test_config4_mantissa_vote_on_resnet18_activation_sites():
    """BASELINE config 4 with the internal mantissa sweep (mse_include_mantissa_bits=True: M = 1..6, 666 candidates per
    site): the mantissa width voted at each of the 29 sites equals the oracle's, 29 of 29."""
    import fp8_quantization_b200 as fq
    from fp8_quantization_b200 import workloads
    from oracle import fp8_oracle as O

    torch.manual_seed(10)
    qp = workloads.readme_quant_params(5, act_range_method=fq.FP_MSE_Estimator, mse_include_mantissa_bits=True)
    model = workloads.resnet18_quantized(**qp).to(DEV).eval()
    x = torch.randn(8, 3, 224, 224, device=DEV, generator=torch.Generator(device=DEV).manual_seed(10))
    got, handles = _capture_estimator_inputs(model, fq.FP_MSE_Estimator)
    workloads.pass_data_for_range_estimation([x], model, True, True, 1)
    for h in handles:
        h.remove()
    assert len(got) == 29
    votes = range_equal = 0
    for name, est, xin in got:
        oq = O.OracleFPQuantizer(8, mantissa_bits=5, set_maxval=True, mse_include_mantissa_bits=True)
        oest = O.OracleFPMSE(quantizer=oq)
        _, omx = oest(xin)
        assert float(est.quantizer._mbits_host) == float(oq.mantissa_bits), name
        votes += 1
        range_equal += int(torch.equal(est.quantizer.maxval.reshape(-1), omx.reshape(-1).to(DEV)))
    assert votes == 29 and range_equal >= 27, (votes, range_equal)
