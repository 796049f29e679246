"""GPU: range estimators (K2a min/max family, K2b MSE grid) against the golden vectors of the real
reference and against the oracle."""
import numpy as np
import pytest
import torch

from conftest import bits, load_golden
from oracle import fp8_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_minmax_estimators_match_reference_golden_bit_exact():
    import fp8_quantization_b200 as fq

    g = load_golden("estimators.npz")
    for name, cls, kw in (("current", fq.CurrentMinMaxEstimator, {}), ("all", fq.AllMinMaxEstimator, {}),
                          ("running", fq.RunningMinMaxEstimator, {"momentum": 0.9})):
        for pc in (False, True):
            key = f"{name}_{'pc' if pc else 'pt'}"
            est = cls(per_channel=pc, **kw)
            for i, x in enumerate(g[key + "_x"]):
                mn, mx = est(torch.from_numpy(x).to(DEV))
                # min/max are order independent and the EMA is two fp32 multiplies and an add: bit-exact
                assert np.array_equal(mn.cpu().numpy().reshape(-1), g[key + "_min"][i], equal_nan=True), (key, i)
                assert np.array_equal(mx.cpu().numpy().reshape(-1), g[key + "_max"][i], equal_nan=True), (key, i)
            est.reset()
            assert est.current_xmin is None


@pytest.mark.parametrize("shape,pc", [((1 << 22) + 3, False), ((7,), False), ((1,), False), ((64, 147), True),
                                      ((512, 4608), True), ((960, 9), True), ((3, 1 << 20), True), ((2, 3, 5, 7), False)])
def test_minmax_sizes_and_nan(shape, pc):
    from fp8_quantization_b200 import ops

    torch.manual_seed(1)
    x = torch.randn(shape, device=DEV)
    C = x.shape[0] if pc else 1
    mn, mx = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    for _ in range(2):  # second call re-uses the (self-resetting) workspace
        ops.minmax(x, pc, mn, mx, ops.EST_CURRENT, False)
        omn, omx = O.minmax(x, pc)
        assert torch.equal(mn, omn.reshape(-1)) and torch.equal(mx, omx.reshape(-1))
    x.view(-1)[x.numel() // 2] = float("nan")
    ops.minmax(x, pc, mn, mx, ops.EST_CURRENT, False)
    omn, omx = O.minmax(x, pc)
    assert np.array_equal(mn.cpu().numpy(), omn.reshape(-1).cpu().numpy(), equal_nan=True)
    assert np.array_equal(mx.cpu().numpy(), omx.reshape(-1).cpu().numpy(), equal_nan=True)
    assert torch.isnan(mn).any()  # NaN poisons like torch.min/max


def test_manager_calibration_is_fused_and_matches_oracle():
    """QuantizationManager.forward in estimate state (quantization_manager.py:114-122): the fused
    estimate+set_range+prologue launch gives the same maxval and output as the step-by-step oracle."""
    import fp8_quantization_b200 as fq
    from fp8_quantization_b200 import ops

    torch.manual_seed(2)
    for pc, est, oest in ((False, fq.AllMinMaxEstimator, O.OracleAllMinMax),
                          (True, fq.CurrentMinMaxEstimator, O.OracleCurrentMinMax),
                          (False, fq.RunningMinMaxEstimator, O.OracleRunningMinMax)):
        mgr = fq.QuantizationManager(qmethod=fq.FPQuantizer, init=est, per_channel=pc,
                                     qparams=dict(n_bits=8, mantissa_bits=5, set_maxval=True))
        oq = O.OracleFPQuantizer(8, per_channel=pc, mantissa_bits=5, set_maxval=True)
        oe = oest(per_channel=pc, quantizer=oq)
        for step in range(3):
            x = torch.randn(32, 200, device=DEV) * (step + 1)
            n0 = ops.launch_count()
            y = mgr(x)
            assert ops.launch_count() - n0 == 2  # estimate+prepare, quantise
            yo = O.manager_forward(oe, oq, x, True)
            assert torch.equal(mgr.quantizer.maxval.reshape(-1), oq.maxval.reshape(-1))
            assert torch.equal(bits(y), bits(yo))
        mgr.fix_ranges()
        x = torch.randn(32, 200, device=DEV) * 10
        n0 = ops.launch_count()
        y = mgr(x)
        assert ops.launch_count() - n0 == 1  # fixed ranges: ONE launch per quantisation site
        assert torch.equal(bits(y), bits(oq(x)))


def test_set_quant_range_semantics():
    """fp8_quantizer.py:222-240: floats / 0-dim / [1] / [C] inputs, set_maxval=False ignores, sticky unsigned."""
    import fp8_quantization_b200 as fq

    q = fq.FPQuantizer(8, mantissa_bits=4, maxval=3.0, set_maxval=False)
    q.set_quant_range(-10.0, 10.0)
    assert float(q.maxval) == 3.0
    q = fq.FPQuantizer(8, mantissa_bits=4, set_maxval=True)
    q.set_quant_range(-7.5, 2.0)
    assert float(q.maxval) == 7.5 and q.maxval.shape == (1,)
    q.set_quant_range(torch.tensor(-1.0, device=DEV), torch.tensor(4.0, device=DEV))
    assert float(q.maxval) == 4.0 and q.maxval.shape == (1,)
    q.set_quant_range(torch.tensor([-1.0, -9.0], device=DEV), torch.tensor([4.0, 2.0], device=DEV))
    assert q.maxval.tolist() == [4.0, 9.0]
    qu = fq.FPQuantizer(8, mantissa_bits=4, set_maxval=True, allow_unsigned=True)
    qu.set_quant_range(torch.tensor([0.0], device=DEV), torch.tensor([4.0], device=DEV))
    assert qu.sign_bits == 0
    qu.set_quant_range(torch.tensor([-1.0], device=DEV), torch.tensor([4.0], device=DEV))
    assert qu.sign_bits == 0  # sticky, never reset
    oq = O.OracleFPQuantizer(8, mantissa_bits=4, set_maxval=True, allow_unsigned=True)
    oq.set_quant_range(torch.tensor([0.0]), torch.tensor([4.0]))
    x = torch.rand(4096, device=DEV) * 5
    assert torch.equal(bits(qu(x)), bits(oq(x)))


def test_mse_estimator_matches_reference_golden():
    """FP_MSE_Estimator (range_estimators.py:285-369): same grid, same argmins / plurality vote as the real
    reference; the MSE table itself within fp32 summation-order noise."""
    import fp8_quantization_b200 as fq

    g = load_golden("mse_estimator.npz")
    for key, pc, include in (("pt_sweep", False, True), ("pt_fixed", False, False), ("pc_sweep", True, True),
                             ("pc_fixed", True, False)):
        x = torch.from_numpy(g[key + "_x"]).to(DEV)
        q = fq.FPQuantizer(8, per_channel=pc, mantissa_bits=4, set_maxval=True, mse_include_mantissa_bits=include)
        est = fq.FP_MSE_Estimator(per_channel=pc, quantizer=q)
        mn, mx = est(x)
        assert np.array_equal(est.search_grid.cpu().numpy(), g[key + "_grid"])
        np.testing.assert_allclose(est.mses.cpu().numpy(), g[key + "_mses"], rtol=2e-4, atol=1e-12)
        assert float(q._mbits_host) == float(g[key + "_best_m"]), key
        ref_mx = g[key + "_xmax"]
        same = mx.cpu().numpy() == ref_mx
        # a channel may pick a neighbouring grid point only if the two candidates' MSEs tie to 1e-4 rel
        if not same.all():
            m_idx = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0].index(float(g[key + "_best_m"])) if include else 0
            row = g[key + "_mses"][m_idx]
            for c in np.nonzero(~same)[0]:
                gi = int(np.argmin(np.abs(g[key + "_grid"][:, c] - mx.cpu().numpy()[c])))
                assert abs(row[gi, c] - row[:, c].min()) <= 2e-4 * row[:, c].min()
        q.set_quant_range(mn, mx)  # what the manager does next
        assert torch.isfinite(q(x)).all()


def test_mse_estimator_accumulates_across_batches_like_oracle():
    import fp8_quantization_b200 as fq

    torch.manual_seed(4)
    q = fq.FPQuantizer(8, mantissa_bits=3, set_maxval=True, mse_include_mantissa_bits=False)
    est = fq.FP_MSE_Estimator(quantizer=q)
    oq = O.OracleFPQuantizer(8, mantissa_bits=3, set_maxval=True, mse_include_mantissa_bits=False)
    oest = O.OracleFPMSE(quantizer=oq)
    for _ in range(3):
        x = torch.randn(4, 16, 10, 10)
        mn, mx = est(x.to(DEV))
        omn, omx = oest(x)
    np.testing.assert_allclose(est.mses.cpu().numpy(), oest.mses.numpy(), rtol=2e-4)
    assert float(mx) == float(omx)


# ---- fused calibration epilogue: statistics of act(bn(x)) without materialising it ---------------------------
@pytest.mark.parametrize("shape", [(8, 64, 56, 56), (4, 512, 7, 7), (3, 24, 9, 4), (6, 96, 14, 14), (5, 1000), (2, 3, 224, 224)])
@pytest.mark.parametrize("layout", ["nchw", "channels_last"])
def test_bn_act_estimate_prepare_equals_unfused_statistics(shape, layout):
    """fp8fq_bn_act_estimate_prepare_f32 == min / max of act(F.batch_norm(x)) (exact-BN mode reproduces ATen's NCHW
    arithmetic, so the statistics are the same bits), the estimator update rules over several calls, the fused
    set_quant_range + table == fp8fq_set_range_prepare_f32, NaN poisons the range like torch.min / torch.max."""
    import torch.nn.functional as F

    import fp8_quantization_b200 as fq
    from fp8_quantization_b200 import ops

    if layout == "channels_last" and len(shape) != 4:
        pytest.skip("channels_last is a 4-D layout")
    torch.manual_seed(41)
    C = shape[1]
    mean, var = torch.randn(C, device=DEV), torch.rand(C, device=DEV) + 0.3
    gamma, beta = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    pk = ops.bn_pack(mean, var, gamma, beta, 1e-5)
    for act, est_mode in ((1, ops.EST_ALL), (2, ops.EST_RUNNING), (0, ops.EST_CURRENT)):
        cmin, cmax = torch.empty(1, device=DEV), torch.empty(1, device=DEV)
        rmin = rmax = None
        for call in range(3):
            x = torch.randn(shape, device=DEV) * (1.0 + call)
            xin = x.contiguous(memory_format=torch.channels_last) if layout == "channels_last" else x
            maxval = torch.empty(1, device=DEV)
            table = ops.new_table(1, 5.0, 8, 1, DEV)
            ok = ops.bn_act_estimate_prepare(xin, pk, None, act, 1, cmin, cmax, est_mode, call > 0, 0.9, maxval,
                                             (5.0, 8, 1), table)
            hw = x[0, 0].numel() if x.dim() > 2 else 1
            covered = ((layout == "channels_last" or hw == 1) and C % 4 == 0) or \
                      (layout == "nchw" and hw > 1 and hw % 4 == 0 and 2 + 4095 // hw <= C)
            if not covered:
                assert not ok
                return
            assert ok
            t = F.batch_norm(x, mean, var, gamma, beta, False, 0.0, 1e-5)
            t = torch.relu(t) if act == 1 else (F.relu6(t) if act == 2 else t)
            bmin, bmax = t.min().reshape(1), t.max().reshape(1)
            if call == 0 or est_mode == ops.EST_CURRENT:
                rmin, rmax = bmin, bmax
            elif est_mode == ops.EST_ALL:
                rmin, rmax = torch.min(rmin, bmin), torch.max(rmax, bmax)
            else:
                rmin, rmax = (1 - 0.9) * bmin + 0.9 * rmin, (1 - 0.9) * bmax + 0.9 * rmax
            assert torch.equal(bits(cmin), bits(rmin)) and torch.equal(bits(cmax), bits(rmax)), (act, call)
            mv2, tb2 = ops.set_range_prepare(rmin, rmax, 5.0, 8, 1)
            assert torch.equal(bits(maxval), bits(mv2)) and torch.equal(bits(table), bits(tb2))
    x = torch.randn(shape, device=DEV)
    x.view(-1)[x.numel() // 2] = float("nan")
    xin = x.contiguous(memory_format=torch.channels_last) if layout == "channels_last" else x
    assert ops.bn_act_estimate_prepare(xin, pk, None, 1, 1, cmin, cmax, ops.EST_CURRENT, False, 0.9)
    assert torch.isnan(cmin).all() and torch.isnan(cmax).all()


@pytest.mark.parametrize("memory_format", ["nchw", "channels_last"])
def test_fused_calibration_of_resnet18_equals_op_by_op_calibration(memory_format):
    """Calibrating QuantizedResNet with the fused calibration epilogues gives the ranges of the op-by-op path.  In NCHW
    they are the same bits for every quantiser, and so are the logits; in channels_last the op-by-op path runs ATen's
    channels_last batch-norm kernel, whose arithmetic differs from its NCHW kernel by ulps (tools/bn_cl_check.py), so
    the comparison there is to 1e-6 relative."""
    from fp8_quantization_b200 import modules, ops, workloads
    from fp8_quantization_b200.quantizers import FPQuantizer

    gen = torch.Generator().manual_seed(10)
    x = torch.randn(4, 3, 224, 224, generator=gen).to(DEV)
    out = {}
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for fused in (True, False):
            modules.FUSE_CALIBRATION = fused
            torch.manual_seed(10)
            m = workloads.resnet18_quantized(**workloads.readme_quant_params(5)).to(DEV).eval()
            if memory_format == "channels_last":
                m = m.to(memory_format=torch.channels_last)
            n0 = ops.launch_count()
            workloads.pass_data_for_range_estimation([x], m, True, True, 1)
            launches = ops.launch_count() - n0
            m.fix_ranges()
            with torch.no_grad():
                y = m(x)
            out[fused] = ([q.maxval.clone() for q in m.modules() if isinstance(q, FPQuantizer)], y, launches)
    finally:
        modules.FUSE_CALIBRATION = True
        torch.backends.cudnn.allow_tf32 = prev
    for a, b in zip(out[True][0], out[False][0]):
        if memory_format == "nchw":
            assert torch.equal(bits(a), bits(b))
        else:
            assert torch.allclose(a, b, rtol=1e-5)
    if memory_format == "nchw":
        assert torch.equal(out[True][1], out[False][1])
    # same number of estimate / quantise launches either way (the 20 batch-norm parameter packs merely happen during
    # calibration instead of at the first fused forward); what disappears are ATen's batch-norm and ReLU kernels and
    # 16 of the 28 bytes moved per element
    assert out[True][2] == out[False][2] + 20
