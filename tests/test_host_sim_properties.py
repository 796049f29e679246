"""CPU: property-based tests (hypothesis, derandomised) of the kernels' index arithmetic on the host simulation
(tests/host_sim): random shapes, channel counts, alignments, formats and activations for the fused epilogues in both
memory layouts, the per-channel row kernel and the reductions, each checked bit for bit against the direct-formula
C oracle.  The fixed shape lists of tests/test_host_sim.py cover the layout classes by construction; this searches
for the ones nobody thought of (exact unsigned divisions by run-time constants, multiply-high row look-ups, CTA sizes
fitted to the channel count, ragged tails)."""
import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from test_host_sim import (ACT_NONE, ACT_RELU, ACT_RELU6, EST_CURRENT, L, P, aligned, bn_params, np_minmax, rand,  # noqa: F401
                           ref, ref_bn_act, ref_quant, same_bits, sim, table_for, workspace)

import os

# FP8FQ_HYP_SCALE=20 multiplies the example budgets and draws fresh random examples (a bug hunt); the default run is
# small and derandomised so that the suite is reproducible.
SCALE = int(os.environ.get("FP8FQ_HYP_SCALE", "1"))
COMMON = dict(deadline=None, derandomize=SCALE == 1, suppress_health_check=[HealthCheck.function_scoped_fixture,
                                                                            HealthCheck.too_slow, HealthCheck.data_too_large])


def _apply_act(v, act):
    if act >= ACT_RELU:
        v = np.where(np.isnan(v), v, np.maximum(v, 0))
    if act == ACT_RELU6:
        v = np.where(np.isnan(v), v, np.minimum(v, 6))
    return v.astype(np.float32)


@settings(max_examples=120 * SCALE, **COMMON)
@given(N=st.integers(1, 5), C=st.integers(1, 70), hw=st.one_of(st.integers(1, 40), st.integers(41, 5000)),
       mode=st.integers(0, 1), act=st.integers(0, 2), M=st.sampled_from([5, 4, 3, 2]), off=st.integers(0, 3),
       seed=st.integers(0, 2**16))
def test_bn_act_quant_nchw_random_geometry(sim, ref, N, C, hw, mode, act, M, off, seed):
    rng = np.random.default_rng(seed)
    n = N * C * hw
    p0, p1 = bn_params(sim, rng, C, mode)
    x, y = aligned(n, offset_elems=off), aligned(n, offset_elems=off)
    x[:] = rand(rng, n)
    mv = np.array([3.0], np.float32)
    tab = table_for(sim, mv, M)
    assert sim.fp8fq_bn_act_quant_f32(P(x), P(y), P(p0), P(p1), N * C, hw, C, act, mode, P(tab), M, 8, 1, None) == 0
    v = ref_bn_act(ref, x, hw, C, 0, mode, p0, p1, act)
    assert same_bits(y, ref_quant(ref, v, mv, M)[0])


@settings(max_examples=120 * SCALE, **COMMON)
@given(pixels=st.integers(1, 700), C=st.one_of(st.integers(1, 64), st.sampled_from([96, 144, 192, 320, 384, 576, 960, 1000, 1024,
                                                                                 1028, 1280, 2052])),
       mode=st.integers(0, 1), act=st.integers(0, 2), M=st.sampled_from([5, 4, 3]), off=st.integers(0, 3),
       tail=st.booleans(), seed=st.integers(0, 2**16))
def test_channel_innermost_epilogues_random_geometry(sim, ref, pixels, C, mode, act, M, off, tail, seed):
    rng = np.random.default_rng(seed)
    pixels = min(pixels, max(1, 60000 // C))
    n = pixels * C
    p0, p1 = bn_params(sim, rng, C, mode)
    x, y = aligned(n, offset_elems=off), aligned(n, offset_elems=off)
    x[:] = rand(rng, n)
    mv = np.array([3.0], np.float32)
    tab = table_for(sim, mv, M)
    if not tail:
        assert sim.fp8fq_bn_act_quant_nhwc_f32(P(x), P(y), P(p0), P(p1), pixels, C, act, mode, P(tab), M, 8, 1, None) == 0
        v = ref_bn_act(ref, x, 1, C, 1, mode, p0, p1, act)
        assert same_bits(y, ref_quant(ref, v, mv, M)[0])
        return
    res = aligned(n, offset_elems=off)
    res[:] = np.maximum(rand(rng, n, specials=False), 0)
    mvo = np.array([4.1], np.float32)
    to = table_for(sim, mvo, 5)
    assert sim.fp8fq_bn_quant_add_act_quant_nhwc_f32(P(x), P(res), P(y), P(p0), P(p1), pixels, C, act, mode, P(tab), M, 8, 1,
                                                     P(to), 5, 8, 1, None) == 0
    inner = ref_quant(ref, ref_bn_act(ref, x, 1, C, 1, mode, p0, p1, ACT_NONE), mv, M)[0]
    assert same_bits(y, ref_quant(ref, _apply_act(inner + res, act), mvo, 5)[0])


@settings(max_examples=80 * SCALE, **COMMON)
@given(C=st.integers(1, 40), inner=st.one_of(st.integers(1, 64), st.integers(65, 3000)), M=st.sampled_from([5, 4, 2, 7]),
       sb=st.integers(0, 1), off=st.integers(0, 3), seed=st.integers(0, 2**16))
def test_rows_kernel_random_geometry(sim, ref, C, inner, M, sb, off, seed):
    rng = np.random.default_rng(seed)
    n = C * inner
    x, y = aligned(n, offset_elems=off), aligned(n, offset_elems=off)
    x[:] = rand(rng, n, scale=0.7)
    mv = (np.abs(rng.standard_normal(C)) + 0.05).astype(np.float32)
    tab = table_for(sim, mv, M, 8, sb)
    assert sim.fp8fq_fake_quant_f32(P(x), P(y), P(tab), n, C, inner, M, 8, sb, None) == 0
    assert same_bits(y, ref_quant(ref, x, mv, M, 8, sb, per_channel=True)[0])


@settings(max_examples=60 * SCALE, **COMMON)
@given(nhwc=st.booleans(), N=st.integers(1, 4), C4=st.integers(1, 80), hw4=st.integers(1, 600), mode=st.integers(0, 1),
       act=st.integers(0, 2), seed=st.integers(0, 2**16))
def test_fused_calibration_statistics_random_geometry(sim, ref, nhwc, N, C4, hw4, mode, act, seed):
    """fp8fq_bn_act_estimate_prepare_f32: where the fused kernel accepts the shape its min / max equal those of the
    materialised act(bn(x)); where it does not it says so (FP8FQ_ERR_UNSUPPORTED) and touches nothing."""
    rng = np.random.default_rng(seed)
    ws = workspace(sim)
    if nhwc:
        C, pixels = 4 * C4, min(N * hw4, max(1, 40000 // (4 * C4)))
        hw, outer, n = 1, pixels, pixels * C
        supported = True
    else:
        C, hw = C4, 4 * hw4
        outer, n = N * C, N * C * hw
        supported = 2 + 4095 // hw <= C
    p0, p1 = bn_params(sim, rng, C, mode)
    x = aligned(n)
    x[:] = rand(rng, n, specials=False)
    cmin, cmax = aligned(1), aligned(1)
    cmin[0], cmax[0] = 123.0, -123.0
    code = sim.fp8fq_bn_act_estimate_prepare_f32(P(x), outer, hw, C, int(nhwc), P(p0), P(p1), mode, act, P(cmin), P(cmax),
                                                 EST_CURRENT, 0, 0.9, None, 0.0, 0, 0, None, P(ws), None)
    if not supported:
        assert code == -2 and cmin[0] == 123.0 and cmax[0] == -123.0
        return
    assert code == 0
    v = ref_bn_act(ref, x, hw, C, int(nhwc), mode, p0, p1, act)
    lo, hi = np_minmax(v)
    assert same_bits(np.array([cmin[0], cmax[0]]), np.array([lo, hi])) and ws[0] == 0
