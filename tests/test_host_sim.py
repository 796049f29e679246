"""CPU: the CUDA kernels' OWN source (fp8_quantization_b200/csrc/fp8fq_kernels.cu), compiled by g++ against a CUDA shim
and executed by a cooperative grid simulator (tests/host_sim), checked against the C oracle -- the direct per-element
evaluation of the reference formula (oracle/fp8_oracle_c.c, fp8_quantizer.py:105-133) -- through the SAME C ABI
(include/fp8fq.h) the GPU tests call.  Both sides use the host libm, so everything is compared bit for bit.

What this covers that tests/test_host_emul.py (the arithmetic core alone) does not: the kernels' tile / row / channel
index arithmetic in every layout class (NCHW rows with H*W % 4 == 0 or not, flat-division variant, channel-innermost
with C dividing the pass stride, CTA size fitted to C, per-vector channels, scalar accesses), vector and scalar tails,
misaligned pointers, the multi-tensor work-item mapping, the two-stage / last-CTA reductions with the estimator update
rules and the fused prologue, the MSE grid kernel's staging, the STE backward's accumulators, the host-buffer entry
point's chunking.  It says nothing about the GPU's libdevice, memory model or speed: those are the -m gpu tests.

TEST ONLY: the simulation library is never loaded by the product package (which has no CPU path)."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

F, I, L, D, VP = ctypes.c_float, ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 2
EST_CURRENT, EST_ALL, EST_RUNNING = 0, 1, 2


def P(a):
    return a.ctypes.data_as(VP) if a is not None else None


def bits(a):
    return np.ascontiguousarray(a).view(np.int32)


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(bits(a), bits(b))


@pytest.fixture(scope="module")
def sim(built):
    """libfp8fq_sim.so with the product's own ctypes signature table applied (so the table is exercised too)."""
    from fp8_quantization_b200._lib import SIGNATURES

    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libfp8fq_sim.so"))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    lib.fp8fq_sim_ctas.restype = L
    lib.fp8fq_sim_launches.restype = L
    return lib


@pytest.fixture(scope="module")
def ref(oracle_c):
    oracle_c.oracle_c_fake_quant.restype = None
    oracle_c.oracle_c_bn_act.restype = None
    return oracle_c


def aligned(n, dtype=np.float32, offset_elems=0):
    """A float32 array of n elements whose address is 16-byte aligned + 4 * offset_elems bytes."""
    raw = np.zeros(n + 8 + offset_elems, dtype=dtype)
    start = ((-raw.ctypes.data) % 16) // 4 + offset_elems
    return raw[start:start + n]


def rand(rng, shape, scale=2.0, specials=True):
    x = (rng.standard_normal(shape) * scale).astype(np.float32)
    flat = x.reshape(-1)
    if specials and flat.size >= 16:
        idx = rng.choice(flat.size, size=min(8, flat.size), replace=False)
        vals = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-30, -1e-38, 65504.0], np.float32)
        flat[idx] = vals[:idx.size]
    return x


def table_for(sim, maxval, M, nb=8, sb=1):
    maxval = np.ascontiguousarray(maxval, np.float32).reshape(-1)
    stride = sim.fp8fq_table_stride(M, nb, sb)
    assert stride > 0
    tab = aligned(stride * maxval.size)
    assert sim.fp8fq_prepare_f32(P(maxval), maxval.size, M, nb, sb, P(tab), None) == 0
    return tab


def ref_quant(ref, x, maxval, M, nb=8, sb=1, per_channel=False):
    """The direct formula (C oracle) over x viewed as [C, inner]."""
    x = np.ascontiguousarray(x, np.float32)
    maxval = np.ascontiguousarray(maxval, np.float32).reshape(-1)
    C = maxval.size if per_channel else 1
    y = np.empty_like(x)
    e = np.empty_like(x)
    q = np.empty_like(x)
    ref.oracle_c_fake_quant(P(x), P(y), P(e), P(q), P(maxval), L(C), L(x.size // C), F(M), I(nb), I(sb))
    return y, e, q


# ---------------------------------------------------------------------------------------------------------------------
# prologue
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sb", [1, 0])
def test_prepare_and_set_range_prepare_equal_the_serial_prologue(sim, host_emul, sb):
    """prepare_kernel (one CTA per channel, one thread per exponent code, barriers) builds the tables the serial
    host emulation builds; set_range_prepare additionally applies fp8_quantizer.py:236-237."""
    rng = np.random.default_rng(1)
    for M in range(1, 8):
        C = 37
        mv = np.abs(rng.standard_normal(C)).astype(np.float32) * 3 + 1e-3
        mv[:3] = [0.0, np.inf, 1e-30]            # degenerate ranges
        tab = table_for(sim, mv, M, 8, sb)
        tab2 = np.zeros_like(tab)
        assert host_emul.emul_prepare(P(mv), L(C), F(M), 8, sb, P(tab2)) == 0
        assert same_bits(tab, tab2), (M, sb)
        xmin = -np.abs(rng.standard_normal(C)).astype(np.float32) * 4
        xmax = rng.standard_normal(C).astype(np.float32) * 4
        xmax[5] = np.nan
        mv_out, tab3 = aligned(C), aligned(tab.size)
        assert sim.fp8fq_set_range_prepare_f32(P(xmin), P(xmax), C, P(mv_out), M, 8, sb, P(tab3), None) == 0
        a = np.abs(xmin)
        want = np.abs(np.where(np.isnan(a) | np.isnan(xmax), np.float32(np.nan), np.maximum(a, xmax)))
        assert same_bits(mv_out, want.astype(np.float32))
        tab4 = np.zeros_like(tab)
        host_emul.emul_prepare(P(np.ascontiguousarray(mv_out)), L(C), F(M), 8, sb, P(tab4))
        assert same_bits(tab3, tab4)


# ---------------------------------------------------------------------------------------------------------------------
# K1 per tensor (fq_stream_kernel<PRE_PLAIN>)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,sb", [(5, 1), (4, 1), (2, 1), (7, 1), (3, 0), (6, 0)])
def test_stream_kernel_per_tensor_equals_direct_formula(sim, ref, M, sb):
    """Every size class: below one vector, tile boundaries (4096 elements per CTA) +-1, ragged vector tails; 16-byte
    aligned (128-bit accesses) and misaligned (scalar accesses) pointers; y aliasing x; the code-plane variant."""
    rng = np.random.default_rng(100 + M)
    mv = np.array([2.5], np.float32)
    tab = table_for(sim, mv, M, 8, sb)
    for n in (1, 2, 3, 4, 5, 1023, 1024, 4095, 4096, 4097, 8191, 3 * 4096 + 2, 50001):
        for off in (0, 1):
            x = aligned(n, offset_elems=off)
            x[:] = rand(rng, n)
            y = aligned(n, offset_elems=off)
            assert sim.fp8fq_fake_quant_f32(P(x), P(y), P(tab), n, 1, n, M, 8, sb, None) == 0
            yr, er, qr = ref_quant(ref, x, mv, M, 8, sb)
            assert same_bits(y, yr), (M, sb, n, off)
            # codes: sign << 31 | e << 16 | |q| (NaN -> 0x7fffffff), the reference's intermediate integers
            codes = aligned(n, np.int32, offset_elems=off)
            y2 = aligned(n, offset_elems=off)
            assert sim.fp8fq_fake_quant_codes_f32(P(x), P(y2), P(codes), P(tab), n, 1, n, M, 8, sb, None) == 0
            assert same_bits(y2, yr)
            ok = ~np.isnan(yr)
            with np.errstate(invalid="ignore"):
                want = ((bits(yr) & np.int32(-2**31)).astype(np.int64) & 0xFFFFFFFF) | (er.astype(np.int64) << 16) | \
                    np.abs(qr).astype(np.int64)
            assert np.array_equal(codes[ok].astype(np.int64) & 0xFFFFFFFF, want[ok])
            assert np.all(codes[~ok] == 0x7FFFFFFF)
            # in place
            assert sim.fp8fq_fake_quant_f32(P(x), P(x), P(tab), n, 1, n, M, 8, sb, None) == 0
            assert same_bits(x, yr)


def test_stream_kernel_grid_stride_path(sim, ref):
    """More tiles than CTAs: FP8FQ's launcher caps the grid at 2^31-1, so the kernel's tile loop normally runs once;
    the multi-tile loop is what the BN variants use (tiles per CTA 2..4) -- covered by the fused tests below.  Here:
    a tensor of many tiles, plain kernel, results independent of tiling."""
    rng = np.random.default_rng(7)
    n = 40 * 4096 + 17
    x = aligned(n)
    x[:] = rand(rng, n)
    mv = np.array([np.abs(x[np.isfinite(x)]).max()], np.float32)
    tab = table_for(sim, mv, 5)
    y = aligned(n)
    c0 = sim.fp8fq_sim_ctas()
    assert sim.fp8fq_fake_quant_f32(P(x), P(y), P(tab), n, 1, n, 5, 8, 1, None) == 0
    assert sim.fp8fq_sim_ctas() - c0 == 41     # one 4096-element tile per CTA
    assert same_bits(y, ref_quant(ref, x, mv, 5)[0])


# ---------------------------------------------------------------------------------------------------------------------
# K1 per channel (fq_rows_kernel): warp per (tensor, row, 1024-chunk) work item, multi-tensor launches
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M", [5, 4, 2])
def test_rows_kernel_per_channel_equals_direct_formula(sim, ref, M):
    rng = np.random.default_rng(200 + M)
    # the reference's weight shapes: inner 9 (depthwise), 27, 147, 576, 1024-chunk edges, 4608
    for C, inner in ((5, 1), (32, 9), (16, 27), (64, 147), (7, 576), (3, 1023), (3, 1024), (3, 1025), (2, 4608), (1000, 4)):
        for off in (0, 1):
            x = aligned(C * inner, offset_elems=off)
            x[:] = rand(rng, C * inner, scale=0.5)
            mv = np.abs(np.nan_to_num(x.reshape(C, inner), nan=0.0, posinf=1.0, neginf=1.0)).max(1).astype(np.float32)
            if C > 2:
                x.reshape(C, inner)[1] = 0.0
                mv[1] = 0.0                      # all-zero row: the whole channel is NaN in the reference
            tab = table_for(sim, mv, M)
            y = aligned(C * inner, offset_elems=off)
            assert sim.fp8fq_fake_quant_f32(P(x), P(y), P(tab), C * inner, C, inner, M, 8, 1, None) == 0
            yr, _, _ = ref_quant(ref, x, mv, M, per_channel=True)
            assert same_bits(y, yr), (M, C, inner, off)
            if C > 2:
                assert np.all(np.isnan(y.reshape(C, inner)[1]))


def test_multi_tensor_launch_maps_work_items_to_the_right_rows(sim, ref):
    """fp8fq_fake_quant_multi_f32: ResNet-18-like weight list (+ an empty tensor, + more than 48 tensors so that the
    launch is split): every tensor equals its own single-tensor result."""
    from fp8_quantization_b200._lib import TensorDesc

    rng = np.random.default_rng(3)
    shapes = [(64, 147), (64, 576), (128, 576), (128, 1152), (10, 2304), (0, 16), (6, 5), (1, 1)] + [(3, 7 + i) for i in range(50)]
    xs, ys, tabs, mvs = [], [], [], []
    for C, inner in shapes:
        x = aligned(max(C * inner, 1))[:C * inner]
        x[:] = rand(rng, C * inner, scale=0.3, specials=False)
        mv = np.abs(x.reshape(C, inner)).max(1).astype(np.float32) if C else np.zeros(0, np.float32)
        xs.append(x), ys.append(aligned(max(C * inner, 1))[:C * inner]), mvs.append(mv)
        tabs.append(table_for(sim, mv, 5) if C else aligned(4))
    descs = (TensorDesc * len(shapes))()
    for d, x, y, t, (C, inner) in zip(descs, xs, ys, tabs, shapes):
        d.x, d.y, d.table, d.C, d.inner = x.ctypes.data, y.ctypes.data, t.ctypes.data, max(C, 1), inner if C else 0
    l0 = sim.fp8fq_sim_launches()
    assert sim.fp8fq_fake_quant_multi_f32(descs, len(shapes), 5, 8, 1, None) == 0
    assert sim.fp8fq_sim_launches() - l0 == 2      # 57 non-empty tensors -> 48 + 9
    for x, y, mv, (C, inner) in zip(xs, ys, mvs, shapes):
        if C:
            assert same_bits(y, ref_quant(ref, x, mv, 5, per_channel=True)[0]), (C, inner)


# ---------------------------------------------------------------------------------------------------------------------
# fused epilogues
# ---------------------------------------------------------------------------------------------------------------------
def bn_params(sim, rng, C, mode):
    mean = rng.standard_normal(C).astype(np.float32)
    var = (rng.random(C) + 0.3).astype(np.float32)
    gamma = rng.standard_normal(C).astype(np.float32)
    beta = rng.standard_normal(C).astype(np.float32)
    if mode == 1:
        packed = aligned(4 * C)
        assert sim.fp8fq_bn_pack_f32(P(mean), P(var), P(gamma), P(beta), 1e-5, C, P(packed), None) == 0
        pk = packed.reshape(C, 4)
        assert same_bits(pk[:, 0], mean) and same_bits(pk[:, 1], gamma) and same_bits(pk[:, 3], beta)
        np.testing.assert_allclose(pk[:, 2], 1 / np.sqrt(var + np.float32(1e-5)), rtol=3e-7)
        return packed, None
    scale, shift = aligned(C), aligned(C)
    assert sim.fp8fq_bn_fold_f32(P(mean), P(var), P(gamma), P(beta), 1e-5, C, P(scale), P(shift), None) == 0
    inv = np.float32(1) / np.sqrt(var + np.float32(1e-5), dtype=np.float32)
    assert same_bits(scale, gamma * inv) and same_bits(shift, beta - mean * (gamma * inv))
    return scale, shift


def ref_bn_act(ref, x, hw, C, layout, mode, p0, p1, act):
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    ref.oracle_c_bn_act(P(x), P(y), L(x.size), L(hw), L(C), I(layout), I(mode), P(p0), P(p1), I(act))
    return y


NCHW_SHAPES = [
    (2, 64, 56 * 56),    # H*W % 4 == 0: tile-local rows, 128-bit accesses
    (3, 16, 8 * 8),      # many rows per tile (channel wrap inside a tile)
    (2, 24, 9 * 5),      # H*W % 4 != 0, aligned base: per-lane rows
    (3, 5, 17 * 13),     # odd everything
    (1, 1, 20000),       # one channel, rows longer than a tile
    (4, 3, 4096),        # rows exactly one tile
    (2, 2, 70000),       # rows much longer than a tile
    (64, 8, 2),          # tiny rows: more rows per tile than channels -> flat-division variant
    (50, 3, 1),          # H*W == 1 through the NCHW entry point -> flat-division variant
]


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("shape", NCHW_SHAPES)
def test_bn_act_quant_nchw_all_layout_classes(sim, ref, shape, mode):
    """fp8fq_bn_act_quant_f32 = Q(act(bn(x))) (quantized_folded_bn.py:39-55) for every NCHW layout class, both
    batch-norm arithmetic modes, the three activations, K <= 3 and K > 3 formats, aligned and misaligned bases."""
    N, C, hw = shape
    rng = np.random.default_rng(hash(shape) % 2**31 + mode)
    n = N * C * hw
    p0, p1 = bn_params(sim, rng, C, mode)
    for (M, act), off in zip(((5, ACT_RELU), (4, ACT_RELU6), (3, ACT_NONE), (5, ACT_NONE)), (0, 0, 1, 1)):
        x = aligned(n, offset_elems=off)
        x[:] = rand(rng, n)
        mv = np.array([3.0], np.float32)
        tab = table_for(sim, mv, M)
        y = aligned(n, offset_elems=off)
        rows = N * C
        assert sim.fp8fq_bn_act_quant_f32(P(x), P(y), P(p0), P(p1), rows, hw, C, act, mode, P(tab), M, 8, 1, None) == 0
        v = ref_bn_act(ref, x, hw, C, 0, mode, p0, p1, act)
        assert same_bits(y, ref_quant(ref, v, mv, M)[0]), (shape, mode, M, act, off)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("shape", [(2, 64, 28 * 28), (3, 128, 8 * 8), (3, 16, 8 * 8), (2, 96, 9 * 5), (3, 5, 17 * 113),
                                   (1, 2, 9000), (2, 3, 4099), (64, 8, 2)])
def test_block_tail_nchw_equals_composition(sim, ref, shape, mode):
    """fp8fq_bn_quant_add_act_quant_f32 = Q_outer(act(Q_inner(bn(x)) + residual)) (models/resnet_quantized.py:39-46);
    shapes the fused variant does not cover answer FP8FQ_ERR_UNSUPPORTED."""
    N, C, hw = shape
    rng = np.random.default_rng(hash(shape) % 2**31 + 17 * mode)
    n = N * C * hw
    p0, p1 = bn_params(sim, rng, C, mode)
    for (Mi, Mo, act), off in zip(((5, 5, ACT_RELU), (4, 3, ACT_NONE), (5, 4, ACT_RELU6)), (0, 0, 1)):
        x, res, y = aligned(n, offset_elems=off), aligned(n, offset_elems=off), aligned(n, offset_elems=off)
        x[:] = rand(rng, n)
        res[:] = np.maximum(rand(rng, n, specials=False), 0)
        mvi, mvo = np.array([2.7], np.float32), np.array([4.1], np.float32)
        ti, to = table_for(sim, mvi, Mi), table_for(sim, mvo, Mo)
        code = sim.fp8fq_bn_quant_add_act_quant_f32(P(x), P(res), P(y), P(p0), P(p1), N * C, hw, C, act, mode, P(ti), Mi,
                                                    8, 1, P(to), Mo, 8, 1, None)
        # the fused tail exists for the tile-local-rows variants only: at most one channel wrap per 4096-element tile
        if not (hw > 1 and 2 + 4095 // hw <= C):
            assert code == -2, shape
            continue
        assert code == 0, shape
        inner = ref_quant(ref, ref_bn_act(ref, x, hw, C, 0, mode, p0, p1, ACT_NONE), mvi, Mi)[0]
        v = inner + res
        if act >= ACT_RELU:
            v = np.where(np.isnan(v), v, np.maximum(v, 0))
        if act == ACT_RELU6:
            v = np.where(np.isnan(v), v, np.minimum(v, 6))
        assert same_bits(y, ref_quant(ref, v.astype(np.float32), mvo, Mo)[0]), (shape, mode, Mi, Mo, act, off)


NHWC_SHAPES = [
    (56 * 56 * 2, 64),    # C divides the 1024-element pass stride: parameters loaded once per tile
    (200, 4), (300, 8), (77, 256), (33, 1024),
    (14 * 14 * 3, 96),    # MobileNetV2 widths: CTA size fitted to C (240 threads)
    (28 * 28, 144), (100, 24), (50, 192), (20, 384), (30, 576), (9, 960),
    (7 * 7 * 2, 1280),    # C / 4 > 256 lanes: per-vector channel look-up
    (5, 2048), (3, 4100),
    (40, 30),             # C % 4 != 0: scalar accesses, 240 threads
    (221, 3), (64, 1), (1000, 10),
    (128, 1000),          # Linear output [N, C]
]


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("shape", NHWC_SHAPES)
def test_bn_act_quant_and_block_tail_channel_innermost(sim, ref, shape, mode):
    """fp8fq_bn_act_quant_nhwc_f32 / fp8fq_bn_quant_add_act_quant_nhwc_f32 on [pixels, C] memory (channels_last
    activations, Linear outputs): every channel-count class incl. the run-time CTA size, misaligned bases."""
    pixels, C = shape
    rng = np.random.default_rng(hash(shape) % 2**31 + 5 * mode)
    n = pixels * C
    p0, p1 = bn_params(sim, rng, C, mode)
    for (M, act), off in zip(((5, ACT_RELU), (4, ACT_RELU6), (3, ACT_NONE)), (0, 0, 1)):
        x, y = aligned(n, offset_elems=off), aligned(n, offset_elems=off)
        x[:] = rand(rng, n)
        mv = np.array([3.0], np.float32)
        tab = table_for(sim, mv, M)
        assert sim.fp8fq_bn_act_quant_nhwc_f32(P(x), P(y), P(p0), P(p1), pixels, C, act, mode, P(tab), M, 8, 1, None) == 0
        v = ref_bn_act(ref, x, 1, C, 1, mode, p0, p1, act)
        assert same_bits(y, ref_quant(ref, v, mv, M)[0]), (shape, mode, M, act, off)
    for (Mi, Mo, act), off in zip(((5, 5, ACT_RELU), (4, 3, ACT_NONE)), (0, 1)):
        x, res, y = aligned(n, offset_elems=off), aligned(n, offset_elems=off), aligned(n, offset_elems=off)
        x[:] = rand(rng, n)
        res[:] = np.maximum(rand(rng, n, specials=False), 0)
        mvi, mvo = np.array([2.7], np.float32), np.array([4.1], np.float32)
        ti, to = table_for(sim, mvi, Mi), table_for(sim, mvo, Mo)
        assert sim.fp8fq_bn_quant_add_act_quant_nhwc_f32(P(x), P(res), P(y), P(p0), P(p1), pixels, C, act, mode, P(ti), Mi,
                                                         8, 1, P(to), Mo, 8, 1, None) == 0
        inner = ref_quant(ref, ref_bn_act(ref, x, 1, C, 1, mode, p0, p1, ACT_NONE), mvi, Mi)[0]
        v = inner + res
        if act == ACT_RELU:
            v = np.where(np.isnan(v), v, np.maximum(v, 0))
        assert same_bits(y, ref_quant(ref, v.astype(np.float32), mvo, Mo)[0]), (shape, mode, Mi, Mo, act, off)


def test_add_act_quant_equals_composition(sim, ref):
    """fp8fq_add_act_quant_f32 = Q(act(a + b)) (models/resnet_quantized.py:43-46, mobilenet_v2_quantized.py:22-24)."""
    rng = np.random.default_rng(11)
    mv = np.array([3.3], np.float32)
    for M in (5, 4):
        tab = table_for(sim, mv, M)
        for n in (1, 5, 2048, 2049, 9000):
            for off in (0, 1):
                for act in (ACT_NONE, ACT_RELU, ACT_RELU6):
                    a, b, y = aligned(n, offset_elems=off), aligned(n, offset_elems=off), aligned(n, offset_elems=off)
                    a[:], b[:] = rand(rng, n), rand(rng, n)
                    assert sim.fp8fq_add_act_quant_f32(P(a), P(b), P(y), n, act, P(tab), M, 8, 1, None) == 0
                    v = a + b
                    if act >= ACT_RELU:
                        v = np.where(np.isnan(v), v, np.maximum(v, 0))
                    if act == ACT_RELU6:
                        v = np.where(np.isnan(v), v, np.minimum(v, 6))
                    assert same_bits(y, ref_quant(ref, v.astype(np.float32), mv, M)[0]), (M, n, off, act)
