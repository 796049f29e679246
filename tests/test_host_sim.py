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


def same_values(a, b):
    """Bit equality, any NaN equal to any NaN (payload and sign of a NaN are not part of the contract)."""
    return a.shape == b.shape and bool(np.all((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))))


@pytest.fixture(scope="module")
def sim(built):
    """libfp8fq_sim.so with the product's own ctypes signature table applied (so the table is exercised too)."""
    from fp8_quantization_b200._lib import SIGNATURES

    # FP8FQ_SIM_LIB: an alternative build of the same simulation (e.g. -fsanitize=address, see tests/host_sim/README.md)
    lib = ctypes.CDLL(os.environ.get("FP8FQ_SIM_LIB") or os.path.join(ROOT, "oracle", "_build", "libfp8fq_sim.so"))
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
    """An n-element array whose address is 16-byte aligned + 4 * offset_elems bytes and which ENDS where its allocation
    ends (malloc returns 16-byte aligned blocks), so that under the AddressSanitizer build of the simulation an access
    one element past the end -- or before the start -- lands in a redzone."""
    raw = np.zeros(n + offset_elems, dtype=dtype)
    if raw.ctypes.data % 16 == 0:
        return raw[offset_elems:]
    raw = np.zeros(n + 8 + offset_elems, dtype=dtype)      # allocator without 16-byte alignment: align by hand
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
    (7 * 7 * 2, 1280),    # C / 4 = 320 lanes: the fitted 320-thread CTA (round 1: per-vector channel look-up)
    (6, 2048),            # C / 4 > 320 lanes: per-vector channel look-up
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


def test_several_tiles_per_cta_strided_tile_loop(sim, ref):
    """The launcher gives the batch-norm kernels several tiles per CTA (4 for K <= 3 formats, 2 for the fitted-CTA K > 3
    instantiations) as long as >= 12 waves of CTAs remain -- on 148 SMs that needs tens of millions of elements.  With the
    simulated SM count set to 1 the same rule applies at test sizes: the grid is smaller than the tile count and every
    CTA strides over its tiles, incl. a partial last tile, in both layouts and for both element paths."""
    sim.fp8fq_sim_set_sm_count.argtypes = [ctypes.c_int]
    sim.fp8fq_sim_set_sm_count.restype = None
    sim.fp8fq_sim_ctas.restype = ctypes.c_int64
    rng = np.random.default_rng(77)
    try:
        sim.fp8fq_sim_set_sm_count(1)
        # the rule is in force: [1100, 96] with M = 4 is 34 tiles of 192 x 16 elements on 17 CTAs
        x, y = aligned(1100 * 96), aligned(1100 * 96)
        x[:] = rand(rng, x.size)
        p0, p1 = bn_params(sim, rng, 96, 1)
        c0 = sim.fp8fq_sim_ctas()
        assert sim.fp8fq_bn_act_quant_nhwc_f32(P(x), P(y), P(p0), P(p1), 1100, 96, ACT_RELU6, 1,
                                               P(table_for(sim, np.array([3.0], np.float32), 4)), 4, 8, 1, None) == 0
        assert sim.fp8fq_sim_ctas() - c0 == 1 + 18, sim.fp8fq_sim_ctas() - c0   # (1 CTA of the table prologue)
        for (pixels, C) in ((1100, 96), (700, 144), (650, 64)):       # 34 / 25 / 10 tiles: fitted CTA (192, 252) and 256
            n = pixels * C
            for mode in (0, 1):
                p0, p1 = bn_params(sim, rng, C, mode)
                for M, act, mvv in ((4, ACT_RELU6, 3.0), (4, ACT_RELU, 4.0), (5, ACT_RELU, 3.0)):
                    x, y = aligned(n), aligned(n)
                    x[:] = rand(rng, n)
                    mv = np.array([mvv], np.float32)
                    tab = table_for(sim, mv, M)
                    assert sim.fp8fq_bn_act_quant_nhwc_f32(P(x), P(y), P(p0), P(p1), pixels, C, act, mode, P(tab), M, 8, 1,
                                                           None) == 0
                    v = ref_bn_act(ref, x, 1, C, 1, mode, p0, p1, act)
                    assert same_bits(y, ref_quant(ref, v, mv, M)[0]), (pixels, C, mode, M, act)
        # NCHW: [N * C rows, hw], 4 tiles per CTA for M = 5
        N, C, hw = 6, 16, 2052
        n = N * C * hw
        p0, p1 = bn_params(sim, rng, C, 1)
        for M, act in ((5, ACT_RELU), (4, ACT_RELU6)):
            x, y = aligned(n), aligned(n)
            x[:] = rand(rng, n)
            mv = np.array([3.0], np.float32)
            tab = table_for(sim, mv, M)
            assert sim.fp8fq_bn_act_quant_f32(P(x), P(y), P(p0), P(p1), N * C, hw, C, act, 1, P(tab), M, 8, 1, None) == 0
            v = ref_bn_act(ref, x, hw, C, 0, 1, p0, p1, act)
            assert same_bits(y, ref_quant(ref, v, mv, M)[0]), (M, act)
    finally:
        sim.fp8fq_sim_set_sm_count(148)


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


@pytest.mark.parametrize("variant", ["nofoldact", "nofulltile", "fulltilecl", "magicopts", "magicone"])
def test_build_options_are_bit_identical_to_the_default_build(sim, variant):
    """The product's build options (csrc/fp8fq_kernels.cu) -- FP8FQ_FOLD_ACT: ReLU / ReLU6 folded into the quantiser's
    clamp bounds; FP8FQ_FULL_TILE: a second, predicate-free instantiation of the stream kernel's tile body for full
    tiles (the launch below has four full tiles and a partial one); both ON by default since round 2, so the first two
    variants are the round-1 arithmetic; "fulltilecl": the predicate-free body for the channel-innermost variants too
    (+ the PACK2 / PIN_SEL code paths, which the host build evaluates with scalar arithmetic; this build has the
    scaled-domain element path FP8FQ_MAGIC off); "magicopts": that path with its three options flipped (K <= 3 formats on
    it too, two-group tables on it, one loop for one- and two-group tables, tie-guard lanes finished by a division inside the
    path); "magicone": the default build with two-group tables on the look-up path (FP8FQ_MAGIC_TWO=0) --
    against the default build, bit for bit, on inputs made of the cases the equivalence has to survive: +-0 (identity
    batch norm: scale 1, shift -0.0, so that -0.0 reaches the activation), +-inf, NaN, values around 0 / 6 / maxval,
    ranges below and above 6, zero / inf / NaN ranges, signed and unsigned formats, K <= 3 and K > 3, all three fused
    entry points in both layouts.  (The whole of this module also passes with FP8FQ_SIM_LIB pointing at that build.)"""
    from fp8_quantization_b200._lib import SIGNATURES

    fold = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", f"libfp8fq_sim_{variant}.so"))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(fold, name)
        fn.restype, fn.argtypes = res, args
    rng = np.random.default_rng(5)
    C, hw, N = 8, 1028, 2     # (the NCHW block tail needs 2 + 4095 // hw <= C)
    n = N * C * hw
    one, negzero = aligned(C), aligned(C)
    one[:], negzero[:] = 1.0, -0.0
    compared = 0
    for mvv in (3.0, 7.5, 6.0, 0.4, 0.0, np.inf, np.nan, 1e-38):
        mv = np.array([mvv], np.float32)
        sp = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-30, -1e-30, 6.0, np.nextafter(np.float32(6), np.float32(7)),
                       np.nextafter(np.float32(6), np.float32(0)), -6.0, mvv, -mvv, 1e30, -1e30, 5.9999, 1e-45, -1e-45],
                      np.float32)
        for M, sb in ((5, 1), (3, 1), (4, 0), (2, 0), (7, 1)):
            for act in (ACT_NONE, ACT_RELU, ACT_RELU6):
                x, r = aligned(n), aligned(n)
                x[:] = rng.standard_normal(n).astype(np.float32) * 3
                x[:: 7][:sp.size] = sp
                x[1::11][:sp.size] = sp[::-1]
                r[:] = np.where(rng.random(n) < 0.5, np.float32(-0.0), rng.standard_normal(n).astype(np.float32))
                outs = []
                for lib in (sim, fold):
                    tab, tab2 = table_for(lib, mv, M, 8, sb), table_for(lib, np.array([2.5], np.float32), 4, 8, 1)
                    ys = [aligned(n) for _ in range(5)]
                    assert lib.fp8fq_bn_act_quant_f32(P(x), P(ys[0]), P(one), P(negzero), N * C, hw, C, act, 0, P(tab), M, 8,
                                                      sb, None) == 0
                    assert lib.fp8fq_bn_act_quant_nhwc_f32(P(x), P(ys[1]), P(one), P(negzero), N * hw, C, act, 0, P(tab), M,
                                                           8, sb, None) == 0
                    assert lib.fp8fq_add_act_quant_f32(P(x), P(r), P(ys[2]), n, act, P(tab), M, 8, sb, None) == 0
                    assert lib.fp8fq_bn_quant_add_act_quant_f32(P(x), P(r), P(ys[3]), P(one), P(negzero), N * C, hw, C, act,
                                                                0, P(tab2), 4, 8, 1, P(tab), M, 8, sb, None) == 0
                    assert lib.fp8fq_bn_quant_add_act_quant_nhwc_f32(P(x), P(r), P(ys[4]), P(one), P(negzero), N * hw, C,
                                                                     act, 0, P(tab2), 4, 8, 1, P(tab), M, 8, sb, None) == 0
                    outs.append(ys)
                for k, (ya, yb) in enumerate(zip(*outs)):
                    assert same_values(ya, yb), (mvv, M, sb, act, k)    # bit for bit (signed zeros included); NaN == NaN
                    compared += 1
    assert compared == 8 * 5 * 3 * 5


# ---------------------------------------------------------------------------------------------------------------------
# K2a: min/max + estimator update rules (+ fused set_quant_range / prologue)
# ---------------------------------------------------------------------------------------------------------------------
def workspace(sim):
    return np.zeros(sim.fp8fq_minmax_workspace_bytes() // 4, np.int32)


def np_minmax(x):
    """torch.min / torch.max semantics: NaN propagates (numpy's min / max do the same)."""
    with np.errstate(invalid="ignore"):
        return np.float32(np.min(x)), np.float32(np.max(x))


def est_rule(mode, init, cur, new, momentum):
    if not init or mode == EST_CURRENT:
        return new
    if mode == EST_ALL:
        return tuple(np.float32(f(c, v)) if not (np.isnan(c) or np.isnan(v)) else np.float32(np.nan)
                     for f, c, v in ((min, cur[0], new[0]), (max, cur[1], new[1])))
    w_new, w_old = np.float32(1.0 - momentum), np.float32(momentum)     # range_estimators.py:121-123
    return tuple(np.float32(np.float32(w_new * v) + np.float32(w_old * c)) for c, v in zip(cur, new))


@pytest.mark.parametrize("n", [1, 3, 4, 1000, 4099, 70001, 150001])
def test_minmax_per_tensor_two_stage_reduction_and_update_rules(sim, n):
    """minmax_tensor_kernel: grid-stride 128-bit loads, warp shuffles, block reduce, last-CTA finish; the three
    estimator update rules (range_estimators.py:61-125); NaN propagation; misaligned base; the workspace's ticket
    counter is left at zero."""
    rng = np.random.default_rng(n)
    ws = workspace(sim)
    for off in (0, 1):
        for with_nan in (False, True):
            for mode in (EST_CURRENT, EST_ALL, EST_RUNNING):
                cur = None
                cmin, cmax = aligned(1), aligned(1)
                for call in range(3):
                    x = aligned(n, offset_elems=off)
                    x[:] = rand(rng, n, specials=False) * (call + 1)
                    if with_nan and call == 1:
                        x[rng.integers(n)] = np.nan
                    assert sim.fp8fq_minmax_f32(P(x), n, 1, n, P(cmin), P(cmax), mode, int(call > 0), 0.9, P(ws), None) == 0
                    cur = est_rule(mode, call > 0, cur, np_minmax(x), 0.9)
                    assert same_bits(np.array([cmin[0], cmax[0]]), np.array(cur, np.float32)), (n, off, with_nan, mode, call)
                    assert ws[0] == 0


def test_minmax_per_channel_rows_and_fused_prologue(sim, host_emul):
    """minmax_rows_kernel (CTA per row) incl. the fused set_quant_range + table build (fp8fq_estimate_prepare_f32 ==
    fp8fq_minmax_f32 + fp8fq_set_range_prepare_f32), per-channel and per-tensor."""
    rng = np.random.default_rng(5)
    ws = workspace(sim)
    for C, inner in ((1, 5000), (7, 1), (64, 147), (5, 576), (3, 4097), (1000, 12)):
        for M in (5, 3):
            x = aligned(C * inner)
            x[:] = rand(rng, C * inner, specials=False)
            if C > 1:
                x.reshape(C, inner)[C // 2, inner // 2] = np.nan
            cmin, cmax, mv = aligned(C), aligned(C), aligned(C)
            stride = sim.fp8fq_table_stride(M, 8, 1)
            tab = aligned(stride * C)
            assert sim.fp8fq_estimate_prepare_f32(P(x), C * inner, C, inner, P(cmin), P(cmax), EST_CURRENT, 0, 0.9, P(mv),
                                                  M, 8, 1, P(tab), P(ws), None) == 0
            with np.errstate(invalid="ignore"):
                rmin, rmax = x.reshape(C, inner).min(1), x.reshape(C, inner).max(1)
            assert same_bits(cmin, rmin) and same_bits(cmax, rmax)
            with np.errstate(invalid="ignore"):
                want_mv = np.abs(np.maximum(np.abs(rmin), rmax)).astype(np.float32)
            assert same_bits(mv, want_mv)
            tab2 = np.zeros_like(tab)
            host_emul.emul_prepare(P(np.ascontiguousarray(mv)), L(C), F(M), 8, 1, P(tab2))
            assert same_bits(tab, tab2), (C, inner, M)
            # second call in "all" mode keeps the running extremes
            x2 = aligned(C * inner)
            x2[:] = x * 0.5
            assert sim.fp8fq_minmax_f32(P(x2), C * inner, C, inner, P(cmin), P(cmax), EST_ALL, 1, 0.9, P(ws), None) == 0
            with np.errstate(invalid="ignore"):       # np.minimum / np.maximum propagate NaN like torch.min / torch.max
                want_min = np.minimum(rmin, x2.reshape(C, inner).min(1))
                want_max = np.maximum(rmax, x2.reshape(C, inner).max(1))
            assert same_bits(cmin, want_min) and same_bits(cmax, want_max)


@pytest.mark.parametrize("mode", [0, 1])
def test_bn_act_estimate_prepare_statistics_without_materialising(sim, ref, host_emul, mode):
    """minmax_bn_act_kernel: min / max of act(bn(x)) straight from x, NCHW rows and channel-innermost, then estimator
    update + set_quant_range + table; unsupported shapes answer FP8FQ_ERR_UNSUPPORTED (caller composes)."""
    rng = np.random.default_rng(8 + mode)
    ws = workspace(sim)
    cases = [(0, (2, 64, 28 * 28)), (0, (3, 128, 8 * 8)), (0, (1, 2, 9000)), (0, (3, 16, 8 * 8)), (0, (2, 24, 45)),
             (1, (300, 64)), (1, (196, 96)), (1, (49, 1280)), (1, (100, 24)), (1, (40, 30))]
    for nhwc, shape in cases:
        if nhwc:
            pixels, C = shape
            hw, outer, n = 1, pixels, pixels * C
            supported = C % 4 == 0
        else:
            N, C, hw = shape
            outer, n = N * C, N * C * hw
            supported = hw % 4 == 0 and 2 + 4095 // hw <= C
        p0, p1 = bn_params(sim, rng, C, mode)
        for act in (ACT_RELU, ACT_RELU6, ACT_NONE):
            x = aligned(n)
            x[:] = rand(rng, n, specials=False)
            cmin, cmax, mv = aligned(1), aligned(1), aligned(1)
            stride = sim.fp8fq_table_stride(5, 8, 1)
            tab = aligned(stride)
            code = sim.fp8fq_bn_act_estimate_prepare_f32(P(x), outer, hw, C, nhwc, P(p0), P(p1), mode, act, P(cmin), P(cmax),
                                                         EST_CURRENT, 0, 0.9, P(mv), 5, 8, 1, P(tab), P(ws), None)
            if not supported:
                assert code == -2, (nhwc, shape)
                continue
            assert code == 0, (nhwc, shape)
            v = ref_bn_act(ref, x, hw, C, nhwc, mode, p0, p1, act)
            lo, hi = np_minmax(v)
            assert same_bits(np.array([cmin[0], cmax[0]]), np.array([lo, hi])), (nhwc, shape, act)
            assert same_bits(mv, np.array([abs(max(abs(lo), hi))], np.float32))
            tab2 = np.zeros_like(tab)
            host_emul.emul_prepare(P(np.ascontiguousarray(mv)), L(1), F(5), 8, 1, P(tab2))
            assert same_bits(tab, tab2)
            assert ws[0] == 0
            # statistics only (maxval_out / table NULL), running update on top of the previous state
            x2 = aligned(n)
            x2[:] = x * 1.5
            assert sim.fp8fq_bn_act_estimate_prepare_f32(P(x2), outer, hw, C, nhwc, P(p0), P(p1), mode, act, P(cmin),
                                                         P(cmax), EST_RUNNING, 1, 0.9, None, 0.0, 0, 0, None, P(ws),
                                                         None) == 0
            v2 = ref_bn_act(ref, x2, hw, C, nhwc, mode, p0, p1, act)
            want = est_rule(EST_RUNNING, True, (lo, hi), np_minmax(v2), 0.9)
            assert same_bits(np.array([cmin[0], cmax[0]]), np.array(want, np.float32))


# ---------------------------------------------------------------------------------------------------------------------
# K2b: MSE grid (FP_MSE_Estimator's candidate loop, range_estimators.py:337-347)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,inner,G", [(1, 9000, 111), (1, 777, 111), (5, 576, 40), (3, 1025, 37), (2, 1100, 1030)])
def test_mse_grid_kernel_equals_candidate_loop(sim, ref, C, inner, G):
    """mses[m, g, c] += mean((x - Q(x; grid[g, c], M_m))^2) for all candidates in one sweep: register-resident slices
    (16 or 4 elements per thread), candidate tables staged in shared memory a group at a time, [warps][G] partials,
    double atomics; the mantissa sweep M = 1..6 covers K <= 3 and K > 3, and accumulation across calls."""
    rng = np.random.default_rng(C * 1000 + G)
    n = C * inner
    x = aligned(n)
    x[:] = rand(rng, n, specials=False)
    absmax = np.abs(x.reshape(C, inner)).max(1)
    grid = np.ascontiguousarray(np.linspace(0.1 * absmax, 1.2 * absmax, G, dtype=np.float32))     # [G, C]
    mbits = [5.0] if G > 200 else [float(m) for m in range(1, 7)]
    arr = (F * len(mbits))(*mbits)
    tf = sim.fp8fq_mse_table_floats(arr, len(mbits), 8, 1, G, C)
    assert tf > 0
    scratch = np.zeros(tf // 2 + 2, np.float64).view(np.float32)          # 8-byte aligned
    mses = np.zeros((len(mbits), G, C), np.float32)
    for rep in range(2):
        assert sim.fp8fq_mse_grid_f32(P(x), n, C, inner, P(grid), G, arr, len(mbits), 8, 1, P(mses), P(scratch), None) == 0
    want = np.zeros((len(mbits), G, C), np.float64)
    for mi, M in enumerate(mbits):
        for g in range(G):
            y = ref_quant(ref, x, grid[g], M, per_channel=True)[0]
            d = (x - y).astype(np.float32).astype(np.float64).reshape(C, inner)
            want[mi, g] = (d * d).mean(1)
    np.testing.assert_allclose(mses, 2 * want, rtol=2e-5, atol=1e-12)
    # same selection as the reference's argmin over candidates
    assert np.array_equal(mses[-1].argmin(0), (2 * want[-1]).astype(np.float32).argmin(0))


# ---------------------------------------------------------------------------------------------------------------------
# STE backward (autograd through fp8_quantizer.py:112-132)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,sb,pc", [(5, 1, False), (4, 1, True), (3, 0, False), (2, 1, True), (5, 0, True)])
def test_backward_kernel_gradients(sim, ref, M, sb, pc):
    """grad_x = ((g * s) / s) * clamp weight with autograd's roundings; acc[2c] the clipping term and acc[2c+1] the scale
    term of d/dmaxval (include/fp8fq.h); exact clamp ties get weight 1/2; NaN inputs poison their gradients."""
    rng = np.random.default_rng(M * 10 + sb)
    for C, inner in (((6, 4096 + 8), (3, 77)) if pc else ((1, 40000), (1, 13))):
        for off in (0, 1):
            n = C * inner
            x, g = aligned(n, offset_elems=off), aligned(n, offset_elems=off)
            x[:] = rand(rng, n, scale=2.0, specials=False)
            g[:] = rng.standard_normal(n).astype(np.float32)
            mv = (np.full(C, 2.25, np.float32) if C == 1 else (1.0 + rng.random(C) * 2).astype(np.float32))
            xv = x.reshape(C, inner)
            xv[:, 0] = mv                   # exact ties with the clamp bounds
            xv[:, 1 % inner] = -mv if sb else 0.0
            if inner > 4:
                xv[0, 3] = np.nan
            tab = table_for(sim, mv, M, 8, sb)
            gx = aligned(n, offset_elems=off)
            acc = np.zeros(2 * C, np.float64)
            assert sim.fp8fq_fake_quant_backward_f32(P(g), P(x), P(gx), P(tab), n, C, inner, M, 8, sb, P(acc), None) == 0
            # reference, element by element from the documented formula
            y, e, _ = ref_quant(ref, x, mv, M, 8, sb, per_channel=True)
            K = max(1, 2 ** (8 - sb - M) - 1)
            stride, KP = tab.size // C, (K + 2) & ~1
            tabv = tab.reshape(C, stride)
            with np.errstate(invalid="ignore"):
                ei = np.nan_to_num(e.reshape(C, inner), nan=1).astype(np.int64)
            s = np.take_along_axis(tabv[:, 8 + KP:8 + KP + 2 * (K + 1):2], ei, axis=1)
            gv, yv = g.reshape(C, inner), y.reshape(C, inner)
            hi = mv[:, None]
            lo = -hi if sb else np.zeros_like(hi)
            with np.errstate(invalid="ignore", over="ignore"):
                wx = np.where(xv < lo, 0.0, np.where(xv == lo, 0.5, 1.0)).astype(np.float32)
                tt = np.maximum(xv, lo)
                wt = np.where(tt > hi, 0.0, np.where(tt == hi, 0.5, 1.0)).astype(np.float32)
                clip = ((wx - 1.0) if sb else np.zeros_like(wx)) * wt + (1.0 - wt)
                xc = np.minimum(tt, hi)
                want_gx = ((gv * s).astype(np.float32) / s).astype(np.float32) * (wx * wt)
                want_gx = np.where(np.isnan(xv), np.float32(np.nan), want_gx).astype(np.float32)
                a1 = np.where(np.isnan(xv), np.nan, gv.astype(np.float64) * clip)
                a2 = gv.astype(np.float64) * (yv.astype(np.float64) - xc.astype(np.float64))
            assert same_bits(gx.reshape(C, inner), want_gx), (M, sb, C, inner, off)
            accv = acc.reshape(C, 2)
            for c in range(C):
                if np.isnan(xv[c]).any():
                    assert np.isnan(accv[c]).all()
                else:
                    np.testing.assert_allclose(accv[c, 0], a1[c].sum(), rtol=1e-4, atol=1e-4)
                    np.testing.assert_allclose(accv[c, 1], a2[c].sum(), rtol=1e-4, atol=1e-4)


# ---------------------------------------------------------------------------------------------------------------------
# INT uniform quantisers (uniform_quantizers.py:107-164, 224-246, 303-314)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("symmetric", [0, 1])
def test_uniform_quantisers_equal_serial_emulation(sim, host_emul, symmetric):
    rng = np.random.default_rng(40 + symmetric)
    for n_bits in (8, 4, 2):
        for C, inner in ((1, 30001), (1, 5), (9, 147), (4, 1025)):
            x = aligned(C * inner)
            x[:] = rand(rng, C * inner)
            fin = np.where(np.isfinite(x), x, 0).reshape(C, inner)
            xmin, xmax = fin.min(1).astype(np.float32), fin.max(1).astype(np.float32)
            if symmetric and n_bits == 4:
                xmin = np.abs(xmin)                      # unsigned range
            delta, zf, sg = aligned(C), aligned(C), aligned(1)
            tab = aligned(sim.fp8fq_uniform_table_floats(C))
            assert sim.fp8fq_uniform_prepare_f32(P(xmin), P(xmax), C, n_bits, symmetric, 0, 1e-8, P(delta), P(zf), P(sg),
                                                 P(tab), None) == 0
            delta2, tab2 = np.zeros(C, np.float32), np.zeros_like(tab)
            assert host_emul.emul_uq_prepare(P(xmin), P(xmax), L(C), n_bits, symmetric, F(1e-8), P(delta2), P(tab2)) == 0
            assert same_bits(delta, delta2) and same_bits(tab, tab2)
            assert sg[0] == (1.0 if np.minimum(xmin, 0).min() < 0 else 0.0)
            y, y2 = aligned(C * inner), np.zeros(C * inner, np.float32)
            assert sim.fp8fq_uniform_quant_f32(P(x), P(y), P(tab), C * inner, C, inner, None) == 0
            host_emul.emul_uq_quant(P(x), P(y2), P(tab2), L(C), L(inner))
            assert same_bits(y, y2), (symmetric, n_bits, C, inner)


# ---------------------------------------------------------------------------------------------------------------------
# data formats either side of the path
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,C,H,W,k", [(2, 3, 16, 12, 7), (1, 1, 8, 8, 3), (3, 4, 10, 14, 5), (1, 3, 224, 224, 7)])
def test_space_to_depth_gather(sim, N, C, H, W, k):
    """y[n, Y, X, c*4 + p*2 + q] = xpad[n, c, 2Y + p, 2X + q] with `pad` zero rows / columns; channels >= 4C zero."""
    rng = np.random.default_rng(N * 100 + k)
    pad, a = k // 2, (k + 1) // 2
    hs, ws = H // 2 + a - 1, W // 2 + a - 1
    x = aligned(N * C * H * W)
    x[:] = rng.standard_normal(x.size).astype(np.float32)
    y = aligned(N * hs * ws * 16)
    y[:] = 7.0
    assert sim.fp8fq_space_to_depth2_nhwc_f32(P(x), P(y), N, C, H, W, pad, hs, ws, None) == 0
    xp = np.zeros((N, C, 2 * hs, 2 * ws), np.float32)
    xv = x.reshape(N, C, H, W)
    hh, wwid = min(H, 2 * hs - pad), min(W, 2 * ws - pad)
    xp[:, :, pad:pad + hh, pad:pad + wwid] = xv[:, :, :hh, :wwid]
    want = np.zeros((N, hs, ws, 16), np.float32)
    for c in range(C):
        for p in range(2):
            for q in range(2):
                want[..., c * 4 + p * 2 + q] = xp[:, c, p::2, q::2]
    assert same_bits(y.reshape(N, hs, ws, 16), want)


@pytest.mark.parametrize("N,H,W,C,k,s,p", [(2, 12, 10, 8, 3, 2, 1), (1, 7, 7, 4, 2, 2, 0), (2, 9, 11, 12, 3, 1, 1),
                                           (1, 112, 112, 64, 3, 2, 1), (3, 5, 4, 16, 5, 3, 2)])
def test_max_pool_channel_innermost(sim, N, H, W, C, k, s, p):
    """ATen's selection rule (val > max || isnan(val), row-major window order, -inf padding): NaN propagates."""
    rng = np.random.default_rng(H * W + C)
    x = aligned(N * H * W * C)
    x[:] = rng.standard_normal(x.size).astype(np.float32)
    x[rng.choice(x.size, size=5, replace=False)] = np.nan
    ho, wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    y = aligned(N * ho * wo * C)
    assert sim.fp8fq_max_pool2d_nhwc_f32(P(x), P(y), N, H, W, C, k, k, s, s, p, p, None) == 0
    xv = np.full((N, H + 2 * p, W + 2 * p, C), -np.inf, np.float32)
    xv[:, p:p + H, p:p + W] = x.reshape(N, H, W, C)
    want = np.full((N, ho, wo, C), -np.inf, np.float32)
    with np.errstate(invalid="ignore"):
        for i in range(k):
            for j in range(k):
                v = xv[:, i:i + s * ho:s, j:j + s * wo:s][:, :ho, :wo]
                want = np.where((v > want) | np.isnan(v), v, want)
    assert same_bits(y.reshape(N, ho, wo, C), want)


# ---------------------------------------------------------------------------------------------------------------------
# host-buffer entry point: chunked H2D -> quantise -> D2H pipeline (memcpy stands in for the copies)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pinned", [0, 1])
def test_host_entry_point_chunking(sim, ref, pinned):
    """fp8fq_fake_quant_host_f32 splits the tensor into 8 Mi-element chunks (whole rows when per-channel, each chunk
    seeing its own slice of the channel tables) over three rotating staging buffers; pageable and page-locked callers."""
    sim.fp8fq_sim_report_pinned(pinned)
    try:
        rng = np.random.default_rng(50 + pinned)
        # per tensor across chunks; per channel with whole rows per chunk; per channel with rows LONGER than a chunk
        # (cut into pieces that each use their row's table); device 1: every device has its own pipeline
        for C, inner, device in ((1, (8 << 20) + 4099, 0), (5, 3 << 20, 0), (2, (8 << 20) + 1029, 1)):
            n = C * inner
            x = rng.standard_normal(n).astype(np.float32)
            mv = np.abs(x.reshape(C, inner)).max(1).astype(np.float32)
            y = np.zeros(n, np.float32)
            assert sim.fp8fq_fake_quant_host_f32(P(x), P(y), P(mv), n, C, inner, 5, 8, 1, device) == 0
            assert same_bits(y, ref_quant(ref, x, mv, 5, per_channel=True)[0]), (C, inner, pinned)
    finally:
        sim.fp8fq_sim_report_pinned(0)


# ---------------------------------------------------------------------------------------------------------------------
# degenerate ranges and the format limits, through the kernels
# ---------------------------------------------------------------------------------------------------------------------
def test_degenerate_ranges_through_the_kernels(sim, ref):
    """maxval 0 (whole channel NaN), inf, NaN, denormal, tiny and huge: tables flagged irregular / reciprocal-unusable
    take the linear-scan and IEEE-division paths inside the row and stream kernels; results are the direct formula's."""
    rng = np.random.default_rng(77)
    mvs = np.array([0.0, np.inf, np.nan, 1e-45, 1e-39, 1.1754944e-38, 1e-30, 1e30, 3.0e38, 1.0, 2.0 ** -126 * 3],
                   np.float32)
    C, inner = mvs.size, 515
    for M, sb in ((5, 1), (4, 1), (2, 1), (7, 1), (1, 1), (3, 0), (8, 0)):
        x = aligned(C * inner)
        base = rand(rng, (C, inner))
        with np.errstate(invalid="ignore", over="ignore"):
            scale = np.where(np.isfinite(mvs) & (mvs > 0), mvs, np.float32(1.0))[:, None]
            x[:] = (base * scale).reshape(-1)
        tab = table_for(sim, mvs, M, 8, sb)
        y = aligned(C * inner)
        assert sim.fp8fq_fake_quant_f32(P(x), P(y), P(tab), C * inner, C, inner, M, 8, sb, None) == 0
        yr = ref_quant(ref, x, mvs, M, 8, sb, per_channel=True)[0]
        bad = (bits(y) != bits(yr)) & ~(np.isnan(y) & np.isnan(yr))
        assert not bad.any(), (M, sb, [int(c) for c in np.unique(np.nonzero(bad)[0] // inner)])
        assert np.all(np.isnan(y.reshape(C, inner)[[0, 2]]))         # maxval 0 and NaN: the whole channel is NaN
        # the same ranges one at a time through the per-tensor stream kernel
        for c in range(C):
            xt, yt = aligned(inner), aligned(inner)
            xt[:] = x.reshape(C, inner)[c]
            tc = table_for(sim, mvs[c:c + 1], M, 8, sb)
            assert sim.fp8fq_fake_quant_f32(P(xt), P(yt), P(tc), inner, 1, inner, M, 8, sb, None) == 0
            assert same_values(yt, yr.reshape(C, inner)[c]), (M, sb, c)


@pytest.mark.parametrize("nb,M,sb", [(4, 1, 1), (4, 2, 1), (6, 3, 1), (8, 1, 1), (16, 12, 1), (16, 9, 1), (12, 4, 1),
                                      (13, 12, 0), (2, 1, 1), (5, 5, 0)])
def test_format_limits_through_the_stream_kernel(sim, ref, nb, M, sb):
    """Run-time bit widths other than 8: up to E = 7 (127 exponent codes, the largest table) and M = 12."""
    rng = np.random.default_rng(nb * 100 + M)
    E = nb - sb - M
    assert 0 <= E <= 7
    n = 6001
    x = aligned(n)
    # spread the inputs over the format's whole exponent range
    x[:] = (rng.standard_normal(n) * np.exp2(rng.uniform(-min(2 ** E, 60), 2, n))).astype(np.float32)
    mv = np.array([3.7], np.float32)
    tab = table_for(sim, mv, M, nb, sb)
    y = aligned(n)
    assert sim.fp8fq_fake_quant_f32(P(x), P(y), P(tab), n, 1, n, M, nb, sb, None) == 0
    assert same_bits(y, ref_quant(ref, x, mv, M, nb, sb)[0]), (nb, M, sb)
    assert sim.fp8fq_fake_quant_f32(P(x), P(y), P(tab), n, 1, n, 1, 16, 1, None) == -2      # E = 14: unsupported


# ---------------------------------------------------------------------------------------------------------------------
# the kernels' code against the REAL reference's golden vectors
# ---------------------------------------------------------------------------------------------------------------------
def test_simulated_kernels_against_the_real_reference_golden_vectors(sim):
    """tests/golden/fp8_quantizer.npz was written by the real reference (ATen on the CPU: Sleef / glibc log2 and pow).
    The simulated kernels use the host libm for the prologue, so a scale-table entry can differ from ATen's by an ulp
    (the reference's own backends differ from each other the same way); an ulp in a scale moves
    ~1e-5 of the elements across a rounding tie.  Asserted per case: same exponent code and mantissa integer for all
    but <= 2e-3 of the elements, dequantised floats within 1e-5 relative, NaN / zero-range channels identical."""
    from conftest import load_golden

    g = load_golden("fp8_quantizer.npz")
    n_cases = int(g["num_cases"])
    assert n_cases == 98
    total = mism = 0
    for i in range(n_cases):
        name = f"c{i:03d}"
        M, sb, pc = [int(v) for v in g[name + "_meta"]]
        xs, y_ref, q_ref = g[name + "_x"], g[name + "_y"], g[name + "_q"]
        mv = np.ascontiguousarray(g[name + "_maxval"], np.float32).reshape(-1)
        C = mv.size if pc else 1
        n = xs.size
        x, y, codes = aligned(n), aligned(n), aligned(n, np.int32)
        x[:] = xs.reshape(-1)
        tab = table_for(sim, mv, M, 8, sb)
        assert sim.fp8fq_fake_quant_codes_f32(P(x), P(y), P(codes), P(tab), n, C, n // C, M, 8, sb, None) == 0
        y_ref, q_ref = y_ref.reshape(-1), q_ref.reshape(-1)
        nan_ref = np.isnan(y_ref)
        assert np.array_equal(np.isnan(y), nan_ref), name
        ok = ~nan_ref
        with np.errstate(invalid="ignore"):
            q_ours = (codes & 0xFFFF).astype(np.float32)
            diff = ok & (q_ours != np.abs(q_ref))
            rel = np.abs(y[ok].astype(np.float64) - y_ref[ok]) / np.maximum(np.abs(y_ref[ok].astype(np.float64)), 1e-30)
        # a differing mantissa integer is a tie resolved the other way (or the (e, 2^(M+1)) / (e+1, 2^M) double code)
        assert diff.sum() <= max(2, 2e-3 * n), (name, M, sb, pc, int(diff.sum()), n)
        same_q = ok & ~diff
        with np.errstate(invalid="ignore"):
            rel_same = np.abs(y[same_q].astype(np.float64) - y_ref[same_q]) / np.maximum(np.abs(y_ref[same_q].astype(np.float64)), 1e-30)
        assert rel_same.size == 0 or rel_same.max() < 1e-5, (name, rel_same.max())
        assert rel.size == 0 or np.all((rel < 2.0 ** -(M - 1)) | (y_ref[ok] == 0)), name   # never more than one code away
        total += int(ok.sum())
        mism += int(diff.sum())
    assert mism / total < 1e-4, (mism, total)


@pytest.mark.parametrize("M,sb", [(4, 1), (3, 1), (2, 1), (4, 0), (1, 1)])
@pytest.mark.parametrize("which", ["sdouble", "magic", "magicopts", "magicone"])
def test_exact_doubling_scale_tables_take_the_arithmetic_path_and_stay_bit_exact(built, ref, M, sb, which):
    """which = "magic": the DEFAULT build -- tables with FLAG_MAGIC run the element path in the scaled domain
    (quant_magic: add-and-subtract rounding of |xc| / s_1, no look-up at all), the others the look-up; both kinds must
    occur and both must reproduce the direct formula bit for bit.  which = "sdouble":
    (Build option FP8FQ_SDOUBLE, exercised on oracle/_build/libfp8fq_sim_fulltilecl.so, which is built with it.)
    K > 3 formats: when the prologue finds the reference's scales to be exact doublings of each other (FLAG_SDOUBLE,
    the usual case), the element path derives (s, 1/s) from the exponent code by integer arithmetic instead of loading
    them (lookup_scale_fast); otherwise it keeps the table look-up.  Both kinds of table must occur over a sweep of
    ranges and both must reproduce the direct formula bit for bit -- values at every code boundary, rounding ties, +-0,
    denormals, NaN / inf, per tensor (stream kernel) and per channel (row kernel)."""
    from fp8_quantization_b200._lib import SIGNATURES

    sim = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build",
                                   {"sdouble": "libfp8fq_sim_fulltilecl.so", "magic": "libfp8fq_sim.so",
                                    "magicopts": "libfp8fq_sim_magicopts.so", "magicone": "libfp8fq_sim_magicone.so"}[which]))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(sim, name)
        fn.restype, fn.argtypes = res, args
    rng = np.random.default_rng(40 + M)
    FLAG_SDOUBLE, H_FLAGS = (8 if which == "sdouble" else 16), 4
    kinds = {True: 0, False: 0}
    mvs = np.exp(rng.uniform(np.log(0.02), np.log(60.0), 48)).astype(np.float32)
    n = 4096 + 64
    for mvv in mvs:
        mv = np.array([mvv], np.float32)
        tab = table_for(sim, mv, M, 8, sb)
        dbl = bool(int(tab[H_FLAGS:H_FLAGS + 1].view(np.uint32)[0]) & FLAG_SDOUBLE)
        kinds[dbl] += 1
        x = aligned(n)
        x[:] = rng.standard_normal(n).astype(np.float32) * mvv * 0.6
        # every binade edge below maxval +- a few ulps, and half-way points: the code boundaries and rounding ties
        edges = mvv * np.float32(2.0) ** -np.arange(0, 40, dtype=np.float32)
        pts = np.concatenate([edges, np.nextafter(edges, np.float32(0)), np.nextafter(edges, np.float32(np.inf)), edges * 1.5,
                              edges * np.float32(0.75)]).astype(np.float32)
        x[:pts.size] = pts
        x[pts.size:2 * pts.size] = -pts
        x[2 * pts.size:2 * pts.size + 8] = [0.0, -0.0, np.nan, np.inf, -np.inf, 1e-38, -1e-45, 3e38]
        y = aligned(n)
        assert sim.fp8fq_fake_quant_f32(P(x), P(y), P(tab), n, 1, n, M, 8, sb, None) == 0
        assert same_bits(y, ref_quant(ref, x, mv, M, 8, sb)[0]), (M, sb, float(mvv), dbl)
    assert kinds[True] >= 10, kinds                      # the arithmetic path is the common one ...
    if M == 4 and sb == 1 and which == "sdouble":
        assert kinds[False] >= 1, kinds                  # ... and the look-up still has its cases
    if which != "sdouble":
        assert kinds[True] >= 44, kinds                  # one- and two-group scale tables both qualify
    # per channel: rows with and without the flag in one launch
    C, inner = 24, 260
    mv = mvs[:C].copy()
    xr = aligned(C * inner)
    xr[:] = (rng.standard_normal((C, inner)).astype(np.float32) * mv[:, None] * 0.6).reshape(-1)
    tabs = table_for(sim, mv, M, 8, sb)
    yr = aligned(C * inner)
    assert sim.fp8fq_fake_quant_f32(P(xr), P(yr), P(tabs), C * inner, C, inner, M, 8, sb, None) == 0
    assert same_bits(yr, ref_quant(ref, xr, mv, M, 8, sb, per_channel=True)[0])


@pytest.mark.parametrize("C", [1, 37])
def test_data_parallel_statistics_mode_and_finish_kernel(sim, host_emul, C):
    """SURVEY 8e on the kernels themselves: two "ranks" compute their shard's statistics in mode FP8FQ_EST_DP_STATS
    ([-min | max] written straight into the exchange buffer), the buffers are merged with an element-wise MAX (what the
    NCCL all-reduce does), and fp8fq_dp_finish_prepare_f32 applies the estimator rule, set_quant_range and builds the
    table -- equal, bit for bit, to the single-process fused calibration launch on the concatenated batch, for the three
    estimator rules over three calibration batches."""
    EST_DP_STATS = 3
    rng = np.random.default_rng(70 + C)
    ws = workspace(sim)
    inner = 1500
    for mode in (EST_CURRENT, EST_ALL, EST_RUNNING):
        dmin, dmax = aligned(C), aligned(C)          # data-parallel estimator state
        smin, smax = aligned(C), aligned(C)          # single-process estimator state
        for call in range(3):
            shards = [rand(rng, (C, inner), specials=False) * (call + 1 + r) for r in range(2)]
            if call == 2:
                shards[1][C // 2, 7] = np.nan      # one rank sees a NaN in one channel: the global range of it is NaN
            packed = []
            for xs in shards:
                xs = np.ascontiguousarray(xs)
                buf = aligned(3 * C)               # [-min | max | NaN flag]
                assert sim.fp8fq_minmax_f32(P(xs), xs.size, C, inner, P(buf[:C]), P(buf[C:]), EST_DP_STATS, 0, 0.9, P(ws),
                                            None) == 0
                isnan = np.isnan(xs).any(1)
                assert np.array_equal(buf[2 * C:], isnan.astype(np.float32))
                assert same_bits(buf[:C][~isnan], -xs.min(1)[~isnan]) and same_bits(buf[C:2 * C][~isnan], xs.max(1)[~isnan])
                assert np.all(np.isneginf(buf[:C][isnan])) and np.all(np.isneginf(buf[C:2 * C][isnan]))
                packed.append(buf)
            merged = aligned(3 * C)
            merged[:] = np.fmax(packed[0], packed[1])   # a MAX reduction that DROPS NaN, the worst case NCCL may be
            stride = sim.fp8fq_table_stride(5, 8, 1)
            mv_dp, tab_dp = aligned(C), aligned(stride * C)
            assert sim.fp8fq_dp_finish_prepare_f32(P(merged), C, P(dmin), P(dmax), mode, int(call > 0), 0.9, P(mv_dp), 5, 8, 1,
                                                   P(tab_dp), None) == 0
            # single process: the two shards of every channel side by side = the concatenated batch
            whole = np.ascontiguousarray(np.concatenate(shards, axis=1))
            mv_sp, tab_sp = aligned(C), aligned(stride * C)
            assert sim.fp8fq_estimate_prepare_f32(P(whole), whole.size, C, 2 * inner, P(smin), P(smax), mode, int(call > 0),
                                                  0.9, P(mv_sp), 5, 8, 1, P(tab_sp), P(ws), None) == 0
            assert same_bits(dmin, smin) and same_bits(dmax, smax), (mode, call)
            assert same_bits(mv_dp, mv_sp) and same_bits(tab_dp, tab_sp), (mode, call)


def test_peer_memory_exchange_of_the_calibration_statistics(sim):
    """fp8fq_estimate_prepare_p2p_f32 / fp8fq_bn_act_estimate_prepare_p2p_f32: the last CTA exchanges (-min, max, NaN flag)
    with the peers through their exchange buffers (64-bit words: value | epoch << 32) and finishes the calibration step
    itself.  Two "ranks" run one after the other here: rank 1's words are put into rank 0's buffer beforehand (what rank 1's
    kernel does concurrently on a real machine), rank 0's kernel then delivers its own words to both buffers, and rank
    1's kernel finds everything in place.  Both must end with the single-process result on the concatenated batch, over
    three calibration batches (slots rotate with the epoch), for the three estimator rules; a missing peer gives NaN."""
    rng = np.random.default_rng(90)
    ws = workspace(sim)
    world = 2
    words = sim.fp8fq_dp_exchange_words(world)
    assert words == 4 * world * 3 and sim.fp8fq_dp_exchange_words(17) < 0
    U64 = ctypes.c_uint64

    def word(v, epoch):
        return (epoch << 32) | int(np.float32(v).view(np.uint32))

    for mode in (EST_CURRENT, EST_ALL, EST_RUNNING):
        bufs = [np.zeros(words, np.uint64) for _ in range(world)]
        ptrs = np.array([b.ctypes.data for b in bufs], np.uint64)
        state = [(aligned(1), aligned(1)) for _ in range(world)]
        smin, smax = aligned(1), aligned(1)
        stride = sim.fp8fq_table_stride(5, 8, 1)
        for call in range(3):
            epoch = call + 1
            shards = [np.ascontiguousarray(rand(rng, 3000 + 40 * r, specials=False) * (call + 1 + r)) for r in range(world)]
            if call == 2:
                shards[1][11] = np.nan
            s1 = shards[1]
            nan1 = bool(np.isnan(s1).any())
            slot = (epoch % 4) * world * 3
            for j, v in enumerate((-np.inf if nan1 else -np.nanmin(s1), -np.inf if nan1 else np.nanmax(s1), 1.0 if nan1 else 0.0)):
                bufs[0][slot + 1 * 3 + j] = word(v, epoch)
            outs = []
            for r in range(world):
                mv, tab = aligned(1), aligned(stride)
                cmin, cmax = state[r]
                assert sim.fp8fq_estimate_prepare_p2p_f32(P(shards[r]), shards[r].size, P(cmin), P(cmax), mode, int(call > 0), 0.9,
                                                          P(mv), 5, 8, 1, P(tab), P(ws), ptrs.ctypes.data_as(ctypes.c_void_p), r,
                                                          world, epoch, None) == 0
                outs.append((mv.copy(), tab.copy(), cmin.copy(), cmax.copy()))
            whole = np.ascontiguousarray(np.concatenate(shards))
            mv_sp, tab_sp = aligned(1), aligned(stride)
            assert sim.fp8fq_estimate_prepare_f32(P(whole), whole.size, 1, whole.size, P(smin), P(smax), mode, int(call > 0), 0.9,
                                                  P(mv_sp), 5, 8, 1, P(tab_sp), P(ws), None) == 0
            for r in range(world):
                assert same_values(outs[r][0], mv_sp) and same_values(outs[r][1], tab_sp), (mode, call, r)
                assert same_values(outs[r][2], smin) and same_values(outs[r][3], smax), (mode, call, r)
    # a peer that never delivers: the range is NaN (on the device after a ~3 s timeout; the simulation reads once)
    bufs = [np.zeros(words, np.uint64) for _ in range(world)]
    ptrs = np.array([b.ctypes.data for b in bufs], np.uint64)
    x = np.ascontiguousarray(rand(rng, 1000, specials=False))
    mv, tab, cmin, cmax = aligned(1), aligned(sim.fp8fq_table_stride(5, 8, 1)), aligned(1), aligned(1)
    assert sim.fp8fq_estimate_prepare_p2p_f32(P(x), x.size, P(cmin), P(cmax), EST_CURRENT, 0, 0.9, P(mv), 5, 8, 1, P(tab), P(ws),
                                              ptrs.ctypes.data_as(ctypes.c_void_p), 0, world, 1, None) == 0
    assert np.isnan(mv[0]) and np.isnan(cmin[0])
    # argument checks: world < 2, rank out of range, epoch 0, too many ranks
    for rank, w, ep, want in ((0, 1, 1, -1), (2, 2, 1, -1), (0, 2, 0, -1), (0, 17, 1, -2)):
        assert sim.fp8fq_estimate_prepare_p2p_f32(P(x), x.size, P(cmin), P(cmax), EST_CURRENT, 0, 0.9, P(mv), 5, 8, 1, P(tab),
                                                  P(ws), ptrs.ctypes.data_as(ctypes.c_void_p), rank, w, ep, None) == want
