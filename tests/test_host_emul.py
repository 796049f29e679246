"""CPU: the product's table / bucket-lookup / tie-guard algorithm (csrc/fp8fq_core.h compiled for the
host, tests/host_emul/emul.cpp) returns the same BITS as evaluating the reference formula directly
per element (oracle/fp8_oracle_c.c), given the same libm.  This isolates the algorithmic claim --
"no log2 / pow / division per element, identical results" -- from libm differences between backends.
Also cross-checks the C restatement against the torch oracle."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import fp8_oracle as O

FP = ctypes.POINTER(ctypes.c_float)
IP = ctypes.POINTER(ctypes.c_int32)


def P(a):
    return a.ctypes.data_as(FP)


def run_pair(oracle_c, emul, x, maxval, mb, nb, sb, force=0, codes=True):
    """force: 0 = the paths the kernels take (with ``codes=False`` FLAG_MAGIC tables run the scaled-domain path
    quant_magic; the code-plane variant keeps the look-up), 1 = linear threshold scan, 2 = look-up path."""
    C, inner = x.shape
    y0, e0, q0 = np.empty_like(x), np.empty_like(x), np.empty_like(x)
    oracle_c.oracle_c_fake_quant(P(x), P(y0), P(e0), P(q0), P(maxval), ctypes.c_int64(C), ctypes.c_int64(inner),
                                 ctypes.c_float(mb), nb, sb)
    st = emul.emul_table_stride(ctypes.c_float(mb), nb, sb)
    assert st > 0
    tab = np.zeros(st * C, np.float32)
    assert emul.emul_prepare(P(maxval), ctypes.c_int64(C), ctypes.c_float(mb), nb, sb, P(tab)) == 0
    y1 = np.empty_like(x)
    cd = np.empty(x.shape, np.int32)
    slow = ctypes.c_int64(0)
    assert emul.emul_fake_quant(P(x), P(y1), cd.ctypes.data_as(IP) if codes else None, P(tab), ctypes.c_int64(C),
                                ctypes.c_int64(inner), ctypes.c_float(mb), nb, sb, force, ctypes.byref(slow)) == 0
    return y0, e0, q0, y1, cd, slow.value, tab


@pytest.mark.parametrize("M", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("sb", [0, 1])
def test_table_algorithm_equals_direct_formula(oracle_c, host_emul, M, sb):
    if M == 8 and sb == 1:
        pytest.skip("M is clamped to n_bits - sign_bits")
    rng = np.random.default_rng(10 + M)
    for scale in (1e-3, 1.0, 37.0, 1e4):
        for force in (0, 1):
            C, inner = 8, 20000
            x = (rng.standard_normal((C, inner)) * scale).astype(np.float32)
            mv = (np.abs(x).max(1) * rng.uniform(0.3, 1.1, size=C)).astype(np.float32)
            y0, e0, q0, y1, cd, slow, _ = run_pair(oracle_c, host_emul, x, mv, float(M), 8, sb, force)
            same = (y0.view(np.int32) == y1.view(np.int32)) | (np.isnan(y0) & np.isnan(y1))
            assert same.all(), f"float bits differ: M={M} sign={sb} scale={scale} force_irregular={force}"
            e1, q1 = (cd >> 16) & 0x7FFF, cd & 0xFFFF
            ok = ((e1 == e0) & (q1 == np.abs(q0))) | np.isnan(y0)
            assert ok.all(), "codes differ"
            if force == 0:   # the path the stream / row kernels take: scaled-domain rounding for FLAG_MAGIC tables
                y2 = run_pair(oracle_c, host_emul, x, mv, float(M), 8, sb, 0, codes=False)[3]
                same = (y0.view(np.int32) == y2.view(np.int32)) | (np.isnan(y0) & np.isnan(y2))
                assert same.all(), f"scaled-domain path differs: M={M} sign={sb} scale={scale}"


def test_edge_inputs_and_degenerate_ranges(oracle_c, host_emul):
    specials = np.array([0.0, -0.0, np.nan, np.inf, -np.inf, 1e-30, -1e-30, 1e-40, -1e-40, 1.4e-45, 3.0, -3.0, 2.1152,
                         1.0, 2.0, 0.5, 0.25], dtype=np.float32)
    for M in (1, 2, 3, 4, 5, 6, 7):
        for sb in (0, 1):
            for mvv in (2.1152, 1.0, 0.0, np.inf, np.nan, 1e-37, 1e30, 3e38):
                x = np.tile(specials, (1, 4)).astype(np.float32)
                x = np.concatenate([x, x * np.float32(mvv if np.isfinite(mvv) else 1.0)], axis=1)
                mv = np.array([mvv], np.float32)
                for codes in (True, False):
                    y0, e0, q0, y1, cd, _, _ = run_pair(oracle_c, host_emul, x, mv, float(M), 8, sb, codes=codes)
                    same = (y0.view(np.int32) == y1.view(np.int32)) | (np.isnan(y0) & np.isnan(y1))
                    assert same.all(), (M, sb, mvv, codes, x[~same], y0[~same], y1[~same])


def test_binade_edges_and_ties_exhaustive_neighbourhood(oracle_c, host_emul):
    """Every float within +-64 ulps of every switching point and of every rounding tie of one binade."""
    for M, mvv in ((5, 2.1152), (4, 6.3), (3, 240.0), (2, 1.37)):
        mv = np.array([mvv], np.float32)
        st = host_emul.emul_table_stride(ctypes.c_float(M), 8, 1)
        tab = np.zeros(st, np.float32)
        host_emul.emul_prepare(P(mv), ctypes.c_int64(1), ctypes.c_float(M), 8, 1, P(tab))
        K = int(tab[5:6].view(np.int32)[0]) & 0xFF
        kp = (K + 2) & ~1
        thr = tab[8:8 + K + 1]
        sr = tab[8 + kp:8 + kp + 2 * (K + 1)].reshape(K + 1, 2)
        pts = [t for t in thr[1:K] if np.isfinite(t)]
        for e in range(1, K + 1):
            s = sr[e, 0]
            pts += [np.float32((q + 0.5) * s) for q in (0, 1, 2, 2**M - 1, 2**M, 2 ** (M + 1) - 1)]
        xs = []
        for p in pts:
            pi = np.float32(p).view(np.int32).astype(np.int64)
            xs.append((pi + np.arange(-64, 65)).astype(np.int32).view(np.float32))
        x = np.concatenate(xs)[None, :].astype(np.float32)
        x = np.concatenate([x, -x], axis=1)
        for codes in (True, False):
            y0, e0, q0, y1, cd, _, _ = run_pair(oracle_c, host_emul, x, mv, float(M), 8, 1, codes=codes)
            assert (y0.view(np.int32) == y1.view(np.int32)).all(), (M, mvv, codes)


@pytest.mark.parametrize("nb,M", [(8, 1), (8, 2), (8, 3), (8, 4), (8, 5), (8, 6), (8, 7), (6, 2), (12, 6), (16, 10), (16, 12), (10, 3)])
def test_scaled_domain_path_is_the_common_one_and_bit_exact(oracle_c, host_emul, nb, M):
    """FLAG_MAGIC (csrc/fp8fq_core.h prep_finish / quant_magic): over a sweep of ranges (nearly) all tables qualify -- one
    group of exactly doubling scales, or two groups with the switching point between them -- and on those
    the add-and-subtract rounding in the scaled domain returns the reference's bits for random values, for every float
    within +-200 ulps of every code boundary (where the two classifications of the binade differ) and of rounding ties
    in every binade, for zeros of both signs, denormals, NaN and infinities -- signed and unsigned."""
    FLAG_MAGIC = 16
    rng = np.random.default_rng(1000 * nb + M)
    mvs = np.exp(rng.uniform(np.log(1e-3), np.log(3e3), 64)).astype(np.float32)
    magic = two = 0
    for sb in (1, 0):
        for mvv in mvs[: 64 if sb else 16]:
            mv = np.array([mvv], np.float32)
            st = host_emul.emul_table_stride(ctypes.c_float(M), nb, sb)
            tab = np.zeros(st, np.float32)
            assert host_emul.emul_prepare(P(mv), ctypes.c_int64(1), ctypes.c_float(M), nb, sb, P(tab)) == 0
            is_magic = bool(host_emul.emul_table_flags(P(tab), ctypes.c_int64(0), ctypes.c_float(M), nb, sb) & FLAG_MAGIC)
            magic += is_magic and sb == 1
            two += host_emul.emul_table_break(P(tab), ctypes.c_int64(0), ctypes.c_float(M), nb, sb) > 0
            K = int(tab[5:6].view(np.int32)[0]) & 0xFF
            kp = (K + 2) & ~1
            thr = tab[8:8 + K + 1]
            sr = tab[8 + kp:8 + kp + 2 * (K + 1)].reshape(K + 1, 2)
            pts = [t for t in thr[1:K] if np.isfinite(t)]
            for e in sorted(set([1, 2, 3, K // 2, K - 1, K]) & set(range(1, K + 1))):
                s = sr[e, 0]
                pts += [np.float32((q + 0.5) * s) for q in (0, 1, 2, 3, 2 ** M - 1, 2 ** M, 2 ** M + 1, 2 ** (M + 1) - 2,
                                                              2 ** (M + 1) - 1)]
                pts += [np.float32(q * s) for q in (1, 2 ** M, 2 ** (M + 1))]
            pts = [p for p in pts if np.isfinite(p) and p > 0]
            xs = [(np.float32(p).view(np.int32).astype(np.int64) + np.arange(-200, 201)).astype(np.int32).view(np.float32)
                  for p in pts]
            xs.append((rng.standard_normal(20000) * mvv * 0.5).astype(np.float32))
            xs.append(np.exp(rng.uniform(np.log(1e-6), 0.1, 20000)).astype(np.float32) * mvv)
            xs.append(np.array([0.0, -0.0, np.nan, np.inf, -np.inf, 1e-38, -1e-38, 1e-45, -1e-45, 3e38, -3e38], np.float32))
            x = np.concatenate(xs)[None, :].astype(np.float32)
            x = np.ascontiguousarray(np.concatenate([x, -x], axis=1))
            y0, _, _, y1, _, _, _ = run_pair(oracle_c, host_emul, x, mv, float(M), nb, sb, codes=False)
            same = (y0.view(np.int32) == y1.view(np.int32)) | (np.isnan(y0) & np.isnan(y1))
            assert same.all(), (nb, M, sb, float(mvv), is_magic, x[~same][:8], y0[~same][:8], y1[~same][:8])
    assert magic >= (56 if M <= 8 else 32), magic   # of 64 signed tables (the rest: three scale groups at |bias| < 1)
    if (nb, M) in ((8, 4), (8, 3), (8, 5)):
        assert two >= 1, two    # two-group scale tables (e.g. E3M4 with maxval in [2, 8)) occur and were just checked


def test_c_restatement_vs_torch_oracle(oracle_c, host_emul):
    """glibc (C restatement) and Sleef (ATen) differ by <= 1 ulp in log2/pow; canonical codes still agree
    except where a 1-ulp scale difference moves an element across a rounding tie (rare)."""
    torch.manual_seed(5)
    for M in (2, 3, 4, 5):
        x = torch.randn(4, 1 << 14)
        mv = x.abs().max(1)[0] * 0.9
        y0, e0, q0, *_ = run_pair(oracle_c, host_emul, x.numpy(), mv.numpy(), float(M), 8, 1)
        y, e, q = O.fake_quant(x, 8, mv, torch.Tensor([float(M)]), 1, return_codes=True)
        a = O.canonical_codes(torch.from_numpy(y0), torch.from_numpy(e0), torch.from_numpy(q0), M)
        b = O.canonical_codes(y, e, q, M)
        bad = (a[0] != b[0]) | (a[1] != b[1]) | (a[2] != b[2])
        assert bad.float().mean() < 1e-4
        rel = ((torch.from_numpy(y0) - y).abs() / y.abs().clamp_min(1e-30))[~bad]
        assert rel.max() < 2.5e-7


def test_uniform_quantizer_algorithm_equals_reference_golden(host_emul):
    """the INT uniform quantiser path of csrc/fp8fq_core.h (set_quant_range + reciprocal-multiply with tie
    guard + saturation shortcut) reproduces the REAL reference's outputs bit for bit -- its arithmetic is IEEE-exact
    (div, round, clamp, mul), so there is no libm caveat here."""
    from conftest import load_golden

    g = load_golden("uniform_quantizers.npz")
    for i in range(int(g["num_cases"])):
        n = f"u{i:02d}"
        sym, nb, pc = [int(v) for v in g[n + "_meta"]]
        x = g[n + "_x"]
        x2 = np.ascontiguousarray(x.reshape(x.shape[0], -1) if pc else x.reshape(1, -1))
        C, inner = x2.shape
        mn, mx = np.ascontiguousarray(g[n + "_min"]), np.ascontiguousarray(g[n + "_max"])
        tab, d = np.zeros(C * 8, np.float32), np.zeros(C, np.float32)
        assert host_emul.emul_uq_prepare(P(mn), P(mx), ctypes.c_int64(C), nb, sym, ctypes.c_float(1e-8), P(d), P(tab)) == 0
        y = np.empty_like(x2)
        host_emul.emul_uq_quant(P(x2), P(y), P(tab), ctypes.c_int64(C), ctypes.c_int64(inner))
        yr = g[n + "_y"].reshape(C, inner)
        assert np.array_equal(d, g[n + "_delta"]), n
        assert ((y.view(np.int32) == yr.view(np.int32)) | (np.isnan(y) & np.isnan(yr))).all(), n
