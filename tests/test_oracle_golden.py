"""CPU: the oracle restatement against the committed golden vectors (produced by the REAL reference,
tests/golden/make_golden.py).  Bit-exact where ATen takes the same libm path as when the fixtures
were generated (AVX-512 Sleef, one thread); otherwise canonical codes + 1 ulp."""
import numpy as np
import pytest
import torch

from conftest import bits, load_golden, ulp_diff
from oracle import fp8_oracle as O
from oracle.reference_loader import reference_available

STRICT = torch.backends.cpu.get_cpu_capability() == "AVX512"


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def test_fake_quant_matches_reference_golden():
    g = load_golden("fp8_quantizer.npz")
    n = int(g["num_cases"])
    assert n == 98
    for i in range(n):
        name = f"c{i:03d}"
        M, sb, pc = [int(v) for v in g[name + "_meta"]]
        x = torch.from_numpy(g[name + "_x"])
        y_ref = torch.from_numpy(g[name + "_y"])
        maxval = torch.from_numpy(g[name + "_maxval"])
        y, e, q = O.fake_quant(x, 8, maxval, torch.Tensor([float(M)]), sb, return_codes=True)
        if STRICT:
            same = (bits(y) == bits(y_ref)) | (torch.isnan(y) & torch.isnan(y_ref))
            assert bool(same.all()), f"case {i} (M={M}, sign={sb}, per_channel={pc})"
            assert np.array_equal(q.numpy(), g[name + "_q"], equal_nan=True)
        else:
            assert int(ulp_diff(y, y_ref).max()) <= 1


def test_default_maxval_and_ideal_grid():
    g = load_golden("fp8_format.npz")
    for M in range(1, 8):
        assert np.float32(O.default_maxval(8, M)) == g[f"M{M}"][0]
    # with the default (integer-bias) maxval the quantiser's outputs lie on the
    # reference's own enumerated grid exactly for M in {2, 3, 5} and within 1e-6 rel for the others
    torch.manual_seed(0)
    for M in (2, 3, 4, 5):
        grid = g[f"grid_M{M}"]
        mv = float(g[f"M{M}"][0])
        x = torch.randn(4096) * mv / 3
        y = O.fake_quant(x, 8, torch.Tensor([mv]), torch.Tensor([float(M)]), 1).double().numpy()
        idx = np.clip(np.searchsorted(grid, y), 1, len(grid) - 1)
        near = np.where(np.abs(grid[idx] - y) < np.abs(grid[idx - 1] - y), grid[idx], grid[idx - 1])
        rel = np.abs(near - y) / np.maximum(np.abs(y), 1e-30)
        rel[y == 0] = 0
        assert rel.max() < (1e-12 if M in (2, 3, 5) else 1e-6), (M, rel.max())


def test_estimators_match_reference_golden():
    g = load_golden("estimators.npz")
    for name, cls, kw in (("current", O.OracleCurrentMinMax, {}), ("all", O.OracleAllMinMax, {}),
                          ("running", O.OracleRunningMinMax, {"momentum": 0.9})):
        for pc in (False, True):
            key = f"{name}_{'pc' if pc else 'pt'}"
            est = cls(per_channel=pc, **kw)
            for i, x in enumerate(g[key + "_x"]):
                mn, mx = est(torch.from_numpy(x))
                assert np.array_equal(mn.reshape(-1).numpy(), g[key + "_min"][i], equal_nan=True)
                assert np.array_equal(mx.reshape(-1).numpy(), g[key + "_max"][i], equal_nan=True)


def test_mse_estimator_matches_reference_golden():
    g = load_golden("mse_estimator.npz")
    for key, pc, include in (("pt_sweep", False, True), ("pt_fixed", False, False), ("pc_sweep", True, True),
                             ("pc_fixed", True, False)):
        x = torch.from_numpy(g[key + "_x"])
        q = O.OracleFPQuantizer(8, per_channel=pc, mantissa_bits=4, set_maxval=True, mse_include_mantissa_bits=include)
        est = O.OracleFPMSE(per_channel=pc, quantizer=q)
        mn, mx = est(x)
        assert float(q.mantissa_bits) == float(g[key + "_best_m"])
        assert np.array_equal(est.search_grid.numpy(), g[key + "_grid"])
        if STRICT:
            assert np.array_equal(est.mses.numpy(), g[key + "_mses"])
            assert np.array_equal(mx.numpy(), g[key + "_xmax"])
        else:
            assert np.allclose(est.mses.numpy(), g[key + "_mses"], rtol=1e-4)


def test_line_search_matches_reference_golden():
    g = load_golden("line_search.npz")
    for key in ("pt", "pc", "pt_onesided"):
        ncand, M, pc = [int(v) for v in g[key + "_meta"]]
        q = O.OracleFPQuantizer(8, mantissa_bits=M, set_maxval=True)
        est = O.OracleLineSearch(quantizer=q, per_channel=bool(pc), num_candidates=ncand)
        mn, mx = est(torch.from_numpy(g[key + "_x"]))
        if STRICT:
            assert np.array_equal(est.loss_array, g[key + "_loss"])
            assert np.array_equal(mx.numpy(), g[key + "_xmax"]) and np.array_equal(mn.numpy(), g[key + "_xmin"])
        else:
            assert np.allclose(est.loss_array[:, 1:], g[key + "_loss"][:, 1:], rtol=1e-4)


def test_uniform_quantizers_match_reference_golden():
    g = load_golden("uniform_quantizers.npz")
    for i in range(int(g["num_cases"])):
        n = f"u{i:02d}"
        sym, nb, pc = [int(v) for v in g[n + "_meta"]]
        q = (O.OracleSymmetricUniform if sym else O.OracleAsymmetricUniform)(nb, per_channel=bool(pc))
        mn, mx = torch.from_numpy(g[n + "_min"]), torch.from_numpy(g[n + "_max"])
        if not pc:
            mn, mx = mn.reshape(()), mx.reshape(())
        q.set_quant_range(mn, mx)
        y = q(torch.from_numpy(g[n + "_x"]))
        yr = torch.from_numpy(g[n + "_y"])
        assert bool(((bits(y) == bits(yr)) | (torch.isnan(y) & torch.isnan(yr))).all()), n  # IEEE-exact everywhere


@pytest.mark.skipif(not reference_available(), reason="reference checkout not present")
def test_oracle_equals_live_reference():
    """In the build container the oracle is also compared with the live reference on fresh inputs."""
    from oracle.reference_loader import load_reference

    R = load_reference()
    torch.manual_seed(123)
    for M in (2, 3, 4, 5, 6):
        for pc in (False, True):
            x = torch.randn(32, 96)
            q = R.FPQuantizer(8, per_channel=pc, mantissa_bits=M, set_maxval=True)
            mn, mx = O.minmax(x, pc)
            q.set_quant_range(mn, mx)
            y_ref = q(x)
            y = O.fake_quant(x, 8, q.maxval, q.mantissa_bits, 1)
            assert torch.equal(bits(y), bits(y_ref))
            oq = O.OracleFPQuantizer(8, per_channel=pc, mantissa_bits=M, set_maxval=True)
            oq.set_quant_range(mn, mx)
            assert torch.equal(oq.maxval, q.maxval)
            assert torch.equal(bits(oq(x)), bits(y_ref))


# ---- oracle-composed networks (oracle/fp8_oracle_models.py) against the REAL reference's ranges and logits -------------
def _check_composed(golden, digits, sites, logits):
    names = [str(n) for n in golden["names"]]
    mv = sites.maxvals()
    assert list(mv.keys()) != [] and set(mv.keys()) == set(names)
    for i, n in enumerate(names):
        ref = golden[f"maxval_{i:0{digits}d}"]
        ours = mv[n].numpy()
        if STRICT:
            assert np.array_equal(ours, ref), n
        else:
            np.testing.assert_allclose(ours, ref, rtol=1e-5, err_msg=n)
    if STRICT:
        assert np.array_equal(logits.numpy(), golden["logits"])
    else:
        np.testing.assert_allclose(logits.numpy(), golden["logits"], rtol=1e-3, atol=1e-3)


def test_oracle_composed_resnet18_equals_reference_golden():
    """BASELINE config 2 in miniature: the oracle-composed quantised ResNet-18 (plain F.conv2d / F.batch_norm /
    oracle quantisers and estimators, no product module) calibrated on the golden batch reproduces all 50 ranges and the
    logits of the real reference's QuantizedResNet bit for bit.  This pins the checker of tests/test_gpu_model_parity.py."""
    from torchvision.models import resnet18

    from oracle import fp8_oracle_models as OM

    g = load_golden("resnet18_m5.npz")
    torch.manual_seed(10)
    net = resnet18().eval()
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(10))
    S = OM.OracleSites(5)
    OM.resnet_forward(S, net, x)
    S.fix_ranges()
    _check_composed(g, 2, S, OM.resnet_forward(S, net, x))
    assert len(S.sites) == 50


def test_oracle_composed_mobilenetv2_equals_reference_golden():
    """BASELINE config 3 in miniature: same for QuantizedMobileNetV2 (M=4): 123 quantisers (116 used + 7 that keep
    their default range), logits bit for bit."""
    from fp8_quantization_b200 import workloads
    from oracle import fp8_oracle_models as OM

    g = load_golden("mobilenetv2_m4.npz")
    torch.manual_seed(10)
    net = workloads.MobileNetV2().eval()
    sd = net.state_dict()
    sums = [int(sd[k].float().contiguous().view(torch.int32).to(torch.int64).sum()) for k in sd.keys()]
    assert sums == list(g["w_checksums"])   # same fp32 network as the reference built under this seed
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(10))
    S = OM.OracleSites(4)
    OM.mobilenetv2_forward(S, net, x)
    S.fix_ranges()
    _check_composed(g, 3, S, OM.mobilenetv2_forward(S, net, x))
    assert len(S.sites) == 123


# ---- golden set at the survey's sizes, digests written by the real reference ------------------
def _survey_module():
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden_survey_sizes.py")
    spec = importlib.util.spec_from_file_location("make_golden_survey_sizes", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not STRICT, reason="digests need ATen's AVX-512 (Sleef) path, as when the fixtures were written")
def test_oracle_reproduces_the_reference_digests_at_survey_sizes():
    """252 quantiser cases (M 1..7 x sign x sigma {1e-3, 1, 1e3} x {per tensor 2^20, per channel with inner 1 / 9 / 27 /
    147 / 576}): the oracle's outputs, exponent codes and mantissa integers hash to what the REAL reference produced."""
    S = _survey_module()
    g = load_golden("survey_sizes.npz")
    n = 0
    for M in range(1, 8):
        for sb in (0, 1):
            for sigma in S.SIGMAS:
                for inner in (0,) + S.INNERS:
                    x, pc = S.quantizer_case(M, sb, sigma, inner)
                    key = f"q_M{M}_s{sb}_sig{S.SIGMAS.index(sigma)}_in{inner}"
                    assert tuple(g[key + "_shape"]) == tuple(x.shape)
                    q = O.OracleFPQuantizer(8, per_channel=pc, mantissa_bits=M, set_maxval=True)
                    q.sign_bits = sb
                    mn, mx = O.minmax(x, pc)
                    q.set_quant_range(mn * 0.9, mx * 0.9)
                    assert S.digest(q.maxval) == str(g[key + "_maxval_digest"])
                    y, e, qq = O.fake_quant(x, 8, q.maxval, q.mantissa_bits, sb, return_codes=True)
                    assert [S.digest(y), S.digest(e), S.digest(qq)] == [str(v) for v in g[key]], key
                    n += 1
    assert n == 252


@pytest.mark.skipif(not STRICT, reason="bit-exact MSE tables need the AVX-512 path")
def test_oracle_mse_estimator_at_survey_sizes_weight_case():
    """FP_MSE_Estimator with the mantissa sweep on the per-channel [128,64,3,3] weight (the [8,64,56,56] activation case
    takes minutes on one CPU thread; it is covered on the GPU, tests/test_gpu_parity.py)."""
    S = _survey_module()
    g = load_golden("survey_sizes.npz")
    x, pc = S.mse_case("weight_128x64x3x3")
    oq = O.OracleFPQuantizer(8, per_channel=pc, mantissa_bits=4, set_maxval=True, mse_include_mantissa_bits=True)
    oest = O.OracleFPMSE(per_channel=pc, quantizer=oq)
    _, omx = oest(x)
    assert np.array_equal(oest.mses.numpy(), g["mse_weight_128x64x3x3_mses"])
    assert float(oq.mantissa_bits) == float(g["mse_weight_128x64x3x3_best_m"])
    assert np.array_equal(omx.float().numpy(), g["mse_weight_128x64x3x3_xmax"])
