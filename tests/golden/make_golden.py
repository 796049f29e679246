"""Generates tests/golden/*.npz by running the REAL reference (/root/reference) on the CPU.

Run in the build container only (the reference checkout does not exist on the GPU box):

    python tests/golden/make_golden.py

For every case the oracle restatement (oracle/fp8_oracle.py) is run on the same input and must
match the reference BIT FOR BIT before anything is written -- this is what pins the oracle.
The reference ships no tests or golden vectors of its own, so these files are
the only pin there is; they are produced by the reference's own code, not by ours.

Determinism notes: torch.set_num_threads(1) and sizes that are multiples of 32 keep every element
on ATen's vectorised (Sleef) path -- ATen's scalar loop tails call glibc instead, whose pow/log2
differ from Sleef's in the last bit for ~2 % of arguments (measured).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import fp8_oracle as O  # noqa: E402
from oracle.reference_loader import load_reference, load_reference_models  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def bits(t):
    return t.contiguous().view(torch.int32)


def same_bits(a, b):
    nan = torch.isnan(a) & torch.isnan(b)
    return bool(((bits(a) == bits(b)) | nan).all())


def edge_vector(maxval: float, M: int, bias: float):
    """+-0, NaN, +-inf, tiny, denormals, binade edges 2^(k-bias) +- 1ulp, ties, +-maxval(1 +- ulp)."""
    vals = [0.0, -0.0, float("nan"), float("inf"), float("-inf"), 1e-30, -1e-30, 1e-40, -1e-40, 1.4e-45,
            maxval, -maxval, np.nextafter(np.float32(maxval), np.float32(np.inf)),
            np.nextafter(np.float32(maxval), np.float32(0)), -np.nextafter(np.float32(maxval), np.float32(0))]
    for k in range(-3, 6):
        edge = np.float32(2.0 ** (k - bias))
        vals += [edge, np.nextafter(edge, np.float32(0)), np.nextafter(edge, np.float32(np.inf)), -edge]
        s = np.float32(2.0 ** (k - M - bias))
        for q in (0.5, 1.5, 2.5, 2**M + 0.5, 2 ** (M + 1) - 0.5):
            vals += [np.float32(q) * s, -np.float32(q) * s]
    v = torch.tensor(np.array(vals, dtype=np.float32))
    pad = (-len(v)) % 32
    return torch.cat([v, torch.zeros(pad)])


def gen_quantizer(R):
    cases = {}
    g = torch.Generator().manual_seed(10)  # README.md:64 uses --seed 10
    idx = 0
    for M in range(1, 8):
        for sb in (0, 1):
            for sigma in (1e-3, 1.0, 1e3):
                for per_channel in (False, True):
                    shape = (16, 128) if per_channel else (2048,)
                    x = torch.randn(shape, generator=g) * sigma
                    if per_channel:
                        x = x * torch.linspace(0.25, 4.0, shape[0]).view(-1, 1)
                        x[3] = 0.0  # all-zero channel -> maxval 0 -> whole channel NaN
                    q = R.FPQuantizer(8, per_channel=per_channel, mantissa_bits=M, set_maxval=True)
                    q.sign_bits = sb
                    mn, mx = O.minmax(x, per_channel)
                    q.set_quant_range(mn * 0.9, mx * 0.9)  # some clipping
                    y_ref = q(x)
                    y, e, qq = O.fake_quant(x, 8, q.maxval, q.mantissa_bits, sb, return_codes=True)
                    assert same_bits(y_ref, y), ("oracle != reference", M, sb, sigma, per_channel)
                    name = f"c{idx:03d}"
                    cases[name + "_x"] = x.numpy()
                    cases[name + "_y"] = y_ref.numpy()
                    cases[name + "_e"] = e.to(torch.int16).numpy() if not torch.isnan(e).any() else e.numpy()
                    cases[name + "_q"] = qq.numpy()
                    cases[name + "_maxval"] = q.maxval.numpy()
                    cases[name + "_meta"] = np.array([M, sb, int(per_channel)], dtype=np.int32)
                    idx += 1
            # edge vector, per-tensor, maxval not a power of two
            mv = torch.Tensor([2.1152])
            mb = torch.Tensor([float(M)])
            Mt, Et = O.mantissa_exponent_split(mb, 8, sb)
            bias = float(2**Et - torch.log2(mv) + torch.log2(2 - 2 ** (-Mt)) - 1)
            x = edge_vector(2.1152, M, bias)
            q = R.FPQuantizer(8, mantissa_bits=M, maxval=2.1152)
            q.sign_bits = sb
            y_ref = q(x)
            y, e, qq = O.fake_quant(x, 8, q.maxval, q.mantissa_bits, sb, return_codes=True)
            assert same_bits(y_ref, y), ("oracle != reference (edge)", M, sb)
            name = f"c{idx:03d}"
            cases[name + "_x"] = x.numpy()
            cases[name + "_y"] = y_ref.numpy()
            cases[name + "_e"] = e.numpy()
            cases[name + "_q"] = qq.numpy()
            cases[name + "_maxval"] = q.maxval.numpy()
            cases[name + "_meta"] = np.array([M, sb, 0], dtype=np.int32)
            idx += 1
    cases["num_cases"] = np.array(idx)
    np.savez_compressed(os.path.join(OUT, "fp8_quantizer.npz"), **cases)
    print("fp8_quantizer.npz:", idx, "cases")


def gen_default_maxval(R):
    out = {}
    for M in range(1, 8):
        q = R.FPQuantizer(8, mantissa_bits=M, maxval=None)
        out[f"M{M}"] = q.maxval.numpy()
        assert np.float32(O.default_maxval(8, M)) == np.float32(float(q.maxval))
        # independent known-answer sets of the reference for the ideal grid (fp8_quantizer.py:13-41,82-88)
        E = 7 - M
        grid = R.fp8_quantizer.generate_all_values_fp(8, E, 2 ** (E - 1)) if E >= 1 else None
        if grid is not None:
            out[f"grid_M{M}"] = np.asarray(grid, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "fp8_format.npz"), **out)
    print("fp8_format.npz")


def gen_estimators(R):
    out = {}
    g = torch.Generator().manual_seed(11)
    RE = R.range_estimators
    for name, cls, ocls, kw in (("current", RE.CurrentMinMaxEstimator, O.OracleCurrentMinMax, {}),
                                ("all", RE.AllMinMaxEstimator, O.OracleAllMinMax, {}),
                                ("running", RE.RunningMinMaxEstimator, O.OracleRunningMinMax, {"momentum": 0.9})):
        for pc in (False, True):
            est = cls(per_channel=pc, **kw)
            oest = ocls(per_channel=pc, **kw)
            xs = [torch.randn(12, 40, generator=g) * (1 + i) for i in range(4)]
            xs[2][5, 7] = float("nan") if name == "all" and not pc else xs[2][5, 7]
            mins, maxs = [], []
            for x in xs:
                mn, mx = est(x)
                omn, omx = oest(x)
                assert same_bits(mn.float().reshape(-1), omn.float().reshape(-1))
                assert same_bits(mx.float().reshape(-1), omx.float().reshape(-1))
                mins.append(mn.reshape(-1).numpy().copy())
                maxs.append(mx.reshape(-1).numpy().copy())
            key = f"{name}_{'pc' if pc else 'pt'}"
            out[key + "_x"] = np.stack([x.numpy() for x in xs])
            out[key + "_min"] = np.stack(mins)
            out[key + "_max"] = np.stack(maxs)
    np.savez_compressed(os.path.join(OUT, "estimators.npz"), **out)
    print("estimators.npz")


def gen_mse(R):
    out = {}
    g = torch.Generator().manual_seed(12)
    RE = R.range_estimators
    for key, shape, pc, include in (("pt_sweep", (4, 8, 14, 14), False, True), ("pt_fixed", (4, 8, 14, 14), False, False),
                                    ("pc_sweep", (16, 8, 3, 3), True, True), ("pc_fixed", (16, 8, 3, 3), True, False)):
        x = torch.randn(shape, generator=g)
        if not pc:
            x = torch.relu(x) + 0.05 * torch.randn(shape, generator=g)
        q = R.FPQuantizer(8, per_channel=pc, mantissa_bits=4, set_maxval=True, mse_include_mantissa_bits=include)
        est = RE.FP_MSE_Estimator(per_channel=pc, quantizer=q)
        oq = O.OracleFPQuantizer(8, per_channel=pc, mantissa_bits=4, set_maxval=True, mse_include_mantissa_bits=include)
        oest = O.OracleFPMSE(per_channel=pc, quantizer=oq)
        mn, mx = est(x)
        omn, omx = oest(x)
        assert same_bits(mx.float(), omx.float()) and same_bits(est.mses, oest.mses)
        assert float(q.mantissa_bits) == float(oq.mantissa_bits)
        out[key + "_x"] = x.numpy()
        out[key + "_mses"] = est.mses.numpy()
        out[key + "_grid"] = est.search_grid.numpy()
        out[key + "_xmin"] = mn.numpy()
        out[key + "_xmax"] = mx.numpy()
        out[key + "_best_m"] = np.array(float(q.mantissa_bits))
    np.savez_compressed(os.path.join(OUT, "mse_estimator.npz"), **out)
    print("mse_estimator.npz")


def gen_modules(R):
    """QuantLinear (config 1 of BASELINE.json: Linear(1024,1024) E2M5 per-channel) and BNQConv."""
    out = {}
    torch.manual_seed(10)
    aq = R.autoquant_utils
    RE = R.range_estimators
    qp = dict(method=R.FPQuantizer, n_bits=8, per_channel_weights=True, weight_range_method=RE.CurrentMinMaxEstimator,
              act_range_method=RE.AllMinMaxEstimator,
              fp8_kwargs=dict(mantissa_bits=5, set_maxval=True, maxval=None, mse_include_mantissa_bits=False))
    lin = aq.QuantLinear(1024, 1024, **qp)
    x = torch.randn(32, 1024)
    lin.quantized()
    with torch.no_grad():
        y_cal = lin(x)  # calibration state
        lin.fix_ranges()
        y_fix = lin(x)
        wq = lin.weight_quantizer(lin.weight)
    out["lin_w"] = lin.weight.detach().numpy()
    out["lin_b"] = lin.bias.detach().numpy()
    out["lin_x"] = x.numpy()
    out["lin_wq"] = wq.numpy()
    out["lin_y"] = y_fix.numpy()
    out["lin_w_maxval"] = lin.weight_quantizer.quantizer.maxval.numpy()
    out["lin_a_maxval"] = lin.activation_quantizer.quantizer.maxval.numpy()
    assert torch.equal(y_cal, y_fix)

    conv = aq.BNQConv(8, 16, 3, padding=1, activation=torch.nn.ReLU(), **qp)
    conv.running_mean.normal_()
    conv.running_var.uniform_(0.5, 2.0)
    conv.gamma.data.normal_(1.0, 0.2)
    conv.beta.data.normal_(0.0, 0.2)
    conv.eval()
    conv.quantized()
    xc = torch.randn(4, 8, 12, 12)
    with torch.no_grad():
        yc = conv(xc)
        conv.fix_ranges()
        yc2 = conv(xc)
    assert torch.equal(yc, yc2)
    for k, v in (("w", conv.weight), ("mean", conv.running_mean), ("var", conv.running_var), ("gamma", conv.gamma),
                 ("beta", conv.beta)):
        out["conv_" + k] = v.detach().numpy()
    out["conv_x"] = xc.numpy()
    out["conv_y"] = yc.numpy()
    out["conv_a_maxval"] = conv.activation_quantizer.quantizer.maxval.numpy()
    np.savez_compressed(os.path.join(OUT, "modules.npz"), **out)
    print("modules.npz")


def gen_resnet18(R):
    """Reference QuantizedResNet(resnet18()) under seed 10, README quant params, M=5: calibrate on one
    batch, fix ranges, record the ranges of every quantiser and the logits."""
    Rm = load_reference_models()
    from torchvision.models import resnet18

    RE = R.range_estimators
    torch.manual_seed(10)
    net = resnet18()
    qp = dict(method=R.FPQuantizer, act_method=R.FPQuantizer, n_bits=8, n_bits_act=None, per_channel_weights=True,
              quant_setup="all", weight_range_method=RE.CurrentMinMaxEstimator, weight_range_options={},
              act_range_method=RE.AllMinMaxEstimator, act_range_options={}, quantize_input=False,
              fp8_kwargs=dict(maxval=None, mantissa_bits=5, set_maxval=True, learn_maxval=False,
                              learn_mantissa_bits=False, mse_include_mantissa_bits=False, allow_unsigned=False))
    model = Rm.resnet_quantized.QuantizedResNet(net, **qp)
    model.eval()
    g = torch.Generator().manual_seed(10)
    x = torch.randn(2, 3, 224, 224, generator=g)
    model.set_quant_state(True, True)
    with torch.no_grad():
        model(x)
        model.fix_ranges()
        logits = model(x)
    maxvals = []
    names = []
    for name, mod in model.named_modules():
        if isinstance(mod, R.FPQuantizer):
            names.append(name)
            maxvals.append(mod.maxval.reshape(-1).numpy().copy())
    out = {"logits": logits.numpy(), "x_seed": np.array(10), "names": np.array(names)}
    for i, mv in enumerate(maxvals):
        out[f"maxval_{i:02d}"] = mv
    np.savez_compressed(os.path.join(OUT, "resnet18_m5.npz"), **out)
    print("resnet18_m5.npz", len(names), "quantizers")


def gen_uniform(R):
    """Asymmetric / SymmetricUniformQuantizer (uniform_quantizers.py) on seeded tensors; the
    arithmetic is IEEE-exact (div, round, clamp, mul), so these vectors must be reproduced bit for bit everywhere."""
    import quantization.quantizers.uniform_quantizers as uq

    out = {}
    g = torch.Generator().manual_seed(13)
    idx = 0
    for cls, ocls, tag in ((uq.AsymmetricUniformQuantizer, O.OracleAsymmetricUniform, "asym"),
                           (uq.SymmetricUniformQuantizer, O.OracleSymmetricUniform, "sym")):
        for n_bits in (8, 4, 2):
            for pc in (False, True):
                for kind in ("signed", "unsigned"):
                    shape = (16, 96) if pc else (1536,)
                    x = torch.randn(shape, generator=g) * 3
                    if kind == "unsigned":
                        x = x.abs()
                    if pc:
                        x = x * torch.linspace(0.1, 5.0, shape[0]).view(-1, 1)
                    x.view(-1)[:6] = torch.tensor([0.0, -0.0, float("inf"), float("-inf"), float("nan"), 1e-30])
                    q = cls(n_bits=n_bits, per_channel=pc)
                    oq = ocls(n_bits, per_channel=pc)
                    mn, mx = O.minmax(torch.nan_to_num(x, nan=0.0, posinf=9.0, neginf=-9.0 if kind == "signed" else 0.0), pc)
                    q.set_quant_range(mn * 0.8, mx * 0.8)
                    oq.set_quant_range(mn * 0.8, mx * 0.8)
                    y = q(x)
                    yo = oq(x)
                    assert same_bits(y, yo), (tag, n_bits, pc, kind)
                    name = f"u{idx:02d}"
                    out[name + "_x"] = x.numpy()
                    out[name + "_y"] = y.numpy()
                    out[name + "_min"] = (mn * 0.8).reshape(-1).numpy()
                    out[name + "_max"] = (mx * 0.8).reshape(-1).numpy()
                    out[name + "_delta"] = q.delta.reshape(-1).numpy()
                    out[name + "_meta"] = np.array([tag == "sym", n_bits, pc], dtype=np.int32)
                    idx += 1
    out["num_cases"] = np.array(idx)
    np.savez_compressed(os.path.join(OUT, "uniform_quantizers.npz"), **out)
    print("uniform_quantizers.npz:", idx, "cases")


def gen_line_search(R):
    """LineSearchEstimator's 1-D grid search (range_estimators.py:236-256) with an FP quantiser,
    as compute_quant_error.py uses it (there with 1000 candidates on 5 M samples)."""
    RE = R.range_estimators
    out = {}
    g = torch.Generator().manual_seed(14)
    for key, shape, pc, ncand, M in (("pt", (4096,), False, 200, 4), ("pc", (6, 512), True, 120, 3),
                                     ("pt_onesided", (2048,), False, 150, 5)):
        x = torch.randn(shape, generator=g)
        if key == "pt_onesided":
            x = x.abs()
        q = R.FPQuantizer(8, mantissa_bits=M, set_maxval=True)
        est = RE.LineSearchEstimator(quantizer=q, per_channel=pc, num_candidates=ncand)
        oq = O.OracleFPQuantizer(8, mantissa_bits=M, set_maxval=True)
        oest = O.OracleLineSearch(quantizer=oq, per_channel=pc, num_candidates=ncand)
        mn, mx = est(x)
        omn, omx = oest(x)
        assert np.array_equal(est.loss_array, oest.loss_array) and torch.equal(mx, omx) and torch.equal(mn, omn)
        out[key + "_x"] = x.numpy()
        out[key + "_loss"] = est.loss_array
        out[key + "_xmin"] = mn.numpy()
        out[key + "_xmax"] = mx.numpy()
        out[key + "_meta"] = np.array([ncand, M, int(pc)])
    np.savez_compressed(os.path.join(OUT, "line_search.npz"), **out)
    print("line_search.npz")


def gen_backward(R):
    """gradients of the reference's FPQuantizer (learnable maxval / mantissa_bits,
    fp8_quantizer.py:242-254) w.r.t. x, maxval and mantissa_bits for L = sum(y * w)."""
    out = {}
    g = torch.Generator().manual_seed(15)
    idx = 0
    for M in (2, 3, 4, 5):
        for pc in (False, True):
            for sb in (1, 0):
                shape = (8, 192) if pc else (1536,)
                x = (torch.randn(shape, generator=g) * 1.5)
                if sb == 0:
                    x = x.abs() - 0.1
                w = torch.randn(shape, generator=g)
                q = R.FPQuantizer(8, per_channel=pc, mantissa_bits=M, set_maxval=True)
                q.sign_bits = sb
                mn, mx = O.minmax(x, pc)
                q.set_quant_range(mn * 0.7, mx * 0.7)
                x.view(-1)[:3] = torch.stack([q.maxval.reshape(-1)[0], -q.maxval.reshape(-1)[0] * sb, torch.tensor(0.0)])
                q.learn_maxval()
                q.learn_mantissa_bits()
                xr = x.clone().requires_grad_(True)
                y = q(xr)
                (y * w).sum().backward()
                mv = q.maxval.detach().clone().requires_grad_(True)
                mb = q.mantissa_bits.detach().clone().requires_grad_(True)
                xo = x.clone().requires_grad_(True)
                yo = O.fake_quant_ste(xo, 8, mv, mb, sb)
                (yo * w).sum().backward()
                assert same_bits(y.detach(), yo.detach()) and same_bits(xr.grad, xo.grad)
                assert same_bits(q.maxval.grad, mv.grad) and same_bits(q.mantissa_bits.grad, mb.grad)
                n = f"b{idx:02d}"
                out[n + "_x"] = x.numpy()
                out[n + "_w"] = w.numpy()
                out[n + "_maxval"] = q.maxval.detach().numpy()
                out[n + "_gx"] = xr.grad.numpy()
                out[n + "_gmaxval"] = q.maxval.grad.numpy()
                out[n + "_gmbits"] = q.mantissa_bits.grad.numpy()
                out[n + "_meta"] = np.array([M, sb, int(pc)])
                idx += 1
    out["num_cases"] = np.array(idx)
    np.savez_compressed(os.path.join(OUT, "backward.npz"), **out)
    print("backward.npz:", idx, "cases")


def gen_bn_reestimate(R):
    """the reference's reestimate_BN_stats (utils/qat_utils.py:45-90) on a small quantised
    conv-BN-ReLU-conv-BN stack with fixed ranges, 3 batches."""
    import importlib

    qat = importlib.import_module("utils.qat_utils") if False else None
    # utils/qat_utils.py imports the ImageNet data loaders (torchvision ok); import the function directly
    import utils.qat_utils as qat_utils

    aq = R.autoquant_utils
    RE = R.range_estimators
    torch.manual_seed(10)
    qp = dict(method=R.FPQuantizer, n_bits=8, per_channel_weights=True, weight_range_method=RE.CurrentMinMaxEstimator,
              act_range_method=RE.AllMinMaxEstimator,
              fp8_kwargs=dict(mantissa_bits=5, set_maxval=True, maxval=None, mse_include_mantissa_bits=False))
    seq = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1, bias=False), torch.nn.BatchNorm2d(8), torch.nn.ReLU(),
                              torch.nn.Conv2d(8, 6, 1, bias=False), torch.nn.BatchNorm2d(6))
    seq[1].running_mean.normal_()
    seq[1].running_var.uniform_(0.5, 2.0)
    seq[4].running_mean.normal_()
    seq[4].running_var.uniform_(0.5, 2.0)
    state = {k: v.clone() for k, v in seq.state_dict().items()}
    model = aq.quantize_model(seq, **qp)
    from quantization.base_quantized_model import QuantizedModel

    class Wrap(QuantizedModel):
        def __init__(self, f):
            super().__init__((1, 3, 16, 16))
            self.f = f

        def forward(self, x):
            return self.f(x)

    wm = Wrap(model)
    wm.eval()
    g = torch.Generator().manual_seed(11)
    xs = [torch.randn(8, 3, 16, 16, generator=g) for _ in range(3)]
    wm.set_quant_state(True, True)
    with torch.no_grad():
        wm(xs[0])
    wm.fix_ranges()
    qat_utils.reestimate_BN_stats(wm, [(x, None) for x in xs], num_batches=3)
    out = {"x": torch.stack(xs).numpy()}
    for k, v in state.items():
        out["init_" + k.replace(".", "_")] = v.numpy()
    for i in (0, 1):
        out[f"mean_{i}"] = model[i].running_mean.numpy()
        out[f"var_{i}"] = model[i].running_var.numpy()
    np.savez_compressed(os.path.join(OUT, "bn_reestimate.npz"), **out)
    print("bn_reestimate.npz")


def gen_mobilenetv2(R):
    """BASELINE config 3: reference QuantizedMobileNetV2(MobileNetV2()) under seed 10, README parameters with M=4:
    calibrate on one batch, fix ranges, record every quantiser's range and the logits."""
    Rm = load_reference_models()
    RE = R.range_estimators
    torch.manual_seed(10)
    net = Rm.mobilenet_v2.MobileNetV2()
    qp = dict(method=R.FPQuantizer, act_method=R.FPQuantizer, n_bits=8, n_bits_act=None, per_channel_weights=True,
              quant_setup="all", weight_range_method=RE.CurrentMinMaxEstimator, weight_range_options={},
              act_range_method=RE.AllMinMaxEstimator, act_range_options={}, quantize_input=False,
              fp8_kwargs=dict(maxval=None, mantissa_bits=4, set_maxval=True, learn_maxval=False,
                              learn_mantissa_bits=False, mse_include_mantissa_bits=False, allow_unsigned=False))
    model = Rm.mobilenet_v2_quantized.QuantizedMobileNetV2(net, **qp)
    model.eval()
    g = torch.Generator().manual_seed(10)
    x = torch.randn(2, 3, 224, 224, generator=g)
    model.set_quant_state(True, True)
    with torch.no_grad():
        model(x)
        model.fix_ranges()
        logits = model(x)
    names, maxvals = [], []
    for name, mod in model.named_modules():
        if isinstance(mod, R.FPQuantizer):
            names.append(name)
            maxvals.append(mod.maxval.reshape(-1).numpy().copy())
    out = {"logits": logits.numpy(), "names": np.array(names),
           "state_keys": np.array([k for k in net.state_dict().keys()])}
    # checksums of the fp32 parameters: the test rebuilds the network under the same seed with this package's
    # MobileNetV2 (same construction and initialisation order) and verifies it got bit-identical weights
    sd = net.state_dict()
    out["w_checksums"] = np.array([int(sd[k].float().contiguous().view(torch.int32).to(torch.int64).sum())
                                   for k in sd.keys()], dtype=np.int64)  # exact: sum of the fp32 bit patterns
    for i, mv in enumerate(maxvals):
        out[f"maxval_{i:03d}"] = mv
    np.savez_compressed(os.path.join(OUT, "mobilenetv2_m4.npz"), **out)
    print("mobilenetv2_m4.npz", len(names), "quantizers")


if __name__ == "__main__":
    torch.set_num_threads(1)
    R = load_reference()
    gen_quantizer(R)
    gen_default_maxval(R)
    gen_estimators(R)
    gen_mse(R)
    gen_modules(R)
    gen_resnet18(R)
    gen_uniform(R)
    gen_line_search(R)
    gen_backward(R)
    gen_bn_reestimate(R)
    if "--mobilenet" in sys.argv:
        gen_mobilenetv2(R)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
