"""GPU: the remaining BASELINE.json configurations as parity cases.
  config 3: MobileNetV2-quantized, M=4, per-channel + BN-fused modules  (vs golden of the real reference)
  config 4: mantissa sweep M in {2..7} with the MSE grid estimator on ResNet-18-like activations (vs oracle)
  config 5: batch-sharded calibration: ranks' ranges == single-process ranges on the concatenated batch
            (2 GPUs when available, see also tests/test_dist_gloo.py for the CPU/gloo version)
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import bits, load_golden, ulp_diff
from oracle import fp8_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config3_mobilenetv2_m4_vs_reference_golden():
    from fp8_quantization_b200 import modules, ops, workloads
    from fp8_quantization_b200.quantizers import FPQuantizer

    g = load_golden("mobilenetv2_m4.npz")
    torch.manual_seed(10)
    net = workloads.MobileNetV2()
    sd = net.state_dict()
    assert list(sd.keys()) == list(g["state_keys"])
    mine = np.array([int(sd[k].float().contiguous().view(torch.int32).to(torch.int64).sum()) for k in sd.keys()],
                    dtype=np.int64)
    assert np.array_equal(mine, g["w_checksums"])  # same fp32 network as the reference built under seed 10
    model = workloads.QuantizedMobileNetV2(net, **workloads.readme_quant_params(4)).to(DEV).eval()
    gen = torch.Generator().manual_seed(10)
    x = torch.randn(2, 3, 224, 224, generator=gen).to(DEV)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        workloads.pass_data_for_range_estimation([x], model, True, True, 1)
        model.fix_ranges()
        with torch.no_grad():
            model(x)  # folds the batch norms once
            n0 = ops.launch_count()
            logits = model(x)
            n_fused = ops.launch_count() - n0
            modules.FUSE_BLOCK_TAIL = False
            modules.BATCH_WEIGHT_QUANT = False
            n0 = ops.launch_count()
            logits_layerwise = model(x)
            n_layerwise = ops.launch_count() - n0
            modules.FUSE_BLOCK_TAIL = True
            modules.BATCH_WEIGHT_QUANT = True
            modules.FUSE_EPILOGUES = False
            logits_unfused = model(x)
            modules.FUSE_EPILOGUES = True
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    # 117 quantiser calls of the reference (53 weights + 64 activations) -> 1 weight launch + 52 BN epilogues
    # (10 of them fused with the residual add and its quantiser) + avgpool + fc output
    assert n_layerwise == 53 + 52 + 10 + 2
    assert n_fused == 2 + 42 + 10 + 2  # 53 weight tensors = 2 multi-tensor launches (48 descriptors per launch)
    assert torch.equal(logits, logits_layerwise)  # fusing launches changes no bit
    assert torch.equal(logits, logits_unfused)    # nor does fusing batch norm / ReLU6 / add into the quantiser
    names = [n for n, m in model.named_modules() if isinstance(m, FPQuantizer)]
    assert names == list(g["names"])
    mods = dict(model.named_modules())
    for i, n in enumerate(names):
        ours = mods[n].maxval.reshape(-1).cpu().numpy()
        ref = g[f"maxval_{i:03d}"]
        if ours.size > 1:
            assert np.array_equal(ours, ref), n       # per-channel weight ranges: exact
        elif ref[0] != 3.0:                           # 3.0 = never-calibrated default of unused quantisers
            np.testing.assert_allclose(ours, ref, rtol=3e-2, err_msg=n)
    ref_logits = torch.from_numpy(g["logits"])
    cos = F.cosine_similarity(logits.cpu().flatten(), ref_logits.flatten(), dim=0).item()
    assert cos > 0.97, cos
    assert F.cosine_similarity(logits.flatten(), logits_unfused.flatten(), dim=0).item() > 0.99


@pytest.mark.parametrize("M", [2, 3, 4, 5, 6, 7])
def test_config4_mse_grid_sweep_vs_oracle(M):
    """FP_MSE_Estimator at every mantissa width on a post-ReLU activation tensor: identical search grid, MSE table
    within fp32 summation noise of the oracle's 111-iteration loop, same selected range."""
    import fp8_quantization_b200 as fq

    torch.manual_seed(20 + M)
    x = torch.relu(torch.randn(4, 16, 14, 14)) * 1.7 + 0.02 * torch.randn(4, 16, 14, 14)
    q = fq.FPQuantizer(8, mantissa_bits=M, set_maxval=True, mse_include_mantissa_bits=False)
    est = fq.FP_MSE_Estimator(quantizer=q)
    mn, mx = est(x.to(DEV))
    oq = O.OracleFPQuantizer(8, mantissa_bits=M, set_maxval=True, mse_include_mantissa_bits=False)
    oest = O.OracleFPMSE(quantizer=oq)
    omn, omx = oest(x)
    assert torch.equal(est.search_grid.cpu(), oest.search_grid)
    np.testing.assert_allclose(est.mses.cpu().numpy(), oest.mses.numpy(), rtol=3e-4, atol=1e-12)
    row = oest.mses[0, :, 0]
    gi = int((est.search_grid[:, 0].cpu() - mx.cpu()).abs().argmin())
    assert float(row[gi]) <= float(row.min()) * (1 + 3e-4)  # same argmin, or a tie within summation noise
    assert float(q._mbits_host) == float(oq.mantissa_bits) == float(M)
    # whole-model: ResNet-18 calibrated with the MSE estimator on its activations runs and gives finite logits
    if M == 5:
        from fp8_quantization_b200 import workloads

        torch.manual_seed(10)
        qp = workloads.readme_quant_params(5, act_range_method=fq.FP_MSE_Estimator)
        model = workloads.resnet18_quantized(**qp).to(DEV).eval()
        xb = torch.randn(4, 3, 224, 224, device=DEV)
        workloads.pass_data_for_range_estimation([xb], model, True, True, 1)
        model.fix_ranges()
        with torch.no_grad():
            assert torch.isfinite(model(xb)).all()


def test_next_row_f1_bn_reestimation_vs_reference_golden():
    """reestimate_BN_stats of the quantised network (utils/qat_utils.py:45-90) vs the real reference run
    on the CPU: same flow (momentum 1, BN-train mode on the fused layers only, mean over batches)."""
    from fp8_quantization_b200 import modules, workloads

    g = load_golden("bn_reestimate.npz")
    seq = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1, bias=False), torch.nn.BatchNorm2d(8), torch.nn.ReLU(),
                              torch.nn.Conv2d(8, 6, 1, bias=False), torch.nn.BatchNorm2d(6))
    sd = {k: torch.from_numpy(g["init_" + k.replace(".", "_")]) for k in seq.state_dict().keys()}
    seq.load_state_dict(sd)
    qp = workloads.readme_quant_params(5)
    qp.pop("quant_setup")

    class Wrap(modules.QuantizedModel):
        def __init__(self, f):
            super().__init__((1, 3, 16, 16))
            self.f = f

        def forward(self, x):
            return self.f(x)

    model = Wrap(modules.quantize_model(seq, **qp)).to(DEV).eval()
    xs = [torch.from_numpy(x).to(DEV) for x in g["x"]]
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        workloads.pass_data_for_range_estimation([xs[0]], model, True, True, 1)
        model.fix_ranges()
        n = workloads.reestimate_BN_stats(model, xs, num_batches=3)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert n == 3 and not model.f[0].training and model.f[0].momentum == 0.1
    for i in (0, 1):
        np.testing.assert_allclose(model.f[i].running_mean.cpu().numpy(), g[f"mean_{i}"], rtol=2e-3, atol=2e-4)
        np.testing.assert_allclose(model.f[i].running_var.cpu().numpy(), g[f"var_{i}"], rtol=2e-3, atol=2e-5)
    with torch.no_grad():   # the fused epilogue picks up the new statistics (cached fold is invalidated)
        y_fused = model(xs[1])
        modules.FUSE_EPILOGUES = False
        y_unfused = model(xs[1])
        modules.FUSE_EPILOGUES = True
    assert torch.equal(y_fused, y_unfused)


def test_next_row_f2_line_search_estimator_vs_reference_golden():
    """LineSearchEstimator's 1-D grid search vs the real reference (range_estimators.py:236-256): loss
    array within fp32 summation noise, same argmin (or a tie within that noise), same returned range."""
    import fp8_quantization_b200 as fq

    g = load_golden("line_search.npz")
    for key in ("pt", "pc", "pt_onesided"):
        ncand, M, pc = [int(v) for v in g[key + "_meta"]]
        x = torch.from_numpy(g[key + "_x"]).to(DEV)
        q = fq.FPQuantizer(8, mantissa_bits=M, set_maxval=True)
        est = fq.LineSearchEstimator(quantizer=q, per_channel=bool(pc), num_candidates=ncand)
        mn, mx = est(x)
        ref_loss = g[key + "_loss"]
        np.testing.assert_allclose(est.loss_array[:, 1:], ref_loss[:, 1:], rtol=3e-4)
        assert np.isinf(est.loss_array[:, 0]).all()
        for c in range(ref_loss.shape[0]):
            ours_i = int(round(float(mx[c]) / est.step_size))
            assert ref_loss[c, ours_i] <= ref_loss[c].min() * (1 + 3e-4)
        agree = (mx.cpu().numpy() == g[key + "_xmax"]).mean()
        assert agree >= 0.8 and np.array_equal(mn.cpu().numpy() == 0, g[key + "_xmin"] == 0)
        # accumulation over a second call, and the helper
        est(x)
        np.testing.assert_allclose(est.loss_array[:, 1:], 2 * ref_loss[:, 1:], rtol=3e-4)
    mn, mx = fq.estimate_range_line_search(torch.from_numpy(g["pt_x"]).to(DEV),
                                           fq.FPQuantizer(8, mantissa_bits=4, set_maxval=True), num_candidates=200)
    assert float(mx) == float(g["pt_xmax"][0])


def test_next_row_f4_ste_backward_vs_reference_golden():
    """gradients through FPQuantizer with learnable maxval / mantissa_bits vs the real reference's autograd
    (CPU): grad_x = ((g * s) / s) * {0, 0.5, 1} within 2 ulp (s is ATen's pow result, an exact power of two on
    neither backend every time, and the two backends' pow differ -- DESIGN.md section 3) with the clamp mask
    identical, grad_maxval and grad_mantissa_bits within fp32 summation noise; and grad_x value-identical to the
    oracle's autograd graph evaluated by ATen on the same GPU (parity statement P1), NaN inputs included."""
    import fp8_quantization_b200 as fq

    g = load_golden("backward.npz")
    for i in range(int(g["num_cases"])):
        n = f"b{i:02d}"
        M, sb, pc = [int(v) for v in g[n + "_meta"]]
        x = torch.from_numpy(g[n + "_x"]).to(DEV).requires_grad_(True)
        w = torch.from_numpy(g[n + "_w"]).to(DEV)
        q = fq.FPQuantizer(8, per_channel=bool(pc), mantissa_bits=M, set_maxval=True)
        q.sign_bits = sb
        q.maxval = torch.from_numpy(g[n + "_maxval"]).to(DEV)
        q.learn_maxval()
        q.learn_mantissa_bits()
        assert isinstance(q.maxval, torch.nn.Parameter) and len(list(q.parameters())) == 2
        y = q(x)
        (y * w).sum().backward()
        gx_ref = torch.from_numpy(g[n + "_gx"])
        assert torch.equal(x.grad.cpu() == 0, gx_ref == 0), n               # same clamp mask
        assert int(ulp_diff(x.grad.cpu(), gx_ref)[gx_ref != 0].max()) <= 2, n
        np.testing.assert_allclose(q.maxval.grad.cpu().numpy().reshape(-1), g[n + "_gmaxval"].reshape(-1), rtol=2e-3,
                                   atol=2e-3)
        np.testing.assert_allclose(q.mantissa_bits.grad.cpu().numpy().reshape(-1), g[n + "_gmbits"].reshape(-1),
                                   rtol=2e-3, atol=5e-2)
        # same graph evaluated by ATen autograd on the GPU
        xo = torch.from_numpy(g[n + "_x"]).to(DEV).requires_grad_(True)
        mv = q.maxval.detach().clone().requires_grad_(True)
        mb = torch.tensor([float(M)], device=DEV, requires_grad=True)
        yo = O.fake_quant_ste(xo, 8, mv, mb, sb)
        (yo * w).sum().backward()
        assert torch.equal(bits(y.detach()), bits(yo.detach())) and torch.equal(x.grad, xo.grad)
        np.testing.assert_allclose(q.maxval.grad.cpu().numpy(), mv.grad.cpu().numpy(), rtol=2e-3, atol=2e-3)
    # scalar-access variant (4-byte aligned view, ragged length) == vector variant, E2M5 and E4M3 tables
    from fp8_quantization_b200 import ops
    gen = torch.Generator(DEV).manual_seed(11)
    for M in (5, 3):
        qq = fq.FPQuantizer(8, mantissa_bits=M, maxval=2.5)
        big = torch.randn(3 * 70001 + 4, device=DEV, generator=gen) * 2
        gbig = torch.randn(3 * 70001 + 4, device=DEV, generator=gen)
        tb, _ = qq.table_for(big)
        for C, n0 in ((1, 70001 * 3), (1, 210000), (3, 70001 * 3), (1, 70000), (4, 4 * 16388)):
            xa, ga = big[1:1 + n0], gbig[1:1 + n0]                         # unaligned
            xb, gb = xa.clone(), ga.clone()                                # aligned copies
            tbc = tb if C == 1 else ops.prepare(torch.full((C,), 2.5, device=DEV), float(M), 8, 1)
            gx_a, acc_a = ops.fake_quant_backward(ga, xa, tbc, C, float(M), 8, 1)
            gx_b, acc_b = ops.fake_quant_backward(gb, xb, tbc, C, float(M), 8, 1)
            assert torch.equal(bits(gx_a), bits(gx_b))
            np.testing.assert_allclose(acc_a.cpu().numpy(), acc_b.cpu().numpy(), rtol=1e-4, atol=1e-3)
            xo = xb.clone().requires_grad_(True)
            mv = torch.full((C, 1) if C > 1 else (1,), 2.5, device=DEV, requires_grad=True)
            yo = O.fake_quant_ste(xo.reshape(C, -1) if C > 1 else xo, 8, mv, torch.tensor([float(M)], device=DEV), 1)
            (yo.reshape(-1) * gb).sum().backward()
            assert torch.equal(gx_b, xo.grad)
            gmv = (acc_b[:, 0] + acc_b[:, 1] / 2.5).float().cpu().numpy()
            np.testing.assert_allclose(gmv, mv.grad.reshape(-1).cpu().numpy(), rtol=2e-3, atol=2e-2)
    # a NaN input poisons its own grad_x entry and both parameter gradients, as in the reference
    q = fq.FPQuantizer(8, mantissa_bits=3, maxval=3.0)
    q.learn_maxval()
    xn = torch.tensor([float("nan"), 1.0, 5.0, -5.0, float("inf")], device=DEV, requires_grad=True)
    wn = torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0], device=DEV)
    (q(xn) * wn).sum().backward()
    xo = xn.detach().clone().requires_grad_(True)
    mv = torch.tensor([3.0], device=DEV, requires_grad=True)
    (O.fake_quant_ste(xo, 8, mv, torch.tensor([3.0], device=DEV), 1) * wn).sum().backward()
    assert torch.equal(torch.isnan(xn.grad), torch.isnan(xo.grad)) and bool(torch.isnan(xn.grad[0]))
    assert torch.equal(xn.grad[1:], xo.grad[1:]) and bool(torch.isnan(q.maxval.grad).all() and torch.isnan(mv.grad).all())
    # an optimiser step on maxval changes the table (Parameter version bump) and the output
    q = fq.FPQuantizer(8, mantissa_bits=5, maxval=2.0, learn_maxval=True)
    q.make_range_trainable()
    x = torch.randn(4096, device=DEV, generator=torch.Generator(DEV).manual_seed(4)) * 3
    y0 = q(x).detach().clone()
    opt = torch.optim.SGD(q.parameters(), lr=1e-3)
    q(x).sum().backward()
    opt.step()
    assert 0.5 < float(q.maxval.detach()) < 8.0 and float(q.maxval.detach()) != 2.0
    assert not torch.equal(q(x).detach(), y0)
    q.fix_ranges()
    assert not isinstance(q.maxval, torch.nn.Parameter) and len(list(q.parameters())) == 0
    with torch.no_grad():
        assert torch.equal(q(x), O.fake_quant(x, 8, q.maxval, torch.tensor([5.0], device=DEV), 1))


def test_next_row_f3_uniform_quantizers_bit_exact_vs_reference_golden():
    """Asymmetric / SymmetricUniformQuantizer classes on the GPU vs the real reference (CPU): delta,
    zero-point and every output bit identical (IEEE-exact arithmetic), incl. +-0, +-inf, NaN, 2/4/8 bits,
    per-tensor and per-channel; and through the QuantizationManager with a min/max estimator, bit-identical to the
    reference's op sequence on the same GPU (default mode: ATen-CUDA's reciprocal-multiply for `range / int_max`)."""
    import fp8_quantization_b200 as fq

    g = load_golden("uniform_quantizers.npz")
    for i in range(int(g["num_cases"])):
        n = f"u{i:02d}"
        sym, nb, pc = [int(v) for v in g[n + "_meta"]]
        cls = fq.SymmetricUniformQuantizer if sym else fq.AsymmetricUniformQuantizer
        q = cls(n_bits=nb, per_channel=bool(pc))
        q.aten_cuda_scalar_div = False  # the golden vectors come from the reference run on the CPU (true division)
        assert not q.is_initialized
        q.set_quant_range(torch.from_numpy(g[n + "_min"]).to(DEV), torch.from_numpy(g[n + "_max"]).to(DEV))
        assert q.is_initialized and q.symmetric == bool(sym)
        assert np.array_equal(q.delta.reshape(-1).cpu().numpy(), g[n + "_delta"])
        x = torch.from_numpy(g[n + "_x"]).to(DEV)
        y = q(x).cpu()
        yr = torch.from_numpy(g[n + "_y"])
        assert bool(((bits(y) == bits(yr)) | (torch.isnan(y) & torch.isnan(yr))).all()), n
        # unaligned view -> scalar-access variant
        if not pc:
            y2 = q(x[1:]).cpu()
            assert bool(((bits(y2) == bits(yr[1:])) | (torch.isnan(y2) & torch.isnan(yr[1:]))).all())
    # big tensor vs the oracle on the same device, and the manager flow with the reference's default methods
    torch.manual_seed(5)
    x = torch.randn(64, 64, 56, 56, device=DEV) * 2
    for cls, ocls in ((fq.AsymmetricUniformQuantizer, O.OracleAsymmetricUniform),
                      (fq.SymmetricUniformQuantizer, O.OracleSymmetricUniform)):
        mgr = fq.QuantizationManager(qmethod=cls, init=fq.CurrentMinMaxEstimator, qparams=dict(n_bits=8))
        y = mgr(x)
        oq = ocls(8)
        oq.set_quant_range(x.min(), x.max())
        assert torch.equal(bits(y), bits(oq(x)))
        mgr.fix_ranges()
        assert torch.equal(bits(mgr(x * 0.5)), bits(oq(x * 0.5)))
    with pytest.raises(fq.QuantizerNotInitializedError):
        fq.QuantizationManager(qmethod=fq.AsymmetricUniformQuantizer, qparams=dict(n_bits=8)).fix_ranges()


_DP_SCRIPT = r"""
import os, sys, json, torch
sys.path.insert(0, os.environ["FQ_ROOT"])
import fp8_quantization_b200 as fq
from fp8_quantization_b200 import dist as fq_dist, workloads
from fp8_quantization_b200.quantizers import FPQuantizer
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
assert fq_dist.init_from_env("nccl")
torch.manual_seed(10)
model = workloads.resnet18_quantized(**workloads.readme_quant_params(5)).to(dev).eval()
gen = torch.Generator().manual_seed(10)
gx = torch.randn(4 * world, 3, 224, 224, generator=gen)
workloads.pass_data_for_range_estimation([fq_dist.shard_batch(gx).to(dev)], model, True, True, 1)
model.fix_ranges()
ranges = [m.maxval.reshape(-1).cpu() for m in model.modules() if isinstance(m, FPQuantizer)]
# the two data-parallel routes -- exchange over NVLink peer memory inside the statistics kernel (default when symmetric
# memory is available) and NCCL all-reduce + finishing launch -- must give the same ranges bit for bit
route = "peer-memory" if fq_dist.peer_exchange(dev) is not None else "nccl"
saved = fq_dist._peer_exchange
fq_dist._peer_exchange = None
torch.manual_seed(10)
model_nccl = workloads.resnet18_quantized(**workloads.readme_quant_params(5)).to(dev).eval()
workloads.pass_data_for_range_estimation([fq_dist.shard_batch(gx).to(dev)], model_nccl, True, True, 1)
fq_dist._peer_exchange = saved
ranges_nccl = [m.maxval.reshape(-1).cpu() for m in model_nccl.modules() if isinstance(m, FPQuantizer)]
routes_equal = all(torch.equal(a.view(torch.int32), b.view(torch.int32)) for a, b in zip(ranges, ranges_nccl))
assert routes_equal, "peer-memory route and NCCL route disagree on rank %d" % rank
del model_nccl
fq_dist.enable(False)
with torch.no_grad():
    logits = model(fq_dist.shard_batch(gx).to(dev))
stats = workloads.validate(model, [fq_dist.shard_batch(gx).to(dev)])
if rank == 0:
    # single-process run on the concatenated batch
    torch.manual_seed(10)
    ref = workloads.resnet18_quantized(**workloads.readme_quant_params(5)).to(dev).eval()
    workloads.pass_data_for_range_estimation([gx.to(dev)], ref, True, True, 1)
    ref.fix_ranges()
    ref_ranges = [m.maxval.reshape(-1).cpu() for m in ref.modules() if isinstance(m, FPQuantizer)]
    exact = sum(int(torch.equal(a, b)) for a, b in zip(ranges, ref_ranges))
    close = all(torch.allclose(a, b, rtol=2e-2) for a, b in zip(ranges, ref_ranges))
    with torch.no_grad():
        ref_logits = ref(gx.to(dev))[: logits.shape[0]]
    cos = torch.nn.functional.cosine_similarity(logits.flatten(), ref_logits.flatten(), dim=0).item()
    print("DPRESULT " + json.dumps({"world": world, "quantizers": len(ranges), "bit_equal_ranges": exact,
                                    "all_close": bool(close), "logits_cos": cos, "count": stats["count"],
                                    "route": route, "routes_equal": bool(routes_equal)}))
fq_dist.barrier()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (run under gpurun --gpus 2)")
def test_config5_sharded_calibration_matches_single_process(tmp_path):
    """Weight ranges and the first layer's activation range are bit-identical to the single-process run (min/max
    are order independent); deeper activation ranges agree to cuDNN's batch-size-dependent conv algorithm choice."""
    script = tmp_path / "dp.py"
    script.write_text(_DP_SCRIPT)
    env = dict(os.environ, FQ_ROOT=ROOT)
    world = min(torch.cuda.device_count(), 8)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], env=env,
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("DPRESULT ")][-1]
    import json
    r = json.loads(line[len("DPRESULT "):])
    print(r)
    assert r["quantizers"] == 50 and r["all_close"] and r["count"] == 4 * world
    assert r["routes_equal"] and r["route"] in ("peer-memory", "nccl")
    assert r["bit_equal_ranges"] >= 22  # 21 weight quantisers + the stem's activation range at least
    assert r["logits_cos"] > 0.98


@pytest.mark.parametrize("memory_format", ["nchw", "channels_last"])
def test_graphed_forward_equals_eager_and_pipelines_host_batches(memory_format):
    """workloads.GraphedForward: the validate forward as one CUDA graph gives the eager forward's logits bit for
    bit (both layouts), refuses to capture while ranges are being estimated, and run_pipelined() (pinned host
    batches in, pinned host logits out, copies overlapped with the forwards) returns the same logits per batch."""
    from fp8_quantization_b200 import workloads

    torch.manual_seed(10)
    model = workloads.resnet18_quantized(**workloads.readme_quant_params(5)).to(DEV).eval()
    if memory_format == "channels_last":
        model = model.to(memory_format=torch.channels_last)
    gen = torch.Generator().manual_seed(3)
    xs = [torch.randn(4, 3, 224, 224, generator=gen) for _ in range(5)]
    x0 = xs[0].to(DEV)
    model.set_quant_state(True, True)
    with pytest.raises(RuntimeError):
        workloads.GraphedForward(model, x0)          # still estimating ranges
    workloads.pass_data_for_range_estimation([x0], model, True, True, 1)
    model.fix_ranges()
    with torch.no_grad():
        eager = [model(x.to(DEV)).clone() for x in xs]
    gf = workloads.GraphedForward(model, x0)
    for x, e in zip(xs, eager):
        assert torch.equal(gf(x.to(DEV)), e)
    with pytest.raises(ValueError):
        gf(torch.zeros(2, 3, 224, 224, device=DEV))
    host = [x.pin_memory() for x in xs]
    out = torch.empty(len(xs), 4, 1000).pin_memory()
    gf.run_pipelined(host, out)
    for i, e in enumerate(eager):
        assert torch.equal(out[i], e.cpu()), i
    # validate() accepts the graphed forward in place of the model
    stats = workloads.validate(gf, [x.to(DEV) for x in xs])
    ref = workloads.validate(model, [x.to(DEV) for x in xs])
    assert stats == ref


def test_config3_mobilenetv2_channels_last_fused_equals_layerwise():
    """MobileNetV2 M=4 in channels_last (channel counts 16..1280, none dividing the 1024-element pass stride except
    16/32/64: exercises the per-vector channel path of the channel-innermost kernels, depthwise weights, ReLU6, the
    10 residual tails): same launch counts as NCHW, fused == layer-wise bit for bit, logits track the NCHW network."""
    from fp8_quantization_b200 import modules, ops, workloads

    torch.manual_seed(10)
    m_cl = workloads.QuantizedMobileNetV2(workloads.MobileNetV2(), **workloads.readme_quant_params(4)).to(DEV).eval()
    m_cl = m_cl.to(memory_format=torch.channels_last)
    torch.manual_seed(10)
    m_nchw = workloads.QuantizedMobileNetV2(workloads.MobileNetV2(), **workloads.readme_quant_params(4)).to(DEV).eval()
    gen = torch.Generator().manual_seed(10)
    x = torch.randn(2, 3, 224, 224, generator=gen).to(DEV)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for m in (m_cl, m_nchw):
            workloads.pass_data_for_range_estimation([x], m, True, True, 1)
            m.fix_ranges()
        with torch.no_grad():
            m_cl(x)
            n0 = ops.launch_count()
            y = m_cl(x)
            assert ops.launch_count() - n0 == 2 + 42 + 10 + 2 + 1   # + the space-to-depth gather of the stem
            modules.FUSE_BLOCK_TAIL = False
            modules.BATCH_WEIGHT_QUANT = False
            y_layerwise = m_cl(x)
            modules.FUSE_BLOCK_TAIL = True
            modules.BATCH_WEIGHT_QUANT = True
            y_nchw = m_nchw(x)
    finally:
        modules.FUSE_BLOCK_TAIL = True
        modules.BATCH_WEIGHT_QUANT = True
        torch.backends.cudnn.allow_tf32 = prev
    assert torch.equal(y, y_layerwise)
    assert F.cosine_similarity(y.flatten(), y_nchw.flatten(), dim=0).item() > 0.99
