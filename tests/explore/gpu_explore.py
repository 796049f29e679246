"""First-contact GPU script: parity statistics, libm questions and first throughput numbers.
Lives under tests/ because it uses the oracle (test infrastructure) as the checker.
Writes gpurun_out/explore.json.  Diagnostic only (bench.py is the measured artefact)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import fp8_quantization_b200 as fq  # noqa: E402
from fp8_quantization_b200 import ops  # noqa: E402
from oracle import fp8_oracle as O  # noqa: E402

out = {}
dev = torch.device("cuda:0")
out["gpu"] = torch.cuda.get_device_name(0)
out["cpu_count"] = os.cpu_count()
out["reference_on_box"] = os.path.exists("/root/reference")
out["torch"] = torch.__version__
out["build"] = fq.lib().fp8fq_build_info().decode()
out["cpu_capability"] = torch.backends.cpu.get_cpu_capability()


def bits(t):
    return t.contiguous().view(torch.int32)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


# ---- 1. device tables vs oracle tables (CUDA eager and CPU) ------------------------------------------
tab_stats = {}
g = torch.Generator().manual_seed(1)
for M in range(1, 8):
    for sb in (0, 1):
        C = 512
        mv = (torch.rand(C, generator=g) * 8 + 0.01).float()
        mv[:8] = torch.tensor([1.0, 2.0, 0.5, 3.0, 240.0, 15.5, 3.9375, 57344.0])
        mvd = mv.to(dev)
        table = ops.prepare(mvd, float(M), 8, sb)
        _, _, K = ops.format_split(float(M), 8, sb)
        stride = ops.table_stride(float(M), 8, sb)
        t = table.view(C, stride).cpu()
        kp = (K + 2) & ~1
        bias_dev = t[:, 3]
        sc_dev = t[:, 8 + kp:8 + kp + 2 * (K + 1)].reshape(C, K + 1, 2)[:, 1:, 0]
        b_cu, s_cu = O.quant_tables(8, mvd, torch.tensor([float(M)], device=dev), sb)
        b_cpu, s_cpu = O.quant_tables(8, mv, torch.Tensor([float(M)]), sb)
        fin = torch.isfinite(s_cpu) & (s_cpu > 0)
        tab_stats[f"M{M}_s{sb}"] = dict(
            K=K,
            bias_ne_cuda=int((bits(bias_dev) != bits(b_cu.cpu())).sum()),
            bias_ne_cpu=int((bits(bias_dev) != bits(b_cpu)).sum()),
            scale_ne_cuda=int((bits(sc_dev) != bits(s_cu.cpu()))[fin].sum()),
            scale_ne_cpu=int((bits(sc_dev) != bits(s_cpu))[fin].sum()),
            cpu_vs_cuda_scale_ne=int((bits(s_cu.cpu()) != bits(s_cpu))[fin].sum()),
            n_scales=int(fin.sum()),
            irregular=int((t[:, 4].view(torch.int32) & 1).sum()),
        )
out["tables"] = tab_stats

# ---- 2. K1 parity vs oracle on CUDA and on CPU --------------------------------------------------------
par = {}
for M in range(1, 8):
    for sb in (0, 1):
        for pc in (False, True):
            torch.manual_seed(100 + M)
            x = torch.randn(64, 4099 if not pc else 4608) * (1.0 if M > 2 else 20.0)
            xd = x.to(dev)
            q = fq.FPQuantizer(8, per_channel=pc, mantissa_bits=M, set_maxval=True)
            q.sign_bits = sb
            if pc:
                mn, mx = xd.min(1)[0], xd.max(1)[0]
            else:
                mn, mx = xd.min().reshape(1), xd.max().reshape(1)
            q.set_quant_range(mn * 0.8, mx * 0.8)
            y, codes = q.quantize_with_codes(xd)
            mvd = q.maxval.view(-1) if pc else q.maxval
            yo, eo, qo = O.fake_quant(xd, 8, mvd, torch.tensor([float(M)], device=dev), sb, return_codes=True)
            yc, ec, qc = O.fake_quant(x, 8, mvd.cpu(), torch.Tensor([float(M)]), sb, return_codes=True)
            e_dev = ((codes >> 16) & 0x7fff).float()
            q_dev = (codes & 0xffff).float()
            so, eo_c, qo_c = O.canonical_codes(yo, eo, qo, M)
            sd, ed_c, qd_c = O.canonical_codes(y, e_dev, torch.copysign(q_dev, y), M)
            sc, ec_c, qc_c = O.canonical_codes(yc, ec, qc, M)
            par[f"M{M}_s{sb}_{'pc' if pc else 'pt'}"] = dict(
                float_ne_cuda=int((bits(y) != bits(yo)).sum()),
                raw_e_ne_cuda=int((e_dev != eo).sum()),
                canon_ne_cuda=int(((sd != so) | (ed_c != eo_c) | (qd_c != qo_c)).sum()),
                float_ne_cpu=int((bits(y.cpu()) != bits(yc)).sum()),
                canon_ne_cpu=int(((sd.cpu() != sc) | (ed_c.cpu() != ec_c) | (qd_c.cpu() != qc_c)).sum()),
                cuda_eager_vs_cpu_float_ne=int((bits(yo.cpu()) != bits(yc)).sum()),
                cuda_eager_vs_cpu_canon_ne=int(((so.cpu() != sc) | (eo_c.cpu() != ec_c) | (qo_c.cpu() != qc_c)).sum()),
                n=x.numel(),
            )
out["parity"] = par

# ---- 3. estimators ---------------------------------------------------------------------------------------
est = {}
x = torch.randn(3, 1 << 20, device=dev)
cm, cx = torch.empty(1, device=dev), torch.empty(1, device=dev)
ops.minmax(x, False, cm, cx, ops.EST_CURRENT, False)
est["pt_ok"] = bool(cm.item() == x.min().item() and cx.item() == x.max().item())
cm3, cx3 = torch.empty(3, device=dev), torch.empty(3, device=dev)
ops.minmax(x, True, cm3, cx3, ops.EST_CURRENT, False)
est["pc_ok"] = bool(torch.equal(cm3, x.min(1)[0]) and torch.equal(cx3, x.max(1)[0]))
x[1, 77] = float("nan")
ops.minmax(x, False, cm, cx, ops.EST_CURRENT, False)
est["nan_propagates"] = bool(torch.isnan(cm).item() and torch.isnan(cx).item())
out["estimators"] = est

# ---- 4. BN formula experiment: which arithmetic does F.batch_norm(eval) use on this GPU? ----------------
bn = {}
torch.manual_seed(3)
xb = torch.randn(8, 64, 28, 28, device=dev) * 3
mean = torch.randn(64, device=dev)
var = torch.rand(64, device=dev) + 0.3
gamma = torch.randn(64, device=dev)
beta = torch.randn(64, device=dev)
eps = 1e-5
ref = torch.nn.functional.batch_norm(xb, mean, var, gamma, beta, False, 0.0, eps)
with torch.backends.cudnn.flags(enabled=False):
    ref_native = torch.nn.functional.batch_norm(xb, mean, var, gamma, beta, False, 0.0, eps)
bn["cudnn_vs_native_ne"] = int((bits(ref) != bits(ref_native)).sum())
scale, shift = ops.bn_fold(mean, var, gamma, beta, eps)
v = lambda t: t.view(1, -1, 1, 1)
d = lambda t: t.double()
fma0 = (d(xb) * d(v(scale)) + d(v(shift))).float()
bn["fold_fma_vs_cudnn_ne"] = int((bits(fma0) != bits(ref)).sum())
bn["fold_fma_vs_native_ne"] = int((bits(fma0) != bits(ref_native)).sum())
mul_add = (xb * v(scale)) + v(shift)
bn["fold_muladd_vs_cudnn_ne"] = int((bits(mul_add) != bits(ref)).sum())
invstd = 1.0 / torch.sqrt(var + eps)
nat = ((d(v(gamma) * (xb - v(mean))) * d(v(invstd))) + d(v(beta))).float()  # fma(gamma*(x-mean), invstd, beta)
bn["native_formula_fma_vs_native_ne"] = int((bits(nat) != bits(ref_native)).sum())
bn["native_formula_fma_vs_cudnn_ne"] = int((bits(nat) != bits(ref)).sum())
nat2 = (v(gamma) * (xb - v(mean))) * v(invstd) + v(beta)
bn["native_formula_nofma_vs_native_ne"] = int((bits(nat2) != bits(ref_native)).sum())
rs = torch.rsqrt(var + eps)
nat3 = ((d(xb - v(mean)) * d(v(rs * gamma))) + d(v(beta))).float()
bn["xm_times_grs_fma_vs_cudnn_ne"] = int((bits(nat3) != bits(ref)).sum())
bn["xm_times_grs_fma_vs_native_ne"] = int((bits(nat3) != bits(ref_native)).sum())
nat4 = ((d(xb - v(mean)) * d(v(invstd))) * 1.0).float()
nat4 = (d(nat4) * d(v(gamma)) + d(v(beta))).float()
bn["xm_invstd_then_fma_gamma_beta_vs_cudnn_ne"] = int((bits(nat4) != bits(ref)).sum())
bn["xm_invstd_then_fma_gamma_beta_vs_native_ne"] = int((bits(nat4) != bits(ref_native)).sum())
bn["n"] = xb.numel()
# our fused kernel vs its own formula
qz = fq.FPQuantizer(8, mantissa_bits=5, set_maxval=True)
qz.set_quant_range(torch.zeros(1, device=dev), ref.max().reshape(1))
tb, _ = qz.table_for(xb)
yk = ops.bn_act_quant(xb, scale, shift, ops.ACT_RELU, tb, 5.0, 8, 1)
yr = qz(torch.relu(fma0))
bn["fused_kernel_vs_composition_ne"] = int((bits(yk) != bits(yr)).sum())
yr2 = qz(torch.relu(ref))
bn["fused_kernel_vs_cudnn_bn_then_quant_ne"] = int((bits(yk) != bits(yr2)).sum())
out["bn"] = bn

# ---- 5. throughput ------------------------------------------------------------------------------------------
perf = {}
n = 1 << 28
x = torch.randn(n, device=dev)
y = torch.empty_like(x)
for M in (5, 4, 3, 2, 7):
    q = fq.FPQuantizer(8, mantissa_bits=M, set_maxval=True)
    q.set_quant_range(x.min().reshape(1), x.max().reshape(1))
    tb, _ = q.table_for(x)
    med, mn_ = timeit(lambda: ops.fake_quant(x, tb, 1, float(M), 8, 1, out=y))
    perf[f"k1_pt_M{M}"] = dict(ms=med, ms_min=mn_, gelem_s=n / med / 1e6, gbs=8 * n / med / 1e6)
cm, cx = torch.empty(1, device=dev), torch.empty(1, device=dev)
med, mn_ = timeit(lambda: ops.minmax(x, False, cm, cx, ops.EST_CURRENT, False))
perf["minmax_pt"] = dict(ms=med, ms_min=mn_, gbs=4 * n / med / 1e6)
med, mn_ = timeit(lambda: y.copy_(x))
perf["torch_copy"] = dict(ms=med, ms_min=mn_, gbs=8 * n / med / 1e6)
# fused bn+relu+quant on [256,64,56,56]
xa = torch.randn(256, 64, 56, 56, device=dev)
ya = torch.empty_like(xa)
sc64, sh64 = ops.bn_fold(mean, var, gamma, beta, eps)
q = fq.FPQuantizer(8, mantissa_bits=5, set_maxval=True)
q.set_quant_range(torch.zeros(1, device=dev), torch.full((1,), 8.0, device=dev))
tb, _ = q.table_for(xa)
med, mn_ = timeit(lambda: ops.bn_act_quant(xa, sc64, sh64, ops.ACT_RELU, tb, 5.0, 8, 1, out=ya))
perf["bn_relu_quant_256x64x56x56"] = dict(ms=med, ms_min=mn_, gbs=8 * xa.numel() / med / 1e6)
xb2 = torch.randn_like(xa)
med, mn_ = timeit(lambda: ops.add_act_quant(xa, xb2, ops.ACT_RELU, tb, 5.0, 8, 1, out=ya))
perf["add_relu_quant_256x64x56x56"] = dict(ms=med, ms_min=mn_, gbs=12 * xa.numel() / med / 1e6)
# per-channel weights [512, 4608]
w = torch.randn(512, 4608, device=dev)
qw = fq.FPQuantizer(8, per_channel=True, mantissa_bits=5, set_maxval=True)
qw.set_quant_range(w.min(1)[0], w.max(1)[0])
tw, Cw = qw.table_for(w)
yw = torch.empty_like(w)
med, mn_ = timeit(lambda: ops.fake_quant(w, tw, Cw, 5.0, 8, 1, out=yw))
perf["k1_pc_512x4608"] = dict(ms=med, ms_min=mn_, gbs=8 * w.numel() / med / 1e6)
# reference eager on the same GPU (13 kernels)
mvd = q.maxval
mb = torch.tensor([5.0], device=dev)
xs = x[: 1 << 26]
med, mn_ = timeit(lambda: O.fake_quant(xs, 8, mvd, mb, 1), iters=5, warm=2)
perf["reference_eager_cuda_2^26"] = dict(ms=med, gelem_s=xs.numel() / med / 1e6)
# reference on CPU (all threads)
xc = torch.randn(1 << 24)
t0 = time.perf_counter()
O.fake_quant(xc, 8, torch.Tensor([3.0]), torch.Tensor([5.0]), 1)
t1 = time.perf_counter()
perf["reference_cpu_2^24"] = dict(s=t1 - t0, gelem_s=xc.numel() / (t1 - t0) / 1e9, threads=torch.get_num_threads())
out["perf"] = perf

# ---- 6. MSE estimator smoke ------------------------------------------------------------------------------
try:
    xm = torch.relu(torch.randn(8, 64, 56, 56, device=dev))
    qm = fq.FPQuantizer(8, mantissa_bits=4, set_maxval=True, mse_include_mantissa_bits=True)
    em = fq.FP_MSE_Estimator(quantizer=qm)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mn, mx = em(xm)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    oq = O.OracleFPQuantizer(8, mantissa_bits=4, set_maxval=True, mse_include_mantissa_bits=True)
    oe = O.OracleFPMSE(quantizer=oq)
    xs_ = xm[:1].cpu()
    out["mse"] = dict(seconds=t1 - t0, best_m=float(qm._mbits_host), maxval=float(mx.item()))
except Exception as ex:  # noqa: BLE001
    out["mse"] = dict(error=repr(ex))

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "explore.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
