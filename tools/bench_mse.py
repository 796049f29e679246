"""FP_MSE_Estimator timings (BASELINE config 4): per-tensor activations and per-channel weights, mantissa sweep."""
import json, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fp8_quantization_b200 as fq
from fp8_quantization_b200 import ops
dev = torch.device("cuda:0")
out = {}
def run(name, x, pc, include, M=4, reps=5):
    ts = []
    for r in range(reps):
        q = fq.FPQuantizer(8, per_channel=pc, mantissa_bits=M, set_maxval=True, mse_include_mantissa_bits=include)
        est = fq.FP_MSE_Estimator(per_channel=pc, quantizer=q)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        est(x)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    cands = (6 if include else 1) * 111
    t = sorted(ts)[len(ts) // 2]
    out[name] = {"ms": t * 1e3, "candidate_evals_per_s": x.numel() * cands / t, "elements": x.numel(), "candidates": cands}
    # kernel only
    C = x.shape[0] if pc else 1
    mses = torch.zeros(6 if include else 1, 111, C, device=dev)
    ml = [1., 2., 3., 4., 5., 6.] if include else [float(M)]
    grid = est.search_grid
    for _ in range(2): ops.mse_grid(x, pc, grid, ml, 8, 1, mses)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.mse_grid(x, pc, grid, ml, 8, 1, mses); e1.record(); torch.cuda.synchronize()
    out[name]["kernel_ms"] = e0.elapsed_time(e1)
    out[name]["kernel_candidate_evals_per_s"] = x.numel() * cands / (e0.elapsed_time(e1) * 1e-3)
torch.manual_seed(0)
run("act_8x64x56x56_sweep", torch.relu(torch.randn(8, 64, 56, 56, device=dev)), False, True)
run("act_64x64x56x56_sweep", torch.relu(torch.randn(64, 64, 56, 56, device=dev)), False, True)
run("act_64x64x56x56_fixedM", torch.relu(torch.randn(64, 64, 56, 56, device=dev)), False, False)
run("weight_128x64x3x3_sweep", torch.randn(128, 64, 3, 3, device=dev), True, True)
run("weight_512x512x3x3_sweep", torch.randn(512, 512, 3, 3, device=dev), True, True)
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", os.environ.get("MSE_JSON", "mse.json")), "w"), indent=1)
