"""Channel-innermost BN + ReLU6 + E3M4 quantiser (MobileNetV2's dominant kernel, BASELINE config 3) at that network's
site shapes, batch 128: device time per launch from CUDA-graph replays over rotating buffers (> L2).
Writes gpurun_out/cl_shapes.json."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fp8_quantization_b200 as fq
from fp8_quantization_b200 import ops
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except (OSError, ValueError, KeyError):
    PEAK = 6650.0
# CL_MAXVAL: the range decides which element path a K > 3 table takes (FLAG_MAGIC, csrc/fp8fq_core.h prep_finish):
# e.g. 3.0 qualifies for M = 4, 4.0 does not (its scale table is not an exact doubling table)
MAXVAL = float(os.environ.get("CL_MAXVAL", "4.0"))
q4 = fq.FPQuantizer(8, mantissa_bits=4, maxval=MAXVAL)
q5 = fq.FPQuantizer(8, mantissa_bits=5, maxval=MAXVAL)
out = {"batch": B, "peak_gbs": PEAK, "maxval": MAXVAL, "sites": []}
for (C, H) in ((32, 112), (96, 112), (96, 56), (144, 56), (144, 28), (192, 28), (384, 14), (576, 14), (960, 7), (1280, 7), (64, 56)):
    n = B * C * H * H
    nbuf = max(2, min(16, int(600e6 // (n * 4)) + 1))
    xs = [torch.randn(B, C, H, H, device=dev).contiguous(memory_format=torch.channels_last) for _ in range(nbuf)]
    y = torch.empty_like(xs[0])
    mean, var = torch.randn(C, device=dev), torch.rand(C, device=dev) + 0.5
    pk = ops.bn_pack(mean, var, None, None, 1e-5)
    rec = {"shape": [B, C, H, H]}
    for name, q, M, act in (("bn_relu6_quant_M4", q4, 4.0, 2), ("bn_relu_quant_M5", q5, 5.0, 1)):
        tb, _ = q.table_for(xs[0])
        out.setdefault("table_flags", {})[name] = int(tb.view(torch.int32)[4].item()) & 0xff
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for x in xs:
                ops.bn_act_quant(x, pk, None, act, tb, M, 8, 1, bn_mode=1, out=y)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for x in xs:
                ops.bn_act_quant(x, pk, None, act, tb, M, 8, 1, bn_mode=1, out=y)
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / (5 * nbuf) * 1e3
        rec[name] = {"us": us, "gbs": 8 * n / us / 1e3, "frac": 8 * n / us / 1e3 / PEAK}
    out["sites"].append(rec)
    print(rec["shape"], " ".join(f"{k}={v['gbs']:.0f}GB/s({v['us']:.1f}us)" for k, v in rec.items() if isinstance(v, dict)), flush=True)
    del xs, y
    torch.cuda.empty_cache()
# MobileNetV2's 53 per-channel weight tensors (M = 4) in one multi-tensor call of the row kernel, graph-timed
from fp8_quantization_b200 import workloads
torch.manual_seed(10)
ws = [m.weight.detach().to(dev) for m in workloads.MobileNetV2().modules() if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear))]
wt = []
for w in ws:
    qq = fq.FPQuantizer(8, per_channel=True, mantissa_bits=4, set_maxval=True)
    wf = w.reshape(w.shape[0], -1)
    qq.set_quant_range(wf.min(1)[0], wf.max(1)[0])
    wt.append(qq.table_for(w)[0])
wo = [torch.empty_like(w) for w in ws]
call = lambda: ops.fake_quant_multi(ws, wt, [w.shape[0] for w in ws], 4.0, 8, 1, outs=wo)
s_ = torch.cuda.Stream()
s_.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s_):
    call()
torch.cuda.current_stream().wait_stream(s_)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10):
        call()
g.replay()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    g.replay()
b.record()
torch.cuda.synchronize()
nw = sum(w.numel() for w in ws)
out["mobilenetv2_weights_M4"] = {"tensors": len(ws), "elements": nw, "us_per_call": a.elapsed_time(b) / 50 * 1e3,
                                 "note": "L2-resident (14 MB of weights, back to back)"}
print("weights", out["mobilenetv2_weights_M4"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", os.environ.get("CL_JSON", "cl_shapes.json")), "w"), indent=1)
