"""Isolated per-kernel timings at the ResNet-18 site shapes (B per GPU given on the command line).
Each measurement rotates over enough distinct buffers to exceed L2 (126 MB) between reuses.
Writes gpurun_out/kernels.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fp8_quantization_b200 as fq  # noqa: E402
from fp8_quantization_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except (OSError, ValueError, KeyError):
    PEAK = 6650.0
out = {"batch": B, "peak_gbs": PEAK, "sites": []}


def timeit(fns, iters=30, warm=5):
    n = len(fns)
    for i in range(warm):
        fns[i % n]()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i, (a, b) in enumerate(evs):
        a.record()
        fns[i % n]()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


q = fq.FPQuantizer(8, mantissa_bits=5, maxval=4.0)
q4 = fq.FPQuantizer(8, mantissa_bits=4, maxval=4.0)
for (C, H) in ((64, 112), (64, 56), (128, 28), (256, 14), (512, 7)):
    shape = (B, C, H, H)
    n = B * C * H * H
    nbuf = max(2, int(400e6 // (n * 4)) + 1)
    nbuf = min(nbuf, 24)
    xs = [torch.randn(shape, device=dev) for _ in range(nbuf)]
    rs = [torch.relu(torch.randn(shape, device=dev)) for _ in range(min(nbuf, 8))]
    y = torch.empty(shape, device=dev)
    mean, var = torch.randn(C, device=dev), torch.rand(C, device=dev) + 0.5
    sc, sh = ops.bn_fold(mean, var, None, None, 1e-5)
    pk = ops.bn_pack(mean, var, None, None, 1e-5)
    tb, _ = q.table_for(xs[0])
    tb4, _ = q4.table_for(xs[0])
    rec = {"shape": list(shape), "elems": n, "buffers": nbuf}
    CL = torch.channels_last
    xs_cl = [x.contiguous(memory_format=CL) for x in xs[:min(nbuf, 12)]]
    rs_cl = [r.contiguous(memory_format=CL) for r in rs[:4]]
    y_cl = torch.empty_like(xs_cl[0])
    for name, bytes_pe, fns in (
        ("plain_M5", 8, [lambda x=x: ops.fake_quant(x, tb, 1, 5.0, 8, 1, out=y) for x in xs]),
        ("plain_M4", 8, [lambda x=x: ops.fake_quant(x, tb4, 1, 4.0, 8, 1, out=y) for x in xs]),
        ("bn_relu_quant", 8, [lambda x=x: ops.bn_act_quant(x, sc, sh, 1, tb, 5.0, 8, 1, out=y) for x in xs]),
        ("bn_relu6_quant_M4", 8, [lambda x=x: ops.bn_act_quant(x, sc, sh, 2, tb4, 4.0, 8, 1, out=y) for x in xs]),
        ("add_relu_quant", 12, [lambda x=x, r=rs[i % len(rs)]: ops.add_act_quant(x, r, 1, tb, 5.0, 8, 1, out=y)
                                for i, x in enumerate(xs)]),
        ("block_tail", 12, [lambda x=x, r=rs[i % len(rs)]: ops.bn_quant_add_act_quant(x, r, sc, sh, 1, tb, (5.0, 8, 1), tb,
                                                                                      (5.0, 8, 1), out=y)
                            for i, x in enumerate(xs)]),
        ("bn_relu_quant_exact", 8, [lambda x=x: ops.bn_act_quant(x, pk, None, 1, tb, 5.0, 8, 1, bn_mode=1, out=y) for x in xs]),
        ("block_tail_exact", 12, [lambda x=x, r=rs[i % len(rs)]: ops.bn_quant_add_act_quant(
            x, r, pk, None, 1, tb, (5.0, 8, 1), tb, (5.0, 8, 1), bn_mode=1, out=y) for i, x in enumerate(xs)]),
        ("bn_relu_quant_cl", 8, [lambda x=x: ops.bn_act_quant(x, sc, sh, 1, tb, 5.0, 8, 1, out=y_cl) for x in xs_cl]),
        ("bn_relu_quant_exact_cl", 8, [lambda x=x: ops.bn_act_quant(x, pk, None, 1, tb, 5.0, 8, 1, bn_mode=1, out=y_cl)
                                       for x in xs_cl]),
        ("bn_relu6_quant_M4_exact_cl", 8, [lambda x=x: ops.bn_act_quant(x, pk, None, 2, tb4, 4.0, 8, 1, bn_mode=1, out=y_cl)
                                           for x in xs_cl]),
        ("block_tail_exact_cl", 12, [lambda x=x, r=rs_cl[i % len(rs_cl)]: ops.bn_quant_add_act_quant(
            x, r, pk, None, 1, tb, (5.0, 8, 1), tb, (5.0, 8, 1), bn_mode=1, out=y_cl) for i, x in enumerate(xs_cl)]),
        ("torch_copy", 8, [lambda x=x: y.copy_(x) for x in xs]),
    ):
        med, mn = timeit(fns)
        rec[name] = {"us": med * 1e3, "us_min": mn * 1e3, "gbs": bytes_pe * n / (med * 1e-3) / 1e9,
                     "frac": bytes_pe * n / (med * 1e-3) / 1e9 / PEAK}
    cm, cx = torch.empty(1, device=dev), torch.empty(1, device=dev)
    med, mn = timeit([lambda x=x: ops.minmax(x, False, cm, cx, ops.EST_CURRENT, False) for x in xs])
    rec["minmax"] = {"us": med * 1e3, "gbs": 4 * n / (med * 1e-3) / 1e9, "frac": 4 * n / (med * 1e-3) / 1e9 / PEAK}
    gy = torch.randn(shape, device=dev)
    med, mn = timeit([lambda x=x: ops.fake_quant_backward(gy, x, tb, 1, 5.0, 8, 1) for x in xs])
    rec["backward_M5"] = {"us": med * 1e3, "gbs": 12 * n / (med * 1e-3) / 1e9, "frac": 12 * n / (med * 1e-3) / 1e9 / PEAK}
    med, mn = timeit([lambda x=x: ops.fake_quant_backward(gy, x, tb4, 1, 4.0, 8, 1) for x in xs])
    rec["backward_M4"] = {"us": med * 1e3, "gbs": 12 * n / (med * 1e-3) / 1e9, "frac": 12 * n / (med * 1e-3) / 1e9 / PEAK}
    out["sites"].append(rec)
    del xs, rs, y, gy, xs_cl, rs_cl, y_cl
    torch.cuda.empty_cache()

# weights: all 21 ResNet-18 tensors in one launch vs one launch each
from torchvision.models import resnet18  # noqa: E402
torch.manual_seed(10)
ws = [m.weight.detach().to(dev) for m in resnet18().modules() if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear))]
qs = []
for w in ws:
    qq = fq.FPQuantizer(8, per_channel=True, mantissa_bits=5, set_maxval=True)
    wf = w.reshape(w.shape[0], -1)
    qq.set_quant_range(wf.min(1)[0], wf.max(1)[0])
    qs.append(qq)
tables = [qq.table_for(w)[0] for qq, w in zip(qs, ws)]
outs = [torch.empty_like(w) for w in ws]
nw = sum(w.numel() for w in ws)
Cs = [w.shape[0] for w in ws]


def graph_time(fn, reps=20, flush=None):
    """Device time of fn() alone: captured in a CUDA graph (no host launch gaps), an L2-flushing write before each
    replay when `flush` is given (its time is measured separately and subtracted)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)

    def cap(with_fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            if flush is not None:
                flush.fill_(1.0)
            if with_fn:
                fn()
        return g

    def run(g):
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    t = run(cap(True))
    return t - (run(cap(False)) if flush is not None else 0.0)


flush = torch.empty(64 << 20, device=dev)  # 256 MB > L2
t_multi = graph_time(lambda: ops.fake_quant_multi(ws, tables, Cs, 5.0, 8, 1, outs=outs), flush=flush)
out["weights_multi"] = {"us": t_multi * 1e3, "elems": nw, "gbs": 8 * nw / (t_multi * 1e-3) / 1e9,
                        "timing": "CUDA graph replay, L2 flushed before each replay (flush time subtracted)"}
t_multi_hot = graph_time(lambda: ops.fake_quant_multi(ws, tables, Cs, 5.0, 8, 1, outs=outs))
out["weights_multi_l2_resident"] = {"us": t_multi_hot * 1e3, "gbs": 8 * nw / (t_multi_hot * 1e-3) / 1e9}
t_each = graph_time(lambda: [ops.fake_quant(w, t, w.shape[0], 5.0, 8, 1, out=o) for w, t, o in zip(ws, tables, outs)],
                    flush=flush)
out["weights_21_launches"] = {"us": t_each * 1e3, "gbs": 8 * nw / (t_each * 1e-3) / 1e9}

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", os.environ.get("KERNELS_JSON", "kernels.json")), "w"), indent=1)
for s in out["sites"]:
    print(s["shape"], " ".join(f"{k}={v['gbs']:.0f}({v['us']:.1f}us)" for k, v in s.items() if isinstance(v, dict)))
print("weights", out["weights_multi"], out["weights_multi_l2_resident"], out["weights_21_launches"])
