"""Small launches of every kernel family that synchronises between threads or CTAs -- last-CTA tickets (min/max),
shared-memory partials and double atomics (MSE grid, STE backward), block reductions (rows min/max, uniform prepare),
the data-parallel finish kernel, the uint8 table staging -- for

    compute-sanitizer --tool racecheck python tools/sanitize_targets.py
    compute-sanitizer --tool memcheck  python tools/sanitize_targets.py

(sizes are small: the sanitizer slows the kernels down 10-100x).  Prints "done" when every launch returned."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("FP8FQ_BINDING", "ctypes")
import fp8_quantization_b200 as fq
from fp8_quantization_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(4, 16, 28, 28, device=dev)
xc = x.contiguous(memory_format=torch.channels_last)
w = torch.randn(24, 147, device=dev)
cm, cx = torch.empty(1, device=dev), torch.empty(1, device=dev)
for n in (1, 1000, 70001):
    ops.minmax(torch.randn(n, device=dev), False, cm, cx, ops.EST_ALL, n > 1)
cmr, cxr = torch.empty(24, device=dev), torch.empty(24, device=dev)
ops.minmax(w, True, cmr, cxr, ops.EST_CURRENT, False)
q = fq.FPQuantizer(8, mantissa_bits=5, set_maxval=True)
mv, tb = torch.empty(1, device=dev), ops.new_table(1, 5.0, 8, 1, dev)
ops.estimate_prepare(x, False, cm, cx, ops.EST_ALL, True, 0.9, mv, 5.0, 8, 1, tb)
mvr, tbr = torch.empty(24, device=dev), ops.new_table(24, 4.0, 8, 1, dev)
ops.estimate_prepare(w, True, cmr, cxr, ops.EST_CURRENT, False, 0.9, mvr, 4.0, 8, 1, tbr)
pk = ops.bn_pack(torch.randn(16, device=dev), torch.rand(16, device=dev) + 0.5, None, None, 1e-5)
for t in (x, xc):
    ops.bn_act_estimate_prepare(t, pk, None, 1, 1, cm, cx, ops.EST_ALL, True, 0.9, mv, (5.0, 8, 1), tb)
packed = torch.empty(3, device=dev)
ops.minmax(x, False, packed[:1], packed[1:], ops.EST_DP_STATS, False)
ops.dp_finish_prepare(packed, cm, cx, ops.EST_ALL, True, 0.9, mv, (5.0, 8, 1), tb)
for M in (5.0, 3.0):
    grid = (torch.linspace(0.1, 1.2, 111, device=dev) * x.abs().max()).reshape(111, 1).contiguous()
    ops.mse_grid(x, False, grid, [M, 2.0], 8, 1, torch.zeros(2, 111, 1, device=dev))
gridw = (torch.linspace(0.1, 1.2, 111, device=dev).view(-1, 1) * w.abs().max(1)[0].view(1, -1)).contiguous()
ops.mse_grid(w, True, gridw, [5.0], 8, 1, torch.zeros(1, 111, 24, device=dev))
for M, C, t, tab in ((5.0, 1, x, tb), (4.0, 24, w, tbr)):
    ops.fake_quant_backward(torch.randn_like(t), t, tab, C, M, 8, 1)
ops.uniform_prepare(cmr, cxr, 8, True, 1e-8)
ops.uniform_prepare(cmr, cxr, 8, False, 1e-8)
u8 = torch.randint(0, 256, (2, 3, 16, 16), dtype=torch.uint8, device=dev)
ops.normalize_u8(u8, ops.normalize_lut((0.485, 0.456, 0.406), (0.229, 0.224, 0.225), dev))
ops.fake_quant(x, tb, 1, 5.0, 8, 1)
ops.bn_act_quant(xc, pk, None, 2, tb, 5.0, 8, 1, bn_mode=1)
torch.cuda.synchronize()
print("done")
