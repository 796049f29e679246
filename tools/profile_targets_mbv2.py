"""Launches the kernels of BASELINE config 3 (MobileNetV2, M=4 = E3M4, channels_last and NCHW) at that network's site
shapes, twice each (second round = warm instruction cache), for `ncu --set full`:
    ncu --set full --clock-control none --profile-from-start off -k regex:fq_ -o gpurun_out/prof_mbv2 python tools/profile_targets_mbv2.py
(only the second round of launches is profiled: cudaProfilerStart after the warm-up round)
"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fp8_quantization_b200 as fq
from fp8_quantization_b200 import ops, workloads
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
CL = torch.channels_last
q4 = fq.FPQuantizer(8, mantissa_bits=4, maxval=float(os.environ.get("CL_MAXVAL", "4.0")))   # 3.0: a FLAG_MAGIC table
sites = []
for (C, H) in ((96, 56), (144, 56), (384, 14), (32, 112)):
    x = torch.randn(B, C, H, H, device=dev)
    r = torch.relu(torch.randn(B, C, H, H, device=dev))
    mean, var = torch.randn(C, device=dev), torch.rand(C, device=dev) + 0.5
    sites.append((x, x.contiguous(memory_format=CL), r, r.contiguous(memory_format=CL), ops.bn_pack(mean, var, None, None, 1e-5)))
t4, _ = q4.table_for(sites[0][0])
torch.manual_seed(10)
ws = [m.weight.detach().to(dev) for m in workloads.MobileNetV2().modules() if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear))]
wt = []
for w in ws:
    qq = fq.FPQuantizer(8, per_channel=True, mantissa_bits=4, set_maxval=True)
    wf = w.reshape(w.shape[0], -1)
    qq.set_quant_range(wf.min(1)[0], wf.max(1)[0])
    wt.append(qq.table_for(w)[0])
wo = [torch.empty_like(w) for w in ws]
for i in range(2):
    if i == 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    for x, xc, r, rc, pk in sites[:3] if i == 1 else sites:
        y, yc = torch.empty_like(x), torch.empty_like(xc)
        ops.bn_act_quant(xc, pk, None, 2, t4, 4.0, 8, 1, bn_mode=1, out=yc)             # CL BN + ReLU6 + E3M4
        ops.bn_quant_add_act_quant(xc, rc, pk, None, 0, t4, (4.0, 8, 1), t4, (4.0, 8, 1), bn_mode=1, out=yc)  # CL tail
        ops.bn_act_quant(x, pk, None, 2, t4, 4.0, 8, 1, bn_mode=1, out=y)               # NCHW
    ops.fake_quant_multi(ws, wt, [w.shape[0] for w in ws], 4.0, 8, 1, outs=wo)          # 53 weight tensors
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
