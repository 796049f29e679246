"""Does ATen's eval-mode F.batch_norm on a channels_last tensor use the same fp32 arithmetic as on NCHW (which the
fused exact-BN epilogue reproduces bit for bit)?  Prints mismatch counts."""
import json, os, sys, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fp8_quantization_b200 as fq
from fp8_quantization_b200 import ops
dev = "cuda:0"
torch.manual_seed(5)
res = {}
for shape in ((16, 64, 56, 56), (8, 96, 14, 14), (4, 24, 9, 5)):
    C = shape[1]
    x = torch.randn(shape, device=dev) * 2
    x_cl = x.contiguous(memory_format=torch.channels_last)
    mean, var = torch.randn(C, device=dev), torch.rand(C, device=dev) + 0.3
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    y = F.batch_norm(x, mean, var, g, b, False, 0.0, 1e-5)
    y_cl = F.batch_norm(x_cl, mean, var, g, b, False, 0.0, 1e-5)
    exact = torch.addcmul(b.view(1, C, 1, 1), g.view(1, C, 1, 1) * (x - mean.view(1, C, 1, 1)), torch.rsqrt(var + 1e-5).view(1, C, 1, 1))
    q = fq.FPQuantizer(8, mantissa_bits=5, maxval=3.0)
    pk = ops.bn_pack(mean, var, g, b, 1e-5)
    t, _ = q.table_for(x)
    fused_cl = ops.bn_act_quant(x_cl, pk, None, 0, t, 5.0, 8, 1, bn_mode=1)
    res[str(shape)] = {"aten_cl_vs_aten_nchw_mismatches": int((y_cl.view(torch.int32) != y.view(torch.int32)).sum()),
                       "fused_cl_vs_Q_aten_cl_mismatches": int((fused_cl.view(torch.int32) != q(y_cl).view(torch.int32)).sum()),
                       "fused_cl_vs_Q_aten_nchw_mismatches": int((fused_cl.view(torch.int32) != q(y).view(torch.int32)).sum()),
                       "n": x.numel()}
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bn_cl_check.json"), "w"), indent=1)
