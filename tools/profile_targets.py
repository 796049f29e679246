"""Launches each kernel family twice on a ResNet-18 stem-sized tensor ([128,64,112,112]) for `ncu --set full`."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fp8_quantization_b200 as fq
from fp8_quantization_b200 import ops
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
shape = (B, 64, 112, 112)
xs = [torch.randn(shape, device=dev) for _ in range(3)]
r = torch.relu(torch.randn(shape, device=dev))
y = torch.empty(shape, device=dev)
mean, var = torch.randn(64, device=dev), torch.rand(64, device=dev) + 0.5
pk = ops.bn_pack(mean, var, None, None, 1e-5)
q5 = fq.FPQuantizer(8, mantissa_bits=5, maxval=4.0)
q4 = fq.FPQuantizer(8, mantissa_bits=4, maxval=4.0)
t5, _ = q5.table_for(xs[0])
t4, _ = q4.table_for(xs[0])
cm, cx = torch.empty(1, device=dev), torch.empty(1, device=dev)
for i in range(2):
    x = xs[i]
    ops.fake_quant(x, t5, 1, 5.0, 8, 1, out=y)                                   # fq_stream_kernel<0,0,4,0,0>
    ops.fake_quant(x, t4, 1, 4.0, 8, 1, out=y)                                   # <1,0,4,0,0>
    ops.bn_act_quant(x, pk, None, 1, t5, 5.0, 8, 1, bn_mode=1, out=y)            # <0,1,4,0,1>
    ops.add_act_quant(x, r, 1, t5, 5.0, 8, 1, out=y)                             # <0,2,4,0,0>
    ops.bn_quant_add_act_quant(x, r, pk, None, 1, t5, (5.0, 8, 1), t5, (5.0, 8, 1), bn_mode=1, out=y)  # <0,4,4,0,1>
    ops.minmax(x, False, cm, cx, ops.EST_CURRENT, False)
torch.cuda.synchronize()
print("done")
