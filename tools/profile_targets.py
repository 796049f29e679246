"""Launches each kernel family twice on a ResNet-18 stem-sized tensor ([128,64,112,112], NCHW and channels_last), the
multi-tensor weight launch and one MSE sweep, for
    ncu --set full --clock-control none --profile-from-start off -o gpurun_out/prof_targets python tools/profile_targets.py
(the second round of launches is the profiled one: warm instruction cache)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fp8_quantization_b200 as fq
from fp8_quantization_b200 import ops
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
shape = (B, 64, 112, 112)
CL = torch.channels_last
xs = [torch.randn(shape, device=dev) for _ in range(2)]
xs_cl = [x.contiguous(memory_format=CL) for x in xs]
r = torch.relu(torch.randn(shape, device=dev))
r_cl = r.contiguous(memory_format=CL)
y = torch.empty(shape, device=dev)
y_cl = torch.empty_like(xs_cl[0])
mean, var = torch.randn(64, device=dev), torch.rand(64, device=dev) + 0.5
pk = ops.bn_pack(mean, var, None, None, 1e-5)
q5 = fq.FPQuantizer(8, mantissa_bits=5, maxval=4.0)
q4 = fq.FPQuantizer(8, mantissa_bits=4, maxval=4.0)
t5, _ = q5.table_for(xs[0])
t4, _ = q4.table_for(xs[0])
cm, cx = torch.empty(1, device=dev), torch.empty(1, device=dev)
# the 21 ResNet-18 weight tensors (one multi-tensor launch)
from torchvision.models import resnet18
torch.manual_seed(10)
ws = [m.weight.detach().to(dev) for m in resnet18().modules() if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear))]
wt = []
for w in ws:
    qq = fq.FPQuantizer(8, per_channel=True, mantissa_bits=5, set_maxval=True)
    wf = w.reshape(w.shape[0], -1)
    qq.set_quant_range(wf.min(1)[0], wf.max(1)[0])
    wt.append(qq.table_for(w)[0])
wo = [torch.empty_like(w) for w in ws]
xm = torch.relu(torch.randn(8, 64, 56, 56, device=dev))
grid = (torch.linspace(0.1, 1.2, 111, device=dev) * xm.max()).reshape(111, 1).contiguous()
mses = torch.zeros(2, 111, 1, device=dev)
for i in range(2):
    if i == 1:   # only the second round is profiled (ncu --profile-from-start off): warm instruction cache, small report
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    x, xc = xs[i], xs_cl[i]
    ops.fake_quant(x, t5, 1, 5.0, 8, 1, out=y)                                   # fq_stream_kernel<0,0,4,0,0>
    ops.fake_quant(x, t4, 1, 4.0, 8, 1, out=y)                                   # <1,0,4,0,0>
    ops.bn_act_quant(x, pk, None, 1, t5, 5.0, 8, 1, bn_mode=1, out=y)            # <0,1,4,0,1>
    ops.add_act_quant(x, r, 1, t5, 5.0, 8, 1, out=y)                             # <0,2,4,0,0>
    ops.bn_quant_add_act_quant(x, r, pk, None, 1, t5, (5.0, 8, 1), t5, (5.0, 8, 1), bn_mode=1, out=y)  # <0,4,4,0,1>
    ops.bn_act_quant(xc, pk, None, 1, t5, 5.0, 8, 1, bn_mode=1, out=y_cl)        # <0,7,4,0,1>  channels_last
    ops.bn_quant_add_act_quant(xc, r_cl, pk, None, 1, t5, (5.0, 8, 1), t5, (5.0, 8, 1), bn_mode=1, out=y_cl)  # <0,8,4,0,1>
    ops.minmax(x, False, cm, cx, ops.EST_CURRENT, False)
    ops.fake_quant_multi(ws, wt, [w.shape[0] for w in ws], 5.0, 8, 1, outs=wo)   # fq_rows_kernel<0,0>
    ops.mse_grid(xm, False, grid, [5.0, 3.0], 8, 1, mses)                        # mse_grid_kernel<0,16>, <1,16>
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
