"""Runs ONE eager validate forward of the quantised ResNet-18 (channels_last, batch 128, ranges fixed) after warm-up,
between cudaProfilerStart/Stop, for `ncu --profile-from-start off --metrics gpu__time_duration.sum`."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fp8_quantization_b200 import workloads
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
name = sys.argv[2] if len(sys.argv) > 2 else "resnet18"
torch.manual_seed(10)
qp = workloads.readme_quant_params(5 if name == "resnet18" else 4)
m = (workloads.resnet18_quantized(**qp) if name == "resnet18" else workloads.mobilenetv2_quantized(**qp)).to(dev).eval()
m = m.to(memory_format=torch.channels_last)
x = torch.randn(B, 3, 224, 224, device=dev)
workloads.pass_data_for_range_estimation([x], m, True, True, 1)
m.fix_ranges()
with torch.no_grad():
    for _ in range(3):
        m(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    m(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done")
