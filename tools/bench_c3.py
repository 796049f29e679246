"""BASELINE config 3 alone (bench.py's c3 leg: MobileNetV2 M=4 hot-path step + whole forward, both layouts), for A/B runs of
build variants (FP8FQ_LIB=...).  Writes gpurun_out/$C3_JSON."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from fp8_quantization_b200 import ops, workloads
args = argparse.Namespace(batch=int(sys.argv[1]) if len(sys.argv) > 1 else 128, memory_format="nchw", steps=30)
dev = torch.device("cuda:0")
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except (OSError, ValueError, KeyError):
    peak = 6538.0
out = bench.leg_c3_mobilenetv2(args, dev, ops, workloads, peak, 30)
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", os.environ.get("C3_JSON", "c3.json")), "w"), indent=1)
