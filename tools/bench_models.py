"""Whole-network validate forward (quantised ResNet-18 M=5 and MobileNetV2 M=4, BASELINE configs 2 and 3) in both
memory layouts, batch B per GPU, one CUDA graph per forward, device-resident images.  Also times the unquantised fp32
network (plain torch modules, separate BN / ReLU kernels) for scale.  Writes gpurun_out/models.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import time  # noqa: E402

from fp8_quantization_b200 import modules, ops, workloads  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")


def graph_ms(fn, iters=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g), torch.no_grad():
        out = fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, out


def build(name, M, fmt):
    torch.manual_seed(10)
    qp = workloads.readme_quant_params(M)
    if name == "resnet18":
        m = workloads.resnet18_quantized(**qp)
    else:
        m = workloads.mobilenetv2_quantized(**qp)
    m = m.to(dev).eval()
    if fmt == "channels_last":
        m = m.to(memory_format=torch.channels_last)
    return m


def plain(name, fmt):
    torch.manual_seed(10)
    if name == "resnet18":
        from torchvision.models import resnet18
        m = resnet18()
    else:
        m = workloads.MobileNetV2()
    m = m.to(dev).eval()
    if fmt == "channels_last":
        m = m.to(memory_format=torch.channels_last)
    return m


out = {"batch": B}
x = torch.randn(B, 3, 224, 224, device=dev, generator=torch.Generator(device=dev).manual_seed(10))
for name, M in (("resnet18", 5), ("mobilenet_v2", 4)):
    rec = {}
    logits = {}
    for fmt in ("nchw", "channels_last"):
        m = build(name, M, fmt)
        workloads.pass_data_for_range_estimation([x], m, True, True, 1)
        # calibration forward (estimate_ranges state, eager: every site updates its range), fused vs op-by-op epilogues
        calib = {}
        for fused in (True, False):
            modules.FUSE_CALIBRATION = fused
            best = None
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                workloads.pass_data_for_range_estimation([x], m, True, True, 1)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            calib["fused_ms" if fused else "op_by_op_ms"] = best * 1e3
        modules.FUSE_CALIBRATION = True
        m.fix_ranges()
        with torch.no_grad():
            m(x)
            n0 = ops.launch_count()
            m(x)
            launches = ops.launch_count() - n0
        ms, y = graph_ms(lambda: m(x))
        logits[fmt] = y.clone()
        pms, _ = graph_ms(lambda p=plain(name, fmt): p(x))
        rec[fmt] = {"ms_per_forward": ms, "img_per_s": B / (ms * 1e-3), "fp8fq_launches_per_forward": launches, "calibration_forward": calib,
                    "unquantised_fp32_ms": pms, "unquantised_fp32_img_per_s": B / (pms * 1e-3)}
        del m
        torch.cuda.empty_cache()
    rec["logit_cos_nchw_vs_channels_last"] = torch.nn.functional.cosine_similarity(
        logits["nchw"].flatten(), logits["channels_last"].flatten(), dim=0).item()
    out[f"{name}_M{M}"] = rec
    print(name, json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "models.json"), "w"), indent=1)
