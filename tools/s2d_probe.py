"""Probe (GPU box): the ResNet stem (7x7, stride 2, pad 3, 3 -> 64) rewritten as a 4x4 stride-1 convolution over the
2x2 space-to-depth input (12 channels, optionally zero-padded to 16): exact re-indexing of the same sum.  Times
cuDNN on it, the input transform, and checks the result against the direct convolution."""
import json
import os
import sys

import torch
import torch.nn.functional as F

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
CL = torch.channels_last


def gtime(fn, iters=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g), torch.no_grad():
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def s2d_weight(w, cpad):
    O, C = w.shape[:2]
    w8 = F.pad(w, (0, 1, 0, 1))                                  # [O, C, 8, 8], zero row / column 7
    w8 = w8.reshape(O, C, 4, 2, 4, 2).permute(0, 1, 3, 5, 2, 4)   # [O, C, p, q, a, b]
    w4 = w8.reshape(O, C * 4, 4, 4)
    if cpad > C * 4:
        w4 = F.pad(w4, (0, 0, 0, 0, 0, cpad - C * 4))
    return w4.contiguous(memory_format=CL)


def s2d_input(x, cpad):
    s = F.pixel_unshuffle(F.pad(x, (3, 3, 3, 3)), 2)             # [N, C*4 (c, p, q), 115, 115]
    if cpad > s.shape[1]:
        s = F.pad(s, (0, 0, 0, 0, 0, cpad - s.shape[1]))
    return s.contiguous(memory_format=CL)


out = {}
torch.manual_seed(0)
x = torch.randn(B, 3, 224, 224, device=dev)
w = torch.randn(64, 3, 7, 7, device=dev) * 0.05
torch.backends.cudnn.allow_tf32 = False
ref32 = F.conv2d(x, w, stride=2, padding=3)
y32 = F.conv2d(s2d_input(x, 12), s2d_weight(w, 12))
out["fp32_max_abs_diff"] = (ref32 - y32).abs().max().item()
out["fp32_ref_absmax"] = ref32.abs().max().item()
torch.backends.cudnn.allow_tf32 = True
ref = F.conv2d(x.contiguous(memory_format=CL), w.contiguous(memory_format=CL), stride=2, padding=3)
out["direct_cl_us"] = gtime(lambda: F.conv2d(x.contiguous(memory_format=CL), w.contiguous(memory_format=CL), stride=2, padding=3))
for cpad in (12, 16):
    xs, ws = s2d_input(x, cpad), s2d_weight(w, cpad)
    y = F.conv2d(xs, ws)
    assert y.shape == ref.shape, (y.shape, ref.shape)
    out[f"s2d_c{cpad}"] = {"conv_us": gtime(lambda: F.conv2d(xs, ws)), "input_transform_us": gtime(lambda: s2d_input(x, cpad)),
                           "weight_transform_us": gtime(lambda: s2d_weight(w, cpad)),
                           "tf32_max_abs_diff_vs_direct": (y - ref).abs().max().item(),
                           "out_is_channels_last": y.is_contiguous(memory_format=CL)}
    print(cpad, out[f"s2d_c{cpad}"], flush=True)
print(out)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/s2d_probe.json", "w"), indent=1)
