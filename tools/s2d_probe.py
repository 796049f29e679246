"""Probe (GPU box): stride-2 stem convolutions (ResNet: 7x7 pad 3, 3 -> 64; MobileNetV2: 3x3 pad 1, 3 -> 32) rewritten
as stride-1 convolutions over the 2x2 space-to-depth input (12 channels, zero-padded to 16): exact re-indexing of the
same sum.  Times cuDNN on both forms and checks the result against the direct convolution."""
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
    O, C, k, _ = w.shape
    w8 = F.pad(w, (0, 1, 0, 1))                                            # [O, C, k+1, k+1], zero last row / column
    a = (k + 1) // 2
    w8 = w8.reshape(O, C, a, 2, a, 2).permute(0, 1, 3, 5, 2, 4)             # [O, C, p, q, a, b]
    w4 = w8.reshape(O, C * 4, a, a)
    if cpad > C * 4:
        w4 = F.pad(w4, (0, 0, 0, 0, 0, cpad - C * 4))
    return w4.contiguous(memory_format=CL)


def s2d_input(x, k, cpad):
    p = k // 2
    s = F.pixel_unshuffle(F.pad(x, (p, p + (k + 1) % 2 + 1 - 1 + (1 if (x.shape[-1] + 2 * p) % 2 else 0), p,
                                    p + (1 if (x.shape[-2] + 2 * p) % 2 else 0))), 2)
    if cpad > s.shape[1]:
        s = F.pad(s, (0, 0, 0, 0, 0, cpad - s.shape[1]))
    return s.contiguous(memory_format=CL)


out = {}
torch.manual_seed(0)
x = torch.randn(B, 3, 224, 224, device=dev)
for name, cout, k in (("resnet_7x7", 64, 7), ("mobilenet_3x3", 32, 3)):
    w = torch.randn(cout, 3, k, k, device=dev) * 0.05
    torch.backends.cudnn.allow_tf32 = False
    ref32 = F.conv2d(x, w, stride=2, padding=k // 2)
    y32 = F.conv2d(s2d_input(x, k, 16), s2d_weight(w, 16))
    rec = {"fp32_max_abs_diff": (ref32 - y32[..., :ref32.shape[-2], :ref32.shape[-1]]).abs().max().item(),
           "shapes": [list(ref32.shape), list(y32.shape)]}
    torch.backends.cudnn.allow_tf32 = True
    xc, wc = x.contiguous(memory_format=CL), w.contiguous(memory_format=CL)
    rec["direct_cl_us"] = gtime(lambda: F.conv2d(xc, wc, stride=2, padding=k // 2))
    rec["direct_cl_nchw_image_us"] = gtime(lambda: F.conv2d(x, wc, stride=2, padding=k // 2))
    for cpad in (12, 16):
        xs, ws = s2d_input(x, k, cpad), s2d_weight(w, cpad)
        rec[f"s2d_c{cpad}_conv_us"] = gtime(lambda: F.conv2d(xs, ws))
    out[name] = rec
    print(name, rec, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/s2d_probe.json", "w"), indent=1)
