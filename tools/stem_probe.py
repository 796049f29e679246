"""Probe (GPU box): cuDNN time of the ResNet stem convolution (7x7, stride 2, 3 -> 64 channels, batch 128) with the
input channels zero-padded to 4 / 8, in channels_last and NCHW; and of the other convolutions of ResNet-18, to see
where the forward's time goes.  Writes gpurun_out/stem_probe.json."""
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


out = {}
torch.manual_seed(0)
x3 = torch.randn(B, 3, 224, 224, device=dev)
w3 = torch.randn(64, 3, 7, 7, device=dev) * 0.05
ref = F.conv2d(x3, w3, stride=2, padding=3)
for cin in (3, 4, 8):
    x = torch.zeros(B, cin, 224, 224, device=dev)
    x[:, :3] = x3
    w = torch.zeros(64, cin, 7, 7, device=dev)
    w[:, :3] = w3
    for fmt in ("nchw", "channels_last"):
        xx = x.contiguous(memory_format=CL) if fmt == "channels_last" else x
        ww = w.contiguous(memory_format=CL) if fmt == "channels_last" else w
        for bench in (False, True):
            torch.backends.cudnn.benchmark = bench
            us = gtime(lambda: F.conv2d(xx, ww, stride=2, padding=3))
            y = F.conv2d(xx, ww, stride=2, padding=3)
            out[f"stem_cin{cin}_{fmt}_bench{int(bench)}"] = {"us": us, "max_abs_diff_vs_cin3_nchw": (y - ref).abs().max().item()}
            print(f"stem cin={cin} {fmt} benchmark={bench}: {us:.1f} us", flush=True)
torch.backends.cudnn.benchmark = False
# pad + layout conversion cost of the image
xp = torch.empty(B, 4, 224, 224, device=dev).contiguous(memory_format=CL)


def pad_in():
    xp[:, :3].copy_(x3)


out["pad_image_to_4ch_cl_us"] = gtime(pad_in)
print("pad image:", out["pad_image_to_4ch_cl_us"])
# the other ResNet-18 convolutions (channels_last)
for (cin, cout, hw, k, s) in ((64, 64, 56, 3, 1), (64, 128, 56, 3, 2), (128, 128, 28, 3, 1), (128, 256, 28, 3, 2),
                              (256, 256, 14, 3, 1), (256, 512, 14, 3, 2), (512, 512, 7, 3, 1), (64, 128, 56, 1, 2)):
    x = torch.randn(B, cin, hw, hw, device=dev).contiguous(memory_format=CL)
    w = (torch.randn(cout, cin, k, k, device=dev) * 0.05).contiguous(memory_format=CL)
    us = gtime(lambda: F.conv2d(x, w, stride=s, padding=k // 2))
    flop = 2.0 * B * (hw // s) ** 2 * cout * cin * k * k
    byts = 4.0 * (x.numel() + B * cout * (hw // s) ** 2)
    out[f"conv_{cin}_{cout}_{hw}_k{k}s{s}_cl"] = {"us": us, "tflops": flop / us / 1e6, "hbm_bound_us": byts / 6.5e6}
    print(f"conv {cin}->{cout} {hw}x{hw} k{k} s{s}: {us:.1f} us  {flop / us / 1e6:.0f} TFLOP/s  (HBM bound {byts / 6.5e6:.0f} us)", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/stem_probe.json", "w"), indent=1)
