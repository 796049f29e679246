"""Probe (GPU box): does cuDNN run the fp32/TF32 convolutions of the validate forward faster in channels_last
(NHWC) than in the reference's NCHW?  Plain torchvision networks, no quantisers -- this measures only the library
calls either side of the hot path, to decide which activation layout the fused epilogues should be fed with.
Writes gpurun_out/conv_layout_probe.json."""
import json
import os
import sys

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def time_graph(model, x, iters=20):
    with torch.no_grad():
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                model(x)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            y = model(x)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            g.replay()
        b.record()
        torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, y


def main():
    out = {"torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
           "allow_tf32_cudnn": torch.backends.cudnn.allow_tf32}
    B = int(os.environ.get("PROBE_BATCH", "128"))
    for name, ctor in (("resnet18", torchvision.models.resnet18), ("mobilenet_v2", torchvision.models.mobilenet_v2)):
        torch.manual_seed(10)
        m = ctor().cuda().eval()
        x = torch.randn(B, 3, 224, 224, device="cuda")
        for bench in (False, True):
            torch.backends.cudnn.benchmark = bench
            ms_nchw, y0 = time_graph(m, x)
            m_cl = ctor().cuda().eval()
            m_cl.load_state_dict(m.state_dict())
            m_cl = m_cl.to(memory_format=torch.channels_last)
            x_cl = x.contiguous(memory_format=torch.channels_last)
            ms_cl, y1 = time_graph(m_cl, x_cl)
            ms_cl_in, _ = time_graph(m_cl, x)  # NCHW images in, channels_last inside
            cos = torch.nn.functional.cosine_similarity(y0.flatten(), y1.flatten(), dim=0).item()
            out[f"{name}_cudnn_benchmark_{int(bench)}"] = {
                "batch": B, "nchw_ms": ms_nchw, "channels_last_ms": ms_cl, "channels_last_nchw_input_ms": ms_cl_in,
                "logit_cos": cos, "max_abs_diff": (y0 - y1).abs().max().item()}
            print(name, bench, out[f"{name}_cudnn_benchmark_{int(bench)}"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/conv_layout_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
