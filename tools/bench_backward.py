"""STE-backward kernel timing at the two largest ResNet-18 activation shapes, sweeping the chunk size knob
(FP8FQ_BWD_CHUNK, read per call).  Buffers rotate over > L2."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fp8_quantization_b200 as fq  # noqa: E402
from fp8_quantization_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 6570.3
for shape in ((128, 64, 112, 112), (128, 64, 56, 56)):
    n = 1
    for d in shape:
        n *= d
    nbuf = min(12, max(2, int(400e6 // (n * 4)) + 1))
    xs = [torch.randn(shape, device=dev) for _ in range(nbuf)]
    gs = [torch.randn(shape, device=dev) for _ in range(min(nbuf, 4))]
    for M in (5.0, 4.0):
        q = fq.FPQuantizer(8, mantissa_bits=M, maxval=4.0)
        tb, _ = q.table_for(xs[0])
        for chunk in (0, 8192, 32768):
            if chunk:
                os.environ["FP8FQ_BWD_CHUNK"] = str(chunk)
            else:
                os.environ.pop("FP8FQ_BWD_CHUNK", None)
            for i in range(5):
                ops.fake_quant_backward(gs[i % len(gs)], xs[i % nbuf], tb, 1, M, 8, 1)
            torch.cuda.synchronize()
            evs = []
            for i in range(30):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ops.fake_quant_backward(gs[i % len(gs)], xs[i % nbuf], tb, 1, M, 8, 1)
                b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) for a, b in evs)
            med = ts[len(ts) // 2]
            print(shape, "M", M, "chunk", chunk or "default", f"{med * 1e3:.1f} us  {12 * n / (med * 1e-3) / 1e9:.0f} GB/s "
                  f"({12 * n / (med * 1e-3) / 1e9 / PEAK:.3f})")
