"""GPU: C-ABI behaviours that need a device: host-buffer end-to-end entry point, error codes with real
pointers, stream semantics / CUDA-graph capture."""

import pytest
import torch

from conftest import bits
from oracle import fp8_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_host_buffer_entry_point_matches_device_path():
    import fp8_quantization_b200 as fq
    from fp8_quantization_b200 import ops

    torch.manual_seed(0)
    x = torch.randn((1 << 24) + 5)          # > one 8 Mi-element chunk: exercises the 3-stream pipeline
    mv = torch.tensor([2.7])
    y = ops.fake_quant_host(x, mv, 5.0, 8, 1)
    q = fq.FPQuantizer(8, mantissa_bits=5, maxval=2.7)
    assert torch.equal(bits(y), bits(q(x.to(DEV)).cpu()))
    w = torch.randn(1000, 512)
    mvw = w.abs().max(1)[0]
    yw = ops.fake_quant_host(w, mvw, 4.0, 8, 1, per_channel=True)
    qw = fq.FPQuantizer(8, per_channel=True, mantissa_bits=4, maxval=1.0)
    qw.maxval = mvw.to(DEV)
    assert torch.equal(bits(yw), bits(qw(w.to(DEV)).cpu()))
    # page-locked caller buffers are DMA'd directly (no staging copies); mixed pinned / pageable also works
    xp, yp = x.pin_memory(), torch.empty_like(x).pin_memory()
    assert torch.equal(bits(ops.fake_quant_host(xp, mv, 5.0, 8, 1, out=yp)), bits(y))
    assert torch.equal(bits(ops.fake_quant_host(xp, mv, 5.0, 8, 1)), bits(y))
    assert torch.equal(bits(ops.fake_quant_host(x, mv, 5.0, 8, 1, out=yp)), bits(y))
    wp = w.pin_memory()
    assert torch.equal(bits(ops.fake_quant_host(wp, mvw, 4.0, 8, 1, per_channel=True, out=torch.empty_like(w).pin_memory())),
                       bits(yw))


def test_error_codes_with_device_pointers():
    from fp8_quantization_b200._lib import lib

    L = lib()
    x = torch.randn(64, device=DEV)
    t = torch.empty(64, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    assert L.fp8fq_prepare_f32(x.data_ptr(), 1, 5.0, 8, 1, t.data_ptr(), st) == 0
    assert L.fp8fq_fake_quant_f32(x.data_ptr() + 2, x.data_ptr(), t.data_ptr(), 8, 1, 8, 5.0, 8, 1, st) == -3
    assert L.fp8fq_fake_quant_f32(x.data_ptr(), x.data_ptr(), t.data_ptr(), 8, 1, 8, 5.0, 8, 2, st) == -1
    assert L.fp8fq_minmax_f32(x.data_ptr(), 64, 1, 64, t.data_ptr(), t.data_ptr() + 4, 0, 0, 0.9, None, st) == -4
    torch.cuda.synchronize()


def test_side_stream_and_cuda_graph_capture():
    """Kernels honour the caller's stream and never synchronise: the whole estimate -> quantise chain is
    capturable in a CUDA graph and replays with new data."""
    import fp8_quantization_b200 as fq

    x = torch.randn(1 << 20, device=DEV)
    q = fq.FPQuantizer(8, mantissa_bits=5, maxval=3.0)
    q(x)  # builds the table outside capture
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        y_side = q(x)
    torch.cuda.current_stream().wait_stream(s)
    assert torch.equal(y_side, q(x))
    static_x = x.clone()
    mgr = fq.QuantizationManager(qmethod=fq.FPQuantizer, init=fq.CurrentMinMaxEstimator,
                                 qparams=dict(n_bits=8, mantissa_bits=4, set_maxval=True))
    mgr(static_x)  # warm-up allocates state
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        static_y = mgr(static_x)  # minmax + set_range + prologue + quantise, no host sync
    for scale in (0.5, 7.0):
        static_x.copy_(x * scale)
        g.replay()
        torch.cuda.synchronize()
        mv = (x * scale).abs().max().reshape(1)
        ref = O.fake_quant(x * scale, 8, mv, torch.tensor([4.0], device=DEV), 1)
        assert torch.equal(bits(static_y), bits(ref))
