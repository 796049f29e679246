"""GPU: cases added late in a round (their first run on a device used to be the driver's round-end run; all of them have
since run green on a B200).  Kept in their own module, collected after the other GPU modules.
  * next row f5 -- the empirical flow of compute_quant_error.py (workloads.compute_quant_error_empirical) against the
    real reference's golden vectors (tests/golden/quant_error.npz, made by tests/golden/make_golden_quant_error.py);
  * row a1 by name -- quantize_to_fp8_ste_MM(x, n_bits, maxval, num_mantissa_bits, sign_bits), the functional form;
  * round 2 -- the uint8 input path (ToTensor + Normalize on the device) and the torch-registered operator binding.
The first two also run on the host simulation in tests/test_host_sim_models.py."""
import numpy as np
import pytest
import torch

from conftest import bits, load_golden
from oracle import fp8_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_next_row_f5_empirical_quant_error_flow_vs_reference_golden():
    """compute_quant_error.py:18-57, empirical half (workloads.compute_quant_error_empirical), vs the real reference run
    on the CPU (tests/golden/make_golden_quant_error.py): four sample distributions x the script's five formats.  The
    loss curves are flat near their minimum, so the chosen threshold may be another point of the same plateau: it must
    be optimal for the REFERENCE's loss; the empirical errors agree to the backends' ulp-level differences."""
    from fp8_quantization_b200 import workloads

    g = load_golden("quant_error.npz")
    ncand = int(g["num_candidates"])
    for name in g["names"]:
        x, y = torch.from_numpy(g[f"{name}_x"]).to(DEV), torch.from_numpy(g[f"{name}_y"]).to(DEV)
        rows = workloads.compute_quant_error_empirical(x, y, n_bits=8, num_candidates=ncand)
        assert [r["exp_bits"] for r in rows] == list(g["exp_bits"])
        for r in rows:
            key = f"{name}_e{r['exp_bits']}"
            ref_loss = g[key + "_loss"][0]
            step = float(g[key + "_xmax"][0]) / max(int(np.argmin(ref_loss)), 1)
            ours_i = int(round(r["range_max"] / step))
            assert 1 <= ours_i <= ncand and ref_loss[ours_i] <= ref_loss.min() * (1 + 1e-3), (key, ours_i)
            assert (r["range_min"] == 0.0) == (float(g[key + "_xmin"][0]) == 0.0), key
            np.testing.assert_allclose(r["mse"], float(g[key + "_mse"]), rtol=5e-3, err_msg=key)
            np.testing.assert_allclose(r["dot_prod_mse"], float(g[key + "_dot"]), rtol=5e-3, err_msg=key)


@pytest.mark.parametrize("M,sb,pc", [(5, 1, False), (4, 1, True), (3, 0, False), (2, 1, True), (7, 1, False)])
def test_functional_entry_point_quantize_to_fp8_ste_MM(M, sb, pc):
    """fp8_quantizer.py:91-133 under its own name and signature: the bits of the module (same two launches), hence of
    the reference's op sequence run by ATen on this GPU (the parity bar of tests/test_gpu_parity.py)."""
    import fp8_quantization_b200 as fq

    gen = torch.Generator().manual_seed(10 + M)
    x = (torch.randn(64, 3, 37, generator=gen) * 1.5).to(DEV)
    if sb == 0:
        x = x.abs()
    mv = x.reshape(64, -1).abs().max(1)[0] if pc else x.abs().max().reshape(1)
    mb = torch.tensor([float(M)], device=DEV)
    y = fq.quantize_to_fp8_ste_MM(x, 8, mv, mb, sb)
    assert y.shape == x.shape and y.dtype == torch.float32 and y.device == x.device
    y_view = fq.quantize_to_fp8_ste_MM(x, 8, mv.reshape(-1, 1, 1) if pc else mv, float(M), sb)   # :108-109
    assert torch.equal(bits(y), bits(y_view))
    qz = fq.FPQuantizer(8, per_channel=pc, mantissa_bits=M, maxval=1.0)
    qz.sign_bits = sb
    qz.maxval = mv.clone()
    assert torch.equal(bits(y), bits(qz(x)))
    y_ref = O.fake_quant(x, 8, mv, mb, sb)
    assert bool(((bits(y) == bits(y_ref)) | (torch.isnan(y) & torch.isnan(y_ref))).all())


def test_uint8_normalisation_on_the_gpu_and_in_the_graph():
    """ops.normalize_u8 == torchvision's ToTensor + Normalize bit for bit on the device, and GraphedForward with the
    uint8 pre-processing inside the graph gives the logits of the fp32-fed forward on the same images."""
    import torch
    from fp8_quantization_b200 import workloads

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randint(0, 256, (8, 3, 224, 224), dtype=torch.uint8, device=dev, generator=g)
    norm = workloads.U8Normalize(device=dev)
    def torchvision_cpu(u8):     # ToTensor + Normalize as the reference's loader workers compute them: on the HOST
        t = u8.cpu().to(torch.float32).div(255)   # (ATen-CUDA would multiply by the fp32 reciprocal of 255 instead)
        return t.sub_(torch.tensor(workloads.IMAGENET_MEAN).view(1, 3, 1, 1)).div_(
            torch.tensor(workloads.IMAGENET_STD).view(1, 3, 1, 1)).to(dev)

    want = torchvision_cpu(x)
    got = norm(x)
    assert torch.equal(got.view(torch.int32), want.view(torch.int32))
    odd = torch.randint(0, 256, (2, 3, 7, 9), dtype=torch.uint8, device=dev, generator=g)     # scalar path
    assert torch.equal(norm(odd).view(torch.int32), torchvision_cpu(odd).view(torch.int32))
    torch.manual_seed(10)
    model = workloads.resnet18_quantized(**workloads.readme_quant_params(5)).to(dev).eval().to(memory_format=torch.channels_last)
    workloads.pass_data_for_range_estimation([want], model, True, True, 1)
    model.fix_ranges()
    with torch.no_grad():
        ref = model(want)
    gf = workloads.GraphedForward(model, x, preprocess=norm)
    assert torch.equal(gf(x), ref)
    host = x.cpu().pin_memory()
    out = torch.empty(3, 8, 1000).pin_memory()
    gf.run_pipelined([host] * 3, out)
    assert torch.equal(out[0], ref.cpu()) and torch.equal(out[2], ref.cpu())


def test_torch_registered_ops_are_the_default_binding_and_equal_ctypes():
    """libfp8fq_torch.so (TORCH_LIBRARY(fp8fq, ...), csrc/fp8fq_torch.cpp) is what the package calls for CUDA tensors;
    the ctypes binding of the same C ABI gives the same bits; argument errors surface as Fp8fqError from both."""
    import torch
    import fp8_quantization_b200 as fq
    from fp8_quantization_b200 import ops

    assert ops.torch_binding() is not None and ops.torch_binding().abi_version() == fq.lib().fp8fq_version()
    dev = torch.device("cuda:0")
    torch.manual_seed(6)
    x = torch.randn(4, 32, 28, 28, device=dev)
    r = torch.relu(torch.randn_like(x))
    xc, rc = x.contiguous(memory_format=torch.channels_last), r.contiguous(memory_format=torch.channels_last)
    q5, q4 = fq.FPQuantizer(8, mantissa_bits=5, maxval=3.0), fq.FPQuantizer(8, mantissa_bits=4, maxval=2.0)
    t5, _ = q5.table_for(x)
    t4, _ = q4.table_for(x)
    pk = ops.bn_pack(torch.randn(32, device=dev), torch.rand(32, device=dev) + 0.5, None, None, 1e-5)
    w = torch.randn(32, 288, device=dev)
    qw = fq.FPQuantizer(8, per_channel=True, mantissa_bits=5, set_maxval=True)
    qw.set_quant_range(w.min(1)[0], w.max(1)[0])
    tw, _ = qw.table_for(w)

    def run_all():
        cm, cx = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
        ops.minmax(x, False, cm, cx, ops.EST_CURRENT, False)
        return [ops.fake_quant(x, t5, 1, 5.0, 8, 1), ops.fake_quant(xc, t4, 1, 4.0, 8, 1),
                ops.bn_act_quant(x, pk, None, 1, t5, 5.0, 8, 1, bn_mode=1), ops.bn_act_quant(xc, pk, None, 2, t4, 4.0, 8, 1, bn_mode=1),
                ops.add_act_quant(x, r, 1, t5, 5.0, 8, 1),
                ops.bn_quant_add_act_quant(x, r, pk, None, 1, t4, (4.0, 8, 1), t5, (5.0, 8, 1), bn_mode=1),
                ops.bn_quant_add_act_quant(xc, rc, pk, None, 0, t5, (5.0, 8, 1), t4, (4.0, 8, 1), bn_mode=1),
                ops.fake_quant_multi([w, w * 0.5], [tw, tw], [32, 32], 5.0, 8, 1)[1], cm, cx]

    a = run_all()
    saved = ops._torch_ops
    ops._torch_ops = None
    try:
        b = run_all()
        with pytest.raises(fq.Fp8fqError):
            ops.add_act_quant(x, rc, 1, t5, 5.0, 8, 1)          # mixed layouts, ctypes binding
    finally:
        ops._torch_ops = saved
    for u, v in zip(a, b):
        assert torch.equal(u.contiguous().view(torch.int32), v.contiguous().view(torch.int32))
    with pytest.raises(fq.Fp8fqError):
        ops.add_act_quant(x, rc, 1, t5, 5.0, 8, 1)              # mixed layouts, torch binding
    with pytest.raises(fq.Fp8fqError):
        ops.fake_quant(x.double(), t5, 1, 5.0, 8, 1)
    with pytest.raises(fq.Fp8fqError):
        ops.fake_quant(x, t5[:4], 1, 5.0, 8, 1)                  # table too small for the format
