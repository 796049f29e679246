"""CPU: bench.py's record / replay machinery and workload accounting, driven on the host simulation of the kernels
(fixture ``simdev``) -- the numbers the bench line is built from: launches per step, quantiser applications per image
(3,237,864 activation elements + 11.7 M weight elements for ResNet-18), algorithmic bytes, and
that replaying the recorded plan reproduces the recorded outputs.  The timing legs need a GPU and are not covered."""
import importlib.util
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_record_replay_and_workload_accounting(simdev):
    from fp8_quantization_b200 import ops, workloads

    bench = _bench()
    torch.manual_seed(10)
    B = 1
    for fmt in ("nchw", "channels_last"):
        model = workloads.resnet18_quantized(**workloads.readme_quant_params(5)).eval()
        if fmt == "channels_last":
            model = model.to(memory_format=torch.channels_last)
        x = torch.randn(B, 3, 224, 224)
        workloads.pass_data_for_range_estimation([x], model, True, True, 1)
        model.fix_ranges()
        with torch.no_grad():
            model(x)
            with bench.Recorder(ops) as rec:
                model(x)
        recorded = [(name, res) for name, _, _, res in rec.calls]
        plan = bench.build_replay(rec.calls, ops)
        st = bench.plan_stats(plan)
        # 1 multi-tensor weight launch + 12 BN epilogues + 8 block tails + 2 plain quantisers (avgpool, fc output)
        assert st["launches"] == 23 and st["weight_launches"] == 1 and st["stream_launches"] == 22, (fmt, st)
        assert st["weight_elems"] == 11_678_912            # the 20 convolution weights + the fc weight
        act_elems = st["elems"] - st["weight_elems"]
        # the reference applies 30 activation quantisers per image = 3,237,864 elements; the fused block tail applies
        # two of them per element it touches, and the count is per quantiser APPLICATION
        assert act_elems == 3_237_864 * B, (fmt, act_elems)
        # algorithmic bytes: 8 B per element of a BN+act+quant / plain site, 12 B per element of a residual tail
        tails = sum(bench._numel(sh) for n_, a_, _ in plan if n_ == "bn_quant_add_act_quant"
                    for sh in bench.data_shapes(n_, a_))
        plain = sum(bench._numel(sh) for n_, a_, _ in plan if bench.is_stream_call(n_, a_) and n_ != "bn_quant_add_act_quant"
                    for sh in bench.data_shapes(n_, a_))
        assert st["stream_bytes"] == 12 * tails + 8 * plain
        assert st["stream_elems"] == 2 * tails + plain == act_elems
        # replaying the plan on the recorded inputs reproduces every recorded output, bit for bit
        with torch.no_grad():
            results = bench.run_plan(plan, ops)
        for (name, res), out in zip(recorded, results):
            a = res if isinstance(res, (list, tuple)) else [res]
            b = out if isinstance(out, (list, tuple)) else [out]
            for u, v in zip(a, b):
                assert torch.equal(u.contiguous().view(torch.int32), v.contiguous().view(torch.int32)), (fmt, name)


def test_reference_arm_line_has_the_contract_keys():
    """`bench.py --impl reference` (the oracle timed on the host cores) prints the same config keys as our arm."""
    bench = _bench()
    cfg = bench.base_config(5)
    assert cfg["workload"] == bench.WORKLOAD and cfg["mantissa_bits"] == 5 and cfg["n_bits"] == 8
    threads = torch.get_num_threads()
    try:
        cb = bench.run_cpu_reference(1, 0, 2, 5)   # probes thread counts and keeps the fastest for its own process
    finally:
        torch.set_num_threads(threads)
    for k in ("value", "unit", "cores", "kind", "sample", "ms_per_step", "steps", "batch"):
        assert k in cb
    assert cb["kind"] == "port" and cb["unit"] == "Gelem/s" and cb["steps"] == 1 and cb["batch"] == 2 and cb["value"] > 0


def test_ab_build_options_tool_hash_leg_on_the_simulation(simdev):
    """tools/ab_build_options.py (the GPU A/B of the kernels' build options): its parity leg -- output hashes of 384 fused
    calls -- runs here on the host simulation, so that the tool is known to work before it is given GPU time."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ab_build_options", os.path.join(ROOT, "tools", "ab_build_options.py"))
    ab = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ab)
    h = ab.hash_leg("cpu")
    assert len(h) == 384 and not any(v.startswith("error") for v in h.values())
    assert sum(v == "unsupported" for v in h.values()) < 40       # block tails of the two small NCHW shapes
    assert len({v for v in h.values() if len(v) == 64}) >= 90      # distinct outputs per (shape, format, activation, op)
    for k, v in h.items():                                          # the two layouts give the same bits
        if "/nchw/" in k and len(v) == 64 and len(h[k.replace("/nchw/", "/channels_last/")]) == 64:
            assert h[k.replace("/nchw/", "/channels_last/")] == v, k
    assert "default" in ab.VARIANTS and ab.VARIANTS["default"] == []


def test_extra_legs_plumbing_on_the_simulation(simdev):
    """The round-2 legs of bench.py that can run without a GPU: the calibration leg's accounting, the replay of the
    reference's eager op sequence over a recorded plan (its outputs must be the product's, up to the two libm's), the
    config-3 recorder on MobileNetV2 (launch / element accounting of SURVEY section 8a)."""
    from fp8_quantization_b200 import dist as fq_dist
    from fp8_quantization_b200 import ops, workloads

    bench = _bench()
    torch.manual_seed(10)
    model = workloads.resnet18_quantized(**workloads.readme_quant_params(5)).eval()
    x = torch.randn(1, 3, 224, 224)
    workloads.pass_data_for_range_estimation([x], model, True, True, 1)
    model.fix_ranges()
    cal = bench.calibration_leg(model, x, workloads, fq_dist, 1, reps=1)
    assert cal["allreduces"] == 0 and cal["ms"] > 0 and cal["batches"] == 1
    assert not any(m.estimating() for m in model.modules() if hasattr(m, "estimating"))   # ranges fixed again
    plan, st = bench.record_hot_path(model, x, ops)
    assert st["launches"] == 23
    step = bench.leg_reference_gpu_eager(plan, st, torch.device("cpu"), steps=0)
    with torch.no_grad():
        ref_outs = step()
        ours = bench.run_plan(plan, ops)
    checked = 0
    for (name, _, _), a, b in zip(plan, ref_outs, ours):
        for u, v in zip(a if isinstance(a, (list, tuple)) else [a], b if isinstance(b, (list, tuple)) else [b]):
            bad = ((u - v).abs() > 1e-5 * v.abs().clamp_min(1e-3)).float().mean().item()
            assert bad < 2e-2, (name, bad)     # tie flips from ulp-level batch-norm / scale-table differences only
            checked += 1
    assert checked == 21 + 22
    # config 3 accounting: 53 weight tensors (2 multi-tensor launches), 6,896,776 activation elements per image at 224^2
    m3 = workloads.mobilenetv2_quantized(**workloads.readme_quant_params(4)).eval()
    x3 = torch.randn(1, 3, 224, 224)
    workloads.pass_data_for_range_estimation([x3], m3, True, True, 1)
    m3.fix_ranges()
    _, st3 = bench.record_hot_path(m3, x3, ops)
    assert st3["weight_launches"] == 1 and st3["weight_elems"] == 3_469_760   # one call = 2 kernel launches (48 + 5 tensors)
    assert st3["elems"] - st3["weight_elems"] == 6_896_776
    assert st3["launches"] == 1 + 42 + 10 + 2
    # which element path the calibrated tables take (the flags word the prologue wrote)
    paths = bench.table_paths(m3)
    pt, pc = paths["per_tensor"], paths["per_channel_rows"]
    assert sum(pt.values()) >= 50 and sum(pc.values()) > 1000
    assert pt["scaled_one_group"] + pt["scaled_two_groups"] >= 0.9 * sum(pt.values()), paths
    assert pc["scaled_one_group"] + pc["scaled_two_groups"] >= 0.9 * sum(pc.values()), paths
