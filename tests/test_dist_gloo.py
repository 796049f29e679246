"""CPU, world_size 2, gloo: the data-parallel plumbing.  The per-rank statistics come from
the oracle here (the CUDA kernels are covered by -m gpu); what is tested is that after the collectives every
rank holds exactly the range / metrics a single process computes on the concatenated batch."""
import os
import sys

import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import fp8_quantization_b200 as fq
    from fp8_quantization_b200 import dist as fq_dist
    from oracle import fp8_oracle as O

    assert fq_dist.init_from_env(backend="gloo") and fq_dist.active()
    assert fq_dist.world_size() == world and fq_dist.rank() == rank
    torch.manual_seed(0)
    ok = True
    for cls, ocls, pc in ((fq.AllMinMaxEstimator, O.OracleAllMinMax, False),
                          (fq.CurrentMinMaxEstimator, O.OracleCurrentMinMax, True),
                          (fq.RunningMinMaxEstimator, O.OracleRunningMinMax, False)):
        est = cls(per_channel=pc)
        single = ocls(per_channel=pc)
        for step in range(3):
            gx = torch.randn(8, 6, 10) * (step + 1)           # the global batch, identical on every rank
            if pc:
                gx_shards = gx.reshape(6 * 8 // 6, 6, 10) if False else gx  # weights are replicated, not sharded
                local = gx
            else:
                local = fq_dist.shard_batch(gx)
            mn, mx = O.minmax(local, pc)
            packed = torch.cat([mn.reshape(-1), mx.reshape(-1)]).clone()
            cur_min, cur_max = est.dp_merge(packed)
            smn, smx = single(gx)
            ok &= torch.equal(cur_min.reshape(-1), smn.reshape(-1)) and torch.equal(cur_max.reshape(-1), smx.reshape(-1))
    # MSE table: mean over ranks of per-shard means == global mean (equal shards)
    gx = torch.randn(8, 4, 5, 5)
    local = fq_dist.shard_batch(gx)
    inc = ((local - local.round()) ** 2).mean().reshape(1)
    fq_dist.all_reduce_mean(inc)
    ok &= torch.allclose(inc, ((gx - gx.round()) ** 2).mean().reshape(1), rtol=1e-6)
    # validation counters
    stats = torch.tensor([1.0 + rank, 2.0, 3.0, 4.0])
    fq_dist.all_reduce_sum(stats)
    ok &= stats.tolist() == [3.0, 4.0, 6.0, 8.0]
    # BN re-estimation: ranks see half the batch each, statistics must equal the single-process ones on
    # the concatenated batch (quantisers off here -- the CUDA quantisers are covered by -m gpu)
    from fp8_quantization_b200 import modules, workloads
    torch.manual_seed(1)
    qp = workloads.readme_quant_params(5)
    qp.pop("quant_setup")
    seq = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, bias=False), torch.nn.BatchNorm2d(8), torch.nn.ReLU(),
                              torch.nn.Conv2d(8, 4, 1, bias=False), torch.nn.BatchNorm2d(4))

    class Wrap(modules.QuantizedModel):
        def __init__(self, f):
            super().__init__((1, 3, 8, 8))
            self.f = f

        def forward(self, x):
            return self.f(x)

    import copy
    dp_model = Wrap(modules.quantize_model(copy.deepcopy(seq), **qp))
    sp_model = Wrap(modules.quantize_model(copy.deepcopy(seq), **qp))
    dp_model.full_precision()
    sp_model.full_precision()
    gxs = [torch.randn(8, 3, 8, 8) for _ in range(3)]
    workloads.reestimate_BN_stats(dp_model, [fq_dist.shard_batch(g) for g in gxs], 3)   # global-batch statistics
    fq_dist.enable(False)
    workloads.reestimate_BN_stats(sp_model, gxs, 3)                                      # single process, whole batch
    fq_dist.enable(True)
    for a, b in zip(dp_model.f, sp_model.f):
        ok &= torch.allclose(a.running_mean, b.running_mean, rtol=1e-5, atol=1e-6)
        ok &= torch.allclose(a.running_var, b.running_var, rtol=1e-4, atol=1e-6)
    with pytest.raises(ValueError):
        fq_dist.shard_batch(torch.zeros(3, 2))
    fq_dist.barrier()
    q.put((rank, bool(ok)))
    td.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def _sim_worker(rank, world, port, q):
    """BASELINE config 5 in miniature, through the product's real path on the host simulation of the kernels: the
    quantised ResNet-18 calibrated on a batch sharded over the ranks (one MAX all-reduce of [-min, max] per activation
    quantiser, incl. the fused calibration epilogue's statistics-only launch), validate counters summed across ranks,
    and the MSE estimator's MAX / MEAN exchanges -- against a single process on the concatenated batch."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "host_sim"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    from harness import install_simulation
    install_simulation()
    import numpy as np
    from torchvision.models import resnet18

    import fp8_quantization_b200 as fq
    from fp8_quantization_b200 import dist as fq_dist
    from fp8_quantization_b200 import workloads

    assert fq_dist.init_from_env(backend="gloo") and fq_dist.active()
    report = {}

    def build():
        torch.manual_seed(10)
        return workloads.QuantizedResNet(resnet18(), **workloads.readme_quant_params(5)).eval()

    gen = torch.Generator().manual_seed(10)
    gx = torch.randn(4, 3, 64, 64, generator=gen)        # the global batch, identical on every rank
    labels = torch.tensor([1, 2, 3, 4])
    local = fq_dist.shard_batch(gx)
    dp = build()
    workloads.pass_data_for_range_estimation([local], dp, True, True, 1)   # collectives inside
    dp.fix_ranges()
    m_dp = workloads.validate(dp, [local], [fq_dist.shard_batch(labels)])  # SUM of the four counters
    fq_dist.enable(False)
    sp = build()
    workloads.pass_data_for_range_estimation([gx], sp, True, True, 1)      # single process, whole batch
    sp.fix_ranges()
    m_sp = workloads.validate(sp, [gx], [labels])
    fq_dist.enable(True)
    qs_dp = [(n, m) for n, m in dp.named_modules() if isinstance(m, fq.FPQuantizer)]
    qs_sp = dict((n, m) for n, m in sp.named_modules() if isinstance(m, fq.FPQuantizer))
    exact = close = 0
    for n, m in qs_dp:
        a, b = m.maxval.reshape(-1), qs_sp[n].maxval.reshape(-1)
        if torch.equal(a, b):
            exact += 1
        elif torch.allclose(a, b, rtol=2e-2):
            close += 1      # the convolution library may sum in another order at another batch size -> a tie flips
        else:
            report["range_mismatch"] = n
    report["ranges"] = (exact, close, len(qs_dp))
    # validate() sums its counters whenever a process group exists: the sharded run counts the 4 images once, the
    # "single process" leg (every rank evaluating the whole batch) counts them once per rank; the means must agree
    report["metrics"] = (m_dp["count"] == 4 and m_sp["count"] == 4 * world
                         and abs(m_dp["loss"] - m_sp["loss"]) < 1e-3 * abs(m_sp["loss"])
                         and m_dp["top_1_accuracy"] == m_sp["top_1_accuracy"]
                         and m_dp["top_5_accuracy"] == m_sp["top_5_accuracy"])
    report["metric_values"] = (m_dp, m_sp)
    # every rank ends with the same ranges (they came out of the same all-reduces)
    flat = torch.cat([m.maxval.reshape(-1) for _, m in qs_dp])
    other = flat.clone()
    fq_dist.all_reduce_max(other)
    report["ranks_agree"] = bool(torch.equal(flat, other))

    # FP_MSE_Estimator under DP: MAX of absmax defines the grid, MEAN of the per-shard MSE tables
    g2 = torch.Generator().manual_seed(3)
    gt = torch.randn(8, 16, 6, 6, generator=g2) * 2
    q_dp = fq.FPQuantizer(8, mantissa_bits=4, set_maxval=True, mse_include_mantissa_bits=True)
    e_dp = fq.FP_MSE_Estimator(quantizer=q_dp)
    mn_dp, mx_dp = e_dp(fq_dist.shard_batch(gt))
    fq_dist.enable(False)
    q_sp = fq.FPQuantizer(8, mantissa_bits=4, set_maxval=True, mse_include_mantissa_bits=True)
    e_sp = fq.FP_MSE_Estimator(quantizer=q_sp)
    mn_sp, mx_sp = e_sp(gt)
    fq_dist.enable(True)
    report["mse"] = (bool(torch.equal(e_dp.search_grid, e_sp.search_grid))
                     and bool(np.allclose(e_dp.mses.numpy(), e_sp.mses.numpy(), rtol=1e-5, atol=1e-12))
                     and q_dp._mbits_host == q_sp._mbits_host and bool(torch.equal(mx_dp, mx_sp)))
    # bench.py's data-parallel legs (the only multi-GPU code the driver executes): calibration timing + accounting,
    # and dp_parity -- all-gathered ranges bit-equal across ranks, rank 0 re-calibrating on the gathered batch
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    cal = bench.calibration_leg(dp, local, workloads, fq_dist, world, reps=1)
    par = bench.dp_parity_leg(dp, local, lambda fmt: build(), "nchw", fq, workloads, fq_dist, world, rank)
    report["bench_calibration"] = cal
    report["bench_dp_parity"] = par
    fq_dist.barrier()
    q.put((rank, report))
    td.destroy_process_group()


def test_world_size_2_gloo_sharded_calibration_through_the_simulated_kernels(built):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_sim_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        r = res[rank]
        assert "range_mismatch" not in r, r
        exact, close, total = r["ranges"]
        # 21 weight + 29 activation quantisers (the tied one is shared); measured here: all 50 bit-identical
        assert total == 50 and exact + close == total and exact >= 40, r["ranges"]
        assert r["metrics"] and r["ranks_agree"] and r["mse"], r
        cal, par = r["bench_calibration"], r["bench_dp_parity"]
        assert cal["allreduces"] == 29 and cal["ms"] > 0 and cal["us_per_allreduce"] is not None, cal
        assert par["ranges"] == 50 and par["bit_equal_across_ranks"], par
        if rank == 0:
            vs = par["vs_single_process_on_gathered_batch"]
            assert vs["weights_bit_equal"] and vs["first_layer_bit_equal"] and vs["weight_ranges"] == 21, vs
            assert vs["all_bit_equal"] >= 40 and vs["max_rel_diff"] < 2e-2 and vs["global_batch"] == 4, vs
