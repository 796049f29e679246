"""GPU A/B of the product's build options against the default build (run on the B200 box, one GPU):

    python tools/ab_build_options.py [--batch 128] [--dry-run]

For the default build and each variant in VARIANTS / VARIANTS_R2 (build options of csrc/fp8fq_kernels.cu) it
  1. builds the variant next to the default library (gpurun_out/libfp8fq_<variant>.so; nvcc is on the box),
  2. hashes the outputs of a fixed set of fused calls (both layouts, three activations, K <= 3 and K > 3 formats, special
     values) in a subprocess with FP8FQ_LIB pointing at the variant -- every variant must give the default's hashes,
  3. times the kernels (tools/bench_kernels.py) and the bench step in both layouts (bench.py --no-cpu --no-e2e
     --no-model),
and writes gpurun_out/ab_build_options.json.  Each leg is its own process, so a failing variant cannot take the others
down.  CPU-side evidence for the same options: tests/test_host_sim.py (bit equality on the host simulation),
profiles/static_build_options_r01.json (static SASS)."""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PREBUILT = os.path.join(ROOT, "build_variants")
sys.path.insert(0, ROOT)
from fp8_quantization_b200 import build as _build  # noqa: E402
VARIANTS = {"default": [], "round1": ["FP8FQ_FOLD_ACT=0", "FP8FQ_FULL_TILE=0"]}
# round 2 (FOLD_ACT and FULL_TILE became the defaults after the first A/B, profiles/ab_build_options_r02a.json): operands
# of the code select pinned in vector registers, two-wide fp32 arithmetic, the predicate-free tile body for the
# channel-innermost variants too (at 5 and at 4 resident CTAs per SM)
VARIANTS_R2 = {"pin_sel": ["FP8FQ_PIN_SEL=1"], "pack2": ["FP8FQ_PACK2=1"], "pin_pack": ["FP8FQ_PIN_SEL=1", "FP8FQ_PACK2=1"],
               "full_cl": ["FP8FQ_FULL_TILE_CL=1"], "full_cl_minb4": ["FP8FQ_FULL_TILE_CL=1", "FQ_MINB_CL=4"],
               "all": ["FP8FQ_PIN_SEL=1", "FP8FQ_PACK2=1", "FP8FQ_FULL_TILE_CL=1", "FQ_MINB_CL=4"]}
# round 2, third A/B (profiles/ab_build_options_r02i.json): resident CTAs per SM of the channel-innermost variants -- 6 for
# the constant-CTA-size instantiations became the default; "cl5" is the previous setting, "cl6_dyn" also the DYN ones at 6
VARIANTS_R2C = {"cl5": ["FQ_MINB_CL=5"], "cl6_dyn": ["FQ_MINB_CL_DYN=6"]}
# round 2, fourth A/B: the scaled-domain element path of the K > 3 formats (FP8FQ_MAGIC, default on) against the look-up
VARIANTS_R2D = {"nomagic": ["FP8FQ_MAGIC=0"], "magic_k0": ["FP8FQ_MAGIC_K0=1"],   # magic_k0: the K <= 3 formats too
                "onepath": ["FP8FQ_MAGIC_ONEPATH=1"], "selfslow": ["FP8FQ_MAGIC_SELFSLOW=1"],
                "onepath_selfslow": ["FP8FQ_MAGIC_ONEPATH=1", "FP8FQ_MAGIC_SELFSLOW=1"],
                # the element path decided once per launch (default) vs per vector; the K > 3 row kernel at 95 registers
                "hoist": ["FP8FQ_MAGIC_HOIST=1"], "rows_minb1": ["FQ_ROWS_MINB=1"],
                # everything but the scaled-domain loop out of line (default) vs inlined into every vector body
                "cold": ["FP8FQ_COLD_CALL=1"],
                # two-group tables on the look-up path (one instantiation of the scaled-domain loop in the kernels; default: both)
                "two0": ["FP8FQ_MAGIC_TWO=0"]}
FULL_BENCH = {"magic_k0", "cl5", "cl6_dyn", "default", "pin_pack", "full_cl_minb4", "all"}   # the others: kernel-level timings only


def hash_leg(device="cuda:0"):
    """Runs inside the subprocess of one variant: returns {"case": sha256 of the output}.  (``device`` exists for the CPU
    test of this tool, which runs it on the host simulation: tests/test_bench_plumbing.py.)"""
    import torch

    sys.path.insert(0, ROOT)
    import fp8_quantization_b200 as fq
    from fp8_quantization_b200 import ops

    dev = torch.device(device)
    gen = torch.Generator().manual_seed(10)
    sp = torch.tensor([0.0, -0.0, float("inf"), -float("inf"), float("nan"), 1e-30, -1e-30, 6.0, -6.0, 5.9999995, 1e30])
    res = {}

    def digest(t):
        t = t.contiguous()
        t = torch.where(torch.isnan(t), torch.full_like(t, float("nan")), t)   # NaN payloads are not part of the contract
        return hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()

    for shape in ((4, 64, 56, 56), (3, 24, 9, 5), (2, 96, 14, 14), (5, 8, 33, 33)):
        x = torch.randn(shape, generator=gen) * 3
        x.view(-1)[::97][:sp.numel()] = sp[: x.view(-1)[::97].numel()]
        r = torch.relu(torch.randn(shape, generator=gen))
        C = shape[1]
        mean, var = torch.randn(C, generator=gen).to(dev), (torch.rand(C, generator=gen) + 0.5).to(dev)
        gamma, beta = torch.randn(C, generator=gen).to(dev), torch.randn(C, generator=gen).to(dev)
        for layout in ("nchw", "channels_last"):
            fmt = torch.channels_last if layout == "channels_last" else torch.contiguous_format
            xd, rd = x.to(dev).contiguous(memory_format=fmt), r.to(dev).contiguous(memory_format=fmt)
            for M, mv in ((5, 3.0), (3, 7.5), (4, 0.4)):
                q = fq.FPQuantizer(8, mantissa_bits=M, maxval=mv)
                qi = fq.FPQuantizer(8, mantissa_bits=4, maxval=2.5)
                tb, _ = q.table_for(xd)
                ti, _ = qi.table_for(xd)
                for mode in (0, 1):
                    if mode == 1:
                        p0, p1 = ops.bn_pack(mean, var, gamma, beta, 1e-5), None
                    else:
                        p0, p1 = ops.bn_fold(mean, var, gamma, beta, 1e-5)
                    for act in (ops.ACT_NONE, ops.ACT_RELU, ops.ACT_RELU6):
                        key = f"{shape}/{layout}/M{M}/bn{mode}/act{act}"
                        res[key + "/bn_act_quant"] = digest(ops.bn_act_quant(xd, p0, p1, act, tb, float(M), 8, 1, bn_mode=mode))
                        try:
                            y = ops.bn_quant_add_act_quant(xd, rd, p0, p1, act, ti, (4.0, 8, 1), tb, (float(M), 8, 1),
                                                           bn_mode=mode)
                            res[key + "/block_tail"] = digest(y) if y is not None else "unsupported"
                        except fq.Fp8fqError as e:   # shapes the fused tail does not cover
                            res[key + "/block_tail"] = "error: " + str(e)[:40]
                for act in (ops.ACT_NONE, ops.ACT_RELU, ops.ACT_RELU6):
                    res[f"{shape}/{layout}/M{M}/act{act}/add_act_quant"] = digest(ops.add_act_quant(xd, rd, act, tb, float(M), 8, 1))
                res[f"{shape}/{layout}/M{M}/plain"] = digest(ops.fake_quant(xd, tb, 1, float(M), 8, 1))
    return res


def run(cmd, env, dry, timeout=1200):
    print("+", " ".join(cmd), flush=True)
    if dry:
        return 0, ""
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    return p.returncode, p.stdout + p.stderr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--dry-run", action="store_true")
    ap.add_argument("--round1-only", action="store_true", help="only the default build and the round-1 arithmetic")
    ap.add_argument("--build-only", action="store_true", help="cross-compile every variant into build_variants/ and stop")
    ap.add_argument("--only", default="", help="comma-separated variant names to run (default: all)")
    ap.add_argument("--hash-leg", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.hash_leg:
        print("HASHES " + json.dumps(hash_leg()))
        return
    os.makedirs(OUT, exist_ok=True)
    summary = {}
    variants = dict(VARIANTS)
    if not args.round1_only:
        variants.update(VARIANTS_R2)
        variants.update(VARIANTS_R2C)
        variants.update(VARIANTS_R2D)
    if args.only:
        variants = {k: v for k, v in variants.items() if k in set(args.only.split(",")) | {"default"}}
    for name, defines in variants.items():
        rec = summary[name] = {"defines": defines}
        env = dict(os.environ)
        if defines:
            # variants cross-compiled beforehand (python tools/ab_build_options.py --build-only, in the build container;
            # build_variants/*.so travel with the snapshot) are used as they are, if newer than the sources
            lib = os.path.join(PREBUILT, f"libfp8fq_{name}.so")
            if not (os.path.exists(lib) and os.path.getmtime(lib) >= max(os.path.getmtime(d) for d in _build.DEPS)):
                # (not into gpurun_out/: only 64 MiB of it travel back from the GPU box)
                where = PREBUILT if args.build_only else os.path.join(tempfile.gettempdir(), "fp8fq_variants")
                os.makedirs(where, exist_ok=True)
                lib = os.path.join(where, f"libfp8fq_{name}.so")
                code, log = run([sys.executable, "-m", "fp8_quantization_b200.build"] + [f"-D{d}" for d in defines]
                                + ["--out", lib], env, args.dry_run)
                if code != 0:
                    rec["build_error"] = log[-500:]
                    continue
            rec["lib"] = os.path.relpath(lib, ROOT)
            env["FP8FQ_LIB"] = lib
        if args.build_only:
            continue
        code, log = run([sys.executable, os.path.abspath(__file__), "--hash-leg"], env, args.dry_run)
        line = next((ln for ln in log.split("\n") if ln.startswith("HASHES ")), None)
        rec["hashes"] = json.loads(line[7:]) if line else None
        if line is None and not args.dry_run:
            rec["hash_error"] = log[-500:]
        env["KERNELS_JSON"] = f"kernels_{name}.json"
        code, log = run([sys.executable, "tools/bench_kernels.py", str(args.batch)], env, args.dry_run)
        rec["bench_kernels"] = "ok" if code == 0 else log[-300:]
        env["CL_JSON"] = f"cl_shapes_{name}.json"
        code, log = run([sys.executable, "tools/bench_cl_shapes.py", str(args.batch)], env, args.dry_run)
        rec["bench_cl_shapes"] = "ok" if code == 0 else log[-300:]
        for layout in (("channels_last", "nchw") if name in FULL_BENCH else ()):
            code, log = run([sys.executable, "bench.py", "--steps", "30", "--warmup", "5", "--no-cpu", "--no-e2e",
                             "--no-model", "--no-configs", "--memory-format", layout, "--batch", str(args.batch)], env, args.dry_run)
            line = next((ln for ln in reversed(log.split("\n")) if ln.startswith("{")), None)
            try:
                d = json.loads(line)
                rec[f"bench_{layout}"] = {"ms_per_step": d["ms_per_step"], "value": d["value"],
                                          "roofline_frac": d["roofline"]["frac"],
                                          "largest_launch_frac": d["roofline"]["largest_launch"]["frac"]}
            except (TypeError, ValueError, KeyError):
                rec[f"bench_{layout}"] = {"error": log[-300:]}
    if args.build_only:
        print(json.dumps(summary, indent=1))
        return
    base = summary["default"].get("hashes")
    for name, rec in summary.items():
        h = rec.pop("hashes", None)
        if base and h:
            bad = sorted(k for k in base if h.get(k) != base[k])
            rec["parity_vs_default"] = {"cases": len(base), "mismatches": len(bad), "first": bad[:5]}
    json.dump(summary, open(os.path.join(OUT, "ab_build_options.json"), "w"), indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
