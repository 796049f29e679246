"""Static comparison (no GPU) of the default build with the product's build options (csrc/fp8fq_kernels.cu):
  fold_act  : -DFP8FQ_FOLD_ACT=1   ReLU / ReLU6 folded into the quantiser's clamp bounds
  full_tile : -DFP8FQ_FULL_TILE=1  second, predicate-free instantiation of the stream kernel's tile body for full tiles
  both      : the two together (the default build since round 2);  neither: the round-1 default
(executed fast-path instruction counts: tools/sass_fast_path.py)
Per fq_stream_kernel instantiation: registers, spill bytes, static SASS instruction count, FMNMX count, and the size of
each TILE BODY (from the first 128-bit data load of a body to its last 128-bit store; a build with full_tile has two
bodies, the first being the predicate-free one that every tile but the last executes).  Static sizes include the
never-taken IEEE-division slow paths (~96 instructions per vector), so differences between bodies are the executed
differences.  Writes profiles/static_build_options_<tag>.json.

    python tools/static_build_options.py [tag]
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fp8_quantization_b200 import build as b  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
env = dict(os.environ)
env.pop("CC", None)


def build(extra, out):
    res = subprocess.run([b.find_nvcc()] + b.NVCC_FLAGS + extra + ["-Xptxas", "-v", "-o", out, b.SRC],
                         capture_output=True, text=True, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    regs, spills, cur = {}, {}, None
    for line in res.stderr.split("\n"):
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
        if m:
            cur = m.group(1)
        m = re.search(r"Used (\d+) registers", line)
        if m and cur:
            regs[cur] = int(m.group(1))
        m = re.search(r"(\d+) bytes spill stores", line)
        if m and cur:
            spills[cur] = int(m.group(1))
    sass = subprocess.run(["cuobjdump", "-sass", out], capture_output=True, text=True).stdout
    kernels = {}
    for blk in re.split(r"\n\s*Function : ", sass)[1:]:
        name = blk.split("\n", 1)[0].strip()
        ops, seq = collections.Counter(), []
        for line in blk.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                ops[m.group(2).split(".")[0]] += 1
                seq.append(m.group(2))
        # tile bodies: first LDG.E.EF.128 (evict-first data load) after the previous body's last STG.E.128
        bodies, start_i, last_st = [], None, None
        for i, op in enumerate(seq):
            if op.startswith("LDG.E.EF.128"):
                if start_i is not None and last_st is not None:
                    bodies.append(last_st - start_i + 1)
                    start_i, last_st = None, None
                if start_i is None:
                    start_i = i
            elif op.startswith("STG.E.128") and start_i is not None:
                last_st = i
        if start_i is not None and last_st is not None:
            bodies.append(last_st - start_i + 1)
        kernels[name] = {"ops": ops, "bodies": bodies, "registers": regs.get(name), "spill_store_bytes": spills.get(name, 0)}
    return kernels


def demangle(names):
    p = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return [re.sub(r"\(anonymous namespace\)::|\(.*$", "", n) for n in p.stdout.split("\n")]


# explicit flags for every variant, so that the table means the same whatever the product's current defaults are
# (round 1: both off; since round 2: both on)
variants = {"default": [], "neither": ["-DFP8FQ_FOLD_ACT=0", "-DFP8FQ_FULL_TILE=0"],
            "fold_act": ["-DFP8FQ_FOLD_ACT=1", "-DFP8FQ_FULL_TILE=0"], "full_tile": ["-DFP8FQ_FOLD_ACT=0", "-DFP8FQ_FULL_TILE=1"],
            "both": ["-DFP8FQ_FOLD_ACT=1", "-DFP8FQ_FULL_TILE=1"]}
built = {v: build(flags, f"/tmp/libfp8fq_opt_{v}.so") for v, flags in variants.items()}
names = sorted(built["default"])
rows = []
for mangled, name in zip(names, demangle(names)):
    if "fq_stream_kernel" not in name or ", 4, false," not in name:   # the 128-bit, no-code-plane instantiations
        continue
    row = {"kernel": name.replace("void ", "")}
    for v in variants:
        k = built[v][mangled]
        row[v] = {"registers": k["registers"], "spill_store_bytes": k["spill_store_bytes"],
                  "instructions": sum(k["ops"].values()), "FMNMX": k["ops"]["FMNMX"], "tile_bodies": k["bodies"]}
    rows.append(row)
out = {"what": "static SASS of the 128-bit fq_stream_kernel<KMODE, PRE, VEC, CODES, BNM, DYN> instantiations per build "
               "variant; tile_bodies = instructions from a body's first data load to its last store (full_tile: "
               "[predicate-free body, bounded body])",
       "variants": {v: " ".join(f) or "(default)" for v, f in variants.items()}, "rows": rows}
path = os.path.join(ROOT, "profiles", f"static_build_options_{tag}.json")
json.dump(out, open(path, "w"), indent=1)
print(path)
for r in rows:
    print(r["kernel"][:52].ljust(52), " | ".join(f"{v}: {r[v]['tile_bodies']} r{r[v]['registers']} s{r[v]['spill_store_bytes']}"
                                                  for v in variants))
