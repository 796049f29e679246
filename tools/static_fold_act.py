"""Static comparison (no GPU) of the default build with -DFP8FQ_FOLD_ACT=1 (csrc/fp8fq_kernels.cu: ReLU / ReLU6 folded
into the quantiser's clamp): registers and SASS instruction counts (total, FMNMX) of every fq_stream_kernel
instantiation that carries an activation.  Static counts, not executed counts -- the loop bodies are fully unrolled
(VEC x kUnroll elements), so the static body size tracks the per-tile work.  Writes profiles/static_fold_act_<tag>.json.

    python tools/static_fold_act.py [tag]
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
    regs, cur = {}, None
    for line in res.stderr.split("\n"):
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
        if m:
            cur = m.group(1)
        m = re.search(r"Used (\d+) registers", line)
        if m and cur:
            regs[cur] = int(m.group(1))
        m = re.search(r"(\d+) bytes spill stores", line)
        if m and cur and int(m.group(1)):
            regs[cur] = (regs.get(cur), f"spill {m.group(1)} B")
    sass = subprocess.run(["cuobjdump", "-sass", out], capture_output=True, text=True).stdout
    kernels = {}
    for blk in re.split(r"\n\s*Function : ", sass)[1:]:
        name = blk.split("\n", 1)[0].strip()
        ops = collections.Counter()
        for line in blk.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                ops[m.group(1).split(".")[0]] += 1
        kernels[name] = ops
    return regs, kernels


def demangle(names):
    p = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return [re.sub(r"\(anonymous namespace\)::|\(.*$", "", n) for n in p.stdout.split("\n")]


r0, k0 = build([], "/tmp/libfp8fq_fold0.so")
r1, k1 = build(["-DFP8FQ_FOLD_ACT=1"], "/tmp/libfp8fq_fold1.so")
names = sorted(k0)
rows = []
for mangled, name in zip(names, demangle(names)):
    a, c = k0[mangled], k1[mangled]
    if "fq_stream_kernel" not in name or (a == c and r0.get(mangled) == r1.get(mangled)):
        continue
    rows.append({"kernel": name, "registers": [r0.get(mangled), r1.get(mangled)],
                 "instructions": [sum(a.values()), sum(c.values())], "FMNMX": [a["FMNMX"], c["FMNMX"]],
                 "FSEL": [a["FSEL"], c["FSEL"]], "ISETP": [a["ISETP"], c["ISETP"]]})
unchanged = sum(1 for m in names if k0[m] == k1[m])
tot0 = sum(r["instructions"][0] for r in rows)
tot1 = sum(r["instructions"][1] for r in rows)
out = {"what": "default build vs -DFP8FQ_FOLD_ACT=1; [default, folded] per changed fq_stream_kernel instantiation "
               "<KMODE, PRE, VEC, CODES, BNM, DYN>",
       "kernels_total": len(names), "kernels_unchanged": unchanged, "kernels_changed": len(rows),
       "static_instructions_changed_kernels": [tot0, tot1], "static_reduction": 1 - tot1 / max(tot0, 1),
       "FMNMX_changed_kernels": [sum(r["FMNMX"][0] for r in rows), sum(r["FMNMX"][1] for r in rows)],
       "rows": rows}
path = os.path.join(ROOT, "profiles", f"static_fold_act_{tag}.json")
json.dump(out, open(path, "w"), indent=1)
print(path)
print({k: v for k, v in out.items() if k != "rows"})
for r in rows:
    print(r["kernel"][:70], r["registers"], r["instructions"], "FMNMX", r["FMNMX"])
