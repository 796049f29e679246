"""Static evidence for the kernels, produced without a GPU (B200_PROFILING.md: check `-Xptxas -v` and `cuobjdump -sass`
before spending GPU time): registers / spills / shared memory per kernel instantiation from ptxas, and the SASS
instruction mix of every kernel in libfp8fq.so (128-bit global accesses, evict-first loads, NaN-propagating min/max,
MUFU / division instructions on or off the streaming path).  Writes profiles/static_sass_<tag>.json.

    python tools/static_report.py [tag]
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
out_so = "/tmp/libfp8fq_static.so"
env = dict(os.environ)
env.pop("CC", None)
res = subprocess.run([b.find_nvcc()] + b.NVCC_FLAGS + ["-Xptxas", "-v", "-o", out_so, b.SRC], capture_output=True,
                     text=True, env=env)
assert res.returncode == 0, res.stderr[-2000:]


def demangle(names):
    p = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return p.stdout.split("\n")


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    return re.sub(r"\(.*$", "", name)


ptxas = {}
cur = None
for line in res.stderr.split("\n"):
    m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", line)
    if m and cur:
        ptxas.setdefault(cur, {}).update(registers=int(m.group(1)), smem_bytes=int(m.group(3) or 0))
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur:
        ptxas.setdefault(cur, {}).update(stack_bytes=int(m.group(1)), spill_store_bytes=int(m.group(2)),
                                         spill_load_bytes=int(m.group(3)))

sass = subprocess.run(["cuobjdump", "-sass", out_so], capture_output=True, text=True).stdout
kernels = {}
for blk in re.split(r"\n\s*Function : ", sass)[1:]:
    name = blk.split("\n", 1)[0].strip()
    ops = collections.Counter()
    for line in blk.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(1)] += 1
    tot = sum(ops.values())

    def count(prefix):
        return sum(v for k, v in ops.items() if k.startswith(prefix))

    def has(prefix, part):
        return sum(v for k, v in ops.items() if k.startswith(prefix) and part in k)

    kernels[name] = {
        "sass_instructions": tot,
        "ldg_total": count("LDG"), "ldg_128": has("LDG", ".128"), "ldg_evict_first": has("LDG", ".EF"),
        "stg_total": count("STG"), "stg_128": has("STG", ".128"),
        "fmnmx_nan": has("FMNMX", ".NAN"),
        "mufu": {k.split(".", 1)[1]: v for k, v in ops.items() if k.startswith("MUFU.")},
        "lds": count("LDS"), "sts": count("STS"), "bar": count("BAR"), "shfl": count("SHFL"),
        "local_memory": count("LDL") + count("STL"),
    }
names = sorted(set(kernels) | set(ptxas))
dm = dict(zip(names, demangle(names)))
rows = []
for n in names:
    r = {"kernel": short(dm[n])}
    r.update(ptxas.get(n, {}))
    r.update(kernels.get(n, {}))
    rows.append(r)
rows.sort(key=lambda r: r["kernel"])
summary = {
    "how": "nvcc " + " ".join(b.NVCC_FLAGS) + " -Xptxas -v; cuobjdump -sass (build container, no GPU)",
    "kernels": len(rows),
    "max_registers": max(r.get("registers", 0) for r in rows),
    "kernels_with_spills": sorted({r["kernel"] for r in rows if r.get("spill_store_bytes", 0) or r.get("spill_load_bytes", 0)
                                   or r.get("local_memory", 0)}),
    "stream_kernels_with_shared_memory_or_barriers": sorted({r["kernel"] for r in rows if "fq_stream_kernel" in r["kernel"]
                                                             and (r.get("lds", 0) or r.get("bar", 0))}),
    # MUFU.RCP is the seed of the IEEE division (__fdiv_rn) that only the tie-guard fallback executes, and the single
    # `MUFU.RSQ R, -QNAN` per kernel is the NaN generator of that division subroutine's 0/0 / inf/inf case: the streaming
    # fast path has no log2 / exp2 / division.  EX2 belongs to the prologue (powf), real RSQ to the
    # batch-norm parameter kernels.
    "mufu_kinds_in_stream_kernels": sorted({k for r in rows if "fq_stream_kernel" in r["kernel"] for k in r.get("mufu", {})}),
    "mufu_kinds_in_prologue_kernels": sorted({k for r in rows if "prepare_kernel" in r["kernel"] for k in r.get("mufu", {})}),
    "stream_kernels_all_use_128bit_access_when_vec4": all(
        r.get("ldg_128", 0) > 0 and r.get("stg_128", 0) > 0 for r in rows
        if re.search(r"fq_stream_kernel<\d, \d, 4,", r["kernel"])),
    "spill_bytes": {r["kernel"]: r.get("spill_store_bytes", 0) for r in rows if r.get("spill_store_bytes", 0)},
}
path = os.path.join(ROOT, "profiles", f"static_sass_{tag}.json")
json.dump({"summary": summary, "per_kernel": rows}, open(path, "w"), indent=1)
print(json.dumps(summary, indent=1))
print("wrote", path, "(", len(rows), "kernels )")
