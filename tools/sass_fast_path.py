"""Static estimate (no GPU) of the instructions one thread EXECUTES on the fast path of a kernel: walks the SASS of a
function from its entry, skipping the never-taken slow paths (forward branches over a region that contains the IEEE
division sequence MUFU.RCP / CALL but no store), taking unconditional branches, stopping at the first backward branch
or EXIT.  Prints, per build (a .so) and kernel-name substring: registers are in the ptxas log, this gives the executed
count and its opcode histogram -- the number that the issue-bound kernels' time follows.

    python tools/sass_fast_path.py lib.so 'fq_stream_kernelILi0ELi1ELi4ELb0ELi1ELb0E' [more substrings...]
"""
import collections
import re
import subprocess
import sys


def functions(so):
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    out = {}
    for blk in re.split(r"\n\s*Function : ", sass)[1:]:
        name = blk.split("\n", 1)[0].strip()
        ins = []
        for line in blk.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        out[name] = ins
    return out


def walk(ins, max_steps=20000, take=(), trace=None):
    """``take``: addresses of conditional branches to follow (e.g. the uniform branch that selects the cl_same path)."""
    addr2i = {a: i for i, (a, _) in enumerate(ins)}
    i, n, hist, seen_back = 0, 0, collections.Counter(), False
    while i < len(ins) and n < max_steps:
        a, text = ins[i]
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*)", text)
        pred, op, rest = m.group(1), m.group(2), m.group(3)
        n += 1
        hist[op.split(".")[0]] += 1
        if trace is not None:
            trace.append((a, text))
        if op.startswith("EXIT") and not pred:
            break
        if op.startswith("BRA"):
            t = re.findall(r"0x([0-9a-f]+)", rest)
            tgt = int(t[-1], 16) if t else None
            cond = bool(pred) or op.startswith("BRA.U") or re.search(r"\bU?P\d", rest.split(",")[0] if "," in rest else "")
            if tgt is None or tgt not in addr2i:
                i += 1
                continue
            if tgt <= a:
                break  # loop back edge: one trip is enough
            region = [x for _, x in ins[i + 1:addr2i[tgt]]]
            slow = any(("MUFU.RCP" in x or re.search(r"\bCALL", x)) for x in region) and not any("STG" in x for x in region)
            if not cond or slow or a in take:
                i = addr2i[tgt]
                continue
        i += 1
    return n, hist


if __name__ == "__main__":
    so, subs = sys.argv[1], sys.argv[2:]
    fs = functions(so)
    for sub in subs:
        for name, ins in fs.items():
            if sub in name:
                n, hist = walk(ins)
                print(f"{sub}: executed~{n} static={len(ins)} " + " ".join(f"{k}={v}" for k, v in hist.most_common(14)))
