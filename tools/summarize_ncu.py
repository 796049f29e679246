"""Turns ncu outputs brought back in gpurun_out/ into the small, tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_r01.csv profiles/launches_r01_summary.json [steps]
    python tools/summarize_ncu.py full gpurun_out/prof_k1_r01.ncu-rep profiles/ncu_full_k1_r01.json
"""
import collections
import csv
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        # round 2: per-pipe utilisation (which unit is the busiest one besides DRAM)
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct"]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("<unnamed>::", "").strip()


def launches(path, out, steps_tail=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, mi, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "Grid Size", "Block Size"))
    data = [(short(row[ki]), float(row[vi].replace(",", "")), row[gi], row[bi]) for row in r
            if len(row) > vi and row[mi] == "gpu__time_duration.sum"]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v, _, _ in data:
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    summary = {"source": path, "note": "ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, "
                                       "serialised launches: compare SHARES, not absolutes",
               "n_launches": len(data), "total_us": tot / 1e3,
               "by_kernel": {k: {"launches": v[0], "total_us": v[1] / 1e3, "share": v[1] / tot}
                             for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}}
    json.dump(summary, open(out, "w"), indent=1)
    print(json.dumps(summary, indent=1))


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {"kernel": short(r[idx["Kernel Name"]]), "grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]]}
        for k in KEEP:
            if k in idx:
                try:
                    d[k] = float(r[idx[k]].replace(",", ""))
                except ValueError:
                    d[k] = r[idx[k]]
                d[k + "__unit"] = units[idx[k]]
        res.append(d)
    json.dump({"source": path, "note": "ncu --set full --clock-control none --import-source on", "launches": res},
              open(out, "w"), indent=1)
    print(len(res), "launches summarised ->", out)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3])
