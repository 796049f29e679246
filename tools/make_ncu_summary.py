"""profiles/ncu_full_targets_<tag>.json (tools/summarize_ncu.py full ...) -> profiles/ncu_summary_<tag>.json: the DRAM
traffic of the largest launch of the bench step in both layouts (bench.py's roofline.traffic).

    python tools/make_ncu_summary.py profiles/ncu_full_targets_r02.json profiles/ncu_summary_r02.json
"""
import json
import sys

src, dst = sys.argv[1], sys.argv[2]
d = json.load(open(src))
out = {"source": f"{src} (ncu --set full --clock-control none, tools/profile_targets.py 128)",
       "algorithmic_bytes_per_launch": 128 * 64 * 112 * 112 * 8, "by_memory_format": {}}
want = {"nchw": "fq_stream_kernel<0, 1, 4, 0, 1", "channels_last": "fq_stream_kernel<0, 7, 4, 0, 1"}
for fmt, key in want.items():
    ls = [l for l in d["launches"] if l["kernel"].startswith(key)]
    if not ls:
        continue
    l = ls[-1]
    mb = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    rd = l["dram__bytes_read.sum"] * mb[l["dram__bytes_read.sum__unit"]]
    wr = l["dram__bytes_write.sum"] * mb[l["dram__bytes_write.sum__unit"]]
    out["by_memory_format"][fmt] = {
        "kernel": l["kernel"] + " = fused exact-BN + ReLU + E2M5 fake-quant on [128,64,112,112] (the largest launch of the bench step)",
        "dram_read_bytes": rd, "dram_write_bytes": wr, "fq_stream_kernel_dram_bytes_per_launch": rd + wr,
        "duration_us_under_ncu": l["gpu__time_duration.sum"], "registers": l["launch__registers_per_thread"],
        "issue_slots_busy_pct": l["smsp__issue_active.avg.pct_of_peak_sustained_active"],
        "warp_instructions": l["smsp__inst_executed.sum"]}
out["note"] = ("DRAM traffic <= algorithmic bytes: every input byte is read once; part of the output is still dirty in the "
               "126 MB L2 when the kernel ends")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
