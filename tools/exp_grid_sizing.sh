#!/bin/bash
# GPU experiment (one B200): launch-shape knobs of fq_stream_kernel that need no rebuild -- tiles per CTA and whole-wave
# grids -- on the MobileNetV2 site shapes (tools/bench_cl_shapes.py) and on the bench step in both layouts.
# Writes gpurun_out/grid_sizing_<tag>.json lines.
cd "$(dirname "$0")/.."
out=gpurun_out/grid_sizing.jsonl
: > $out
for cfg in "default:" "wave:FP8FQ_WAVE_GRID=1" "tpc1:FP8FQ_TILES_PER_CTA=1" "tpc2:FP8FQ_TILES_PER_CTA=2" "tpc8:FP8FQ_TILES_PER_CTA=8" "tpc2wave:FP8FQ_TILES_PER_CTA=2 FP8FQ_WAVE_GRID=1"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs CL_JSON=cl_shapes_$tag.json python tools/bench_cl_shapes.py > /dev/null 2>&1
  for fmt in channels_last nchw; do
    line=$(env $envs python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --no-model --no-configs --memory-format $fmt 2>/dev/null | tail -1)
    python - "$tag" "$fmt" "$line" >> $out <<'PY'
import json, sys
tag, fmt, line = sys.argv[1:4]
try:
    d = json.loads(line)
    print(json.dumps({"tag": tag, "layout": fmt, "ms_per_step": d["ms_per_step"], "roofline_frac": d["roofline"]["frac"],
                      "largest": d["roofline"]["largest_launch"]["frac"]}))
except Exception as e:
    print(json.dumps({"tag": tag, "layout": fmt, "error": repr(e), "line": line[-200:]}))
PY
  done
done
cat $out
