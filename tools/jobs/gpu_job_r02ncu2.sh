#!/bin/bash
# ncu of the same one-group E3M4 launches in the default build and in the build that also carries the two-group loop
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
for v in default two1; do
  if [ $v = default ]; then unset FP8FQ_LIB; else export FP8FQ_LIB=$PWD/build_variants/libfp8fq_$v.so; fi
  CL_MAXVAL=3.0 timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:fq_stream -o gpurun_out/prof_two_$v -f python tools/profile_targets_mbv2.py > gpurun_out/ncu_two_$v.log 2>&1
  python tools/summarize_ncu.py full gpurun_out/prof_two_$v.ncu-rep gpurun_out/ncu_full_two_$v.json > /dev/null 2>&1 && rm -f gpurun_out/prof_two_$v.ncu-rep
done
echo done
