#!/bin/bash
# A/B bench lines (tools/gpu_ab.sh specs) followed by ONE `ncu --set full` capture, with source, of the first launch of a
# kernel in the first timed step of a 1-step bench run.  Read the report here with tools/sass_hotspots.py (instruction
# and stall-sample shares per source line) or `ncu -i ... --page source --csv --print-source sass`.
#   tools/gpu_kernel_prof.sh TAG KERNEL_REGEX "label|ENV=..|bench args" ...
mkdir -p gpurun_out
TAG=$1; KERN=$2; shift 2
bash tools/gpu_ab.sh $TAG "$@"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-steps-api --no-other-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KERN" -s 3 -c 1 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/prof_$TAG.out 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/prof_$TAG.out | cut -c1-200
