#!/bin/bash
# A/B of the backward blend + one `ncu --set full` capture of it (first timed step of a 1-step bench run).
mkdir -p gpurun_out
TAG=$1; shift
bash tools/gpu_ab.sh $TAG "$@"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-steps-api --no-other-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_bwd" -s 3 -c 1 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/prof_$TAG.out 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/prof_$TAG.out | cut -c1-200
