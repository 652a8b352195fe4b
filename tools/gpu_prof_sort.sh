#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --views 1 --no-cpu-baseline --no-steps-api"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"onesweep|duplicate|keygen" -s 24 -c 8 -o gpurun_out/prof_sort -f $CMD > gpurun_out/prof_sort.out 2>&1
echo "rc=$?"; tail -2 gpurun_out/prof_sort.out
