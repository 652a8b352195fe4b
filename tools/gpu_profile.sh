#!/bin/bash
# ncu evidence for round 1: launch list of one bench command + full capture of the hot kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --views 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/launches_r1.out 2>&1
echo "launch list rc=$?"; tail -3 gpurun_out/launches_r1.out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"blend_bwd|blend_fwd|onesweep|duplicate|sh_kernel|tile_range" -s 44 -c 11 -o gpurun_out/prof_r1 -f $CMD > gpurun_out/prof_r1.out 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/prof_r1.out
ls -la gpurun_out/
