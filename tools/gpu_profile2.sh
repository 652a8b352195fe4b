#!/bin/bash
# ncu evidence, fused path: launch list of one bench command + full capture of the hot kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --views 1 --no-cpu-baseline --no-steps-api"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/launches_r1b.out 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/launches_r1b.out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"blend_bwd|blend_fwd|render_pre|onesweep|duplicate" -s 33 -c 11 -o gpurun_out/prof_r1b -f $CMD > gpurun_out/prof_r1b.out 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/prof_r1b.out
ls -la gpurun_out/ | head -30
