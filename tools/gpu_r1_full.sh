#!/bin/bash
# Round-1 evidence run on one B200: GPU parity tests, both bench arms, ncu launch list of the bench
# command and one full capture of the hot kernels.  Outputs land in gpurun_out/ (copied to profiles/).
mkdir -p gpurun_out
TAG=${1:-r1c}
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_$TAG.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench reference"
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err
echo "rc=$?"; cut -c1-400 gpurun_out/bench_reference_$TAG.json
echo "== bench ours"
timeout 900 python bench.py > gpurun_out/bench_ours_$TAG.json 2> gpurun_out/bench_ours_$TAG.err
echo "rc=$?"; cut -c1-3000 gpurun_out/bench_ours_$TAG.json; tail -3 gpurun_out/bench_ours_$TAG.err
CMD="python bench.py --steps 1 --warmup 3 --views 1 --no-cpu-baseline --no-steps-api"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/launches_$TAG.out 2>&1
echo "launch list rc=$?"; tail -1 gpurun_out/launches_$TAG.out | cut -c1-300
echo "== ncu full capture (4th step = first timed step: 14 matching launches per step)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"blend_bwd|blend_fwd|render_pre|onesweep|duplicate|keygen|scan_offsets|tile_range" -s 42 -c 14 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/prof_$TAG.out 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/prof_$TAG.out | cut -c1-300
ls -la gpurun_out/ | head -40
