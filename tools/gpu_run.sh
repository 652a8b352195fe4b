#!/bin/bash
# One evidence run on a B200 box (gpurun): GPU parity tests, noise floor, both bench arms, ncu launch list
# and one full capture of the hot kernels, channel sweep, gs_2d acceptance, config bench.  Outputs land in
# gpurun_out/ (summaries are copied to profiles/).
#   tools/gpu_run.sh TAG [tests] [noise] [bench] [ncu] [sweep] [gs2d] [configs]
mkdir -p gpurun_out
TAG=${1:-r2}; shift
WHAT=${*:-tests noise bench ncu sweep gs2d}
for w in $WHAT; do
case $w in
tests)
  echo "== pytest gpu"
  timeout 2400 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1
  echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
  cp gpurun_out/tolerance_report.json gpurun_out/tolerance_report_$TAG.json 2>/dev/null
  echo "== smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ;;
noise)
  echo "== noise floor"
  timeout 900 python tools/noise_floor.py --out gpurun_out/noise_floor_$TAG.json 2>&1 | tail -20 ;;
bench)
  echo "== bench reference"
  timeout 600 python bench.py --impl reference > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err
  echo "rc=$?"; cut -c1-300 gpurun_out/bench_reference_$TAG.json
  echo "== bench ours"
  timeout 1200 python bench.py > gpurun_out/bench_ours_$TAG.json 2> gpurun_out/bench_ours_$TAG.err
  echo "rc=$?"; cut -c1-1500 gpurun_out/bench_ours_$TAG.json; tail -5 gpurun_out/bench_ours_$TAG.err ;;
ncu)
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-steps-api --no-other-configs"
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/launches_$TAG.out 2>&1
  echo "launch list rc=$?"
  python tools/ncu_summary.py launches gpurun_out/launches_$TAG.csv 3 > gpurun_out/launches_$TAG.md 2>&1; cat gpurun_out/launches_$TAG.md
  echo "== ncu full capture (4th step = first timed step)"
  # per step: 1 pre_fwd, keygen + 4 depth + offsets + duplicate + 2 tile + tile_range = 10 sort kernels, 1 blend fwd, 1 blend bwd, 1 pre_bwd = 14
  timeout 1800 ncu --set full --clock-control none --import-source on -k regex:"blend_bwd|blend_fwd|render_pre|onesweep|duplicate|keygen|scan_offsets|tile_range" -s 42 -c 14 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/prof_$TAG.out 2>&1
  echo "full capture rc=$?"; tail -2 gpurun_out/prof_$TAG.out | cut -c1-300
  ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
  python tools/ncu_summary.py full gpurun_out/prof_${TAG}_raw.csv > gpurun_out/prof_$TAG.md 2>&1; cat gpurun_out/prof_$TAG.md ;;
sweep)
  echo "== channel sweep"
  timeout 900 python tools/channel_sweep.py --out gpurun_out/channel_sweep_$TAG.json > gpurun_out/channel_sweep_$TAG.log 2>&1; tail -2 gpurun_out/channel_sweep_$TAG.log | cut -c1-200 ;;
gs2d)
  echo "== gs_2d acceptance"
  timeout 600 python tools/gs2d_acceptance.py --out gpurun_out/gs2d_acceptance_$TAG.json 2>&1 | grep "ms/iteration" ;;
configs)
  echo "== config bench"
  timeout 900 python tools/config_bench.py 2 4 5 > gpurun_out/config_bench_$TAG.jsonl 2> gpurun_out/config_bench_$TAG.err; cat gpurun_out/config_bench_$TAG.jsonl | cut -c1-600 ;;
esac
done
ls -la gpurun_out/ | grep $TAG
