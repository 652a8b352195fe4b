#!/bin/bash
# GPU tests, then racecheck over the preprocess / sort / blend tests, then one bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout=1400 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
SEL="render_sh_equals_steps or views_equals_single or grad_sync or sort_views or sort_known or sort_edge or sort_compaction or fused_adam or rasterization_fused"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q --timeout=1400 -p no:cacheprovider -k "$SEL" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_racecheck.log | tail -3
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_ours.json
