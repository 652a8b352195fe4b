#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -60 gpurun_out/pytest_gpu.log
echo "== bench ours (short)"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?"; tail -c 3000 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
echo "== bench reference (short)"; timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
