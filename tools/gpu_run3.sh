#!/bin/bash
# GPU check of the fused render path: new tests first, then the whole gpu suite, then both bench arms.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1; nproc >> gpurun_out/gpu_info.txt
echo "== render_sh tests"; timeout 900 python -m pytest tests/test_gpu_render_sh.py -m gpu -q --timeout=400 -p no:cacheprovider > gpurun_out/pytest_render.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/pytest_render.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu (rest)"; timeout 1200 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider --deselect tests/test_gpu_render_sh.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench ours"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?"; tail -c 4000 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
