#!/bin/bash
# quick check: render_sh tests + fused bench stage table
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render_sh.py tests/test_gpu_parity.py -m gpu -q -x --timeout=400 -p no:cacheprovider -k "${KEXPR:-render_sh or sort or raster}" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_quick.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-steps-api > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "rc=$?"; tail -3 gpurun_out/bench_quick.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_quick.json"))
print(round(d["value"],1), "renders/s e2e", round(d["e2e"]["value"],1), {k:(v["ms"], v.get("hbm_frac")) for k,v in d["stages"].items()})
PY
