#!/bin/bash
# A/B of the blend kernels (v1 butterfly vs v2 deferred smem reduction, K=3 / K=2) + parity tests
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for cfg in "v2k3:MSB_BLEND_V1=0" "v2k2:MSB_BWD_K=2" "v1:MSB_BLEND_V1=1"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  echo "== bench $name"; env $envs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-steps-api > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$name.json"))
print("$name", round(d["value"],1), "renders/s", {k:v["ms"] for k,v in d["stages"].items()})
PY
done
