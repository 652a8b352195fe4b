#!/bin/bash
# A/B of kernel variants selected by environment switches: sort/parity tests first, then one
# bench per config with the per-stage table; optional ncu launch list of the last config.
#   usage: gpu_ab2.sh "<name>:<ENV=VAL ...>" ...
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout=400 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for cfg in "$@"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  echo "== bench $name ($envs)"; env $envs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-steps-api > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "rc=$?"; tail -3 gpurun_out/bench_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$name.json"))
print("$name", round(d["value"],1), "renders/s", {k:v["ms"] for k,v in d["stages"].items()})
PY
done
bash tools/gpu_launches.sh
