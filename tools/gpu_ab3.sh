#!/bin/bash
# A/B with the parity tests run UNDER each variant:  gpu_ab3.sh "<name>:<ENV=VAL ...>" ...
mkdir -p gpurun_out
for cfg in "$@"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  echo "== pytest $name ($envs)"; env $envs timeout 600 python -m pytest tests -m gpu -q -x --timeout=200 -p no:cacheprovider > gpurun_out/pytest_$name.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$name.log
  echo "== bench $name"; env $envs timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-steps-api > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "rc=$?"; tail -3 gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json"))
    print("$name", round(d["value"],1), "renders/s", {k:v["ms"] for k,v in d["stages"].items()})
except Exception as e: print("no result", e)
PY
done
