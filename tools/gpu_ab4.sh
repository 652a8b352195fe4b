#!/bin/bash
# bench-only A/B:  gpu_ab4.sh "<name>:<ENV=VAL ...>" ...
mkdir -p gpurun_out
for cfg in "$@"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-steps-api > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "== $name ($envs) rc=$?"; tail -2 gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json"))
    print("$name", round(d["value"],1), "renders/s  ms/step", round(d["ms_per_step"],2), "serial", round(d["serial_ms_per_step"],2))
    print("   serial ", {k:v["ms"] for k,v in d["stages"].items()})
    print("   2stream", d.get("stage_ms_two_stream"))
except Exception as e: print("no result", e)
PY
done
