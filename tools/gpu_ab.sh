#!/bin/bash
# A/B runs of bench.py on one B200: each line "LABEL | ENV... | bench args"; prints value, e2e and the stage table.
#   tools/gpu_ab.sh TAG "label|ENV=1|--view-chunk 4" ...
mkdir -p gpurun_out
TAG=$1; shift
for spec in "$@"; do
  IFS='|' read -r label envs bargs <<< "$spec"
  out=gpurun_out/ab_${TAG}_${label}.json
  env $envs timeout 600 python bench.py --steps 6 --no-cpu-baseline --no-steps-api --no-other-configs $bargs > $out 2> gpurun_out/ab_${TAG}_${label}.err
  python - "$out" "$label" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    st = {k: v["ms_per_step"] for k, v in d.get("stages", {}).items()}
    print(f"{sys.argv[2]:14s} value {d['value']:8.2f} e2e {d['e2e']['value']:8.2f} ms/step {d['ms_per_step']:7.3f} serial {d.get('serial_ms_per_step', 0):7.3f} | " +
          " ".join(f"{k.replace('render_preprocess', 'pre').replace('_forward', '_f').replace('_backward', '_b')}={v:.3f}" for k, v in st.items()))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
