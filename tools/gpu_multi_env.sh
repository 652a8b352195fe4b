#!/bin/bash
# Multi-GPU A/B with environment variables: tools/gpu_multi_env.sh TAG N "label|ENV=..|bench args" ...
mkdir -p gpurun_out
TAG=$1; N=$2; shift 2
i=0
for spec in "$@"; do
  IFS='|' read -r label envs bargs <<< "$spec"
  i=$((i+1))
  out=gpurun_out/multi_${TAG}_${label}_n${N}.json
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+i)) \
      bench.py --gpus $N --no-cpu-baseline --no-steps-api --no-other-configs $bargs > $out 2> ${out%.json}.err
  python - "$out" "$label" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    st = {k: v["ms_per_step"] for k, v in d.get("stages", {}).items()}
    print(f"{sys.argv[2]:12s} value {d['value']:9.2f} e2e {d['e2e']['value']:9.2f} ms/step {d['ms_per_step']:8.3f} | " +
          " ".join(f"{k.replace('render_preprocess', 'pre').replace('_forward', '_f').replace('_backward', '_b')}={v:.3f}" for k, v in st.items()))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
