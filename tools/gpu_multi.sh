#!/bin/bash
# Multi-GPU bench lines on one box: tools/gpu_multi.sh TAG "N|bench args[|ENV=.. ENV=..]" ...   (N = 1 runs without torchrun)
mkdir -p gpurun_out
TAG=$1; shift
i=0
for spec in "$@"; do
  IFS='|' read -r n bargs envs <<< "$spec"
  i=$((i+1))
  out=gpurun_out/multi_${TAG}_${i}_n${n}.json
  if [ "$n" = "1" ]; then
    env $envs timeout 900 python bench.py --gpus 1 --no-cpu-baseline --no-steps-api --no-other-configs $bargs > $out 2> ${out%.json}.err
  else
    env $envs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+i)) \
      bench.py --gpus $n --no-cpu-baseline --no-steps-api --no-other-configs $bargs > $out 2> ${out%.json}.err
  fi
  python - "$out" "$n" "$bargs" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    st = {k: v["ms_per_step"] for k, v in d.get("stages", {}).items()}
    print(f"N={sys.argv[2]} [{sys.argv[3]}] value {d['value']:9.2f} e2e {d['e2e']['value']:9.2f} ms/step {d['ms_per_step']:8.3f} exch {d.get('grad_exchange')} | " +
          " ".join(f"{k.replace('render_preprocess', 'pre').replace('_forward', '_f').replace('_backward', '_b')}={v:.3f}" for k, v in st.items()))
except Exception as e:
    print("N=", sys.argv[2], sys.argv[3], "FAILED", e)
    import subprocess
    print(subprocess.run(["tail", "-15", sys.argv[1].replace(".json", ".err")], capture_output=True, text=True).stdout)
PY
done
