#!/bin/bash
# multi-GPU A/B of the gradient slab count:  gpu_multi2.sh N chunks...
N=$1; shift
mkdir -p gpurun_out
for c in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-steps-api --grad-chunks $c > gpurun_out/bench_dp${N}_c$c.json 2> gpurun_out/bench_dp${N}_c$c.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_dp${N}_c$c.json").read().strip().splitlines()[-1])
print("N=$N chunks=$c", round(d["value"],1), "renders/s  e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],2))
PY
done
