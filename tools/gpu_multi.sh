#!/bin/bash
# multi-GPU bench (view-batch data parallel): N ranks via torchrun, as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-steps-api > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err
echo "rc=$?"; tail -3 gpurun_out/bench_dp$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_dp$N.json").read().strip().splitlines()[-1])
print("N=$N", round(d["value"],1), "renders/s  e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],2), d["clocks"])
PY
