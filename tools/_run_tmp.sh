mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 6 --no-cpu-baseline --no-steps-api --no-other-configs --trace gpurun_out/trace_r2q.json > gpurun_out/bench_r2q.json 2> gpurun_out/bench_r2q.err
tail -2 gpurun_out/bench_r2q.err | cut -c1-200
python tools/trace_summary.py gpurun_out/trace_r2q.json all > gpurun_out/trace_r2q.txt; head -12 gpurun_out/trace_r2q.txt
grep -n "blend_bwd" gpurun_out/trace_r2q.txt | head -3
gzip -f gpurun_out/trace_r2q.json
bash tools/gpu_multi.sh r2q "2|--steps 6" "1|--steps 6" "2|--config 5 --steps 4" "1|--config 5 --steps 4"
