mkdir -p gpurun_out
bash tools/gpu_multi.sh r2f8 "8|--steps 8" "4|--steps 8" "8|--config 5 --steps 4"
