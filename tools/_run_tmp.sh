mkdir -p gpurun_out
MSB_BWD_SPLIT=0 MSB_DP_COALESCE=0 bash tools/gpu_multi.sh r2h_nc "2|--steps 6" "2|--steps 6 --grad-chunks 1"
MSB_BWD_SPLIT=0 bash tools/gpu_multi.sh r2h_w "2|--steps 6 --warmup 8" 
MSB_DP_COALESCE=0 bash tools/gpu_multi.sh r2h_snc "2|--steps 6" "2|--steps 12 --warmup 6"
