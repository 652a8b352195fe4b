mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider -k "render_sh or config5 or config3" 2>&1 | tail -4
bash tools/gpu_ab.sh r2g "split||" "nosplit|MSB_BWD_SPLIT=0|"
bash tools/gpu_multi.sh r2g "2|--steps 6" "2|--steps 6 --grad-chunks 1" "2|--steps 6 --grad-chunks 2"
MSB_BWD_SPLIT=0 bash tools/gpu_multi.sh r2g0 "2|--steps 6"
