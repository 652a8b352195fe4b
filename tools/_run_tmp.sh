mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider -k "grad_sync or ndc_hook" 2>&1 | tail -2
bash tools/gpu_multi.sh r2s "2|--steps 8 --grad-chunks 3" "2|--steps 8 --grad-chunks 4" "2|--steps 8 --grad-chunks 6" "2|--steps 8 --grad-chunks 2"
