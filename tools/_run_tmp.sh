mkdir -p gpurun_out
timeout 1300 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider 2>&1 | tail -4
bash tools/gpu_ab.sh r2o "pt2||"
