mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider -k "render_sh or config3 or config5" 2>&1 | tail -3
bash tools/gpu_ab.sh r2p "bwd3cta||"
