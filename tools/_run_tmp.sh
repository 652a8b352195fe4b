mkdir -p gpurun_out
timeout 1300 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -8
timeout 600 python tools/gs2d_acceptance.py 2>&1 | tail -22
timeout 600 python tools/gs2d_acceptance.py --points 20000 --iters 120 --out gpurun_out/gs2d_small.json 2>&1 | tail -14
timeout 900 python tools/channel_sweep.py > gpurun_out/channel_sweep.log 2>&1; tail -3 gpurun_out/channel_sweep.log
