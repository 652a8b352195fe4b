#!/bin/bash
# compute-sanitizer over the GPU tests that exercise the kernels written or changed in round 2 (batched preprocess / PT
# kernels with TMA + mbarrier + cp.async staging, batched sort incl. the gather pass, blockIdx.z blends, fused Adam).
mkdir -p gpurun_out
SEL="render_sh_equals_steps or views_equals_single or grad_sync or sort_views or sort_known or sort_edge or sort_compaction or fused_adam or rasterization_fused"
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -m gpu -q -x --timeout=1400 -p no:cacheprovider -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_$tool.log | tail -4
done
