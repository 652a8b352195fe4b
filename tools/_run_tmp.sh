mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-steps-api --no-other-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"render_pre" -s 6 -c 2 -o gpurun_out/prof_r2n -f $CMD > gpurun_out/prof_r2n.out 2>&1
echo "rc=$?"; tail -2 gpurun_out/prof_r2n.out | cut -c1-200
ncu -i gpurun_out/prof_r2n.ncu-rep --page raw --csv > gpurun_out/prof_r2n_raw.csv 2>/dev/null
python tools/ncu_summary.py full gpurun_out/prof_r2n_raw.csv
python tools/sass_hotspots.py gpurun_out/prof_r2n.ncu-rep render_pre_bwd_pt 40
python tools/sass_hotspots.py gpurun_out/prof_r2n.ncu-rep render_pre_fwd_pt 25
