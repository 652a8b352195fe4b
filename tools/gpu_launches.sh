#!/bin/bash
# ncu launch list (per-launch device time) of one fused bench step
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --views 1 --no-cpu-baseline --no-steps-api"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_cur.csv $CMD > gpurun_out/launches_cur.out 2>&1
echo "launch list rc=$?"
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_cur.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); 
L=[(r[ki], float(r[vi].replace(",",""))) for r in rows[1:] if r[vi].replace(",","").replace(".","").isdigit()]
# find the last render_pre_fwd launch -> the timed step
idx=[i for i,(k,v) in enumerate(L) if "render_pre_fwd" in k]
s=idx[3] if len(idx)>3 else idx[-1]
e=[i for i,(k,v) in enumerate(L) if "render_pre_bwd" in k and i>s][0]
tot=sum(v for k,v in L[s:e+1])
for k,v in L[s:e+1]: print(f"{v/1000:9.1f} us  {k[:90]}")
print("total", tot/1000, "us")
PY
