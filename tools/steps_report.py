import json
d=json.load(open("gpurun_out/bench_stepsapi.json"))
print(round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],2), "serial", round(d["serial_ms_per_step"],2))
tot=0
for k,v in d["stages"].items():
    print("  %-32s %.4f ms x%d  hbm_frac %s" % (k, v["ms"], v["calls"], v.get("hbm_frac"))); tot+=v["ms"]
print("sum of our calls per render", round(tot,3))
