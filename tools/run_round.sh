# one GPU call: tests, bench, ncu launch list of the bench command, ncu --set full of one step (forward + backward)
B=${1:-64}
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; tail -2 gpurun_out/bench_tmp.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ls_|camera_prep|inverse4x4" -s 23 -c 9 -o gpurun_out/full -f python tools/prof_step.py --batch $B --steps 3 --backward > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_tmp.json"))
print("value",round(d["value"]),"ms/step",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"]),"launches",d["gpu_launches"])
r=d["roofline"]; print("roofline frac",round(r["frac"],4),"achieved",round(r["achieved"]),"lib ms",round(r["launch_ms"],4),"dominant",r["dominant_kernel"],round(r["dominant_share"],2))
for k,v in sorted(r["kernels"].items(), key=lambda kv:-kv[1]["share"]): print("  %-36s %8.1f us  %.2f"%(k,v["avg_us"],v["share"]))
e=d.get("extra",{})
for k,v in e.items(): print(k, v)
print("clocks", d.get("clocks"), "traffic", r.get("traffic"), "cpu", d.get("cpu_baseline"))
PY
