mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 2>&1 | tail -30) > gpurun_out/t12_all.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench12.json 2> gpurun_out/bench12.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench12_ref.json 2> gpurun_out/bench12_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches12.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/bench12_under_ncu.log 2>&1
python tools/time_kernels.py --shape sgv3d_bsm_r50 --batch 16 --pipeline tile > gpurun_out/t12_time.log 2>&1
tail -6 gpurun_out/t12_all.log; tail -3 gpurun_out/bench12.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench12.json")); r=d["roofline"]
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"train frac",round(r["frac"],4),"fwd frac",round(r["forward_only"]["frac"],4),"cpu",d["cpu_baseline"]["kind"],round(d["cpu_baseline"]["value"],1))
for k,v in d["extra"]["shapes"].items(): print(k, {kk:(round(vv["ms"],3),round(vv["frac_of_measured_peak"],4)) for kk,vv in v.items() if isinstance(vv,dict)})
PY
cat gpurun_out/t12_time.log
