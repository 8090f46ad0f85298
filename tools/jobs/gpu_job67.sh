mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 2>&1 | tail -12) > gpurun_out/t67_all.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench67.json 2> gpurun_out/bench67.err
tail -4 gpurun_out/t67_all.log; tail -3 gpurun_out/bench67.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench67.json")); r=d["roofline"]; e=d["extra"]
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"train frac",round(r["frac"],4),"fwd frac",round(r["forward_only"]["frac"],4), d["config"]["pipeline"])
print("batch1", e["batch1_latency_us"], e["batch1_graph_latency_us"], e["batch1_static_camera_graph_latency_us"], e["frames_per_s_by_batch"])
print("op", e["op_level_forward"])
PY
python - <<PY
import json
d=json.load(open("gpurun_out/bench67.json")); r=d["roofline"]; e=d["extra"]
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"train frac",round(r["frac"],4),"fwd frac",round(r["forward_only"]["frac"],4), d["config"]["pipeline"], "launches", d["gpu_launches"])
print("batch1", e["batch1_latency_us"], e["batch1_graph_latency_us"], e["batch1_static_camera_graph_latency_us"])
for k,v in e["shapes"].items(): print(k, {kk:(round(vv["ms"],3),round(vv["frac_of_measured_peak"],4)) for kk,vv in v.items() if isinstance(vv,dict)})
PY
(timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2)
