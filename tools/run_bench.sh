timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -25; python bench.py --steps 20 --warmup 3 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_tmp.json"))
print("value",round(d["value"]),"ms/step",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"]),"launches",d["gpu_launches"])
r=d["roofline"]; print("roofline frac",round(r["frac"],4),"achieved",round(r["achieved"]),"lib ms",round(r["launch_ms"],4),"dominant",r["dominant_kernel"],round(r["dominant_share"],2))
for k,v in sorted(r["kernels"].items(), key=lambda kv:-kv[1]["share"]): print("  %-36s %8.1f us  %.2f"%(k,v["avg_us"],v["share"]))
e=d.get("extra",{}); print(e.get("phases_ms")); print("cached fwd",e.get("cached_plan_forward")); print("train",e.get("train_step_fwd_bwd")); print("b1 us",e.get("batch1_latency_us"), "b1 graph us", e.get("batch1_graph_latency_us"), "clocks", d.get("clocks"), "traffic", r.get("traffic"), "eager fps", e.get("eager_launch_frames_per_s_per_gpu")); print(e.get("op_level_forward")); print(e.get("reference_kernel_sm100a_forward"))
PY
tail -3 gpurun_out/bench_tmp.err
