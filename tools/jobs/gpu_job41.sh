timeout 900 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench41.err > gpurun_out/bench41.json; tail -3 gpurun_out/bench41.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench41.json").read().strip().splitlines()[-1])
print("value",round(d["value"]),"train frac",round(d["roofline"]["frac"],4),"ms",d["roofline"]["launch_ms"], "fwd ms", d["ms_per_step"])
r=d["extra"]["shapes"]["dair_r50_b64_f32_bev_channels_last"]; print(json.dumps(r)[:900])
PY
