(timeout 1200 python -m pytest tests/test_gpu_integration.py tests/test_gpu_lift_splat.py -q -x --tb=short -p no:cacheprovider --timeout 900 2>&1 | tail -15)
timeout 900 python bench.py --quick 2> gpurun_out/bench40.err > gpurun_out/bench40.json; tail -3 gpurun_out/bench40.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench40.json").read().strip().splitlines()[-1])
print("value",round(d["value"]),"train frac",round(d["roofline"]["frac"],4),"ms",d["roofline"]["launch_ms"])
r=d["extra"]["shapes"]["dair_r50_b64_f32_bev_channels_last"]; print(json.dumps(r)[:700])
PY
