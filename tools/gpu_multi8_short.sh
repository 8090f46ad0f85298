mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus 8 --quick --steps 20 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n8.json")); print("N 8 value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"train frac",round(d["roofline"]["frac"],4), "ms", d["ms_per_step"])
PY
tail -3 gpurun_out/bench_n8.err
