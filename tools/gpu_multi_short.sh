mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus $1 --quick --steps 20 > gpurun_out/bench_n$1.json 2> gpurun_out/bench_n$1.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$1.json")); print("N", d["n_gpus"], "value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"train frac",round(d["roofline"]["frac"],4), "ms", d["ms_per_step"], "train ms", d["roofline"]["launch_ms"])
PY
tail -2 gpurun_out/bench_n$1.err
