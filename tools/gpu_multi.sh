# multi-GPU measurements of one size N (argument): default bench, Rope3D-shaped sweep point, bf16 training config, copy bench
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
$TR bench.py --gpus $N --quick --steps 20 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err
for gb in 8 16 32 64; do
  pb=$((gb / N)); [ $pb -ge 1 ] || continue
  $TR bench.py --gpus $N --quick --steps 20 --shape rope3d_r50 --batch $pb 2>> gpurun_out/bench_n${N}.err | sed "s/^/{\"global_batch\": $gb, \"line\": /; s/$/}/" >> gpurun_out/rope3d_n${N}.jsonl
done
$TR bench.py --gpus $N --quick --steps 20 --batch 8 --ctx bf16 > gpurun_out/train_bf16_n${N}.json 2>> gpurun_out/bench_n${N}.err
$TR bench.py --gpus $N --quick --steps 20 --batch 8 --ctx bf16 --shape sgv3d_bsm_r50 > gpurun_out/train_bf16_bsm_n${N}.json 2>> gpurun_out/bench_n${N}.err
$TR tools/copy_bench.py --steps 20 > gpurun_out/copy_n${N}.jsonl 2>> gpurun_out/bench_n${N}.err
python - <<PY
import json
N=$N
d=json.load(open("gpurun_out/bench_n%d.json"%N)); print("N",N,"value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"train frac",round(d["roofline"]["frac"],4))
for l in open("gpurun_out/rope3d_n%d.jsonl"%N):
    j=json.loads(l); print(" rope3d global", j["global_batch"], round(j["line"]["value"]), "e2e", round(j["line"]["e2e"]["value"]))
for f in ("train_bf16_n%d.json"%N, "train_bf16_bsm_n%d.json"%N):
    t=json.load(open("gpurun_out/"+f)); print(f, "train ms", round(t["roofline"]["launch_ms"],4), "frac", round(t["roofline"]["frac"],4))
for l in open("gpurun_out/copy_n%d.jsonl"%N):
    j=json.loads(l); print(" copy", j["copy_bench"], round(j["ms_per_step_max_over_ranks"],2), "ms", round(j["GBs_aggregate"],1), "GB/s agg")
PY
tail -3 gpurun_out/bench_n${N}.err
