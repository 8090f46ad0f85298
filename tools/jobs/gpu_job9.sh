mkdir -p gpurun_out
python tools/copy_bench.py --steps 20 > gpurun_out/copy_n1.jsonl 2> gpurun_out/copy_n1.err
python tools/copy_bench.py --steps 20 --depth 4 >> gpurun_out/copy_n1.jsonl 2>> gpurun_out/copy_n1.err
cat gpurun_out/copy_n1.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['copy_bench'], 'ms/step', round(d['ms_per_step_max_over_ranks'],3), 'frames/s', round(d['frames_per_s_aggregate']), 'GB/s', round(d['GBs_aggregate'],1))"
tail -3 gpurun_out/copy_n1.err
nvidia-smi topo -m 2>&1 | head -20
