mkdir -p gpurun_out
timeout 1200 python tools/sweep.py --batch 32 --iters 10 > gpurun_out/sweep_b32.jsonl 2> gpurun_out/sweep_b32.err; wc -l gpurun_out/sweep_b32.jsonl; tail -2 gpurun_out/sweep_b32.err
timeout 1200 python tools/sweep.py --batch 8 --iters 10 > gpurun_out/sweep_b8.jsonl 2> gpurun_out/sweep_b8.err; wc -l gpurun_out/sweep_b8.jsonl; tail -2 gpurun_out/sweep_b8.err
