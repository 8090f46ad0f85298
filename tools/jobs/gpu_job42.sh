for b in 64 8; do
timeout 120 python tools/time_kernels.py --shape dair_r50 --batch $b --pipeline tile --iters 30 2>&1 | tail -2
timeout 120 python tools/time_kernels.py --shape dair_r50 --batch $b --pipeline tile --iters 30 --channels-last 2>&1 | tail -2
done
