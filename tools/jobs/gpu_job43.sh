(timeout 1200 python -m pytest tests/test_gpu_integration.py -q -x --tb=short -p no:cacheprovider --timeout 900 2>&1 | tail -8)
for b in 64 8; do
timeout 120 python tools/time_kernels.py --shape dair_r50 --batch $b --pipeline tile --iters 30 --channels-last 2>&1 | tail -2
done
