(timeout 1200 python -m pytest tests/test_gpu_lift_splat.py tests/test_gpu_integration.py -q -x --tb=short -p no:cacheprovider --timeout 900 2>&1 | tail -5)
for cfg in "dair_r50 64" "dair_r50 8" "rope3d_r50 32" "rope3d_native 32" "sgv3d_bsm_r50 16" "dair_r50_256 16"; do set -- $cfg
timeout 120 python tools/time_kernels.py --shape $1 --batch $2 --pipeline tile --iters 30 2>&1 | sed -n 3p
done
