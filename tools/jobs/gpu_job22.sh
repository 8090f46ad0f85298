(timeout 1500 python -m pytest tests/test_gpu_lift_splat.py tests/test_gpu_geometry.py -q -x --tb=short -p no:cacheprovider --timeout 900 2>&1 | tail -8)
for cfg in "dair_r50 64" "dair_r50 1" "sgv3d_bsm_r50 16" "rope3d_r101_256 16"; do set -- $cfg
  for p in tile block; do timeout 120 python tools/time_kernels.py --shape $1 --batch $2 --pipeline $p --iters 30 2>&1 | head -1; done
done
