(timeout 900 python -m pytest tests/test_gpu_lift_splat.py tests/test_gpu_geometry.py -q -x --tb=short -p no:cacheprovider --timeout 900 2>&1 | tail -3)
for occ in 9 8 6; do echo occ $occ
for cfg in "dair_r50 64" "dair_r50 1" "sgv3d_bsm_r50 16"; do set -- $cfg
  SGV3D_PLAN_OCC=$occ timeout 120 python tools/time_kernels.py --shape $1 --batch $2 --pipeline tile --iters 30 2>&1 | head -1
done; done
timeout 120 python tools/time_kernels.py --shape dair_r50 --batch 64 --pipeline block --iters 30 2>&1 | head -1
timeout 120 python tools/time_kernels.py --shape dair_r50 --batch 1 --pipeline block --iters 30 2>&1 | head -1
