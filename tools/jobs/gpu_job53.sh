(timeout 1200 python -m pytest tests/test_gpu_lift_splat.py tests/test_gpu_integration.py -q -x --tb=short -p no:cacheprovider --timeout 900 -k "bsm or BSM" 2>&1 | tail -4)
for bg in -1 0.9; do timeout 200 python tools/time_bsm.py --batch 16 --background $bg --pipeline tile 2>&1 | tail -1; done
