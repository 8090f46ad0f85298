(timeout 1200 python -m pytest tests/test_gpu_integration.py tests/test_gpu_lift_splat.py -q --tb=short -p no:cacheprovider --timeout 900 2>&1 | tail -8)
