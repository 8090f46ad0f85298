(timeout 900 python -m pytest tests/test_gpu_lift_splat.py -q --tb=short -p no:cacheprovider --timeout 900 -k "random_configurations" 2>&1 | tail -30)
