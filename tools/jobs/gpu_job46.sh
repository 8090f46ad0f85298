(timeout 900 python -m pytest tests/test_gpu_integration.py -q -x --tb=short -p no:cacheprovider --timeout 900 -k "channels_last" 2>&1 | tail -12)
