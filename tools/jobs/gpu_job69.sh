(SGV3D_NO_TMA_STORE=1 SGV3D_NO_TMA_LOAD=1 timeout 1200 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider --timeout 900 2>&1 | tail -4)
