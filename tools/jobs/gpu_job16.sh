python tools/time_op.py --batch 4 2>&1 | head -1; python tools/time_op.py --batch 16 2>&1 | head -1
(timeout 900 python -m pytest tests/test_gpu_voxel_pooling.py -q --tb=short -p no:cacheprovider 2>&1 | tail -3)
