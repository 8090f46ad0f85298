mkdir -p gpurun_out
T="python -m pytest tests/test_gpu_lift_splat.py tests/test_gpu_integration.py -q --tb=short -p no:cacheprovider --timeout 600"
(timeout 1200 $T -k "expands or shortcuts or overridden or forward_backward or nan_and_inf" 2>&1 | tail -25) > gpurun_out/t3.log
for p in block tile; do
  timeout 300 python tools/time_kernels.py --shape dair_r50 --batch 64 --pipeline $p 2>&1 | tail -4
  timeout 300 python tools/time_kernels.py --shape sgv3d_bsm_r50 --batch 16 --pipeline $p 2>&1 | head -1
done > gpurun_out/t3_time.log 2>&1
tail -4 gpurun_out/t3.log; cat gpurun_out/t3_time.log
