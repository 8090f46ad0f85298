mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_lift_splat.py tests/test_gpu_integration.py -q --tb=short -p no:cacheprovider --timeout 600 -k "tile or not block" 2>&1 | tail -6) > gpurun_out/t11.log
for s in dair_r50:64 sgv3d_bsm_r50:16 rope3d_r50:32; do
  timeout 300 python tools/time_kernels.py --shape ${s%%:*} --batch ${s##*:} --pipeline tile 2>&1 | tail -1
done > gpurun_out/t11_time.log 2>&1
timeout 300 python tools/time_kernels.py --shape dair_r50 --batch 8 --pipeline tile --bf16 2>&1 | tail -1 >> gpurun_out/t11_time.log
tail -3 gpurun_out/t11.log; cat gpurun_out/t11_time.log
