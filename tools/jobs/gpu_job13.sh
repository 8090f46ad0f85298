mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 -x 2>&1 | tail -12) > gpurun_out/t13_all.log
for s in dair_r50:64 sgv3d_bsm_r50:16; do
  timeout 300 python tools/time_kernels.py --shape ${s%%:*} --batch ${s##*:} --pipeline tile 2>&1 | tail -3
done > gpurun_out/t13_time.log 2>&1
timeout 300 python tools/time_kernels.py --shape dair_r50 --batch 64 --pipeline block 2>&1 | tail -1 >> gpurun_out/t13_time.log
(timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_lift_splat.py -q --tb=line -p no:cacheprovider -k "(forward_backward or expands) and (tiny or small)" 2>&1 | tail -6) > gpurun_out/t13_sanitize.log
tail -5 gpurun_out/t13_all.log; cat gpurun_out/t13_time.log; tail -4 gpurun_out/t13_sanitize.log
