# first GPU job of round 2: block pipeline correctness + phase timings
mkdir -p gpurun_out
T="python -m pytest tests/test_gpu_lift_splat.py -q --tb=short -p no:cacheprovider --timeout 300"
(timeout 600 $T -k "auto and expands" 2>&1 | tail -25) > gpurun_out/t_auto_expand.log
(timeout 900 $T -k "auto and (forward_backward or channel_sweep or bf16 or determin)" 2>&1 | tail -40) > gpurun_out/t_auto_fb.log
(timeout 900 $T -k "auto and not expands and not forward_backward and not channel_sweep and not bf16 and not determin" 2>&1 | tail -40) > gpurun_out/t_auto_rest.log
(timeout 900 $T -k "tile" 2>&1 | tail -25) > gpurun_out/t_tile.log
(timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_lift_splat.py -q --tb=line -p no:cacheprovider -k "auto and (forward_backward or expands) and (tiny or small)" 2>&1 | tail -40) > gpurun_out/t_sanitize.log
for p in block tile; do
  timeout 300 python tools/time_kernels.py --shape dair_r50 --batch 64 --pipeline $p 2>&1 | tail -4
  timeout 300 python tools/time_kernels.py --shape sgv3d_bsm_r50 --batch 16 --pipeline $p 2>&1 | tail -4
  timeout 300 python tools/time_kernels.py --shape rope3d_r101_256 --batch 16 --pipeline $p 2>&1 | tail -4
done > gpurun_out/t_time.log 2>&1
tail -3 gpurun_out/t_auto_expand.log gpurun_out/t_auto_fb.log gpurun_out/t_auto_rest.log gpurun_out/t_tile.log gpurun_out/t_sanitize.log
cat gpurun_out/t_time.log
