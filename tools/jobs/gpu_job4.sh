mkdir -p gpurun_out
T="python -m pytest tests/test_gpu_lift_splat.py -q --tb=short -p no:cacheprovider --timeout 600"
(timeout 1200 $T -k "expands or shortcuts or nan_and_inf or dropped" 2>&1 | tail -25) > gpurun_out/t4.log
for p in block tile; do
  for s in dair_r50:64 sgv3d_bsm_r50:16 rope3d_r101_256:16; do
    timeout 300 python tools/time_kernels.py --shape ${s%%:*} --batch ${s##*:} --pipeline $p 2>&1 | head -1
  done
done > gpurun_out/t4_time.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"plan" -s 3 -c 3 -o gpurun_out/plan_tile -f python tools/time_kernels.py --shape dair_r50 --batch 64 --pipeline tile --iters 1 > gpurun_out/ncu4a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"plan" -s 1 -c 1 -o gpurun_out/plan_blk -f python tools/time_kernels.py --shape dair_r50 --batch 64 --pipeline block --iters 1 > gpurun_out/ncu4b.log 2>&1
tail -4 gpurun_out/t4.log; cat gpurun_out/t4_time.log
