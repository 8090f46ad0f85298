mkdir -p gpurun_out
T="python -m pytest tests/test_gpu_lift_splat.py tests/test_gpu_integration.py -q --tb=short -p no:cacheprovider --timeout 600"
(timeout 1500 $T -k "auto or not tile" 2>&1 | tail -30) > gpurun_out/t5_auto.log
(timeout 1500 $T -k "tile and (expands or shortcuts or forward_backward or determin)" 2>&1 | tail -30) > gpurun_out/t5_tile.log
for p in block tile; do
  for s in dair_r50:64 sgv3d_bsm_r50:16 rope3d_r101_256:16; do
    timeout 300 python tools/time_kernels.py --shape ${s%%:*} --batch ${s##*:} --pipeline $p 2>&1 | tail -3
  done
done > gpurun_out/t5_time.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bp_" -s 5 -c 5 -o gpurun_out/full_blk2 -f python tools/prof_step.py --batch 64 --steps 2 --backward > gpurun_out/ncu5.log 2>&1
tail -4 gpurun_out/t5_auto.log gpurun_out/t5_tile.log; cat gpurun_out/t5_time.log
