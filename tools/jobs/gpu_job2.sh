mkdir -p gpurun_out
T="python -m pytest tests/test_gpu_lift_splat.py tests/test_gpu_integration.py -q --tb=short -p no:cacheprovider --timeout 300"
(timeout 600 $T -k "dropped or plan_cache or load_state or overridden" 2>&1 | tail -25) > gpurun_out/t2_misc.log
ncu --set full --clock-control none --import-source on -k regex:"bp_" -s 5 -c 5 -o gpurun_out/full_blk -f python tools/prof_step.py --batch 64 --steps 2 --backward > gpurun_out/ncu_blk.log 2>&1
tail -3 gpurun_out/t2_misc.log gpurun_out/ncu_blk.log
