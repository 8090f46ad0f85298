timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 6 python -m pytest tests/test_gpu_lift_splat.py -m gpu -q -x -k "block and tiny and expands" -p no:cacheprovider > gpurun_out/race28.log 2>&1
grep -c "Race reported" gpurun_out/race28.log; head -60 gpurun_out/race28.log | cut -c1-330
