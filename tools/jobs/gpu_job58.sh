for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 86 --print-limit 5 \
    python -m pytest tests/test_gpu_lift_splat.py -m gpu -q -x -k "test_forward_backward_vs_fp64_oracle and tile and (rope3d_r50 or sgv3d_bsm_r50)" -p no:cacheprovider 2>&1 | grep -v "Host Frame" | tail -8
  echo "exit code: $?"
done
