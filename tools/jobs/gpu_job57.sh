for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 86 --print-limit 5 \
    python -m pytest tests/test_gpu_integration.py tests/test_gpu_lift_splat.py -m gpu -q -x -k "channels_last or (random_configurations and (3 or 7 or 11 or 19)) or (bit_identical and small)" -p no:cacheprovider 2>&1 | grep -v "Host Frame" | tail -12
  echo "exit code: $?"
done
