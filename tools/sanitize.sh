# compute-sanitizer memcheck + racecheck over the small-shape GPU tests (run on the GPU box)
set -o pipefail
SEL='tiny or small or channel_sweep or static_calibration or bsm'
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 86 --print-limit 5 \
    python -m pytest tests/test_gpu_lift_splat.py tests/test_gpu_voxel_pooling.py tests/test_gpu_geometry.py -m gpu -q -x -k "$SEL" -p no:cacheprovider 2>&1 | tail -8
  echo "exit code: $?"
done
