mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_lift_splat.py -q --tb=short -p no:cacheprovider --timeout 600 -k "expands or shortcuts or nan_and_inf or dropped" 2>&1 | tail -5) > gpurun_out/t8.log
for s in dair_r50:64 sgv3d_bsm_r50:16 rope3d_r101_256:16; do
  timeout 300 python tools/time_kernels.py --shape ${s%%:*} --batch ${s##*:} --pipeline tile 2>&1 | head -1
  timeout 300 python tools/time_kernels.py --shape ${s%%:*} --batch ${s##*:} --pipeline block 2>&1 | head -1
done > gpurun_out/t8_time.log 2>&1
python -c "
import sys; sys.path.insert(0,'.')
from oracle import ref_cpu as R; print('ref bytecode available on the box:', R.available())" >> gpurun_out/t8_time.log 2>&1
tail -3 gpurun_out/t8.log; cat gpurun_out/t8_time.log
