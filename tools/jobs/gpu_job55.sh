(timeout 1200 python -m pytest tests/test_gpu_lift_splat.py tests/test_gpu_integration.py -q -x --tb=short -p no:cacheprovider --timeout 900 -k "bsm or BSM or channel" 2>&1 | tail -3)
for i in 1 2; do timeout 600 python bench.py --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',round(d['value']),'train frac',round(d['roofline']['frac'],4))"; done
timeout 120 python tools/time_kernels.py --shape dair_r50 --batch 64 --pipeline tile --iters 30 2>&1 | sed -n 2p
