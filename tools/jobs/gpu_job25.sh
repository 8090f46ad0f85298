(timeout 900 python -m pytest tests/test_gpu_integration.py -q -x --tb=short -p no:cacheprovider --timeout 900 2>&1 | tail -12)
for ov in 0 1; do echo overlap $ov; SGV3D_OVERLAP_CONTEXT=$ov timeout 600 python bench.py --quick 2> gpurun_out/bench25_$ov.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'train frac',round(d['roofline']['frac'],4),'fwd frac',round(d['roofline']['forward_only']['frac'],4), 'launches', d['gpu_launches'], 'e2e', round(d['e2e']['value']))
print({k:v for k,v in d.get('extra',{}).items() if 'batch1' in k or 'train' in k})
"; done
