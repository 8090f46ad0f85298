for n in 1 3 4 8; do echo ctas $n; SGV3D_CTX_CTAS=$n timeout 600 python bench.py --quick 2> gpurun_out/bench26_$n.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'train frac',round(d['roofline']['frac'],4),'fwd frac',round(d['roofline']['forward_only']['frac'],4))
"; done
