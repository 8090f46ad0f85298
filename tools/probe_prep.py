import torch, sys
sys.path.insert(0,'/root/repo')
from sgv3d_b200 import get_shape
from sgv3d_b200.synthetic import make_mats
from sgv3d_b200.view_transform import _inverse
torch.manual_seed(0)
bad=0
for B in (1,2,8,32,64):
  for seed in range(3):
    for bda in ("identity","random"):
        m=make_mats(get_shape("dair_r50"),B,1,seed=seed,bda=bda)
        ida=m["ida"].unsqueeze(1).cuda(); K=m["intrin"].unsqueeze(1).cuda(); s2v=m["sensor2virtual"].unsqueeze(1).cuda(); s2e=m["sensor2ego"].unsqueeze(1).cuda()
        a0,a1,a2=_inverse(ida),_inverse(K),_inverse(s2v)
        st=_inverse(torch.cat((ida,K,s2v),0))
        b0,b1,b2=st[:B],st[B:2*B],st[2*B:]
        e=[int((x.view(torch.int32)!=y.view(torch.int32)).sum()) for x,y in ((a0,b0),(a1,b1),(a2,b2))]
        mv=s2v.matmul(a1); me=s2e.matmul(a2)
        mm=torch.cat((s2v,s2e),0).matmul(torch.cat((a1,a2),0))
        e2=[int((mv.view(torch.int32)!=mm[:B].view(torch.int32)).sum()), int((me.view(torch.int32)!=mm[B:].view(torch.int32)).sum())]
        if any(e) or any(e2): bad+=1; print("B",B,seed,bda,"inv diff",e,"mm diff",e2)
print("bad",bad)
# count kernels
from torch.profiler import profile, ProfilerActivity
m=make_mats(get_shape("dair_r50"),32,1,seed=1,bda="identity")
ida=m["ida"].unsqueeze(1).cuda(); K=m["intrin"].unsqueeze(1).cuda(); s2v=m["sensor2virtual"].unsqueeze(1).cuda(); s2e=m["sensor2ego"].unsqueeze(1).cuda()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    a0,a1,a2=_inverse(ida),_inverse(K),_inverse(s2v); mv=s2v.matmul(a1); me=s2e.matmul(a2); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=20))
