import sys; sys.path.insert(0,'.')
import torch
from srf_b200 import RoutingStack
from oracle import srf_oracle as o
def rel(a,r): return ((a.double().cpu()-r.double().cpu()).abs().max()/r.double().abs().max()).item()
L,PH,CH,class_n,DIM,lpad,rpad=10,60,30,32,20,2,2
for S in (48,375):
  emb=torch.randn(2,S,PH,DIM,generator=torch.Generator().manual_seed(5)).cuda()
  st={m:RoutingStack(L,PH,CH,class_n,DIM,DIM,DIM,lpad,rpad,1,True,seed=9,uhat_mode=m) for m in ('fp32','tf32','bf16')}
  outs={m:st[m].forward(emb,return_capsules=True) for m in st}
  torch.cuda.synchronize()
  for m in ('tf32','bf16'):
    print('S',S,m,'logits err vs fp32 %.2e'%rel(outs[m][0],outs['fp32'][0]),'per-layer caps err:',['%.1e'%rel(a,b) for a,b in zip(outs[m][1][:-1],outs['fp32'][1][:-1])])
    d=(outs[m][0]-outs['fp32'][0]).abs()
    print('   mean abs logit diff %.2e, 99.9pct %.2e, max %.2e; argmax agreement %.5f'%(d.mean().item(), d.flatten().kthvalue(int(0.999*d.numel())).values.item(), d.max().item(), (outs[m][0].argmax(-1)==outs['fp32'][0].argmax(-1)).float().mean().item()))
