import os, sys; sys.path.insert(0,'.')
import torch
from srf_b200 import routing
def mk(B,S,H,d,O,D,win,seed):
    g=torch.Generator().manual_seed(seed)
    return torch.randn(B,S,H,d,generator=g), torch.randn(win*H,O,D,d,generator=g)*0.1, torch.randn(win*H,O,D,generator=g)*0.1
def rel(a,ref): return ((a.double()-ref.double()).abs().max()/ref.double().abs().max()).item()
cases=[(4,8,60,8,30,8,1,1),(3,5,60,8,30,8,1,1),(2,6,30,8,63,8,1,1),(8,9,30,8,30,8,3,3),(16,12,60,20,30,20,2,2),(5,7,30,20,32,20,2,2),(2,5,7,16,9,16,0,0),(1,6,5,32,6,32,4,4),(64,40,60,20,30,20,2,2)]
tot=0;bad=0
for case in cases:
  B,S,H,d,O,D,lpad,rpad=case
  emb,W,bias=mk(B,S,H,d,O,D,lpad+rpad+1,3)
  for sdr in (True,False):
    for iters in (1,2,3):
      for st in ('0','2'):
        os.environ['SRF_STREAM_STAGES']=st
        os.environ['SRF_NO_STREAM']='0'
        h=routing.Handle()
        a=routing.LayerArgs(W=W.cuda(),bias=bias.cuda(),lpad=lpad,rpad=rpad,iters=iters,sdr=sdr,mask_class0=(iters==3),uhat_mode=os.environ.get('SOAK_MODE','bf16'))
        outs=[routing.route_layer_fwd(emb.cuda(),a,handle=h)[0].clone() for _ in range(10)]
        torch.cuda.synchronize()
        os.environ['SRF_NO_STREAM']='1'
        h2=routing.Handle()
        ref=routing.route_layer_fwd(emb.cuda(),a,handle=h2)[0]
        torch.cuda.synchronize()
        nb=sum((x!=outs[0]).any().item() for x in outs[1:]); e=max(rel(x,ref) for x in outs)
        tot+=1
        if nb or e>1e-5:
            bad+=1; print('BAD',case,sdr,iters,'stages',st,'nondet',nb,'vs nostream %.1e'%e, h.last_kernel[-60:])
        h.close(); h2.close()
print('soak configs',tot,'bad',bad)
