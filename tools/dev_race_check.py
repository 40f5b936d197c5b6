import os, sys; sys.path.insert(0,'.')
import torch
from srf_b200 import routing
def mk(B,S,H,d,O,D,win,seed):
    g=torch.Generator().manual_seed(seed)
    return torch.randn(B,S,H,d,generator=g), torch.randn(win*H,O,D,d,generator=g)*0.1, torch.randn(win*H,O,D,generator=g)*0.1
def rel(a,ref): return ((a.double()-ref.double()).abs().max()/ref.double().abs().max()).item()
B,S,H,d,O,D,lpad,rpad=(4,8,60,8,30,8,1,1)
emb,W,bias=mk(B,S,H,d,O,D,3,17)
os.environ['SRF_FORCE_C']='1'
for st in ('2','3','4','8','12','16'):
  os.environ['SRF_STREAM_STAGES']=st
  for mode in ('tf32','bf16'):
    for iters in (1,2,3):
        h=routing.Handle()
        a=routing.LayerArgs(W=W.cuda(),bias=bias.cuda(),lpad=lpad,rpad=rpad,iters=iters,sdr=True,mask_class0=False,uhat_mode=mode)
        outs=[routing.route_layer_fwd(emb.cuda(),a,handle=h)[0].clone() for _ in range(12)]
        torch.cuda.synchronize()
        nbad=sum(rel(x,outs[0])>0 for x in outs[1:])
        print('stages',st,mode,'iters',iters,'bad calls %d/11'%nbad, h.last_kernel[-44:])
        h.close()
