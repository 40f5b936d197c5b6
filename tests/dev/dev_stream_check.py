import os, sys; sys.path.insert(0,'.')
import torch
from oracle import srf_oracle as o
from srf_b200 import routing
def mk(B,S,H,d,O,D,win,seed):
    g=torch.Generator().manual_seed(seed)
    return torch.randn(B,S,H,d,generator=g), torch.randn(win*H,O,D,d,generator=g)*0.1, torch.randn(win*H,O,D,generator=g)*0.1
def rel(a,ref): return ((a.double().cpu()-ref.double().cpu()).abs().max()/ref.double().abs().max()).item()
case=(3,5,60,8,30,8,1,1)
B,S,H,d,O,D,lpad,rpad=case
emb,W,bias=mk(B,S,H,d,O,D,3,17)
for iters,last in ((1,False),(3,True),(3,False),(2,False)):
  for sdr in (True,False):
    ref=o.route_layer(emb.double(),W.double(),bias.double(),lpad,rpad,iters,sdr,last)
    outs={}
    for mode in ('fp32','tf32','bf16'):
        for ns in (0,1):
            if mode=='fp32' and ns: continue
            os.environ['SRF_NO_STREAM']=str(ns)
            h=routing.Handle()
            a=routing.LayerArgs(W=W.cuda(),bias=bias.cuda(),lpad=lpad,rpad=rpad,iters=iters,sdr=sdr,mask_class0=last,uhat_mode=mode)
            c1,_=routing.route_layer_fwd(emb.cuda(),a,handle=h)
            c2,_=routing.route_layer_fwd(emb.cuda(),a,handle=h)
            torch.cuda.synchronize()
            outs[(mode,ns)]=c1
            print(iters,last,'sdr' if sdr else 'dr',mode,'nostream' if ns else 'stream','err vs oracle %.2e'%rel(c1,ref),'repeat diff %.1e'%rel(c2,c1))
            h.close()
    print('   stream vs nostream tf32 %.2e bf16 %.2e'%(rel(outs[('tf32',0)],outs[('tf32',1)]), rel(outs[('bf16',0)],outs[('bf16',1)])))
