import os, sys; sys.path.insert(0,'.')
import torch
from srf_b200 import routing
from oracle import srf_oracle as o
def mk(B,S,H,d,O,D,win,seed):
    g=torch.Generator().manual_seed(seed)
    return torch.randn(B,S,H,d,generator=g), torch.randn(win*H,O,D,d,generator=g)*0.1, torch.randn(win*H,O,D,generator=g)*0.1
for (B,S,H,d,O,D,lpad,rpad) in ((2,6,30,8,63,8,1,1),(2,6,30,8,30,8,1,1),(4,6,30,8,63,8,1,1)):
  emb,W,bias=mk(B,S,H,d,O,D,lpad+rpad+1,17)
  ref=o.prediction_vectors(o.window_gather(emb.double(),lpad,rpad),W.double(),bias.double())
  for mode in ('tf32','bf16'):
    h=routing.Handle()
    outs=[routing.uhat_fwd(emb.cuda(),W.cuda(),bias.cuda(),lpad,rpad,mode,handle=h).clone() for _ in range(12)]
    torch.cuda.synchronize()
    nbad=sum((x!=outs[0]).any().item() for x in outs[1:])
    errs=[((x.double().cpu()-ref).abs().max()/ref.abs().max()).item() for x in outs]
    print((B,S,H,d,O,D),mode,'calls differing from call0: %d/11'%nbad,'max err %.2e min err %.2e'%(max(errs),min(errs)))
    if nbad:
        for k,x in enumerate(outs[1:],1):
            dd=(x!=outs[0])
            if dd.any():
                idx=dd.nonzero()
                print('  call',k,'n diff',len(idx),'first',idx[:6].cpu().numpy().tolist(),'unique b',idx[:,0].unique().tolist(),'unique s',idx[:,1].unique().tolist(),'n unique i',len(idx[:,2].unique()))
