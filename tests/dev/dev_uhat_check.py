import sys; sys.path.insert(0,'.')
import torch
from oracle import srf_oracle as o
from srf_b200 import routing
torch.manual_seed(0)
def run(B,S,H,d,O,D,lpad,rpad,mode):
    g=torch.Generator().manual_seed(1)
    emb=torch.randn(B,S,H,d,generator=g); win=lpad+rpad+1
    W=torch.randn(win*H,O,D,d,generator=g)*0.1; bias=torch.randn(win*H,O,D,generator=g)*0.1
    ref=o.prediction_vectors(o.window_gather(emb.double(),lpad,rpad),W.double(),bias.double())
    out=routing.uhat_fwd(emb.cuda(),W.cuda(),bias.cuda(),lpad,rpad,mode)
    torch.cuda.synchronize()
    err=((out.double().cpu()-ref).abs().max()/ref.abs().max()).item()
    print((B,S,H,d,O,D,lpad,rpad,mode),'rel err %.3e'%err, flush=True)
    return err
for mode in ('tf32','bf16'):
    run(2,5,6,8,5,8,1,1,mode)
    run(8,75,60,8,30,8,1,1,mode)
    run(3,9,30,8,63,8,3,3,mode)
    run(64,7,60,20,30,20,2,2,mode)
    run(5,11,30,20,32,20,2,2,mode)
    run(2,6,7,16,9,16,0,0,mode)
    run(1,3,5,32,6,32,4,4,mode)
    run(2,4,9,8,100,8,0,1,mode)
    run(70,3,9,4,10,12,1,0,mode)
