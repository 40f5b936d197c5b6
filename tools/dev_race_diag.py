import os, sys; sys.path.insert(0,'.')
import torch
from srf_b200 import routing
g=torch.Generator().manual_seed(17)
B,S,H,d,O,D=(int(sys.argv[2]),int(sys.argv[3]),int(sys.argv[4]),8,30,8)
emb=torch.randn(B,S,H,d,generator=g); W=torch.randn(int(sys.argv[5])*H,O,D,d,generator=g)*0.1; bias=torch.randn(int(sys.argv[5])*H,O,D,generator=g)*0.1
os.environ['SRF_FORCE_C']='1'
os.environ['SRF_STREAM_STAGES']=sys.argv[1] if len(sys.argv)>1 else '2'
a=routing.LayerArgs(W=W.cuda(),bias=bias.cuda(),lpad=(int(sys.argv[5])-1)//2,rpad=int(sys.argv[5])-1-(int(sys.argv[5])-1)//2,iters=1,sdr=True,mask_class0=False,uhat_mode='tf32')
h=routing.Handle()
x=routing.route_layer_fwd(emb.cuda(),a,handle=h)[0]
torch.cuda.synchronize(); print(h.last_kernel)
os.environ['SRF_NO_STREAM']='1'
h2=routing.Handle()
ref=routing.route_layer_fwd(emb.cuda(),a,handle=h2)[0]
torch.cuda.synchronize(); print(h2.last_kernel)
d=(x-ref).abs()
print('max diff',d.max().item())
bad=(d>1e-5).nonzero()
print('n bad',len(bad),'of',x.numel())
print(bad[:40].cpu().numpy().tolist())
print('per (b,s):',(d>1e-5).sum(dim=(2,3)).cpu().numpy())
print('per k:',(d>1e-5).sum(dim=(0,1,2)).cpu().numpy())
print('per j:',(d>1e-5).sum(dim=(0,1,3)).cpu().numpy())
for b in range(B):
  for s in range(S):
    row=[]
    for s2 in range(S):
      row.append('%.1e'%(x[b,s]-ref[b,s2]).abs().max().item())
    print('x[b=%d,s=%d] vs ref[b,s2]:'%(b,s),row, 'mean|x| %.2e'%x[b,s].abs().mean().item())
