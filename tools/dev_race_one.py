import os, sys; sys.path.insert(0,'.')
import torch
from srf_b200 import routing
g=torch.Generator().manual_seed(17)
B,S,H,d,O,D=(3,5,60,8,30,8)
emb=torch.randn(B,S,H,d,generator=g); W=torch.randn(3*H,O,D,d,generator=g)*0.1; bias=torch.randn(3*H,O,D,generator=g)*0.1
os.environ['SRF_FORCE_C']=sys.argv[1] if len(sys.argv)>1 else '1'
h=routing.Handle()
a=routing.LayerArgs(W=W.cuda(),bias=bias.cuda(),lpad=1,rpad=1,iters=2,sdr=True,mask_class0=False,uhat_mode='tf32')
o1=routing.route_layer_fwd(emb.cuda(),a,handle=h)[0]
torch.cuda.synchronize(); print('done', h.last_kernel)
