import os, sys; sys.path.insert(0,'.')
os.environ['SRF_STREAM_FP32']='1'
import torch
from srf_b200 import routing
def mk(B,S,H,d,O,D,win,seed):
    g=torch.Generator().manual_seed(seed)
    return torch.randn(B,S,H,d,generator=g), torch.randn(win*H,O,D,d,generator=g)*0.1, torch.randn(win*H,O,D,generator=g)*0.1
B,S,H,d,O,D,lpad,rpad=(2,6,30,8,63,8,1,1)
emb,W,bias=mk(B,S,H,d,O,D,3,3)
for C in ('1','2','4','8'):
  os.environ['SRF_FORCE_C']=C
  for sdr in (True,False):
    h=routing.Handle()
    a=routing.LayerArgs(W=W.cuda(),bias=bias.cuda(),lpad=lpad,rpad=rpad,iters=1,sdr=sdr,mask_class0=False,uhat_mode='tf32')
    outs=[routing.route_layer_fwd(emb.cuda(),a,handle=h)[0].clone() for _ in range(20)]
    torch.cuda.synchronize()
    nb=sum((x!=outs[0]).any().item() for x in outs[1:])
    bad=[k for k,x in enumerate(outs) if (x!=outs[0]).any()]
    info=''
    if bad:
        dd=(outs[bad[0]]!=outs[0]); idx=dd.nonzero()
        info=' first bad call %d: n=%d b=%s s=%s j-range=%d..%d'%(bad[0],len(idx),idx[:,0].unique().tolist(),idx[:,1].unique().tolist(),idx[:,2].min(),idx[:,2].max())
    print('C',C,'sdr',sdr,'bad %d/19'%nb,h.last_kernel[-50:],info)
    h.close()
