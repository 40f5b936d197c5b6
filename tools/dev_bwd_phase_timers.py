"""Needs a library built with SRF_NVCC_EXTRA=-DSRF_BWD_PHASE_TIMERS (python -m srf_b200.build --force).
clock64 phase timers of the BPTT sweep (SRF_PHASE_TIMERS=1; last pass of the frame only is
meaningful for ITER=1)."""
import ctypes, os, sys
os.environ['SRF_PHASE_TIMERS'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from srf_b200 import routing
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = int(sys.argv[2]) if len(sys.argv) > 2 else 375
mode = sys.argv[3] if len(sys.argv) > 3 else "bf16"
H, d, O, D = 30, 20, 30, 20
g = torch.Generator().manual_seed(0)
emb = torch.randn(B, S, H, d, generator=g).cuda()
W = (torch.randn(5 * H, O, D, d, generator=g) * 0.1).cuda()
bias = (torch.randn(5 * H, O, D, generator=g) * 0.1).cuda()
args = routing.LayerArgs(W=W, bias=bias, lpad=2, rpad=2, iters=1, sdr=True, mask_class0=False,
                         ln_gamma=torch.ones(O * D).cuda(), ln_beta=torch.zeros(O * D).cuda(), uhat_mode=mode)
caps, lg, raw = routing.route_layer_fwd_train(emb, args)
dout = torch.randn(B, S, O, D, generator=g).cuda()
h = routing.default_handle()
lib = h.lib
lib.srf_debug_phase_timers.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
buf = np.zeros((1024, 8), dtype=np.uint64)
for _ in range(2):
  routing.route_layer_bwd(emb, args, raw, d_out=dout)
torch.cuda.synchronize()
lib.srf_debug_phase_timers(h._h, buf.ctypes.data_as(ctypes.c_void_p), 1024)  # clear
routing.route_layer_bwd(emb, args, raw, d_out=dout)
torch.cuda.synchronize()
lib.srf_debug_phase_timers(h._h, buf.ctypes.data_as(ctypes.c_void_p), 1024)
used = buf[buf[:, 7] > 0]
names = ['frame start (gout/vacc, sync, fetch, L2 prefetch)', 'fwd capsule visits', 'fwd reduce + cluster exchange',
         'fwd squash (ITER > 1)', '-', 'g_t (per warp) + bwd capsule visits', 'bwd reduce + cluster exchange']
print('CTAs', len(used), 'frames per CTA', used[:, 7].mean())
tot = 0
for i, n in enumerate(names):
  per = used[:, i] / used[:, 7]
  tot += per.mean()
  print('%-52s mean %7.0f clk  min %7.0f  max %7.0f' % (n, per.mean(), per.min(), per.max()))
print('sum %.0f clk per frame' % tot)
