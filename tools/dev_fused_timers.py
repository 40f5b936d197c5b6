"""Per-phase clock64 timers of the fused routing kernel (thread 0 of team 0 in every CTA).
Needs a library built with SRF_NVCC_EXTRA=-DSRF_FUSED_TIMERS (python -m srf_b200.build --force)
and SRF_PHASE_TIMERS=1 (set here).  usage: dev_fused_timers.py [cfg3|cfg1|cfg2] [mode] [B]"""
import ctypes, os, sys
os.environ['SRF_PHASE_TIMERS'] = '1'
sys.path.insert(0, '.')
import torch
import numpy as np
import bench
from srf_b200 import RoutingStack
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'cfg3']
mode = sys.argv[2] if len(sys.argv) > 2 else 'tf32'
B, S = w['B'], (w['T'] + 3) // 4
if len(sys.argv) > 3:
  B = int(sys.argv[3])
stack = RoutingStack(w['L'], w['PH'], w['CH'], w['class_n'], w['DIM'], w['DIM'], w['DIM'], w['lpad'], w['rpad'],
                     w['iters'], w['sdr'], seed=0, uhat_mode=mode)
emb = torch.randn(B, S, w['PH'], w['DIM'], device='cuda')
for _ in range(3):
  stack.forward(emb)
torch.cuda.synchronize()
lib = stack.handle.lib
lib.srf_debug_phase_timers.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
buf = np.zeros((1024, 8), dtype=np.uint64)
lib.srf_debug_phase_timers(stack.handle._h, buf.ctypes.data_as(ctypes.c_void_p), 1024)   # clear
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
stack.forward(emb)
e1.record()
torch.cuda.synchronize()
lib.srf_debug_phase_timers(stack.handle._h, buf.ctypes.data_as(ctypes.c_void_p), 1024)
import re
ncta = int(re.search(r'grid=(\d+)', stack.handle.last_kernel).group(1))
t = buf.reshape(512, 16)[:ncta]
used = t[t[:, 15] > 0].astype(np.float64)
names = {0: 'pass prologue', 1: 'wait t_full (MMA)', 2: 'team barrier X', 3: 'S(n-1) + A(n)', 13: 'team barrier Y',
         14: 'B(n-1) (+ SA when no B)', 7: 'store partials', 8: 'CTA barrier + flag', 9: 'owner: wait partials',
         10: 'owner: load + sum partials', 4: 'owner: squash, store v, flag', 6: 'owner: rest',
         11: 'wait v', 12: 'load v'}
print(stack.handle.last_kernel)
print('launch %.3f ms; CTAs %d, passes per CTA %.0f' % (e0.elapsed_time(e1), len(used), used[:, 15].mean()))
tot = 0
for i, n in names.items():
  per = used[:, i] / used[:, 15]
  tot += per.mean()
  print('%-48s mean %8.0f clk/pass  min %8.0f  max %8.0f' % (n, per.mean(), per.min(), per.max()))
print('sum %.0f clk per pass' % tot)

tr = buf.reshape(-1).view(np.uint32)
for name, off in (('CTA 0', 6000), ('CTA 80', 6000 + 3072)):
  e = tr[off:off + 3072].reshape(512, 6).astype(np.int64)
  e = e[40:500]   # steady state
  d = lambda a, b: ((e[:, b] - e[:, a]) & 0xffffffff)
  nxt = ((e[1:, 0] - e[:-1, 4]) & 0xffffffff)
  print('MMA issuer trace %s (clk per capsule, median / mean): wait x_full %d / %d; wait t_empty(tile 0) %d / %d; wait w_full %d / %d; issue MMAs %d / %d; commits %d / %d; loop back %d / %d' % (
      name, np.median(d(0, 1)), d(0, 1).mean(), np.median(d(1, 2)), d(1, 2).mean(), np.median(d(2, 5)), d(2, 5).mean(), np.median(d(5, 3)), d(5, 3).mean(),
      np.median(d(3, 4)), d(3, 4).mean(), np.median(nxt), nxt.mean()))
