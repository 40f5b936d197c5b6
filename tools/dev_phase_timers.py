"""Needs a library built with SRF_NVCC_EXTRA=-DSRF_STREAM_PHASE_TIMERS (python -m srf_b200.build --force).
Per-phase clock64 timers of the streaming routing kernel (SRF_PHASE_TIMERS=1)."""
import ctypes, os, sys
os.environ['SRF_PHASE_TIMERS'] = '1'
sys.path.insert(0, '.')
import torch
import numpy as np
import bench
from srf_b200 import RoutingStack
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'cfg3']
B, S = w['B'], (w['T'] + 3) // 4
if len(sys.argv) > 2:
  B = int(sys.argv[2])
stack = RoutingStack(w['L'], w['PH'], w['CH'], w['class_n'], w['DIM'], w['DIM'], w['DIM'], w['lpad'], w['rpad'],
                     w['iters'], w['sdr'], seed=0, uhat_mode='bf16')
emb = torch.randn(B, S, w['PH'], w['DIM'], device='cuda')
for _ in range(3):
  stack.forward(emb)
torch.cuda.synchronize()
lib = stack.handle.lib
lib.srf_debug_phase_timers.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
buf = np.zeros((1024, 8), dtype=np.uint64)
lib.srf_debug_phase_timers(stack.handle._h, buf.ctypes.data_as(ctypes.c_void_p), 1024)   # clear
stack.forward(emb)
lib.srf_debug_phase_timers(stack.handle._h, buf.ctypes.data_as(ctypes.c_void_p), 1024)
used = buf[buf[:, 6] > 0]
names = ['capsule loop', 'store partials + CTA barrier', 'local sum + push', 'wait peers', 'cluster sum + barrier', 'squash + hand-off']
print(stack.handle.last_kernel)
print('CTAs', len(used), 'passes per CTA', used[:, 6].mean())
tot = 0
for i, n in enumerate(names):
  per = used[:, i] / used[:, 6]
  tot += per.mean()
  print('%-30s mean %7.0f clk  min %7.0f  max %7.0f' % (n, per.mean(), per.min(), per.max()))
print('sum %.0f clk per pass' % tot)
