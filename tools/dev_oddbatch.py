import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools.sweep as sw
base = dict(L=4, PH=60, CH=30, cls=32, DIM=20, lpad=2, rpad=2, iters=1, sdr=True)
for B, S in ((43, 375), (16, 1024), (5, 300), (64, 375)):
  r = sw.run(mode="bf16", B=B, S=S, **base)
  print("B=%d S=%d ms %.2f gemm %.2f (hbm %.2f) route %.2f" % (B, S, r["ms"], r["gemm_ms"], r["gemm_hbm"], r["route_ms"]), flush=True)
