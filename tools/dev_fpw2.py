"""Time the two-members-per-warp streaming variants (large per-lane state) for a slot count."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools.sweep as sw
base = dict(L=4, PH=60, CH=30, cls=32, DIM=20, lpad=2, rpad=2, iters=1, sdr=True, B=16, S=1024)
for name, over in (("DIM=32", {"DIM": 32}), ("PH=120,CH=60", {"PH": 120, "CH": 60}), ("DIM=16,CH=60", {"DIM": 16, "CH": 60, "PH": 120})):
  cfg = dict(base); cfg.update(over)
  r = sw.run(mode="bf16", **cfg)
  print("NSLOT_FPW2=%s %-14s ms %.2f gemm %.2f route %.2f (hbm %.2f)" % (
      os.environ.get("SRF_STREAM_NSLOT_FPW2", "8"), name, r["ms"], r["gemm_ms"], r["route_ms"], r["route_hbm"]), flush=True)
