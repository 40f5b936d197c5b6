"""Soak of the fused wavefront kernel: repeated launches of WSJ-shaped stacks must be bit-identical (fixed summation
order; any race in the rings, the L2 exchange or the progress flags shows up as a differing bit), batch sizes with full,
ragged and single frame groups.  Run on a GPU box from the repository root: python tests/dev/dev_fused_soak.py [reps]"""
import sys

sys.path.insert(0, ".")
import torch

import bench
from srf_b200 import RoutingStack

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 25
w = bench.WORKLOADS["cfg3"]
bad = 0
for mode in ("f16", "tf32"):
  stack = RoutingStack(w["L"], w["PH"], w["CH"], w["class_n"], w["DIM"], w["DIM"], w["DIM"], w["lpad"], w["rpad"],
                       w["iters"], w["sdr"], seed=0, uhat_mode=mode)
  for B, S in ((64, 375), (40, 120), (8, 375), (20, 200), (1, 90)):
    emb = torch.randn(B, S, w["PH"], w["DIM"], device="cuda", generator=torch.Generator(device="cuda").manual_seed(B))
    first = stack.forward(emb).clone()
    diff = 0
    for _ in range(reps):
      out = stack.forward(emb)
      diff += int(not torch.equal(out, first))
    torch.cuda.synchronize()
    ok = torch.isfinite(first).all().item() and diff == 0
    bad += int(not ok)
    print("%s B=%d S=%d: %d launches, %d differ, finite %s  %s" % (mode, B, S, reps, diff, torch.isfinite(first).all().item(),
                                                                 stack.handle.last_kernel[:60]), flush=True)
print("fused soak: %d bad" % bad)
sys.exit(1 if bad else 0)
