#!/usr/bin/env python
"""cfg-5 of BASELINE.json: routing sweep (DIM 8-32, PH/CH heights, window 1-9, ITER 1-5) on long
synthetic sequences fed directly to the routing stack.  One factor at a time around the
WSJ-shaped centre (L=4 layers to bound the run).  Prints a markdown table; run on the GPU box:

    python tools/sweep.py > gpurun_out/sweep.md
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from srf_b200 import RoutingStack, layer_shapes  # noqa: E402

P_HBM, P_TENSOR = 6451.8, 1429.9
try:
  pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
  P_HBM, P_TENSOR = pk["hbm_gbs"], pk["bf16_tflops_sustained"]
except Exception:  # pylint: disable=broad-except
  pass


def run(L, PH, CH, cls, DIM, lpad, rpad, iters, sdr, B, S, mode, reps=3):
  st = RoutingStack(L, PH, CH, cls, DIM, DIM, DIM, lpad, rpad, iters, sdr, seed=0, uhat_mode=mode)
  emb = torch.randn(B, S, PH, DIM, device="cuda")
  out = torch.empty(B, S, cls, device="cuda")
  for _ in range(2):
    st.forward(emb, out_logits=out)
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  torch.cuda.synchronize()
  a.record()
  for _ in range(reps):
    st.forward(emb, out_logits=out)
  b.record()
  torch.cuda.synchronize()
  ms = a.elapsed_time(b) / reps
  st.handle.profile_begin()
  st.forward(emb, out_logits=out)
  kp = st.handle.profile_end()
  shapes = layer_shapes(L, PH, CH, cls, DIM, DIM, DIM, lpad + rpad + 1)
  f_uhat = sum(2 * I * O * D * d for I, O, D, d in shapes) * B * S
  esize = {"bf16": 2, "fp32x3": 4}.get(mode, 0)   # f16 / tf32: fused kernel, u_hat is never stored
  uhat_bytes = sum(I * O * D for I, O, D, d in shapes) * B * S * esize
  io_bytes = sum(4 * (I // (lpad + rpad + 1) * d + O * D) for I, O, D, d in shapes) * B * S
  res = {"ms": ms, "fps": B * S / ms * 1e3, "tensor_frac": f_uhat / (ms / 1e3) / 1e12 / P_TENSOR,
         "gemm_ms": kp["uhat_gemm"][0], "route_ms": kp["routing"][0]}
  if esize and kp["uhat_gemm"][0] > 0:   # (0: the library's policy took the fused kernel, nothing is stored)
    res["gemm_hbm"] = uhat_bytes / (kp["uhat_gemm"][0] / 1e3) / 1e9 / P_HBM
    res["route_hbm"] = (uhat_bytes + io_bytes) / (kp["routing"][0] / 1e3) / 1e9 / P_HBM
  del st, emb, out
  torch.cuda.empty_cache()
  return res


def main():
  base = dict(L=4, PH=60, CH=30, cls=32, DIM=20, lpad=2, rpad=2, iters=1, sdr=True, B=16, S=1024)
  sweeps = [("centre (WSJ-shaped, L=4, 16 x 1024 frames)", [{}]),
            ("DIM", [{"DIM": v} for v in (8, 16, 20, 32)]),
            ("PH/CH", [{"PH": a, "CH": b} for a, b in ((30, 15), (60, 30), (120, 60))]),
            ("window", [{"lpad": l, "rpad": r} for l, r in ((0, 0), (1, 1), (2, 2), (3, 3), (4, 4))]),
            ("ITER (SDR)", [{"iters": v} for v in (1, 2, 3, 5)]),
            ("ITER (DR)", [{"iters": v, "sdr": False} for v in (1, 2, 3, 5)]),
            ("S (SDR, frames fixed at 16k)", [{"S": s, "B": 16384 // s} for s in (375, 1024, 4096)])]
  print("| sweep | config | mode | ms/step | k routing frames/s | u_hat GEMM ms (HBM frac) | routing ms (HBM frac) | tensor frac of fused roofline |")
  print("|---|---|---|---|---|---|---|---|")
  for name, variants in sweeps:
    for v in variants:
      cfg = dict(base)
      cfg.update(v)
      for mode in ("f16", "bf16", "fp32x3"):
        t0 = time.time()
        try:
          r = run(mode=mode, **cfg)
        except Exception as e:  # pylint: disable=broad-except
          print("| %s | %s | %s | error: %s | | | | |" % (name, v or "-", mode, str(e)[:60]))
          continue
        print("| %s | %s | %s | %.2f | %.0f | %.2f (%s) | %.2f (%s) | %.4f |" % (
            name, ", ".join("%s=%s" % kv for kv in v.items()) or "-", mode, r["ms"], r["fps"] / 1e3,
            r["gemm_ms"], "%.2f" % r["gemm_hbm"] if "gemm_hbm" in r else "-",
            r["route_ms"], "%.2f" % r["route_hbm"] if "route_hbm" in r else "-", r["tensor_frac"]))
        sys.stdout.flush()
        if time.time() - t0 > 60:
          break


if __name__ == "__main__":
  main()
