"""Timing of the fused kernel under developer environment switches (one process, switches toggled between runs).
usage: dev_exp.py VAR=val,VAR2=val ...   (each argument = one variant; '-' = defaults)"""
import os, sys
sys.path.insert(0, '.')
import torch
import bench
from srf_b200 import RoutingStack


def run(name, mode, B=None, reps=10):
  w = bench.WORKLOADS[name]
  Bw, Sw = (B or w["B"]), (w["T"] + 3) // 4
  st = RoutingStack(w["L"], w["PH"], w["CH"], w["class_n"], w["DIM"], w["DIM"], w["DIM"], w["lpad"], w["rpad"],
                    w["iters"], w["sdr"], seed=0, uhat_mode=mode)
  e = torch.randn(Bw, Sw, w["PH"], w["DIM"], device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
  o = torch.empty(Bw, Sw, w["class_n"], device="cuda")
  for _ in range(3):
    st.forward(e, out_logits=o)
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    st.forward(e, out_logits=o)
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / reps, st.handle.last_kernel, o.clone()


cases = (("cfg3", None), ("cfg3", 8), ("cfg2", None))
ref = {}
for var in sys.argv[1:] or ['-']:
  sets = [kv.split('=') for kv in var.split(',') if '=' in kv]
  for k, v in sets:
    os.environ[k] = v
  for name, B in cases:
    ms, k, out = run(name, "f16", B)
    key = (name, B)
    if key not in ref:
      ref[key] = out
    print("[%s] %s B=%s: %.3f ms  maxdiff vs first %.3e  %s" % (var, name, B, ms, (out - ref[key]).abs().max().item(),
                                                               k[k.find('grid'):]), flush=True)
  for k, v in sets:
    del os.environ[k]
