import sys, os, torch
sys.path.insert(0, '.')
import bench
from srf_b200 import RoutingStack
def run(name, mode, B=None):
    w = bench.WORKLOADS[name]
    Bw, Sw = (B or w["B"]), (w["T"] + 3) // 4
    st = RoutingStack(w["L"], w["PH"], w["CH"], w["class_n"], w["DIM"], w["DIM"], w["DIM"], w["lpad"], w["rpad"], w["iters"], w["sdr"], seed=0, uhat_mode=mode)
    e = torch.randn(Bw, Sw, w["PH"], w["DIM"], device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    o = torch.empty(Bw, Sw, w["class_n"], device="cuda")
    for _ in range(3): st.forward(e, out_logits=o)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): st.forward(e, out_logits=o)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 10, st.handle.last_kernel[:60], o.clone()
for name, B in (("cfg3", None), ("cfg3", 8), ("cfg1", None), ("cfg2", None)):
    ref = None
    for mode in ("tf32", "f16"):
        ms, k, out = run(name, mode, B)
        if ref is None: ref = out
        print("%s B=%s %-5s: %.3f ms  %s  maxdiff vs tf32 %.3e (scale %.2f)" % (name, B, mode, ms, k, (out - ref).abs().max().item(), ref.abs().max().item()), flush=True)
