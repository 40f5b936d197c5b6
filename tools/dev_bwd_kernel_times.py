"""Per-kernel times of one layer backward (cfg-3 mid layer, B x S frames) via profile spans."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srf_b200 import routing
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = int(sys.argv[2]) if len(sys.argv) > 2 else 375
mode = sys.argv[3] if len(sys.argv) > 3 else "bf16"
H, d, O, D, lpad, rpad = 30, 20, 30, 20, 2, 2
g = torch.Generator().manual_seed(0)
emb = torch.randn(B, S, H, d, generator=g).cuda()
W = (torch.randn(5 * H, O, D, d, generator=g) * 0.1).cuda()
bias = (torch.randn(5 * H, O, D, generator=g) * 0.1).cuda()
args = routing.LayerArgs(W=W, bias=bias, lpad=lpad, rpad=rpad, iters=1, sdr=True, mask_class0=False,
                         ln_gamma=torch.ones(O * D).cuda(), ln_beta=torch.zeros(O * D).cuda(), uhat_mode=mode)
caps, lg, raw = routing.route_layer_fwd_train(emb, args)
dout = torch.randn(B, S, O, D, generator=g).cuda()
for it in range(3):
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  got = routing.route_layer_bwd(emb, args, raw, d_out=dout)
  e1.record()
  torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  got = routing.route_layer_bwd(emb, args, raw, d_out=dout)
  torch.cuda.synchronize()
for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:6]:
  print("  %-60s %8.3f ms x%d" % (ev.key[:60], ev.device_time_total / 1e3, ev.count))
print("layer bwd B=%d S=%d %s variant=%s: %.3f ms" % (B, S, mode, os.environ.get("SRF_DWDX_VARIANT", "0"), e0.elapsed_time(e1)))
