"""Soak test of the backward kernels: many shapes x u_hat modes, every launch repeated and compared
bit for bit (dW / dbias / d_emb are atomics-free, so any difference is a race), plus finiteness."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srf_b200 import routing

REPEAT = int(sys.argv[1]) if len(sys.argv) > 1 else 8
bad = total = 0
g = torch.Generator().manual_seed(0)
shapes = [  # B, S, H, d, O, D, lpad, rpad
    (8, 30, 30, 20, 30, 20, 2, 2), (3, 25, 12, 8, 30, 8, 1, 1), (2, 17, 20, 16, 40, 16, 1, 0),
    (5, 12, 9, 8, 100, 8, 0, 1), (1, 40, 60, 20, 32, 20, 1, 1), (7, 9, 16, 32, 20, 32, 1, 1),
    (16, 20, 30, 8, 63, 8, 1, 1), (4, 33, 15, 4, 30, 12, 2, 2)]
for (B, S, H, d, O, D, lpad, rpad), mode, sdr, iters in itertools.product(
    shapes, ("fp32", "tf32", "bf16"), (True, False), (1, 3)):
  I = (lpad + rpad + 1) * H
  emb = torch.randn(B, S, H, d, generator=g).cuda()
  W = (torch.randn(I, O, D, d, generator=g) * 0.1).cuda()
  bias = (torch.randn(I, O, D, generator=g) * 0.1).cuda()
  args = routing.LayerArgs(W=W, bias=bias, lpad=lpad, rpad=rpad, iters=iters, sdr=sdr, mask_class0=True,
                           ln_gamma=torch.ones(O * D).cuda(), ln_beta=torch.zeros(O * D).cuda(),
                           uhat_mode=mode)
  try:
    _, _, raw = routing.route_layer_fwd_train(emb, args)
  except ValueError as e:   # shape not instantiated for this mode
    print("skip", (B, S, H, d, O, D), mode, str(e)[:60])
    continue
  dout = torch.randn(B, S, O, D, generator=g).cuda()
  first = None
  for _ in range(REPEAT):
    got = routing.route_layer_bwd(emb, args, raw, d_out=dout)
    torch.cuda.synchronize()
    cur = [got[k].clone() for k in ("dW", "dbias", "d_emb")]
    if first is None:
      first = cur
    elif not all(torch.equal(a, b) for a, b in zip(cur, first)):
      bad += 1
      print("MISMATCH", (B, S, H, d, O, D, lpad, rpad), mode, sdr, iters)
      break
  if not all(torch.isfinite(t).all() for t in first):
    bad += 1
    print("NONFINITE", (B, S, H, d, O, D, lpad, rpad), mode, sdr, iters)
  total += 1
print("soak_bwd: %d configurations x %d launches, %d bad" % (total, REPEAT, bad))
sys.exit(1 if bad else 0)
