"""Dev aid: fused routing kernel vs the float64 oracle on a list of layer / stack shapes.
Prints the relative error per case instead of asserting (one GPU call = the whole picture)."""
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import srf_oracle as o  # noqa: E402
from srf_b200 import RoutingStack, routing  # noqa: E402


def rel(a, b):
  a, b = a.double().cpu(), b.double().cpu()
  return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def mk(B, S, H, d, O, D, win, seed):
  g = torch.Generator().manual_seed(seed)
  emb = torch.randn(B, S, H, d, generator=g)
  W = torch.randn(win * H, O, D, d, generator=g) * 0.1
  bias = torch.randn(win * H, O, D, generator=g) * 0.1
  return emb, W, bias


LAYERS = [
    (2, 9, 6, 8, 5, 8, 1, 1),
    (3, 5, 60, 8, 30, 8, 1, 1),
    (2, 6, 30, 8, 63, 8, 1, 1),
    (8, 5, 30, 8, 30, 8, 3, 3),
    (64, 3, 60, 20, 30, 20, 2, 2),
    (5, 4, 30, 20, 32, 20, 2, 2),
    (2, 5, 7, 16, 9, 16, 0, 0),
    (3, 4, 9, 4, 10, 12, 1, 0),
    (40, 6, 12, 8, 20, 8, 1, 1),
]


def main():
  modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["tf32", "fp32x3"]
  only = int(sys.argv[2]) if len(sys.argv) > 2 else -1
  h = routing.default_handle(torch.device("cuda:0"))
  for ci, case in enumerate(LAYERS):
    if only >= 0 and ci != only:
      continue
    B, S, H, d, O, D, lpad, rpad = case
    emb, W, bias = mk(B, S, H, d, O, D, lpad + rpad + 1, 17)
    for mode in modes:
      for sdr in (True, False):
        for iters, last in ((1, False), (3, True)):
          ref = o.route_layer(emb.double(), W.double(), bias.double(), lpad, rpad, iters, sdr, last)
          try:
            t0 = time.time()
            a = routing.LayerArgs(W=W, bias=bias, lpad=lpad, rpad=rpad, iters=iters, sdr=sdr,
                                  mask_class0=last, uhat_mode=mode)
            caps, _ = routing.route_layer_fwd(emb.cuda(), a, handle=h)
            torch.cuda.synchronize()
            print("layer %s %s %s it=%d last=%d: rel %.3e  (%.1f ms) %s" % (
                case, mode, "SDR" if sdr else "DR", iters, last, rel(caps, ref), (time.time() - t0) * 1e3,
                h.last_kernel[:60]), flush=True)
          except Exception as e:  # noqa: BLE001
            print("layer %s %s %s it=%d last=%d: EXC %s" % (case, mode, "SDR" if sdr else "DR", iters, last, e),
                  flush=True)
            traceback.print_exc()
  STACKS = [
      ("timit-sdr", 7, 60, 30, 63, 8, 1, 1, 1, True, 3, 10),
      ("timit-dr3", 3, 60, 30, 63, 8, 3, 3, 3, False, 2, 6),
      ("wsj-sdr", 4, 60, 30, 32, 20, 2, 2, 1, True, 40, 9),
      ("sdr-it2", 3, 20, 10, 12, 8, 1, 2, 2, True, 3, 7),
  ]
  for name, L, PH, CH, class_n, DIM, lpad, rpad, iters, sdr, B, S in STACKS:
    if only >= 0:
      break
    shapes = o.layer_shapes(L, PH, CH, class_n, DIM, DIM, DIM, lpad + rpad + 1)
    p32 = o.init_params(shapes, class_n, seed=11, random_ln=True)
    emb = torch.randn(B, S, PH, DIM, generator=torch.Generator().manual_seed(12))
    ref = o.route_stack(emb.double(), p32.to(torch.float64), lpad, rpad, iters, sdr)
    for mode in modes:
      try:
        stack = RoutingStack(L, PH, CH, class_n, DIM, DIM, DIM, lpad, rpad, iters, sdr, seed=0, uhat_mode=mode)
        stack.load_oracle_params(p32)
        t0 = time.time()
        logits = stack.forward(emb.cuda())
        torch.cuda.synchronize()
        lens = [S] + [max(1, S - 3)] * (B - 1)
        same = o.greedy_ctc(logits.cpu(), lens) == o.greedy_ctc(ref, lens)
        print("stack %s %s: rel %.3e ctc_same=%s (%.1f ms) %s" % (name, mode, rel(logits, ref), same,
                                                                  (time.time() - t0) * 1e3, stack.handle.last_kernel),
              flush=True)
      except Exception as e:  # noqa: BLE001
        print("stack %s %s: EXC %s" % (name, mode, e), flush=True)
        traceback.print_exc()


if __name__ == "__main__":
  main()
