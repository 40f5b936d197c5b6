"""Randomised parity fuzz of whole routing stacks (random knobs x u_hat modes) against the CPU
oracle: logits within the mode's tolerance (scaled by the instance's conditioning) and, for the
1e-4 class, identical greedy-CTC strings; training-mode forward with injected dropout masks for a
subset.  Run on a GPU box from the repository root:  python tests/dev/fuzz_stack.py [n] [seed]"""
import random
import sys

sys.path.insert(0, ".")
import torch

from oracle import srf_oracle as o
from srf_b200 import RoutingStack

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
TOL = {"exact": 1e-4, "fp32": 1e-4, "fp32x3": 1e-4, "tf32": 1e-2, "f16": 1e-2, "bf16": 2e-2}


def rel(a, ref):
  return ((a.double().cpu() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-30)).item()


bad = done = skipped = 0
for case in range(N):
  mode = rng.choice(["exact", "fp32", "fp32x3", "tf32", "f16", "bf16"])
  L = rng.randint(1, 3)
  DIM = rng.choice([4, 8, 12, 16, 20])
  PH, CH = rng.randint(2, 24), rng.randint(2, 40)
  class_n = rng.choice([3, 9, 32, 33, 63, 64, 70, 100])
  if DIM > 8 and class_n > 64:
    class_n = 63
  lpad, rpad = rng.randint(0, 3), rng.randint(0, 3)
  iters, sdr = rng.randint(1, 3), rng.random() < 0.6
  B, S = rng.randint(1, 7), rng.randint(1, 12)
  tag = (case, mode, L, PH, CH, class_n, DIM, lpad, rpad, iters, sdr, B, S)
  shapes = o.layer_shapes(L, PH, CH, class_n, DIM, DIM, DIM, lpad + rpad + 1)
  p32 = o.init_params(shapes, class_n, seed=case, random_ln=True)
  g = torch.Generator().manual_seed(1000 + case)
  emb = torch.randn(B, S, PH, DIM, generator=g)
  masks = None
  if case % 4 == 0:
    masks = [((torch.rand(B, S, s[1], s[2], generator=g) < 0.9).float() / 0.9) for s in shapes]
  ref = o.route_stack(emb.double(), p32.to(torch.float64), lpad, rpad, iters, sdr,
                      dropout_masks=None if masks is None else [m.double() for m in masks])
  ref32 = o.route_stack(emb, p32, lpad, rpad, iters, sdr, dropout_masks=masks)
  amp = rel(ref32, ref) / 6e-8
  pert = {"exact": 1e-6, "fp32": 1e-6, "fp32x3": 1e-6, "tf32": 1e-3, "f16": 1e-3, "bf16": 4e-3}[mode]
  tol = max(TOL[mode] * (iters if mode in ("tf32", "f16", "bf16") else 1), 3 * amp * pert)
  stack = RoutingStack(L, PH, CH, class_n, DIM, DIM, DIM, lpad, rpad, iters, sdr, seed=0, uhat_mode=mode)
  stack.load_oracle_params(p32)
  try:
    logits = stack.forward(emb.cuda(), training=masks is not None,
                           dropout_masks=None if masks is None else [m.cuda() for m in masks])
    torch.cuda.synchronize()
  except ValueError as e:
    skipped += 1
    print("skip", tag, str(e)[:80])
    continue
  err = rel(logits, ref)
  ok = err < tol and torch.isfinite(logits).all().item()
  if mode in ("exact", "fp32", "fp32x3") and amp < 100:
    lens = [S] + [max(1, S - 2)] * (B - 1)
    ok = ok and o.greedy_ctc(logits.cpu(), lens) == o.greedy_ctc(ref, lens)
  done += 1
  if not ok:
    bad += 1
    print("FAIL", tag, "err %.2e tol %.2e amp %.0f kernel %s" % (err, tol, amp, stack.handle.last_kernel[:60]), flush=True)
print("fuzz_stack: %d cases, %d skipped, %d bad" % (done, skipped, bad))
sys.exit(1 if bad else 0)
