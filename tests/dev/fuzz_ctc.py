"""Randomised fuzz of the native CTC loss / gradient and greedy decode against torch (float64 CPU)
and the oracle's greedy rule.  python tests/dev/fuzz_ctc.py [n] [seed]"""
import random
import sys

sys.path.insert(0, ".")
import torch

from oracle import srf_oracle as o
from srf_b200 import training

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for case in range(N):
  B, S, C = rng.randint(1, 9), rng.randint(1, 60), rng.choice([2, 3, 5, 9, 32, 63, 100])
  Lmax = rng.randint(1, max(1, min(S, 25)))
  g = torch.Generator().manual_seed(case)
  logits = torch.randn(B, S, C, generator=g, dtype=torch.float64) * rng.choice([0.5, 2.0, 6.0])
  labels = torch.randint(0, max(1, C - 1), (B, Lmax), generator=g)
  if rng.random() < 0.5 and Lmax > 1:
    labels[:, 1] = labels[:, 0]          # repeats need a blank in between
  in_len = torch.tensor([rng.randint(1, S) for _ in range(B)])
  lab_len = torch.tensor([rng.randint(0 if rng.random() < 0.1 else 1, Lmax) for _ in range(B)])
  lt = logits.clone().requires_grad_(True)
  ref = torch.nn.functional.ctc_loss(torch.log_softmax(lt, -1).transpose(0, 1), labels, in_len, lab_len,
                                     blank=C - 1, reduction="none", zero_infinity=True)
  ref.sum().backward()
  loss, d = training.ctc_loss(logits.float().cuda(), labels.cuda(), in_len.cuda(), lab_len.cuda(), blank=C - 1)
  torch.cuda.synchronize()
  ok = torch.allclose(loss.double().cpu(), ref.detach(), rtol=5e-5, atol=2e-4)
  gmax = lt.grad.abs().max().clamp_min(1e-6)
  gerr = ((d.double().cpu() - lt.grad).abs().max() / gmax).item()
  ok = ok and gerr < 5e-4 and torch.isfinite(d).all().item()
  lens = in_len.tolist()
  ok = ok and training.ctc_greedy_decode(logits.float().cuda(), lens, blank=C - 1) == \
      o.greedy_ctc(logits.float(), lens)
  if not ok:
    bad += 1
    print("FAIL", (case, B, S, C, Lmax), "loss", loss.cpu().tolist()[:3], ref.detach().tolist()[:3], "gerr %.2e" % gerr,
          in_len.tolist(), lab_len.tolist(), flush=True)
print("fuzz_ctc: %d cases, %d bad" % (N, bad))
sys.exit(1 if bad else 0)
