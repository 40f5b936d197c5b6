"""CTC decode / CTC loss / Adam kernels against independent CPU
references: the oracle's greedy rule, torch's ctc_loss + autograd in float64, and a numpy
restatement of tf.keras Adam with the reference's warm-up schedule."""
import math

import numpy as np
import pytest
import torch

from oracle import srf_oracle as o

pytestmark = pytest.mark.gpu


def test_greedy_decode_matches_oracle_rule():
  from srf_b200 import training
  g = torch.Generator().manual_seed(0)
  for B, S, C in ((3, 17, 9), (2, 40, 32), (5, 12, 63), (1, 1, 4)):
    logits = torch.randn(B, S, C, generator=g)
    logits[..., 0] = -5.0
    # force repeats and blanks
    logits[0, : S // 2, 2 % C] += 8.0
    logits[0, S // 2:, C - 1] += 8.0
    lens = [S] + [max(1, S - 3)] * (B - 1)
    got = training.ctc_greedy_decode(logits.cuda(), lens)
    assert got == o.greedy_ctc(logits, lens)


def test_greedy_decode_tie_breaks_to_lowest_index():
  from srf_b200 import training
  logits = torch.zeros(1, 3, 40)
  logits[0, :, 37] = 1.0
  logits[0, :, 5] = 1.0
  assert training.ctc_greedy_decode(logits.cuda(), [3]) == [[5]]


CTC_CASES = [
    # B, S, C, Lmax
    (3, 12, 9, 4),
    (4, 30, 32, 9),
    (2, 50, 63, 20),
    (2, 6, 5, 5),
]


@pytest.mark.parametrize("case", CTC_CASES)
def test_ctc_loss_and_gradient_match_torch(case):
  from srf_b200 import training
  B, S, C, Lmax = case
  g = torch.Generator().manual_seed(7)
  logits = torch.randn(B, S, C, generator=g, dtype=torch.float64) * 2
  labels = torch.randint(0, C - 1, (B, Lmax), generator=g)
  labels[0, :2] = labels[0, 0]                      # a repeated label needs the blank between
  in_len = torch.tensor([S] + [max(Lmax * 2 + 1, S - 2 - b) for b in range(1, B)]).clamp(max=S)
  lab_len = torch.tensor([Lmax] + [max(1, Lmax - b) for b in range(1, B)])
  lt = logits.clone().requires_grad_(True)
  ref = torch.nn.functional.ctc_loss(torch.log_softmax(lt, -1).transpose(0, 1), labels, in_len, lab_len,
                                     blank=C - 1, reduction="none", zero_infinity=True)
  (ref.sum() * 0.25).backward()
  loss, d = training.ctc_loss(logits.float().cuda(), labels.cuda(), in_len.cuda(), lab_len.cuda(),
                              grad_scale=0.25)
  torch.cuda.synchronize()
  assert torch.allclose(loss.double().cpu(), ref.detach(), rtol=2e-5, atol=1e-4)
  assert ((d.double().cpu() - lt.grad).abs().max() / lt.grad.abs().max()).item() < 2e-4
  for b in range(B):
    assert torch.count_nonzero(d[b, int(in_len[b]):]) == 0


def test_ctc_infeasible_alignment_is_zeroed():
  from srf_b200 import training
  logits = torch.randn(1, 3, 6)
  labels = torch.tensor([[1, 1, 2, 3]])              # needs >= 5 frames
  loss, d = training.ctc_loss(logits.cuda(), labels.cuda(), torch.tensor([3]).cuda(), torch.tensor([4]).cuda())
  assert loss.item() == 0.0 and torch.count_nonzero(d) == 0


def test_ctc_out_of_range_labels_are_zeroed_and_leave_the_other_utterances_alone():
  """A label outside [0, C) must not index out of bounds: that utterance gets loss 0 / zero gradient, the
  others are unaffected; labels beyond lab_len are never looked at."""
  from srf_b200 import training
  g = torch.Generator().manual_seed(3)
  logits = torch.randn(3, 9, 6, generator=g)
  good = torch.tensor([[1, 2, 3, 0], [2, 2, 4, 1], [0, 1, 0, 0]])
  bad = good.clone()
  bad[1, 2] = 17                                     # >= C inside lab_len
  bad[2, 3] = -5                                     # beyond lab_len[2] = 2: ignored
  in_len, lab_len = torch.tensor([9, 9, 7]), torch.tensor([3, 4, 2])
  l0, d0 = training.ctc_loss(logits.cuda(), good.cuda(), in_len.cuda(), lab_len.cuda())
  l1, d1 = training.ctc_loss(logits.cuda(), bad.cuda(), in_len.cuda(), lab_len.cuda())
  torch.cuda.synchronize()
  assert l1[1].item() == 0.0 and torch.count_nonzero(d1[1]) == 0
  for b in (0, 2):
    assert abs(l1[b].item() - l0[b].item()) < 1e-5 * abs(l0[b].item())
    assert (d1[b] - d0[b]).abs().max().item() < 1e-6


def test_adam_matches_keras_formula_with_warmup_schedule():
  from srf_b200 import training
  rng = np.random.default_rng(0)
  shapes = [(7, 5), (33,), (2, 3, 4)]
  params = [torch.tensor(rng.standard_normal(s), dtype=torch.float32).cuda() for s in shapes]
  opt = training.FlatAdam(params, beta1=0.9, beta2=0.98, eps=1e-9)
  p = np.concatenate([x.cpu().numpy().ravel() for x in params]).astype(np.float64)
  m, v = np.zeros_like(p), np.zeros_like(p)
  for t in range(1, 8):
    gflat = rng.standard_normal(p.size)
    lr = training.warmup_lr(t, k=0.5, d_model=256, warmup_steps=4)
    # CustomSchedule (train_helper.py:52-56)
    assert abs(lr - min(0.5 * 256 ** -0.5 * min(t ** -0.5, t * 4 ** -1.5), 10)) < 1e-12
    opt.step(torch.tensor(gflat, dtype=torch.float32).cuda(), lr)
    m = 0.9 * m + 0.1 * gflat
    v = 0.98 * v + 0.02 * gflat ** 2
    lr_t = lr * math.sqrt(1 - 0.98 ** t) / (1 - 0.9 ** t)
    p = p - lr_t * m / (np.sqrt(v) + 1e-9)
  torch.cuda.synchronize()
  got = opt.flat.double().cpu().numpy()
  assert np.abs(got - p).max() / np.abs(p).max() < 1e-5
  # parameters are views into the flat buffer
  assert torch.equal(opt.views[1], opt.flat[35:68])


def test_native_train_step_reduces_the_loss():
  """fwd + CTC + bwd + Adam entirely in the library: a few steps on one batch must lower the loss."""
  from srf_b200 import RoutingStack, training
  L, PH, CH, cls, DIM, B, S = 2, 10, 6, 9, 8, 4, 16
  stack = RoutingStack(L, PH, CH, cls, DIM, DIM, DIM, 1, 1, 1, True, seed=1, inn_dropout=0.0)
  names = [n for n, _ in stack.named_parameters()]
  opt = training.FlatAdam([t for _, t in stack.named_parameters()])
  g = torch.Generator().manual_seed(3)
  emb = torch.randn(B, S, PH, DIM, generator=g).cuda()
  labels = torch.randint(1, cls - 1, (B, 4), generator=g).cuda()
  in_len, lab_len = torch.full((B,), S).cuda(), torch.full((B,), 4).cuda()
  losses = []
  for step in range(1, 13):
    # parameters live in the optimiser's flat buffer: point the stack at the views
    views = dict(zip(names, opt.views))
    n = len(stack.shapes)
    stack.wgt = [views["W%d" % i] for i in range(n)]
    stack.bias = [views["b%d" % i] for i in range(n)]
    stack.ln_gamma = [views["ln_mid%d/gamma" % (i + 1)] for i in range(n)]
    stack.ln_beta = [views["ln_mid%d/beta" % (i + 1)] for i in range(n)]
    stack.lno_gamma, stack.lno_beta = views["ln_output/gamma"], views["ln_output/beta"]
    stack.mark_weights_changed()
    loss, grads, _ = stack.ctc_train_step_grads(emb, labels, in_len, lab_len)
    flat = torch.cat([(grads[k] / B).reshape(-1) for k in names])
    opt.step(flat, 0.01)
    losses.append(loss.item())
  assert losses[-1] < 0.8 * losses[0], losses


@pytest.mark.parametrize("mode", ["fp32", "f16", "bf16"])
def test_stack_backward_one_call_equals_layer_loop_and_trainstep_buffers(mode):
  """srf_route_stack_bwd (one C call) against the same backward walked layer by layer into the views
  of a flat, layer-ordered gradient buffer (what TrainStep does to launch a layer's all-reduce while
  the layers below are still in their backward): bit-identical; and a TrainStep on top of it equals
  Adam applied to those gradients."""
  from srf_b200 import RoutingStack, training
  L, PH, CH, cls, DIM, B, S = 3, 12, 8, 9, 8, 3, 10
  stack = RoutingStack(L, PH, CH, cls, DIM, DIM, DIM, 1, 1, 1, True, seed=2, inn_dropout=0.1, uhat_mode=mode)
  g = torch.Generator().manual_seed(5)
  emb = torch.randn(B, S, PH, DIM, generator=g).cuda()
  labels = torch.randint(1, cls - 1, (B, 3), generator=g).cuda()
  in_len, lab_len = torch.tensor([S, S - 2, S - 1]).cuda(), torch.full((B,), 3).cuda()
  masks = stack.make_dropout_masks(B, S, generator=None)
  loss_a, grads_a, d_emb_a = stack.ctc_train_step_grads(emb, labels, in_len, lab_len, dropout_masks=masks,
                                                        grad_scale=0.5)
  order = []
  views = {n: torch.zeros_like(t).reshape(-1) for n, t in stack.named_parameters()}
  loss_b, grads_b, d_emb_b = stack.ctc_train_step_grads(emb, labels, in_len, lab_len, dropout_masks=masks,
                                                        grad_scale=0.5, out=views, after_layer=order.append)
  torch.cuda.synchronize()
  assert order == [2, 1, 0]
  assert torch.equal(loss_a, loss_b) and torch.equal(d_emb_a, d_emb_b)
  assert set(grads_a) == set(views)
  for n in views:
    assert grads_b[n].data_ptr() == views[n].data_ptr()                 # written in place
    if n[0] in "Wb":      # dW / dbias are atomics-free; the LayerNorm sums leave with one atomic per CTA
      assert torch.equal(grads_a[n].reshape(-1), views[n]), n
    else:
      assert torch.allclose(grads_a[n].reshape(-1), views[n], rtol=1e-5, atol=1e-6), n
    assert views[n].abs().sum().item() > 0, n
  # TrainStep: flat buffers ordered layer by layer; one step = Adam on the gradient of loss / global_batch
  p0 = {n: t.detach().clone() for n, t in stack.named_parameters()}
  tr = training.TrainStep(stack, global_batch=2, warmup_steps=4.0)
  assert tr.layer_slices[0][0] == 0 and tr.layer_slices[-1][1] == tr.opt.flat.numel()
  assert tr.names[:4] == ["W0", "b0", "ln_mid1/gamma", "ln_mid1/beta"] and tr.names[-1] == "ln_output/beta"
  tr.step(emb, labels, in_len, lab_len, dropout_masks=masks)             # first learning rate is 0 (Keras)
  for n, t in stack.named_parameters():
    assert torch.equal(t, p0[n]), n
  tr.step(emb, labels, in_len, lab_len, dropout_masks=masks)
  torch.cuda.synchronize()
  lr = training.warmup_lr(1, 0.5, 256.0, 4.0)
  for n, t in stack.named_parameters():
    gr = grads_a[n].reshape(t.shape)        # gradient of 0.5 * loss = loss / global_batch, unchanged weights
    m, v = 0.1 * gr, 0.02 * gr * gr
    m = 0.9 * m + 0.1 * gr
    v = 0.98 * v + 0.02 * gr * gr
    want = p0[n] - lr * math.sqrt(1 - 0.98 ** 2) / (1 - 0.9 ** 2) * m / (v.sqrt() + 1e-9)
    assert torch.allclose(t, want, rtol=2e-4, atol=1e-6), n
