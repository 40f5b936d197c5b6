"""CPU tests of the oracle itself: identities the reference's formulation implies
 and the algebra the CUDA kernels rely on."""
import torch

from oracle import srf_oracle as o

torch.manual_seed(1234)
F64 = torch.float64


def _layer(I_h=5, O=6, D=4, d=3, window=3, seed=0, B=2, S=7):
  g = torch.Generator().manual_seed(seed)
  emb = torch.randn(B, S, I_h, d, generator=g, dtype=F64)
  W = torch.randn(window * I_h, O, D, d, generator=g, dtype=F64) * 0.3
  bias = torch.randn(window * I_h, O, D, generator=g, dtype=F64) * 0.3
  return emb, W, bias


def test_window_gather_index_map():
  emb, _, _ = _layer()
  B, S, H, d = emb.shape
  lpad, rpad = 2, 1
  x = o.window_gather(emb, lpad, rpad)
  assert x.shape == (B, S, 4 * H, d)
  for s in range(S):
    for w in range(4):
      src = s - lpad + w
      blk = x[:, s, w * H:(w + 1) * H]
      if 0 <= src < S:
        assert torch.equal(blk, emb[:, src])
      else:
        assert torch.count_nonzero(blk) == 0


def test_zero_padded_frames_still_carry_bias():
  emb, W, bias = _layer()
  x = o.window_gather(emb, 1, 1)
  u = o.prediction_vectors(x, W, bias)
  H = emb.shape[2]
  assert torch.allclose(u[:, 0, :H], bias[:H].expand_as(u[:, 0, :H]))


def test_class0_is_exactly_zero_on_last_layer():
  emb, W, bias = _layer()
  for sdr in (True, False):
    for it in (1, 3):
      v = o.route_layer(emb.float(), W.float(), bias.float(), 1, 1, it, sdr, True)
      assert torch.count_nonzero(v[:, :, 0]) == 0
      assert torch.count_nonzero(v[:, :, 1:]) > 0


def test_mask_equals_dropping_class0_from_softmax():
  """-1e9 on capsule 0 == softmax over j >= 1 (what the kernel does)."""
  emb, W, bias = _layer()
  x = o.window_gather(emb, 1, 1)
  u = o.prediction_vectors(x, W, bias)
  v_ref = o.route_dr(u, 3, True)
  v_sub = o.route_dr(u[:, :, :, 1:], 3, False)
  assert torch.allclose(v_ref[:, :, 1:], v_sub, atol=1e-13)


def test_sdr_single_frame_equals_dr_iter1():
  emb, W, bias = _layer(S=1, window=1)
  a = o.route_layer(emb, W, bias, 0, 0, 1, True, False)
  b = o.route_layer(emb, W, bias, 0, 0, 1, False, False)
  assert torch.allclose(a, b, atol=1e-14)


def test_frame_at_a_time_equals_naive():
  emb, W, bias = _layer()
  for it in (1, 2):
    for last in (False, True):
      a = o.route_layer(emb, W, bias, 1, 1, it, True, last)
      b = o.route_layer_frame_sdr(emb, W, bias, 1, 1, it, last)
      assert torch.allclose(a, b, atol=1e-13)


def test_batch_translation_commutes():
  emb, W, bias = _layer(B=3)
  for sdr in (True, False):
    a = o.route_layer(emb, W, bias, 1, 1, 2, sdr, False)
    b = o.route_layer(emb[[2, 0, 1]], W, bias, 1, 1, 2, sdr, False)
    assert torch.allclose(a[[2, 0, 1]], b, atol=1e-14)


def test_logits_are_linear_in_accumulated_outputs():
  """b_r[i,j] = u[i,j,:] . (v_0 + ... + v_{r-1})[j,:]  -- the kernels keep only the sum of
  squashed outputs instead of the logits (naive:205 / :223 / :240 are linear in v)."""
  emb, W, bias = _layer()
  x = o.window_gather(emb, 1, 1)
  u = o.prediction_vectors(x, W, bias)           # [B,S,I,O,D]
  for sdr in (True, False):
    for iters in (1, 2, 4):
      ref = (o.route_sdr if sdr else o.route_dr)(u, iters, False)
      B, S, I, O, D = u.shape
      out = torch.zeros(B, S, O, D, dtype=F64)
      vprev = torch.zeros(B, O, D, dtype=F64)
      for s in range(S):
        vacc = vprev.clone() if sdr else torch.zeros(B, O, D, dtype=F64)
        for _ in range(iters):
          a = torch.einsum('biok,bok->bio', u[:, s], vacc)
          c = torch.softmax(a, dim=2)
          t = torch.einsum('bio,biok->bok', c, u[:, s])
          v = o.squash(t)
          vacc = vacc + v
        vprev = v
        out[:, s] = v
      assert torch.allclose(out, ref, atol=1e-12), (sdr, iters)


def test_layer_norm_matches_torch():
  x = torch.randn(4, 5, 24, dtype=F64)
  g, b = torch.randn(24, dtype=F64), torch.randn(24, dtype=F64)
  ref = torch.nn.functional.layer_norm(x, (24,), g, b, eps=1e-3)
  assert torch.allclose(o.layer_norm(x, g, b), ref, atol=1e-13)


def test_stack_shapes_and_greedy_ctc():
  shapes = o.layer_shapes(3, 6, 5, 7, 4, 4, 4, 3)
  assert shapes == [(18, 5, 4, 4), (15, 5, 4, 4), (15, 7, 4, 4)]
  p = o.init_params(shapes, 7, dtype=F64)
  emb = torch.randn(2, 9, 6, 4, dtype=F64)
  logits, caps = o.route_stack(emb, p, 1, 1, 1, True, return_capsules=True)
  assert logits.shape == (2, 9, 7) and len(caps) == 3
  # class 0 sits at the floor of the LN input, it can never be the argmax
  assert (logits.argmax(-1) != 0).all()
  hyp = o.greedy_ctc(logits, [9, 5])
  assert len(hyp) == 2 and all(0 < t < 6 for h in hyp for t in h)
  lg = torch.full((1, 6, 4), -1.0)
  for s, t in enumerate([1, 1, 3, 2, 2, 1]):
    lg[0, s, t] = 1.0
  assert o.greedy_ctc(lg, [6]) == [[1, 2, 1]]
  assert o.greedy_ctc(lg, [2]) == [[1]]
