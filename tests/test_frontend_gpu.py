"""Parity of the native capsulation front-end (srf_capsulate_fwd: fbank -> primary capsules,
SURVEY.md 8f next-1) with the oracle restatement of naive:129-142 / sequence_router.py:44-82 and
with the vectors of the reference's own source file (tests/golden).  fp32 tolerance 1e-4 relative."""
import types

import pytest
import torch

from oracle import srf_oracle as o
from tests import golden_util as gu

pytestmark = pytest.mark.gpu


def rel_err(a, ref):
  ref = ref.double()
  return ((a.double().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def make_fe(F, C, PH, PD, seed):
  g = torch.Generator().manual_seed(seed)
  r = lambda *s: torch.randn(*s, generator=g)
  Fq = -(-(-(-F // 2)) // 2)
  fe = {}
  for p in range(2):
    for st in range(2):
      cin = 1 if st == 0 else C
      fe["cnn%d_%d_kernel" % (p, st)] = r(3, 3, cin, C) * (2.0 / (9 * cin)) ** 0.5
      fe["cnn%d_%d_bias" % (p, st)] = r(C) * 0.1
    fe["encaps%d_kernel" % p] = r(3, 3, 1, PD) * 0.4
    fe["encaps%d_bias" % p] = r(PD) * 0.1
  for st in range(2):
    fe["bn%d_gamma" % st] = 1 + 0.2 * r(C)
    fe["bn%d_beta" % st] = 0.1 * r(C)
    fe["bn%d_mean" % st] = 0.3 * r(C)
    fe["bn%d_var" % st] = 0.5 + torch.rand(C, generator=g)
  fe["dense_kernel"] = r(Fq * C, PH) * (1.0 / (Fq * C)) ** 0.5
  fe["dense_bias"] = r(PH) * 0.1
  fe["ln_input_gamma"] = 1 + 0.2 * r(PH * PD)
  fe["ln_input_beta"] = 0.1 * r(PH * PD)
  return fe


def make_batch(B, T, F, seed):
  g = torch.Generator().manual_seed(seed)
  feats = torch.randn(B, T, F, generator=g)
  lens = torch.tensor([T] + [int(T * (0.3 + 0.7 * torch.rand(1, generator=g).item())) for _ in range(B - 1)],
                      dtype=torch.int32)
  for b in range(B):
    feats[b, lens[b]:] = 0
  return feats, lens


@pytest.mark.parametrize("name", gu.golden_names())
def test_native_frontend_matches_reference_vectors(name):
  from srf_b200 import routing
  g = gu.load(name)
  z = g["raw"]
  fe = {n[3:]: torch.from_numpy(z[n]).float().cuda() for n in z.files if n.startswith("fe_")}
  PH, PD = g["emb"].shape[2], g["emb"].shape[3]
  emb = routing.capsulate_fwd(torch.from_numpy(z["feats"]).float().cuda(), z["input_lengths"], fe,
                              fe["cnn0_0_kernel"].shape[-1], PH, PD)
  torch.cuda.synchronize()
  assert emb.shape == g["emb"].shape
  assert rel_err(emb, g["emb"]) < 1e-4


# B, T, F, C, PH, PD: odd sizes exercise both 'same' paddings (pad_before 0 and 1), channel counts
# that are not multiples of 4, several position tiles per thread, the default 123 x 64 geometry
CASES = [
    (2, 16, 17, 4, 12, 8),
    (3, 21, 20, 6, 10, 4),
    (2, 37, 123, 64, 60, 8),
    (1, 5, 9, 3, 7, 20),
    (2, 64, 40, 32, 30, 16),
    (4, 30, 33, 5, 128, 4),
]


@pytest.mark.parametrize("B,T,F,C,PH,PD", CASES)
@pytest.mark.parametrize("einsum_variant", [False, True])
def test_native_frontend_matches_oracle(B, T, F, C, PH, PD, einsum_variant):
  from srf_b200 import routing
  if einsum_variant and PH % 2:
    pytest.skip("positional encoding needs an even PH (model_helper.py:55 concatenates two halves)")
  fe = make_fe(F, C, PH, PD, seed=B * 100 + T)
  feats, lens = make_batch(B, T, F, seed=T)
  if B > 2:
    lens[-1] = 0          # an empty utterance: every frame masked
  ref, _ = o.capsulate(feats.double(), lens, {k: v.double() for k, v in fe.items()},
                       einsum_variant=einsum_variant)
  emb = routing.capsulate_fwd(feats.cuda(), lens, {k: v.cuda() for k, v in fe.items()}, C, PH, PD,
                              pos_enc=einsum_variant)
  torch.cuda.synchronize()
  assert emb.shape == ref.shape
  assert rel_err(emb, ref) < 1e-4
  # masked frames carry exactly ln_input's beta (squash(0) = 0, LayerNorm of a zero vector)
  S = ref.shape[1]
  n = int(-(-int(lens[-1]) // 4))
  if n < S:
    assert torch.allclose(emb[-1, n:].reshape(S - n, -1).cpu(),
                          fe["ln_input_beta"][None, :].expand(S - n, -1), atol=1e-6)


@pytest.mark.parametrize("B,T,F,C,PH,PD", CASES[:3])
def test_native_frontend_training_semantics(B, T, F, C, PH, PD):
  """training=True: injected keep masks where the reference has Dropout layers, BatchNormalization
  with batch statistics and the moving-average update; deterministic."""
  from srf_b200 import routing
  fe = make_fe(F, C, PH, PD, seed=7)
  feats, lens = make_batch(B, T, F, seed=3)
  T1, F1 = -(-T // 2), -(-F // 2)
  S, Fq = -(-T1 // 2), -(-F1 // 2)
  g = torch.Generator().manual_seed(11)
  keep = lambda rate, *s: (torch.rand(*s, generator=g) >= rate).float() / (1 - rate)
  drop = {"inp": keep(0.1, B, S, PH, PD)}
  for p in range(2):
    drop["cnn%d_0" % p] = keep(0.2, B, T1, F1, C)
    drop["cnn%d_1" % p] = keep(0.2, B, S, Fq, C)
    drop["encaps%d" % p] = keep(0.2, B, S, PH, PD)
  ref, moving = o.capsulate(feats.double(), lens, {k: v.double() for k, v in fe.items()}, training=True,
                            dropout={k: v.double() for k, v in drop.items()})
  outs = []
  for _ in range(2):
    dev_fe = {k: v.clone().cuda() for k, v in fe.items()}
    emb = routing.capsulate_fwd(feats.cuda(), lens, dev_fe, C, PH, PD, training=True,
                                dropout={k: v.cuda() for k, v in drop.items()})
    torch.cuda.synchronize()
    outs.append(emb)
  assert rel_err(outs[0], ref) < 1e-4
  assert torch.equal(outs[0], outs[1])
  for k, v in moving.items():
    assert rel_err(dev_fe[k], v) < 1e-5, k
  # without masks and with training=False the same descriptor is the inference function
  emb_inf = routing.capsulate_fwd(feats.cuda(), lens, {k: v.cuda() for k, v in fe.items()}, C, PH, PD)
  ref_inf, _ = o.capsulate(feats.double(), lens, {k: v.double() for k, v in fe.items()})
  assert rel_err(emb_inf, ref_inf) < 1e-4


def test_frontend_argument_errors():
  from srf_b200 import routing
  fe = {k: v.cuda() for k, v in make_fe(17, 4, 12, 8, seed=1).items()}
  feats, lens = make_batch(2, 16, 17, seed=1)
  with pytest.raises(ValueError):
    routing.capsulate_fwd(feats.cuda(), lens[:1], fe, 4, 12, 8)
  bad = dict(fe)
  del bad["dense_bias"]
  with pytest.raises(ValueError):
    routing.capsulate_fwd(feats.cuda(), lens, bad, 4, 12, 8)
  with pytest.raises(ValueError):     # positional encoding with an odd PH
    routing.capsulate_fwd(feats.cuda(), lens, {k: v.cuda() for k, v in make_fe(17, 4, 11, 8, 1).items()},
                          4, 11, 8, pos_enc=True)
  with pytest.raises(ValueError):     # PH beyond the dense kernel's column budget -> -3 from the library
    routing.capsulate_fwd(feats.cuda(), lens, {k: v.cuda() for k, v in make_fe(17, 4, 130, 4, 1).items()},
                          4, 130, 4)


@pytest.mark.parametrize("name", ["sdr_i1_w3", "dr_i3_w7"])
def test_dropin_runs_no_framework_compute_between_fbank_and_logits(name):
  """SequenceRouter.__call__(training=False) = srf_capsulate_fwd + srf_route_stack_fwd: between the
  device-resident fbank tensor and the logits the only torch ops are allocations / views, and every
  kernel is launched by the library (handle launch counter)."""
  from torch.utils._python_dispatch import TorchDispatchMode
  from srf_b200 import SequenceRouter
  g = gu.load(name)
  k, z = g["knobs"], g["raw"]
  cfg = types.SimpleNamespace(
      model_initializer="fan_avg", model_conv_layer_num=2, feat_dim=z["feats"].shape[-1],
      model_conv_filter_num=z["fe_cnn0_0_kernel"].shape[-1], model_encoder_num=k["L"],
      model_caps_iter=k["iters"], model_caps_window_lpad=k["lpad"], model_caps_window_rpad=k["rpad"],
      model_caps_context=k["sdr"], model_caps_primary_num=k["PH"], model_caps_primary_dim=k["DIM"],
      model_caps_convolution_num=k["CH"], model_caps_convolution_dim=k["DIM"],
      model_caps_class_dim=k["DIM"], train_inp_dropout=0.1, train_inn_dropout=0.1)
  model = SequenceRouter(cfg, None, k["class_n"])
  model.load_frontend({n[3:]: z[n] for n in z.files if n.startswith("fe_")})
  model.stack.load_oracle_params(g["params"])
  feats = torch.from_numpy(z["feats"]).float().cuda()
  lens = torch.from_numpy(z["input_lengths"]).int().cuda()
  model(feats, input_lengths=lens)        # warm-up: weight packing, workspaces
  torch.cuda.synchronize()

  seen = []

  class Recorder(TorchDispatchMode):
    def __torch_dispatch__(self, func, types_, args=(), kwargs=None):
      seen.append(str(func))
      return func(*args, **(kwargs or {}))

  before = model.stack.handle.launches
  with Recorder():
    logits = model(feats, input_lengths=lens)
  torch.cuda.synchronize()
  assert model.stack.handle.launches > before
  allowed = ("aten.empty", "aten.detach", "aten.view", "aten.alias", "aten._to_copy", "aten.lift_fresh",
             "aten.contiguous", "aten.reshape", "aten._unsafe_view")
  compute = [s for s in seen if not s.startswith(allowed)]
  assert not compute, compute
  assert rel_err(logits, g["logits"]) < 2e-4
  n = [int(x) // 4 for x in z["input_lengths"]]
  assert o.greedy_ctc(logits.cpu(), n) == o.greedy_ctc(g["logits"], n)
