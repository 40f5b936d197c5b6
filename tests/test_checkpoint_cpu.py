"""Weights interchange / checkpoints / averaging -- host logic, no GPU."""
import os

import numpy as np
import pytest
import torch

from srf_b200 import checkpoint as ck


class FakeStack:
  """Stands in for RoutingStack: canonical W%d / b%d + LayerNorm parameters as CPU tensors."""

  def __init__(self, seed):
    g = torch.Generator().manual_seed(seed)
    self.p = {"W0": torch.randn(6, 5, 4, 3, generator=g), "b0": torch.randn(6, 5, 4, generator=g),
              "W1": torch.randn(15, 7, 4, 4, generator=g), "b1": torch.randn(15, 7, 4, generator=g),
              "ln_mid1/gamma": torch.randn(20, generator=g), "ln_output/beta": torch.randn(7, generator=g)}
    self.changed = 0

  def named_parameters(self):
    return list(self.p.items())

  def mark_weights_changed(self):
    self.changed += 1


@pytest.mark.parametrize("variant", ck.VARIANTS)
def test_variant_layouts_are_reshapes_of_canonical(variant):
  rng = np.random.default_rng(0)
  W = rng.standard_normal((6, 5, 4, 3)).astype(np.float32)
  b = rng.standard_normal((6, 5, 4)).astype(np.float32)
  Wv, bv = ck.from_canonical(W, b, variant)
  # shapes of the reference's variables (naive:88-103, lowmemory:85-101, einsum:82-97)
  want = {"naive": ((1, 1, 6, 5, 4, 3), (1, 1, 6, 5, 4, 1)), "lowmemory": ((1, 6, 5, 4, 3), (1, 6, 5, 4, 1)),
          "einsum": ((6, 5, 4, 3), (1, 1, 6, 5, 4))}[variant]
  assert (Wv.shape, bv.shape) == want
  assert Wv.ravel().tolist() == W.ravel().tolist()      # no transposition
  W2, b2 = ck.to_canonical(Wv, bv)
  assert np.array_equal(W2, W) and np.array_equal(b2, b)


def test_bad_shapes_are_rejected():
  with pytest.raises(ValueError):
    ck.to_canonical(np.zeros((2, 3, 4)), np.zeros((2, 3)))
  with pytest.raises(ValueError):
    ck.to_canonical(np.zeros((2, 3, 4, 5)), np.zeros((2, 3, 5)))
  with pytest.raises(ValueError):
    ck.from_canonical(np.zeros((2, 3, 4, 5)), np.zeros((2, 3, 4)), "dense")


@pytest.mark.parametrize("variant", [None, "naive", "lowmemory", "einsum"])
def test_save_load_round_trip_in_every_variant_layout(tmp_path, variant):
  a, b = FakeStack(1), FakeStack(2)
  ck.save_checkpoint(ck.state_dict(a), str(tmp_path), 3, variant=variant)
  assert ck.load_checkpoint(b, str(tmp_path)) == 3
  for k in a.p:
    assert torch.equal(a.p[k], b.p[k]), k
  assert b.changed == 1


def test_load_checkpoint_follows_the_reference_rules(tmp_path):
  m = FakeStack(0)
  assert ck.load_checkpoint(m, str(tmp_path)) == 0          # nothing there: epoch offset 0
  states = []
  for e in (1, 2, 5):
    s = ck.state_dict(FakeStack(10 + e))
    states.append(s)
    ck.save_checkpoint(s, str(tmp_path), e)
  assert ck.load_checkpoint(m, str(tmp_path)) == 5          # latest
  assert np.array_equal(m.p["W1"].numpy(), states[2]["W1"])
  assert ck.load_checkpoint(m, str(tmp_path), path_ckpt_epoch=2) == 2
  assert np.array_equal(m.p["W1"].numpy(), states[1]["W1"])
  with pytest.raises(FileNotFoundError):
    ck.load_checkpoint(m, str(tmp_path), path_ckpt_epoch=4)
  # strict load of a partial checkpoint fails, the reference's expect_partial() mode passes
  part = {k: v for k, v in states[0].items() if not k.startswith("ln_")}
  ck.save_checkpoint(part, str(tmp_path), 9)
  assert ck.load_checkpoint(m, str(tmp_path)) == 9
  with pytest.raises(KeyError):
    ck.load_checkpoint(m, str(tmp_path), strict=True)


def test_average_mixes_variant_layouts(tmp_path):
  a, b = ck.state_dict(FakeStack(1)), ck.state_dict(FakeStack(2))
  ck.save_checkpoint(a, str(tmp_path), 1, variant="naive")
  ck.save_checkpoint(b, str(tmp_path), 2, variant="einsum")
  avg = ck.read_checkpoint(ck.average_checkpoints(str(tmp_path), 2))
  assert avg["W0"].shape == (6, 5, 4, 3) and avg["b0"].shape == (6, 5, 4)
  assert np.allclose(avg["W1"], (a["W1"] + b["W1"]) / 2, atol=1e-7)
  assert np.allclose(avg["b1"], (a["b1"] + b["b1"]) / 2, atol=1e-7)


def test_max_to_keep_prunes_old_checkpoints(tmp_path):
  s = ck.state_dict(FakeStack(0))
  for e in range(1, 6):
    ck.save_checkpoint(s, str(tmp_path), e, max_to_keep=2)
  assert [e for e, _ in ck.list_checkpoints(str(tmp_path))] == [4, 5]
  for e in range(6, 9):
    ck.save_checkpoint(s, str(tmp_path), e, max_to_keep=-1)   # < 0 keeps everything
  assert [e for e, _ in ck.list_checkpoints(str(tmp_path))] == [4, 5, 6, 7, 8]


def test_average_checkpoints_is_the_mean_of_the_last_n(tmp_path):
  states = [ck.state_dict(FakeStack(s)) for s in range(4)]
  for e, s in enumerate(states, start=1):
    ck.save_checkpoint(s, str(tmp_path), e)
  out = ck.average_checkpoints(str(tmp_path), 3)
  assert out == os.path.join(str(tmp_path), "avg", "ckpt-1.npz")
  avg = ck.read_checkpoint(out)
  for k in states[0]:
    want = np.mean([s[k].astype(np.float64) for s in states[1:]], axis=0)
    assert np.allclose(avg[k], want, rtol=0, atol=1e-6), k
  # a second run replaces the avg directory
  ck.average_checkpoints(str(tmp_path), 2)
  assert len(ck.list_checkpoints(os.path.join(str(tmp_path), "avg"))) == 1
  m = FakeStack(9)
  ck.load_state_dict(m, ck.read_checkpoint(out))
  want2 = np.mean([s["b0"].astype(np.float64) for s in states[2:]], axis=0)
  assert np.allclose(m.p["b0"].numpy(), want2, rtol=0, atol=1e-6)


class FakeRouter:
  """Stands in for SequenceRouter: front-end parameters appear only after load_frontend / the
  first call, like the lazily built Keras layers of the reference."""

  def __init__(self, seed):
    self.stack = FakeStack(seed)
    self.fe = {}

  def load_frontend(self, params):
    self.fe = {k: torch.as_tensor(np.asarray(v), dtype=torch.float32) for k, v in params.items()}

  def named_parameters(self):
    return [("frontend/" + k, v) for k, v in sorted(self.fe.items())] + self.stack.named_parameters()


def test_fresh_model_restores_its_frontend_and_strict_rejects_unknown_entries():
  src = FakeRouter(3)
  src.load_frontend({"dense_kernel": np.arange(6, dtype=np.float32).reshape(2, 3), "dense_bias": np.ones(3)})
  state = ck.state_dict(src)
  dst = FakeRouter(4)           # fresh: no front-end parameters yet
  loaded = ck.load_state_dict(dst, state)
  assert "frontend/dense_kernel" in loaded and "frontend/dense_bias" in loaded
  for (n1, t1), (n2, t2) in zip(src.named_parameters(), dst.named_parameters()):
    assert n1 == n2 and torch.equal(t1, t2)
  state["frontend/not_a_parameter"] = np.zeros(3)
  with pytest.raises(KeyError):
    ck.load_state_dict(FakeRouter(5), {k: v for k, v in state.items() if k != "frontend/dense_bias"} |
                       {"stray": np.zeros(2)})
  # the reference's expect_partial(): extra entries are fine when not strict
  ck.load_state_dict(FakeRouter(6), state | {"stray": np.zeros(2)}, strict=False)


def test_warmup_schedule_starts_at_zero_like_keras():
  from srf_b200 import training
  assert training.warmup_lr(0, k=0.5, d_model=256, warmup_steps=4) == 0.0
  assert training.warmup_lr(1, k=0.5, d_model=256, warmup_steps=4) == 0.5 * 256 ** -0.5 * 4 ** -1.5
