"""TensorFlow-free reader / writer of the reference's checkpoint container (tensor bundle), SURVEY.md
8f next-4: format primitives against published known answers, the committed fixture
tests/golden/tf_ckpt (tests/golden/make_tf_bundle.py), and the checkpoint module on top of it."""
import os
import struct

import numpy as np
import pytest
import torch

from srf_b200 import checkpoint as ck
from srf_b200 import tf_bundle as tb

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_ckpt")


def test_crc32c_known_answers():
  # RFC 3720 appendix B.4 test vectors of CRC32C (Castagnoli)
  assert tb.crc32c(b"123456789") == 0xE3069283
  assert tb.crc32c(b"\x00" * 32) == 0x8A9136AA
  assert tb.crc32c(b"\xff" * 32) == 0x62A8AB43
  assert tb.crc32c(bytes(range(32))) == 0x46DD794E
  # incremental = one shot; LevelDB's mask is a rotation + constant
  assert tb.crc32c(b"6789", tb.crc32c(b"12345")) == 0xE3069283
  assert tb.mask_crc(0) == 0xa282ead8
  assert tb.mask_crc(tb.crc32c(b"foo")) != tb.crc32c(b"foo")


def test_varint_and_proto_helpers():
  for v in (0, 1, 127, 128, 300, 2 ** 32 + 5, 2 ** 63 - 1):
    enc = tb._put_varint(v)
    assert tb._get_varint(enc, 0) == (v, len(enc))
  assert tb._put_varint(300) == b"\xac\x02"
  shape = tb._shape_proto((1, 1, 36, 12, 8, 8))
  assert tb._parse_shape(shape) == (1, 1, 36, 12, 8, 8)
  assert tb._parse_shape(b"") == ()                # scalar
  with pytest.raises(ValueError):
    tb._get_varint(b"\x80", 0)


def test_table_round_trip_many_blocks(tmp_path):
  rng = np.random.default_rng(0)
  pairs = [(b"", b"header")]
  for i in range(300):
    key = ("model/layer_%03d/%s/.ATTRIBUTES/VARIABLE_VALUE" % (i // 3, ("kernel", "bias", "gamma")[i % 3])).encode()
    pairs.append((key, rng.bytes(int(rng.integers(0, 40)))))
  path = str(tmp_path / "t.index")
  tb.write_table(path, pairs, block_size=256)        # dozens of data blocks, prefix-compressed keys
  got = tb.read_table(path)
  assert got == sorted(pairs)
  raw = bytearray(open(path, "rb").read())
  assert struct.unpack_from("<Q", raw, len(raw) - 8)[0] == tb.TABLE_MAGIC
  raw[10] ^= 0x40                                     # flip a bit inside the first data block
  open(path, "wb").write(bytes(raw))
  with pytest.raises(ValueError, match="checksum"):
    tb.read_table(path)
  open(path, "wb").write(b"not a table at all, but longer than the forty-eight byte footer ....")
  with pytest.raises(ValueError, match="magic"):
    tb.read_table(path)


def test_bundle_round_trip_dtypes_and_corruption(tmp_path):
  rng = np.random.default_rng(1)
  tensors = {"a/.ATTRIBUTES/VARIABLE_VALUE": rng.standard_normal((3, 4, 5)).astype(np.float32),
             "b/.ATTRIBUTES/VARIABLE_VALUE": np.asarray(12345678901, dtype=np.int64),
             "c/.ATTRIBUTES/VARIABLE_VALUE": rng.standard_normal(7),
             "d/.ATTRIBUTES/VARIABLE_VALUE": np.zeros((0, 3), dtype=np.float32)}
  prefix = str(tmp_path / "ckpt-1")
  tb.write_bundle(prefix, tensors, object_graph=b"\x0a\x03abc")
  r = tb.BundleReader(prefix)
  assert set(r.keys()) == set(tensors) | {tb.OBJECT_GRAPH_KEY}
  assert r.entry(tb.OBJECT_GRAPH_KEY).dtype == tb.DT_STRING
  v = r.variables()
  assert set(v) == {"a", "b", "c", "d"}
  for k in "abcd":
    want = tensors[k + tb.VAR_SUFFIX]
    assert v[k].dtype == want.dtype and v[k].shape == want.shape and np.array_equal(v[k], want)
  with pytest.raises(ValueError):
    r.tensor(tb.OBJECT_GRAPH_KEY)
  data = prefix + ".data-00000-of-00001"
  raw = bytearray(open(data, "rb").read())
  raw[r.entry("a" + tb.VAR_SUFFIX).offset + 5] ^= 1
  open(data, "wb").write(bytes(raw))
  with pytest.raises(ValueError, match="checksum"):
    tb.BundleReader(prefix).tensor("a" + tb.VAR_SUFFIX)
  open(data, "wb").write(bytes(raw[:20]))
  with pytest.raises(ValueError, match="truncated"):
    tb.BundleReader(prefix).tensor("c" + tb.VAR_SUFFIX)


def test_key_map_covers_every_model_variable_and_inverts():
  names = ["W0", "b3", "ln_mid1/gamma", "ln_mid7/beta", "ln_output/gamma", "frontend/ln_input_beta",
           "frontend/cnn1_0_kernel", "frontend/cnn0_1_bias", "frontend/bn1_mean", "frontend/bn0_var",
           "frontend/bn0_gamma", "frontend/dense_kernel", "frontend/encaps1_bias"]
  for n in names:
    path = tb.inverse_key_map(n)
    assert path is not None and path.startswith("model/"), n
    assert tb.reference_key_map(path) == n
  assert tb.inverse_key_map("ln_mid1/gamma") == "model/ln_m/0/gamma"           # self.ln_m[0], naive:105
  assert tb.inverse_key_map("frontend/cnn1_0_kernel") == "model/conv/conv_layers/1/0/kernel"
  assert tb.inverse_key_map("frontend/bn1_var") == "model/conv/bn_layers/1/moving_variance"
  assert tb.reference_key_map("optimizer/iter") is None
  assert tb.reference_key_map("model/wgt/0/.OPTIMIZER_SLOT/optimizer/m") is None


def test_committed_fixture_reads_back_the_golden_parameters():
  """tests/golden/tf_ckpt was written by make_tf_bundle.py from sdr_i1_w3.npz in the naive variant's
  variable shapes; reading it must give those parameters back under srf_b200's names."""
  z = np.load(os.path.join(os.path.dirname(FIXTURE), "sdr_i1_w3.npz"))
  assert tb.latest_checkpoint(FIXTURE) == os.path.join(FIXTURE, "ckpt-4")
  assert [e for e, _ in ck.list_checkpoints(FIXTURE)] == [3, 4]
  state, other = tb.read_reference_checkpoint(os.path.join(FIXTURE, "ckpt-3"))
  L = sum(1 for k in z.files if k.startswith("W"))
  assert state["W0"].ndim == 6 and state["b0"].shape[-1] == 1                  # naive:88-103
  for i in range(L):
    W, b = ck.to_canonical(state["W%d" % i], state["b%d" % i])
    assert np.array_equal(W, z["W%d" % i].astype(np.float32))
    assert np.array_equal(b, z["b%d" % i].astype(np.float32))
    assert np.array_equal(state["ln_mid%d/gamma" % (i + 1)], z["ln_mid%d_gamma" % i].astype(np.float32))
  for k in z.files:
    if k.startswith("fe_"):
      assert np.array_equal(state["frontend/" + k[3:]], z[k].astype(np.float32)), k
  assert int(other["optimizer/iter"]) == 300 and int(other["save_counter"]) == 3
  assert "model/wgt/%d/.OPTIMIZER_SLOT/optimizer/m" % (L - 1) in other


class FakeModel:
  def __init__(self, shapes):
    self.p = {k: torch.zeros(s) for k, s in shapes.items()}
    self.changed = 0

  def named_parameters(self):
    return list(self.p.items())

  def mark_weights_changed(self):
    self.changed += 1


def _canonical_shapes(state):
  state = ck.canonicalize_state(state)
  return {k: tuple(v.shape) for k, v in state.items()}


def test_checkpoint_module_loads_and_averages_tf_bundles(tmp_path):
  state3, _ = tb.read_reference_checkpoint(os.path.join(FIXTURE, "ckpt-3"))
  model = FakeModel(_canonical_shapes(state3))
  assert ck.load_checkpoint(model, FIXTURE, strict=True) == 4                  # latest
  assert ck.load_checkpoint(model, FIXTURE, path_ckpt_epoch=3, strict=True) == 3
  W0, _ = ck.to_canonical(state3["W0"], state3["b0"])
  assert np.array_equal(model.p["W0"].numpy(), W0)
  with pytest.raises(FileNotFoundError):
    ck.load_checkpoint(model, FIXTURE, path_ckpt_epoch=9)
  # averaging the two epochs (W scaled by 1.0 and 1.5 -> 1.25), written back as a TF bundle the
  # reference could restore (average_ckpt_sr.py:137-179 writes <path_ckpt>/avg)
  work = tmp_path / "ck"
  import shutil
  shutil.copytree(FIXTURE, work)
  out = ck.average_checkpoints(str(work), 2, fmt="tf")
  assert out == os.path.join(str(work), "avg", "ckpt-1") and os.path.exists(out + ".index")
  avg, _ = tb.read_reference_checkpoint(out)
  Wa, _ = ck.to_canonical(avg["W0"], avg["b0"])
  assert np.allclose(Wa, 1.25 * W0, rtol=1e-6)
  assert avg["W0"].ndim == 6                                                   # naive layout on disk
  assert np.array_equal(avg["frontend/dense_kernel"], state3["frontend/dense_kernel"])
  # CheckpointManager pruning across containers
  ck.save_checkpoint(ck.canonicalize_state(state3), str(work), 5, max_to_keep=2, fmt="tf")
  assert [e for e, _ in ck.list_checkpoints(str(work))] == [4, 5]
  assert not os.path.exists(os.path.join(str(work), "ckpt-3.index"))
  assert tb.latest_checkpoint(str(work)) == os.path.join(str(work), "ckpt-5")
