"""The oracle against the golden vectors produced by the reference's own source file
(tests/golden/make_golden.py: /root/reference/tfsr/model/sequence_router_naive.py executed over
the numpy TF emulation).  This is what pins the restatement."""
import pytest
import torch

from oracle import srf_oracle as o
from tests import golden_util as gu


def test_goldens_exist():
  assert len(gu.golden_names()) >= 6


@pytest.mark.parametrize("name", gu.golden_names())
def test_oracle_matches_reference_code(name):
  g = gu.load(name)
  k = g["knobs"]
  logits, caps = o.route_stack(g["emb"], g["params"], k["lpad"], k["rpad"], k["iters"], k["sdr"],
                               dropout_masks=g["masks"], return_capsules=True)
  for i, (c, r) in enumerate(zip(caps, g["caps"])):
    assert c.shape == r.shape
    assert (c - r).abs().max().item() < 1e-10, "layer %d" % i
  assert (logits - g["logits"]).abs().max().item() < 1e-10
  # class 0 of the last layer is exactly zero in the reference run too (before LN it is 0;
  # after ln_mid it is beta-shifted, so check through the logits' floor instead)
  assert (g["logits"].argmax(-1) != 0).all() or k["class_n"] <= 2


@pytest.mark.parametrize("name", gu.golden_names())
def test_oracle_fp32_within_tolerance_of_reference_fp32_run(name):
  g = gu.load(name)
  k = g["knobs"]
  p32 = g["params"].to(torch.float32)
  masks = None if g["masks"] is None else [m.float() for m in g["masks"]]
  logits = o.route_stack(g["emb"].float(), p32, k["lpad"], k["rpad"], k["iters"], k["sdr"],
                         dropout_masks=masks)
  ref = g["logits"]
  assert ((logits.double() - ref).abs().max() / ref.abs().max()).item() < 1e-4
  lens = [int(n) // 4 for n in g["input_lengths"]]
  assert o.greedy_ctc(logits, lens) == o.greedy_ctc(ref, lens)
  assert o.greedy_ctc(g["logits_f32"], lens) == o.greedy_ctc(ref, lens)


@pytest.mark.parametrize("name", gu.golden_names())
def test_oracle_frontend_matches_reference_code(name):
  """The front-end restatement (oracle.capsulate, naive:129-142) against the primary capsules the
  reference's own file produced from the same fbank input and front-end parameters."""
  g = gu.load(name)
  z = g["raw"]
  fe = {n[3:]: z[n] for n in z.files if n.startswith("fe_")}
  # the emulation runs the front-end's Dropout / BatchNormalization layers in inference mode even
  # for the training golden (only the routing stack's masks are injected there)
  emb, _ = o.capsulate(torch.from_numpy(z["feats"]).double(), z["input_lengths"], fe)
  assert emb.shape == g["emb"].shape
  assert (emb - g["emb"]).abs().max().item() < 1e-10
