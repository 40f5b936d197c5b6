"""Generator of tests/golden/tf_ckpt/: a checkpoint directory in TensorFlow's tensor-bundle
container with the reference's object-based variable names (what
`tf.train.CheckpointManager(tf.train.Checkpoint(optimizer=..., model=SequenceRouter)).save()` leaves
behind, tfsr/helper/misc_helper.py:140-147), written WITHOUT TensorFlow by srf_b200.tf_bundle.

    python tests/golden/make_tf_bundle.py

Content: the parameters of the golden case `sdr_i1_w3` (tests/golden/sdr_i1_w3.npz, produced by the
reference's own source file) in the *naive* variant's variable shapes -- W (1,1,I,O,D,d),
bias (1,1,I,O,D,1) -- as epoch 3, the same parameters scaled by 1.5 as epoch 4, plus Adam slots and
the optimizer's iteration counter the way Keras names them.  The directory is a few tens of KB.
TensorFlow itself never ran (it cannot be installed here): the container follows the published
format (LevelDB table + BundleEntryProto), see srf_b200/tf_bundle.py.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from srf_b200 import checkpoint, tf_bundle  # noqa: E402


def main():
  z = np.load(os.path.join(HERE, "sdr_i1_w3.npz"))
  out = os.path.join(HERE, "tf_ckpt")
  if os.path.exists(out):
    shutil.rmtree(out)
  os.makedirs(out)
  L = sum(1 for k in z.files if k.startswith("W"))
  for epoch, scale in ((3, 1.0), (4, 1.5)):
    state, extra = {}, {}
    for i in range(L):
      W, b = checkpoint.from_canonical(z["W%d" % i] * scale, z["b%d" % i] * scale, "naive")
      state["W%d" % i], state["b%d" % i] = W, b
      if i == L - 1:     # Adam slots the way Keras names them (one layer keeps the fixture small)
        extra["model/wgt/%d/.OPTIMIZER_SLOT/optimizer/m" % i] = np.zeros_like(W)
        extra["model/wgt/%d/.OPTIMIZER_SLOT/optimizer/v" % i] = np.full_like(W, 1e-3)
      state["ln_mid%d/gamma" % (i + 1)] = z["ln_mid%d_gamma" % i].astype(np.float32)
      state["ln_mid%d/beta" % (i + 1)] = z["ln_mid%d_beta" % i].astype(np.float32)
    state["ln_output/gamma"] = z["ln_output_gamma"].astype(np.float32)
    state["ln_output/beta"] = z["ln_output_beta"].astype(np.float32)
    for k in z.files:
      if k.startswith("fe_"):
        state["frontend/" + k[3:]] = z[k].astype(np.float32)
    extra["optimizer/iter"] = np.asarray(100 * epoch, dtype=np.int64)
    extra["optimizer/beta_1"] = np.asarray(0.9, dtype=np.float32)
    prefix = tf_bundle.write_reference_checkpoint(out, epoch, state, extra)
    print(prefix, sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out)) >> 10, "KB so far")


if __name__ == "__main__":
  main()
