"""Loader for tests/golden/*.npz (made by tests/golden/make_golden.py from the reference's
own source file)."""
import ast
import glob
import os

import numpy as np
import torch

from oracle import srf_oracle as o

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
  return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
  z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
  knobs = ast.literal_eval(str(z["knobs"]))
  L = knobs["L"]
  t = lambda k: torch.from_numpy(np.asarray(z[k], dtype=np.float64))
  params = o.StackParams(
      W=[t("W%d" % i) for i in range(L)], bias=[t("b%d" % i) for i in range(L)],
      ln_gamma=[t("ln_mid%d_gamma" % i) for i in range(L)],
      ln_beta=[t("ln_mid%d_beta" % i) for i in range(L)],
      lno_gamma=t("ln_output_gamma"), lno_beta=t("ln_output_beta"))
  masks = [t("dropout_mask%d" % i) for i in range(L)] if knobs.get("training") else None
  caps = [t("caps%d" % i) for i in range(L)]
  return dict(knobs=knobs, params=params, emb=t("emb"), caps=caps, logits=t("logits"),
              logits_f32=torch.from_numpy(z["logits_f32"]), masks=masks,
              input_lengths=z["input_lengths"], raw=z)
