#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the REFERENCE'S OWN SOURCE FILE
(/root/reference/tfsr/model/sequence_router_naive.py, unmodified) on top of the numpy TF
emulation in oracle/tf_shim.  Runs only in the build container (needs /root/reference);
the committed .npz files travel to the GPU box.

    python tests/golden/make_golden.py

Each fixture holds: the fbank input, input lengths, every parameter of the routing stack in
canonical layout, the primary capsules `emb` entering the hot path (naive:142), every layer's
output after LayerNorm+dropout (naive:191) and the final logits (naive:193), all produced by the
reference code in float64 (structural pin) -- plus the float32 logits of the same run.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, REF)

import tensorflow as tf  # noqa: E402  (the shim)

CASES = {
    # name: knobs.  Small dims so the fixtures stay a few hundred KB.
    "sdr_i1_w3": dict(L=3, PH=12, CH=6, DIM=8, lpad=1, rpad=1, sdr=True, iters=1, class_n=9, B=2, T=40),
    "sdr_i2_w5": dict(L=3, PH=10, CH=7, DIM=4, lpad=2, rpad=2, sdr=True, iters=2, class_n=11, B=3, T=28),
    "dr_i3_w7": dict(L=3, PH=12, CH=6, DIM=8, lpad=3, rpad=3, sdr=False, iters=3, class_n=9, B=2, T=36),
    "dr_i1_w1": dict(L=2, PH=8, CH=5, DIM=16, lpad=0, rpad=0, sdr=False, iters=1, class_n=6, B=2, T=24),
    "sdr_one_layer": dict(L=1, PH=9, CH=4, DIM=8, lpad=1, rpad=0, sdr=True, iters=3, class_n=7, B=2, T=20),
    "sdr_train_dropout": dict(L=3, PH=12, CH=6, DIM=8, lpad=1, rpad=1, sdr=True, iters=1, class_n=9, B=2, T=32,
                              training=True),
}
FEAT_DIM, NFILT = 20, 4


class _Log:
  def info(self, *a, **k):
    pass


def make_config(c):
  return types.SimpleNamespace(
      model_initializer="fan_avg", model_conv_layer_num=2, feat_dim=FEAT_DIM,
      model_conv_filter_num=NFILT, model_encoder_num=c["L"], model_caps_iter=c["iters"],
      model_caps_window_lpad=c["lpad"], model_caps_window_rpad=c["rpad"],
      model_caps_context=c["sdr"], model_caps_primary_num=c["PH"],
      model_caps_primary_dim=c["DIM"], model_caps_convolution_num=c["CH"],
      model_caps_convolution_dim=c["DIM"], model_caps_class_dim=c["DIM"],
      train_inp_dropout=0.1, train_inn_dropout=0.1)


def run_case(name, c, dtype, seed):
  import importlib
  tf.set_float(dtype)
  tf.set_seed(seed)
  import tfsr.model.sequence_router_naive as naive
  importlib.reload(naive)
  model = naive.SequenceRouter(make_config(c), _Log(), c["class_n"])
  rng = np.random.default_rng(seed + 1)
  B, T = c["B"], c["T"]
  feats = rng.standard_normal((B, T, FEAT_DIM)).astype(dtype)
  lens = np.array([T] + [int(T * f) for f in rng.uniform(0.6, 1.0, B - 1)], dtype=np.int32)
  for b in range(B):
    feats[b, lens[b]:] = 0.0
  training = bool(c.get("training", False))
  # non-trivial LayerNorm affine parameters + (training) injected dropout masks
  _ = model(feats, input_lengths=lens, training=False)      # builds the Keras layers
  for ln in model.ln_m + [model.ln_o, model.ln_i]:
    ln.gamma = (1.0 + 0.2 * rng.standard_normal(ln.gamma.shape)).astype(dtype)
    ln.beta = (0.1 * rng.standard_normal(ln.beta.shape)).astype(dtype)
  masks = []
  if training:
    S = -(-T // 4)
    for i, d in enumerate(model.mid_dropout):
      O, D = model.wgt[i].shape[3], model.wgt[i].shape[4]
      m = ((rng.uniform(size=(B, S, O, D)) < 0.9) / 0.9).astype(dtype)
      d.mask = m
      masks.append(m)
  logits = model(feats, input_lengths=lens, training=training)
  out = {"feats": feats, "input_lengths": lens, "logits": logits,
         "emb": np.asarray(model.inp_dropout.last_output)}
  for i in range(model.enc_num):
    out["W%d" % i] = model.wgt[i].value.reshape(model.wgt[i].shape[2:])
    out["b%d" % i] = model.bias[i].value.reshape(model.bias[i].shape[2:5])
    out["ln_mid%d_gamma" % i] = model.ln_m[i].gamma
    out["ln_mid%d_beta" % i] = model.ln_m[i].beta
    out["caps%d" % i] = np.asarray(model.mid_dropout[i].last_output)
    if training:
      out["dropout_mask%d" % i] = masks[i]
  out["ln_output_gamma"], out["ln_output_beta"] = model.ln_o.gamma, model.ln_o.beta
  # front-end parameters (next-1 row: capsulation), stored for later parity work
  fe = {"dense_kernel": model.proj_pe.kernel, "dense_bias": model.proj_pe.bias,
        "ln_input_gamma": model.ln_i.gamma, "ln_input_beta": model.ln_i.beta}
  for li, pair in enumerate(model.conv.conv_layers):
    for pi, conv in enumerate(pair):
      fe["cnn%d_%d_kernel" % (li, pi)], fe["cnn%d_%d_bias" % (li, pi)] = conv.kernel, conv.bias
  for li, bn in enumerate(model.conv.bn_layers):
    fe["bn%d_gamma" % li], fe["bn%d_beta" % li] = bn.gamma, bn.beta
    fe["bn%d_mean" % li], fe["bn%d_var" % li] = bn.moving_mean, bn.moving_variance
  for pi, conv in enumerate(model.ecs):
    fe["encaps%d_kernel" % pi], fe["encaps%d_bias" % pi] = conv.kernel, conv.bias
  out.update({"fe_" + k: v for k, v in fe.items()})
  return out


def main():
  for idx, (name, c) in enumerate(CASES.items()):
    g64 = run_case(name, c, np.float64, seed=100 + idx)
    g32 = run_case(name, c, np.float32, seed=100 + idx)
    g64["logits_f32"] = g32["logits"]
    knobs = {k: v for k, v in c.items()}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, knobs=np.array(repr(knobs)), **g64)
    print("%-20s logits %s  f32-vs-f64 max diff %.2e  %d KB" % (
        name, g64["logits"].shape, np.abs(g32["logits"] - g64["logits"]).max(), os.path.getsize(path) >> 10))


if __name__ == "__main__":
  main()
