"""CPU oracle for the SRF capsule-routing hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, on the CPU with torch tensors (float64 = truth, float32 =
tolerance calibration, autograd = gradient oracle), the arithmetic of the
reference's routing stack

    /root/reference/tfsr/model/sequence_router_naive.py:145-258

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product path (``srf_b200``) never
does: it fails loudly if the CUDA library is missing.

PARITY STATUS: the reference's own implementation needs TensorFlow >= 2.3
(requirements.txt:1), which cannot be installed in the build container, and the
reference ships no tests / golden vectors for this path.
The restatement is pinned instead by executing the reference's *own source
file* on top of a numpy emulation of the TF ops it calls (``oracle/tf_shim``,
``tests/golden/make_golden.py``); the resulting vectors are committed under
``tests/golden/``.  Because TF itself never ran, this is still declared
"parity unpinned" against real TensorFlow kernels.

Layouts (canonical = the reference's einsum layout, sequence_router_einsum.py:82-97):
    emb   [B, S, H, d]      input capsules of a layer
    W     [I, O, D, d]      I = window * H, i = w * H + h   (naive:88-95, 150-151)
    bias  [I, O, D]         (naive:99-103)
    v     [B, S, O, D]      output capsules
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

SQUASH_EPS = 1e-7     # naive:248
LENGTH_EPS = 1e-7     # naive:256   (the einsum variant uses 1e-9, einsum:238)
MASK_VALUE = -1e9     # naive:174, 219
LN_EPS = 1e-3         # Keras LayerNormalization default epsilon (naive:104-107)


def squash(s: torch.Tensor, axis: int = -1, epsilon: float = SQUASH_EPS) -> torch.Tensor:
  """naive:248-253 / sequence_router.py:29-35."""
  squared_norm = torch.sum(s * s, dim=axis, keepdim=True)
  safe_norm = torch.sqrt(squared_norm + epsilon)
  squash_factor = squared_norm / (1.0 + squared_norm)
  unit_vector = s / safe_norm
  return squash_factor * unit_vector


def length(s: torch.Tensor, axis: int = -1, epsilon: float = LENGTH_EPS) -> torch.Tensor:
  """naive:256-258 / sequence_router.py:38-41."""
  return torch.sqrt(torch.sum(s * s, dim=axis) + epsilon)


def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
               eps: float = LN_EPS) -> torch.Tensor:
  """Keras LayerNormalization(axis=-1): biased variance, eps inside the sqrt
  (naive:104-107, applied at naive:188-191 and naive:193)."""
  mean = x.mean(dim=-1, keepdim=True)
  var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
  return (x - mean) / torch.sqrt(var + eps) * gamma + beta


def window_gather(emb: torch.Tensor, lpad: int, rpad: int) -> torch.Tensor:
  """naive:150-151: ZeroPadding2D((lpad, rpad),(0,0)) then concat of `window`
  shifted slices along the capsule axis: x[b,s,w*H+h,:] = emb[b,s-lpad+w,h,:]."""
  B, S, H, d = emb.shape
  window = lpad + rpad + 1
  pad = torch.zeros(B, S + lpad + rpad, H, d, dtype=emb.dtype)
  pad[:, lpad:lpad + S] = emb
  return torch.cat([pad[:, w:w + S] for w in range(window)], dim=2)


def prediction_vectors(x: torch.Tensor, W: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
  """naive:154-159: u_hat[b,s,i,j,k] = sum_l W[i,j,k,l] x[b,s,i,l] + bias[i,j,k]."""
  return torch.einsum('ijkl,bsil->bsijk', W, x) + bias


def route_dr(u_hat: torch.Tensor, iters: int, is_last: bool) -> torch.Tensor:
  """Classic dynamic routing, all frames at once.  naive:171-185 + _loop_body 200-206.

  u_hat [B,S,I,O,D] -> v [B,S,O,D].  The logits b persist over iterations and the
  -1e9 mask on output capsule 0 (last layer only) is re-added every iteration."""
  B, S, I, O, D = u_hat.shape
  b = torch.zeros(B, S, I, O, 1, dtype=u_hat.dtype)
  masking = torch.zeros(B, S, I, O, 1, dtype=u_hat.dtype)
  if is_last:
    masking[:, :, :, 0] = MASK_VALUE
  v = torch.zeros(B, S, 1, O, D, dtype=u_hat.dtype)
  for _ in range(iters):
    b = b + masking
    c = torch.softmax(b, dim=3)
    s = torch.sum(c * u_hat, dim=2, keepdim=True)
    v = squash(s, axis=-1)
    b = b + torch.sum(u_hat * v, dim=-1, keepdim=True)
  return v[:, :, 0]


def route_sdr(u_hat: torch.Tensor, iters: int, is_last: bool) -> torch.Tensor:
  """Sequential dynamic routing.  naive:162-170 + body_context 232-245 /
  pad_body_context 213-229.

  Per frame the logits restart at 0; each iteration first adds the agreement with
  the carried v (the previous frame's output at iteration 1), then (last layer)
  the -1e9 mask, then softmax over output capsules, weighted sum, squash."""
  B, S, I, O, D = u_hat.shape
  v = torch.zeros(B, 1, O, D, dtype=u_hat.dtype)
  masking = torch.zeros(B, I, O, 1, dtype=u_hat.dtype)
  if is_last:
    masking[:, :, 0] = MASK_VALUE
  outs = []
  for idx in range(S):
    u = u_hat[:, idx]
    b = torch.zeros(B, I, O, 1, dtype=u_hat.dtype)
    for _ in range(iters):
      b = b + torch.sum(u * v, dim=-1, keepdim=True)
      if is_last:
        b = b + masking
      c = torch.softmax(b, dim=2)
      s = torch.sum(c * u, dim=1, keepdim=True)
      v = squash(s, axis=-1)
    outs.append(v[:, 0])
  return torch.stack(outs, dim=1)


def route_layer(emb: torch.Tensor, W: torch.Tensor, bias: torch.Tensor, lpad: int,
                rpad: int, iters: int, sdr: bool, is_last: bool) -> torch.Tensor:
  """One routing layer without the trailing LayerNorm: [B,S,H,d] -> [B,S,O,D]."""
  x = window_gather(emb, lpad, rpad)
  u_hat = prediction_vectors(x, W, bias)
  return route_sdr(u_hat, iters, is_last) if sdr else route_dr(u_hat, iters, is_last)


def route_layer_frame_sdr(emb: torch.Tensor, W: torch.Tensor, bias: torch.Tensor,
                          lpad: int, rpad: int, iters: int, is_last: bool) -> torch.Tensor:
  """Memory-light SDR (u_hat built one frame at a time, the lowmemory variant's
  structure, sequence_router_lowmemory.py:221-254, but with general ITER).  Same
  arithmetic as route_layer(sdr=True); used for the big CPU-baseline shapes."""
  B, S, H, d = emb.shape
  I, O, D, _ = W.shape
  window = lpad + rpad + 1
  pad = torch.zeros(B, S + lpad + rpad, H, d, dtype=emb.dtype)
  pad[:, lpad:lpad + S] = emb
  Wv = W.reshape(window, H, O * D, d)
  v = torch.zeros(B, 1, O, D, dtype=emb.dtype)
  outs = []
  for idx in range(S):
    xs = pad[:, idx:idx + window]                      # [B, window, H, d]
    u = torch.einsum('whnl,bwhl->bwhn', Wv, xs).reshape(B, I, O, D) + bias
    b = torch.zeros(B, I, O, 1, dtype=emb.dtype)
    for _ in range(iters):
      b = b + torch.sum(u * v, dim=-1, keepdim=True)
      if is_last:
        b[:, :, 0] += MASK_VALUE
      c = torch.softmax(b, dim=2)
      s = torch.sum(c * u, dim=1, keepdim=True)
      v = squash(s, axis=-1)
    outs.append(v[:, 0])
  return torch.stack(outs, dim=1)


@dataclass
class StackParams:
  """Parameters of the routing stack in canonical layout (naive:88-114)."""
  W: List[torch.Tensor]           # per layer [I,O,D,d]
  bias: List[torch.Tensor]        # per layer [I,O,D]
  ln_gamma: List[torch.Tensor]    # per layer [O*D]      (ln_mid%d)
  ln_beta: List[torch.Tensor]
  lno_gamma: torch.Tensor         # [class_n]            (ln_output)
  lno_beta: torch.Tensor

  def to(self, dtype):
    return StackParams([w.to(dtype) for w in self.W], [b.to(dtype) for b in self.bias],
                       [g.to(dtype) for g in self.ln_gamma], [b.to(dtype) for b in self.ln_beta],
                       self.lno_gamma.to(dtype), self.lno_beta.to(dtype))


def layer_shapes(enc_num: int, ph: int, ch: int, class_n: int, pd: int, cd: int, vd: int,
                 window: int) -> List[Tuple[int, int, int, int]]:
  """(I, O, D, d) per layer, naive:86-95."""
  if enc_num == 1:
    return [(ph * window, class_n, vd, pd)]
  shapes = [(ph * window, ch, cd, pd)]
  for _ in range(1, enc_num - 1):
    shapes.append((ch * window, ch, cd, cd))
  shapes.append((ch * window, class_n, vd, cd))
  return shapes


def init_params(shapes: Sequence[Tuple[int, int, int, int]], class_n: int, seed: int = 0,
                dtype=torch.float32, random_ln: bool = False) -> StackParams:
  """W, bias ~ N(0, 0.1^2) (naive:97-103); LN gamma=1, beta=0 (Keras default) unless
  random_ln (to exercise the affine part in tests)."""
  g = torch.Generator().manual_seed(seed)
  W = [torch.randn(s, generator=g, dtype=torch.float64).mul(0.1).to(dtype) for s in shapes]
  bias = [torch.randn(s[:3], generator=g, dtype=torch.float64).mul(0.1).to(dtype) for s in shapes]
  if random_ln:
    gam = [(1.0 + 0.2 * torch.randn(s[1] * s[2], generator=g, dtype=torch.float64)).to(dtype) for s in shapes]
    bet = [(0.1 * torch.randn(s[1] * s[2], generator=g, dtype=torch.float64)).to(dtype) for s in shapes]
    og = (1.0 + 0.2 * torch.randn(class_n, generator=g, dtype=torch.float64)).to(dtype)
    ob = (0.1 * torch.randn(class_n, generator=g, dtype=torch.float64)).to(dtype)
  else:
    gam = [torch.ones(s[1] * s[2], dtype=dtype) for s in shapes]
    bet = [torch.zeros(s[1] * s[2], dtype=dtype) for s in shapes]
    og, ob = torch.ones(class_n, dtype=dtype), torch.zeros(class_n, dtype=dtype)
  return StackParams(W, bias, gam, bet, og, ob)


def route_stack(emb: torch.Tensor, p: StackParams, lpad: int, rpad: int, iters: int, sdr: bool,
                dropout_masks: Optional[Sequence[Optional[torch.Tensor]]] = None,
                length_eps: float = LENGTH_EPS, frame_at_a_time: bool = False,
                return_capsules: bool = False):
  """The whole hot path, naive:145-193: for each layer window -> u_hat -> SDR|DR ->
  LayerNorm(O*D) -> dropout; then head = ln_output(length(.)).

  dropout_masks[i], if given, is the already-scaled keep mask (0 or 1/(1-rate)) of
  Dropout `dropout_mid_%d` with shape [B,S,O,D]; None = inference (training=False)."""
  n = len(p.W)
  caps = []
  for i in range(n):
    is_last = i == n - 1
    if sdr and frame_at_a_time:
      v = route_layer_frame_sdr(emb, p.W[i], p.bias[i], lpad, rpad, iters, is_last)
    else:
      v = route_layer(emb, p.W[i], p.bias[i], lpad, rpad, iters, sdr, is_last)
    B, S, O, D = v.shape
    emb = layer_norm(v.reshape(B, S, O * D), p.ln_gamma[i], p.ln_beta[i]).reshape(B, S, O, D)
    if dropout_masks is not None and dropout_masks[i] is not None:
      emb = emb * dropout_masks[i]
    caps.append(emb)
  logits = layer_norm(length(emb, axis=-1, epsilon=length_eps), p.lno_gamma, p.lno_beta)
  if return_capsules:
    return logits, caps
  return logits


def greedy_ctc(logits: torch.Tensor, frame_lengths: Sequence[int]) -> List[List[int]]:
  """Greedy CTC used for the parity criterion: argmax per routing frame
  for s < input_length // 4 (trainer_sr.py:110), collapse repeats, drop blank =
  class_n - 1 (trainer_sr.py:133-134)."""
  blank = logits.shape[-1] - 1
  best = torch.argmax(logits, dim=-1)
  out = []
  for b, n in enumerate(frame_lengths):
    prev, seq = -1, []
    for t in best[b, :int(n)].tolist():
      if t != prev and t != blank:
        seq.append(t)
      prev = t
    out.append(seq)
  return out


# --------------------------------------------------------------------------------------------
# capsulation front-end (SURVEY.md 8f "next-1"): fbank -> primary capsules
# --------------------------------------------------------------------------------------------
def _conv2d_same(x: torch.Tensor, kernel: torch.Tensor, bias: torch.Tensor, stride: int) -> torch.Tensor:
  """tf.keras.layers.Conv2D(padding='same') on NHWC with a [kh,kw,cin,cout] kernel:
  out = ceil(in/stride), pad_total = max((out-1)*stride + k - in, 0), pad_before = pad_total // 2."""
  B, H, W, _ = x.shape
  k = kernel.shape[0]
  oh, ow = -(-H // stride), -(-W // stride)
  ph, pw = max((oh - 1) * stride + k - H, 0), max((ow - 1) * stride + k - W, 0)
  xp = torch.nn.functional.pad(x.permute(0, 3, 1, 2), (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
  y = torch.nn.functional.conv2d(xp, kernel.permute(3, 2, 0, 1).contiguous(), bias, stride=stride)
  return y.permute(0, 2, 3, 1)


def feat_mask(x: torch.Tensor, lengths: torch.Tensor, div: int) -> torch.Tensor:
  """tfsr/helper/model_helper.py:125-140: frames at or beyond ceil(len / div) are zeroed."""
  n = torch.ceil(lengths.to(torch.float64) / div).to(torch.int64)
  mask = (torch.arange(x.shape[1])[None, :] < n[:, None]).to(x.dtype)
  return x * mask[:, :, None, None]


def pos_enc(length_: int, hidden: int) -> torch.Tensor:
  """tfsr/helper/model_helper.py:30-58 (get_pos_enc; computed in float32 there)."""
  nts = hidden // 2
  inc = torch.tensor(math.log(1.0e4), dtype=torch.float32) / (torch.tensor(float(nts)) - 1)
  inv = torch.exp(torch.arange(nts, dtype=torch.float32) * -inc)
  st = torch.arange(length_, dtype=torch.float32)[:, None] * inv[None, :]
  return torch.cat([torch.sin(st), torch.cos(st)], dim=1)


def capsulate(feats: torch.Tensor, lengths, fe: dict, training: bool = False,
              dropout: Optional[dict] = None, einsum_variant: bool = False, bn_momentum: float = 0.99,
              bn_eps: float = 1e-3) -> Tuple[torch.Tensor, dict]:
  """naive:129-142 with CapsulationLayer.call (tfsr/model/sequence_router.py:67-82):
  fbank [B,T,F] -> primary capsules [B,S,PH,PD].  `fe`: parameters by the names of tests/golden
  (TF layouts).  training=True: Keras BatchNormalization in training mode (batch statistics over
  (B,time,freq), biased variance; returns the updated moving statistics) and the already scaled keep
  masks of `dropout` ({"cnn<path>_<stage>", "encaps<path>", "inp"}) where the reference has Dropout
  layers (sequence_router.py:60-61,76-77; naive:81-82,133,142).  einsum_variant: the sqrt(PH) scale and
  positional encoding of sequence_router_einsum.py:130-131.
  Returns (emb, {"bn<stage>_mean"/"bn<stage>_var": updated moving statistics})."""
  dt = feats.dtype
  f = {k: torch.as_tensor(v).to(dt) for k, v in fe.items()}
  lens = torch.as_tensor(lengths)
  dropout = dropout or {}
  drop = lambda t, key: t * torch.as_tensor(dropout[key]).to(dt).reshape(t.shape) \
      if (training and key in dropout) else t
  moving = {}
  x = feats[..., None]                                        # sequence_router.py:69
  for st in range(2):                                         # sequence_router.py:71-81
    x1 = drop(_conv2d_same(x, f["cnn0_%d_kernel" % st], f["cnn0_%d_bias" % st], 2), "cnn0_%d" % st)
    x2 = drop(_conv2d_same(x, f["cnn1_%d_kernel" % st], f["cnn1_%d_bias" % st], 2), "cnn1_%d" % st)
    x = torch.maximum(x1, x2)
    x = feat_mask(x, lens, 2 ** (st + 1))
    mean, var = f["bn%d_mean" % st], f["bn%d_var" % st]
    if training:
      bm, bv = x.mean(dim=(0, 1, 2)), x.var(dim=(0, 1, 2), unbiased=False)
      moving["bn%d_mean" % st] = mean * bn_momentum + bm * (1 - bn_momentum)
      moving["bn%d_var" % st] = var * bn_momentum + bv * (1 - bn_momentum)
      mean, var = bm, bv
    x = (x - mean) / torch.sqrt(var + bn_eps) * f["bn%d_gamma" % st] + f["bn%d_beta" % st]
    x = feat_mask(x, lens, 2 ** (st + 1))
  B, S = x.shape[0], x.shape[1]
  emb = x.reshape(B, S, -1) @ f["dense_kernel"] + f["dense_bias"]      # naive:131-132
  PH = emb.shape[-1]
  if einsum_variant:                                                   # einsum:130-131
    emb = emb * torch.sqrt(torch.tensor(float(PH), dtype=torch.float32)).to(dt) + pos_enc(S, PH).to(dt)
  emb = emb[..., None]
  emb = torch.maximum(drop(_conv2d_same(emb, f["encaps0_kernel"], f["encaps0_bias"], 1), "encaps0"),
                      drop(_conv2d_same(emb, f["encaps1_kernel"], f["encaps1_bias"], 1), "encaps1"))  # naive:133
  emb = feat_mask(emb, lens, 4)                                        # naive:134
  emb = squash(emb, -1)                                                # naive:137
  PD = emb.shape[-1]
  flat = layer_norm(emb.reshape(B, S, PH * PD), f["ln_input_gamma"], f["ln_input_beta"])   # naive:139-141
  emb = drop(flat.reshape(B, S, PH, PD), "inp")                        # naive:142
  return emb, moving
