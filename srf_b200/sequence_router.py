"""SequenceRouter: host-side mirror of the reference's model class for the routing path.

Reference: tfsr/model/sequence_router_naive.py:33-193 (`SequenceRouter(config, logger,
class_n)`, `call(inputs, input_lengths=, training=)`), selected by --model-caps-type
(tfsr/trainer_sr.py:188-199).  The knobs are read from the same config attributes
(tfsr/helper/common_helper.py:395-445):

    model_encoder_num (LAYER)            model_caps_primary_num (PH)
    model_caps_convolution_num (CH)      model_caps_{primary,convolution,class}_dim (DIM)
    model_caps_window_lpad / _rpad       model_caps_context (True = SDR, False = DR)
    model_caps_iter (ITER)               train_inn_dropout

The routing stack (naive:145-193) runs in hand-written CUDA through the C-ABI library;
this class owns the parameters (W%d, b%d, ln_mid%d, ln_output) as device tensors.
"""
from __future__ import annotations

import itertools
from typing import List, Optional, Sequence

import torch

from . import routing

_version_counter = itertools.count(1)


def layer_shapes(enc_num, ph, ch, class_n, pd, cd, vd, window):
  """(I, O, D, d) per layer -- naive:86-95."""
  if enc_num == 1:
    return [(ph * window, class_n, vd, pd)]
  shapes = [(ph * window, ch, cd, pd)]
  for _ in range(1, enc_num - 1):
    shapes.append((ch * window, ch, cd, cd))
  shapes.append((ch * window, class_n, vd, cd))
  return shapes


class RoutingStack:
  """Parameters + forward of the routing stack (everything after `emb` in naive:145-193)."""

  def __init__(self, enc_num: int, ph: int, ch: int, class_n: int, pd: int, cd: int, vd: int,
               lpad: int, rpad: int, iters: int, sdr: bool, inn_dropout: float = 0.1,
               device=None, seed: Optional[int] = None, uhat_mode: str = "exact",
               length_eps: float = routing.LENGTH_EPS, bwd_uhat_mode: Optional[str] = None):
    if enc_num < 1:
      raise ValueError("model_encoder_num must be >= 1")
    self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
        else torch.device(device)
    self.lpad, self.rpad, self.window = lpad, rpad, lpad + rpad + 1
    self.iters, self.sdr = int(iters), bool(sdr)
    self.class_n = class_n
    self.inn_dropout = float(inn_dropout)
    self.uhat_mode = uhat_mode
    # mixed-precision training: the backward recomputes u_hat in this mode (None = the forward's).
    # Nothing of size I*O*D is saved by the forward, so the two are independent: e.g. forward "f16" (the
    # fused wavefront kernel) with backward "bf16" (bf16 u_hat stream for the BPTT sweep); the gradient then
    # carries the backward mode's rounding class.
    self.bwd_uhat_mode = bwd_uhat_mode
    self.length_eps = length_eps
    self.shapes = layer_shapes(enc_num, ph, ch, class_n, pd, cd, vd, self.window)
    g = torch.Generator(device="cpu")
    if seed is not None:
      g.manual_seed(seed)
    # W, bias ~ N(0, 0.1^2) (naive:97-103); LayerNorm gamma = 1, beta = 0 (Keras default)
    self.wgt = [(torch.randn(s, generator=g) * 0.1).to(self.device) for s in self.shapes]
    self.bias = [(torch.randn(s[:3], generator=g) * 0.1).to(self.device) for s in self.shapes]
    self.ln_gamma = [torch.ones(s[1] * s[2], device=self.device) for s in self.shapes]
    self.ln_beta = [torch.zeros(s[1] * s[2], device=self.device) for s in self.shapes]
    self.lno_gamma = torch.ones(class_n, device=self.device)
    self.lno_beta = torch.zeros(class_n, device=self.device)
    self._version = next(_version_counter)
    self.handle = routing.default_handle(self.device)

  # -- parameter access (checkpoint averaging contract, utils/average_ckpt_sr.py:137-170) --
  def named_parameters(self):
    out = []
    for i, (w, b) in enumerate(zip(self.wgt, self.bias)):
      out.append(("W%d" % i, w))
    for i, b in enumerate(self.bias):
      out.append(("b%d" % i, b))
    for i, (g, b) in enumerate(zip(self.ln_gamma, self.ln_beta)):
      out.append(("ln_mid%d/gamma" % (i + 1), g))
      out.append(("ln_mid%d/beta" % (i + 1), b))
    out.append(("ln_output/gamma", self.lno_gamma))
    out.append(("ln_output/beta", self.lno_beta))
    return out

  def mark_weights_changed(self):
    """Call after modifying W/bias in place: invalidates the packed-weight cache tag."""
    self._version = next(_version_counter)

  def parameters(self):
    return [t for _, t in self.named_parameters()]

  def requires_grad_(self, flag: bool = True):
    """Make the parameter tensors autograd leaves (for srf_b200.autograd.route_stack + a torch
    optimiser)."""
    for t in self.parameters():
      t.requires_grad_(flag)
    return self

  def load_oracle_params(self, p):
    """Copy parameters from an oracle StackParams (tests / bench)."""
    for i in range(len(self.shapes)):
      self.wgt[i] = p.W[i].to(torch.float32).to(self.device).contiguous()
      self.bias[i] = p.bias[i].to(torch.float32).to(self.device).contiguous()
      self.ln_gamma[i] = p.ln_gamma[i].to(torch.float32).to(self.device).contiguous()
      self.ln_beta[i] = p.ln_beta[i].to(torch.float32).to(self.device).contiguous()
    self.lno_gamma = p.lno_gamma.to(torch.float32).to(self.device).contiguous()
    self.lno_beta = p.lno_beta.to(torch.float32).to(self.device).contiguous()
    self.mark_weights_changed()

  def layer_args(self, dropout_masks: Optional[Sequence[Optional[torch.Tensor]]] = None
                 ) -> List[routing.LayerArgs]:
    n = len(self.shapes)
    args = []
    for i in range(n):
      last = i == n - 1
      # the packed-weight cache tag follows torch's in-place version counters, so an optimiser
      # step (same storage, new values) can never be served from a stale packed copy
      tag = hash((self._version, self.wgt[i]._version, self.bias[i]._version,
                  self.wgt[i].data_ptr(), self.bias[i].data_ptr())) & 0x7FFFFFFFFFFFFFFF
      args.append(routing.LayerArgs(
          W=self.wgt[i], bias=self.bias[i], lpad=self.lpad, rpad=self.rpad, iters=self.iters,
          sdr=self.sdr, mask_class0=last, ln_gamma=self.ln_gamma[i], ln_beta=self.ln_beta[i],
          dropout_mask=None if dropout_masks is None else dropout_masks[i],
          head_gamma=self.lno_gamma if last else None, head_beta=self.lno_beta if last else None,
          uhat_mode=self.uhat_mode, length_eps=self.length_eps, weights_version=tag or 1))
    return args

  def make_dropout_masks(self, B: int, S: int, generator=None):
    """Scaled keep masks of dropout_mid_%d (naive:112-114, 191): 0 or 1/(1-rate)."""
    keep = 1.0 - self.inn_dropout
    return [(torch.rand((B, S, s[1], s[2]), device=self.device, generator=generator) < keep)
            .to(torch.float32) / keep for s in self.shapes]

  def forward(self, emb, training: bool = False, dropout_masks=None, return_capsules=False,
              out_logits=None):
    """emb [B,S,PH,PD] (primary capsules, naive:139-142) -> logits [B,S,class_n] (naive:193)."""
    emb = routing.as_device_tensor(emb, self.device)
    if training and dropout_masks is None and self.inn_dropout > 0:
      dropout_masks = self.make_dropout_masks(emb.shape[0], emb.shape[1])
    return routing.route_stack_fwd(emb, self.layer_args(dropout_masks), self.handle,
                                   return_capsules=return_capsules, out_logits=out_logits)

  __call__ = forward

  # -- training step pieces (BASELINE.json cfg-4) ------------------------------------------
  def forward_train(self, emb, dropout_masks=None):
    """Forward that keeps what the backward needs (layer inputs, pre-LayerNorm capsules,
    dropout masks): one srf_route_stack_fwd call with every layer's outputs requested (the fused
    layer-wavefront kernel where the mode and shape allow it).  Returns logits [B,S,class_n]."""
    emb = routing.as_device_tensor(emb, self.device)
    if dropout_masks is None and self.inn_dropout > 0:
      dropout_masks = self.make_dropout_masks(emb.shape[0], emb.shape[1])
    args = self.layer_args(dropout_masks)
    logits, caps, raws = routing.route_stack_fwd_train(emb, args, self.handle)
    self._saved = [(emb if i == 0 else caps[i - 1], a, raws[i]) for i, a in enumerate(args)]
    self._saved_caps = caps
    return logits

  def backward(self, d_logits, need_d_emb=True, out=None, after_layer=None):
    """Backward through the routing stack of the last forward_train.  Returns
    ({parameter name: gradient}, d_emb).  Default: ONE srf_route_stack_bwd call.  With `out`
    ({parameter name: zeroed tensor to accumulate into}, e.g. views of a flat gradient buffer) and /
    or `after_layer(i)` (called once layer i's gradients are enqueued, last layer first -- the
    data-parallel step launches that layer's all-reduce there) the layers are walked one
    srf_route_layer_bwd call at a time."""
    n = len(self._saved)
    if self.bwd_uhat_mode is not None and self.bwd_uhat_mode != self.uhat_mode:
      import dataclasses
      self._saved = [(x, dataclasses.replace(a, uhat_mode=self.bwd_uhat_mode), r) for x, a, r in self._saved]
    names = lambda i: {"dW": "W%d" % i, "dbias": "b%d" % i, "dgamma": "ln_mid%d/gamma" % (i + 1),
                       "dbeta": "ln_mid%d/beta" % (i + 1), "dhead_gamma": "ln_output/gamma",
                       "dhead_beta": "ln_output/beta"}
    grads, d_out = {}, None
    if out is None and after_layer is None:
      per_layer = routing.route_stack_bwd(self._saved[0][0], [a for _, a, _ in self._saved], self._saved_caps,
                                          [r for _, _, r in self._saved], d_logits, need_d_emb=need_d_emb,
                                          handle=self.handle)
      for i, g in enumerate(per_layer):
        for k, nm in names(i).items():
          if g[k] is not None:
            grads[nm] = g[k]
      d_out = per_layer[0]["d_emb"]
    else:
      for i in reversed(range(n)):
        x, a, raw = self._saved[i]
        o_i = None if out is None else {k: out.get(nm) for k, nm in names(i).items()}
        g = routing.route_layer_bwd(x, a, raw, d_out=d_out, d_logits=d_logits if i == n - 1 else None,
                                    need_d_emb=need_d_emb or i > 0, handle=self.handle, out=o_i)
        for k, nm in names(i).items():
          if g[k] is not None:
            grads[nm] = g[k]
        d_out = g["d_emb"]
        if after_layer is not None:
          after_layer(i)
    self._saved, self._saved_caps = [], []
    return grads, d_out

  def ctc_train_step_grads(self, emb, labels, input_lengths, label_lengths, dropout_masks=None,
                           grad_scale: float = 1.0, out=None, after_layer=None):
    """fwd + CTC loss + bwd (tfsr/trainer_sr.py:56-71), all in the CUDA library (CTC loss:
    srf_ctc_loss, blank = class_n - 1).  Returns (summed loss, grads, d_emb); the gradients are
    those of grad_scale * loss (1/global_batch = tf.nn.compute_average_loss, trainer_sr.py:67-68)."""
    from . import training
    logits = self.forward_train(emb, dropout_masks)
    loss, d_logits = training.ctc_loss(logits, labels, input_lengths, label_lengths,
                                       blank=self.class_n - 1, grad_scale=grad_scale, handle=self.handle)
    grads, d_emb = self.backward(d_logits, out=out, after_layer=after_layer)
    return loss.sum(), grads, d_emb


# ----------------------------------------------------------------------------------------
# Drop-in model class
# ----------------------------------------------------------------------------------------
def _conv2d_same_nhwc(x, kernel, bias, stride):
  """Keras Conv2D(padding='same') on NHWC input with a TF-layout kernel [kh,kw,cin,cout]:
  out = ceil(in/stride), pad_total = max((out-1)*stride + k - in, 0), pad_before = pad_total//2."""
  import torch.nn.functional as F
  B, H, W, _ = x.shape
  k = kernel.shape[0]
  oh, ow = -(-H // stride), -(-W // stride)
  ph, pw = max((oh - 1) * stride + k - H, 0), max((ow - 1) * stride + k - W, 0)
  xp = F.pad(x.permute(0, 3, 1, 2), (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
  y = F.conv2d(xp, kernel.permute(3, 2, 0, 1).contiguous(), bias, stride=stride)
  return y.permute(0, 2, 3, 1)


def _pos_enc(length, hidden, device):
  """tfsr/helper/model_helper.py:30-58 (float32 throughout)."""
  import math
  nts = hidden // 2
  inc = torch.tensor(math.log(1.0e4), dtype=torch.float32) / (torch.tensor(float(nts)) - 1)
  inv = torch.exp(torch.arange(nts, dtype=torch.float32) * -inc)
  st = torch.arange(length, dtype=torch.float32)[:, None] * inv[None, :]
  return torch.cat([torch.sin(st), torch.cos(st)], dim=1).to(device)


def _feat_mask(x, lengths, div):
  """tfsr/helper/model_helper.py:125-140: zero the frames at or beyond ceil(len / div)."""
  n = torch.ceil(lengths.to(torch.float32) / div).to(torch.int64)
  mask = (torch.arange(x.shape[1], device=x.device)[None, :] < n[:, None]).to(x.dtype)
  return x * mask[:, :, None, None]


class SequenceRouter:
  """Drop-in for tfsr.model.sequence_router_naive.SequenceRouter (naive:33-193): same
  constructor `(config, logger, class_n)`, same call `model(inputs, input_lengths=...,
  training=...)` -> logits [B, ceil(T/4), class_n] (ignored extra kwargs such as `mask`,
  `att_mask` are accepted, utils/average_ckpt_sr.py:127-128).

  The routing stack (naive:145-193) -- the hot path -- runs in the CUDA library.  The
  capsulation front-end (naive:129-142, sequence_router.py:44-82; SURVEY.md 8f "next-1") is
  expressed with torch ops on the same device for now (inference semantics: dropouts off,
  BatchNorm with moving statistics); it feeds the hot path and is not part of the measured
  path.
  """

  def __init__(self, config, logger, class_n, device=None, seed=None, uhat_mode="exact", caps_type=None):
    self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
        else torch.device(device)
    import math
    # --model-caps-type (tfsr/trainer_sr.py:188-199): "naive" and "lowmemory" compute the same function;
    # "einsum" scales the projected frame by sqrt(PH), adds the positional encoding
    # (sequence_router_einsum.py:130-131) and uses 1e-9 inside the head's length (einsum:238)
    self.caps_type = caps_type or getattr(config, "model_caps_type", None) or "naive"
    if self.caps_type not in ("naive", "lowmemory", "einsum"):
      raise ValueError("model_caps_type must be naive, lowmemory or einsum, got %r" % (self.caps_type,))
    self.stride = 2
    self.cnn_n = config.model_conv_layer_num
    self.feat_dim = math.ceil(config.feat_dim / (self.stride * self.cnn_n))   # naive:50
    self.nfilt = config.model_conv_filter_num
    self.class_n = class_n
    self.enc_num = config.model_encoder_num
    self.lpad, self.rpad = config.model_caps_window_lpad, config.model_caps_window_rpad
    self.is_context = bool(config.model_caps_context)
    self.iter = config.model_caps_iter
    self.caps_inp_n, self.caps_inp_d = config.model_caps_primary_num, config.model_caps_primary_dim
    self.stack = RoutingStack(
        self.enc_num, config.model_caps_primary_num, config.model_caps_convolution_num, class_n,
        config.model_caps_primary_dim, config.model_caps_convolution_dim, config.model_caps_class_dim,
        self.lpad, self.rpad, self.iter, self.is_context,
        inn_dropout=getattr(config, "train_inn_dropout", 0.1), device=self.device, seed=seed,
        uhat_mode=uhat_mode, length_eps=1e-9 if self.caps_type == "einsum" else routing.LENGTH_EPS)
    self.inp_dropout = float(getattr(config, "train_inp_dropout", 0.1))
    self.fe = {}   # front-end parameters, TF layouts; filled by load_frontend / first call
    self._fe_seed = seed
    if logger is not None:
      logger.info("Layer x %d, Iter x %d, Init %s, Win %d (l:%d, r:%d), "
                  % (self.enc_num, self.iter, "SDR" if self.is_context else "DR",
                     self.lpad + self.rpad + 1, self.lpad, self.rpad))

  # -- parameters ------------------------------------------------------------------------
  def _init_frontend(self, in_feat):
    g = torch.Generator().manual_seed(0 if self._fe_seed is None else self._fe_seed)

    def glorot(shape):
      rf = 1
      for s in shape[:-2]:
        rf *= s
      limit = (6.0 / (shape[-2] * rf + shape[-1] * rf)) ** 0.5
      return ((torch.rand(shape, generator=g) * 2 - 1) * limit).to(self.device)

    # the reference indexes conv_layers[0][conv_idx] / conv_layers[1][conv_idx]
    # (sequence_router.py:76-77), so the SECOND index is the conv stage
    for li in range(2):
      for pi in range(self.cnn_n):
        cin = 1 if pi == 0 else self.nfilt
        self.fe["cnn%d_%d_kernel" % (li, pi)] = glorot((3, 3, cin, self.nfilt))
        self.fe["cnn%d_%d_bias" % (li, pi)] = torch.zeros(self.nfilt, device=self.device)
    for li in range(self.cnn_n):
      for n, v in (("gamma", 1.0), ("beta", 0.0), ("mean", 0.0), ("var", 1.0)):
        self.fe["bn%d_%s" % (li, n)] = torch.full((self.nfilt,), v, device=self.device)
    self.fe["dense_kernel"] = glorot((self.feat_dim * self.nfilt, self.caps_inp_n))
    self.fe["dense_bias"] = torch.zeros(self.caps_inp_n, device=self.device)
    for pi in range(2):
      self.fe["encaps%d_kernel" % pi] = glorot((3, 3, 1, self.caps_inp_d))
      self.fe["encaps%d_bias" % pi] = torch.zeros(self.caps_inp_d, device=self.device)
    n = self.caps_inp_n * self.caps_inp_d
    self.fe["ln_input_gamma"] = torch.ones(n, device=self.device)
    self.fe["ln_input_beta"] = torch.zeros(n, device=self.device)

  def load_frontend(self, params: dict):
    """params: name -> array in the reference's (TF) layouts, names as in tests/golden."""
    self.fe = {k: torch.as_tensor(v, dtype=torch.float32).to(self.device) for k, v in params.items()}

  def named_parameters(self):
    return [("frontend/" + k, v) for k, v in sorted(self.fe.items())] + self.stack.named_parameters()

  def get_weights(self):
    return [v.detach().cpu().numpy() for _, v in self.named_parameters()]

  def set_weights(self, weights):
    named = self.named_parameters()
    if len(weights) != len(named):
      raise ValueError("expected %d arrays, got %d" % (len(named), len(weights)))
    with torch.no_grad():
      for (name, t), w in zip(named, weights):
        t.copy_(torch.as_tensor(w, dtype=torch.float32).reshape(t.shape))
    self.stack.mark_weights_changed()

  @property
  def trainable_variables(self):
    return [v for _, v in self.named_parameters()]

  # -- forward ---------------------------------------------------------------------------
  def capsulate(self, inputs, input_lengths, training: bool = False, dropout: Optional[dict] = None):
    """fbank [B,T,feat] -> primary capsules emb [B,S,PH,PD] (naive:129-142) in the CUDA library
    (srf_capsulate_fwd: no framework op between the features and the routing stack).
    training=True: BatchNormalization with batch statistics + moving-average update, and the keep
    masks in `dropout` (routing.capsulate_fwd) where the reference has Dropout layers; forward only
    -- the differentiable training path is `capsulate_autograd`."""
    x = routing_as_tensor(inputs, self.device)
    if not self.fe:
      self._init_frontend(x.shape[-1])
    if self.cnn_n != 2:
      raise ValueError("the reference's CapsulationLayer only works with model_conv_layer_num == 2 "
                       "(sequence_router.py:52-63,76-77)")
    return routing.capsulate_fwd(x, input_lengths, self.fe, self.nfilt, self.caps_inp_n, self.caps_inp_d,
                                 training=training, dropout=dropout, pos_enc=self.caps_type == "einsum",
                                 handle=self.stack.handle)

  def capsulate_autograd(self, inputs, input_lengths, training: bool = False):
    """The same front-end as differentiable torch ops, for training through torch autograd
    (tfsr/trainer_sr.py:62-71 differentiates it with tf.GradientTape).  training=True:
    Dropout(0.2) after every front-end convolution (sequence_router.py:60-61,76-77; naive:81-82,
    132), BatchNormalization with batch statistics and moving-average update (momentum 0.99),
    input dropout `train_inp_dropout` (naive:142)."""
    import torch.nn.functional as F
    x = routing_as_tensor(inputs, self.device)
    lens = torch.as_tensor(input_lengths).to(self.device)
    if not self.fe:
      self._init_frontend(x.shape[-1])
    f = self.fe
    drop = (lambda t: F.dropout(t, 0.2, True)) if training else (lambda t: t)
    x = x[..., None]
    for li in range(self.cnn_n):                       # sequence_router.py:71-81
      x1 = drop(_conv2d_same_nhwc(x, f["cnn0_%d_kernel" % li], f["cnn0_%d_bias" % li], self.stride))
      x2 = drop(_conv2d_same_nhwc(x, f["cnn1_%d_kernel" % li], f["cnn1_%d_bias" % li], self.stride))
      x = torch.maximum(x1, x2)
      x = _feat_mask(x, lens, self.stride ** (li + 1))
      mean, var = f["bn%d_mean" % li], f["bn%d_var" % li]
      if training:                                     # Keras BatchNormalization(axis=-1), training
        bm = x.mean(dim=(0, 1, 2))
        bv = x.var(dim=(0, 1, 2), unbiased=False)
        with torch.no_grad():
          mean.mul_(0.99).add_(bm.detach(), alpha=0.01)
          var.mul_(0.99).add_(bv.detach(), alpha=0.01)
        mean, var = bm, bv
      x = (x - mean) / torch.sqrt(var + 1e-3) * f["bn%d_gamma" % li] + f["bn%d_beta" % li]
      x = _feat_mask(x, lens, self.stride ** (li + 1))
    B, S = x.shape[0], x.shape[1]
    emb = x.reshape(B, S, self.feat_dim * self.nfilt) @ f["dense_kernel"] + f["dense_bias"]
    if self.caps_type == "einsum":                     # einsum:130-131
      emb = emb * float(self.caps_inp_n) ** 0.5 + _pos_enc(S, self.caps_inp_n, emb.device)
    emb = emb[..., None]                               # [B,S,PH,1]
    emb = torch.maximum(drop(_conv2d_same_nhwc(emb, f["encaps0_kernel"], f["encaps0_bias"], 1)),
                        drop(_conv2d_same_nhwc(emb, f["encaps1_kernel"], f["encaps1_bias"], 1)))
    emb = _feat_mask(emb, lens, self.stride ** 2)
    n2 = (emb * emb).sum(-1, keepdim=True)             # squash, naive:248-253
    emb = (n2 / (1.0 + n2)) * emb / torch.sqrt(n2 + 1e-7)
    flat = emb.reshape(B, S, self.caps_inp_n * self.caps_inp_d)
    flat = torch.nn.functional.layer_norm(flat, (flat.shape[-1],), f["ln_input_gamma"],
                                          f["ln_input_beta"], eps=1e-3)
    emb = flat.reshape(B, S, self.caps_inp_n, self.caps_inp_d)
    if training and self.inp_dropout > 0:
      emb = F.dropout(emb, self.inp_dropout, True)     # naive:142
    return emb.contiguous()

  # parameters that receive gradients (BatchNorm moving statistics do not)
  def parameters(self):
    if not self.fe:
      raise RuntimeError("front-end parameters are created at the first call (or load_frontend)")
    return [v for k, v in sorted(self.fe.items()) if not k.endswith(("_mean", "_var"))] + \
        self.stack.parameters()

  def requires_grad_(self, flag: bool = True):
    for t in self.parameters():
      t.requires_grad_(flag)
    return self

  def __call__(self, inputs, input_lengths=None, training=False, **kwargs):
    """training=False: inference (no autograd).  training=True (tfsr/trainer_sr.py:63): dropouts
    on, BatchNorm batch statistics, and the result is attached to the torch autograd graph --
    torch differentiates the front-end, the CUDA library the routing stack
    (srf_b200.autograd.route_stack); call `requires_grad_()` once to make the parameters leaves."""
    if input_lengths is None:
      raise ValueError("input_lengths is required (naive:121)")
    if training:
      from . import autograd
      emb = self.capsulate_autograd(inputs, input_lengths, training=True)
      return autograd.route_stack(self.stack, emb)
    with torch.no_grad():
      emb = self.capsulate(inputs, input_lengths)
      return self.stack.forward(emb, training=False)

  call = __call__


def routing_as_tensor(x, device):
  if not isinstance(x, torch.Tensor):
    if hasattr(x, "__dlpack__"):
      x = torch.from_dlpack(x)
    else:
      x = torch.as_tensor(x)
  return x.to(device=device, dtype=torch.float32)


class HostPipeline:
  """Host-buffer front door of RoutingStack for serving loops: pinned host primary capsules in,
  pinned host logits out, with the host->device copy of batch n+1 and the device->host copy of
  batch n-1 overlapping the routing of batch n (two staging slots, three CUDA streams).

      pipe = HostPipeline(stack, B, S)
      for n, host_emb in enumerate(batches):        # host_emb: pinned [B,S,PH,PD] fp32
        pipe.submit(host_emb)
        if n: use(pipe.result())                     # logits of batch n-1, pinned host tensor
      use(pipe.result())
  """

  def __init__(self, stack: "RoutingStack", B: int, S: int):
    self.stack, dev = stack, stack.device
    ph, pd = stack.shapes[0][0] // stack.window, stack.shapes[0][3]
    self.dev_in = [torch.empty(B, S, ph, pd, device=dev) for _ in range(2)]
    self.dev_out = [torch.empty(B, S, stack.class_n, device=dev) for _ in range(2)]
    self.host_out = [torch.empty(B, S, stack.class_n).pin_memory() for _ in range(2)]
    self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    self.ev_in = [torch.cuda.Event() for _ in range(2)]
    self.ev_c = [torch.cuda.Event() for _ in range(2)]
    self.ev_out = [torch.cuda.Event() for _ in range(2)]
    self.n_sub, self.n_res = 0, 0

  def submit(self, host_emb: torch.Tensor):
    k = self.n_sub & 1
    compute = torch.cuda.current_stream(self.stack.device)
    with torch.cuda.stream(self.s_in):
      if self.n_sub >= 2:
        self.s_in.wait_event(self.ev_c[k])          # slot's previous routing has consumed dev_in[k]
      self.dev_in[k].copy_(host_emb, non_blocking=True)
      self.ev_in[k].record(self.s_in)
    compute.wait_event(self.ev_in[k])
    if self.n_sub >= 2:
      compute.wait_event(self.ev_out[k])            # dev_out[k] has been copied out
    self.stack.forward(self.dev_in[k], out_logits=self.dev_out[k])
    self.ev_c[k].record(compute)
    with torch.cuda.stream(self.s_out):
      self.s_out.wait_event(self.ev_c[k])
      self.host_out[k].copy_(self.dev_out[k], non_blocking=True)
      self.ev_out[k].record(self.s_out)
    self.n_sub += 1

  def result(self) -> torch.Tensor:
    if self.n_res >= self.n_sub:
      raise RuntimeError("no batch in flight")
    k = self.n_res & 1
    self.ev_out[k].synchronize()
    self.n_res += 1
    return self.host_out[k]
