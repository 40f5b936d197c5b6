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
               device=None, seed: Optional[int] = None, uhat_mode: str = "fp32",
               length_eps: float = routing.LENGTH_EPS):
    if enc_num < 1:
      raise ValueError("model_encoder_num must be >= 1")
    self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
        else torch.device(device)
    self.lpad, self.rpad, self.window = lpad, rpad, lpad + rpad + 1
    self.iters, self.sdr = int(iters), bool(sdr)
    self.class_n = class_n
    self.inn_dropout = float(inn_dropout)
    self.uhat_mode = uhat_mode
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
      args.append(routing.LayerArgs(
          W=self.wgt[i], bias=self.bias[i], lpad=self.lpad, rpad=self.rpad, iters=self.iters,
          sdr=self.sdr, mask_class0=last, ln_gamma=self.ln_gamma[i], ln_beta=self.ln_beta[i],
          dropout_mask=None if dropout_masks is None else dropout_masks[i],
          head_gamma=self.lno_gamma if last else None, head_beta=self.lno_beta if last else None,
          uhat_mode=self.uhat_mode, length_eps=self.length_eps, weights_version=self._version))
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
