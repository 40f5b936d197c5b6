"""Training-step pieces around the routing stack, all running in the
CUDA library: greedy CTC decode, CTC loss + gradient, Adam with the reference's warm-up schedule.

Reference: tfsr/trainer_sr.py:56-71 (train step), :133-134 (blank = class_n - 1),
tfsr/helper/train_helper.py:32-68 (CustomSchedule, Adam)."""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch

from . import _lib, routing


def _stream(dev):
  return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _i32(x, dev):
  return torch.as_tensor(x).to(device=dev, dtype=torch.int32).contiguous()


def ctc_greedy_decode(logits, frame_lengths, blank: Optional[int] = None,
                      handle: Optional[routing.Handle] = None) -> List[List[int]]:
  """logits [B,S,C] -> label id lists (argmax, collapse repeats, drop blank = C-1 by default)."""
  logits = routing.as_device_tensor(logits)
  h = handle or routing.default_handle(logits.device)
  B, S, C = logits.shape
  blank = C - 1 if blank is None else blank
  lens = _i32(frame_lengths, logits.device)
  ids = torch.empty((B, S), dtype=torch.int32, device=logits.device)
  n = torch.empty((B,), dtype=torch.int32, device=logits.device)
  rc = h.lib.srf_ctc_greedy_decode(h._h, routing._ptr(logits), routing._ptr(lens), B, S, C, blank,
                                   routing._ptr(ids), routing._ptr(n), _stream(logits.device))
  _lib.check(h.lib, h._h, rc, "srf_ctc_greedy_decode")
  ids, n = ids.cpu(), n.cpu()
  return [ids[b, :int(n[b])].tolist() for b in range(B)]


def ctc_loss(logits, labels, input_lengths, label_lengths, blank: Optional[int] = None,
             grad_scale: float = 1.0, handle: Optional[routing.Handle] = None):
  """-> (per-utterance loss [B], d_logits [B,S,C] = grad_scale * dloss/dlogits)."""
  logits = routing.as_device_tensor(logits)
  h = handle or routing.default_handle(logits.device)
  dev = logits.device
  B, S, C = logits.shape
  blank = C - 1 if blank is None else blank
  labels = _i32(labels, dev)
  if labels.dim() != 2 or labels.shape[0] != B:
    raise ValueError("labels must be [B, Lmax]")
  Lmax = max(1, labels.shape[1])
  if labels.shape[1] == 0:
    labels = torch.zeros((B, 1), dtype=torch.int32, device=dev)
  in_l, lab_l = _i32(input_lengths, dev), _i32(label_lengths, dev)
  loss = torch.empty((B,), dtype=torch.float32, device=dev)
  d_logits = torch.empty((B, S, C), dtype=torch.float32, device=dev)
  rc = h.lib.srf_ctc_loss(h._h, routing._ptr(logits), routing._ptr(labels), routing._ptr(in_l),
                          routing._ptr(lab_l), B, S, C, Lmax, blank, float(grad_scale),
                          routing._ptr(loss), routing._ptr(d_logits), _stream(dev))
  _lib.check(h.lib, h._h, rc, "srf_ctc_loss")
  return loss, d_logits


def warmup_lr(step: int, k: float, d_model: float, warmup_steps: float, max_lr: float = 10.0) -> float:
  """CustomSchedule.__call__ (train_helper.py:52-56): min(k * d_model^-0.5 * min(step^-0.5,
  step * warmup^-1.5), max_lr)."""
  step = float(step)
  if step <= 0.0:
    return 0.0   # tf: rsqrt(0) = inf, min(inf, 0 * warmup^-1.5) = 0
  return min(k * d_model ** -0.5 * min(step ** -0.5, step * warmup_steps ** -1.5), max_lr)


class FlatAdam:
  """tf.keras Adam over one flat fp32 buffer (parameters are views into it), so that a training
  step is one gradient all-reduce + one fused update."""

  def __init__(self, params: Sequence[torch.Tensor], beta1=0.9, beta2=0.98, eps=1e-9,
               handle: Optional[routing.Handle] = None):
    dev = params[0].device
    self.h = handle or routing.default_handle(dev)
    self.sizes = [p.numel() for p in params]
    self.flat = torch.cat([p.detach().reshape(-1).float() for p in params]).contiguous()
    self.m = torch.zeros_like(self.flat)
    self.v = torch.zeros_like(self.flat)
    self.beta1, self.beta2, self.eps, self.step_count = beta1, beta2, eps, 0
    self.views, off = [], 0
    for p, n in zip(params, self.sizes):
      self.views.append(self.flat[off:off + n].view(p.shape))
      off += n

  def step(self, flat_grad: torch.Tensor, lr: float):
    if flat_grad.numel() != self.flat.numel():
      raise ValueError("gradient has %d elements, parameters %d" % (flat_grad.numel(), self.flat.numel()))
    self.step_count += 1
    g = flat_grad.contiguous()
    rc = self.h.lib.srf_adam_step(self.h._h, routing._ptr(self.flat), routing._ptr(g), routing._ptr(self.m),
                                  routing._ptr(self.v), self.flat.numel(), float(lr), self.beta1, self.beta2,
                                  self.eps, self.step_count, _stream(self.flat.device))
    _lib.check(self.h.lib, self.h._h, rc, "srf_adam_step")


class TrainStep:
  """One data-parallel training step of the routing stack, as `tfsr/trainer_sr.py:56-71` does it:
  forward (training mode) + CTC loss scaled by 1/global_batch + backward + sum-all-reduce of the
  flat gradient over the ranks (NCCL / gloo through torch.distributed, only when a process group
  is initialised) + Adam with the warm-up schedule.  The stack's parameters become views into
  the optimiser's flat buffer, so the update is one fused kernel."""

  def __init__(self, stack, global_batch: int, k: float = 0.5, d_model: float = 256.0,
               warmup_steps: float = 1200.0, group=None):
    self.stack, self.global_batch, self.group = stack, int(global_batch), group
    self.k, self.d_model, self.warmup_steps = k, d_model, warmup_steps
    n = len(stack.shapes)
    # flat buffers ordered LAYER BY LAYER (W, bias, LayerNorm of layer i contiguous; the head's
    # LayerNorm rides with the last layer), so that the gradient of a layer is one contiguous slice
    # whose all-reduce starts as soon as that layer's backward is enqueued
    named = dict(stack.named_parameters())
    self.names, self.layer_slices, off = [], [], 0
    for i in range(n):
      group_names = ["W%d" % i, "b%d" % i, "ln_mid%d/gamma" % (i + 1), "ln_mid%d/beta" % (i + 1)]
      if i == n - 1:
        group_names += ["ln_output/gamma", "ln_output/beta"]
      size = sum(named[nm].numel() for nm in group_names)
      self.layer_slices.append((off, off + size))
      off += size
      self.names += group_names
    if set(self.names) != set(named):
      raise ValueError("unexpected parameters: %s" % sorted(set(named) ^ set(self.names)))
    self.opt = FlatAdam([named[nm] for nm in self.names], handle=stack.handle)
    views = dict(zip(self.names, self.opt.views))
    stack.wgt = [views["W%d" % i] for i in range(n)]
    stack.bias = [views["b%d" % i] for i in range(n)]
    stack.ln_gamma = [views["ln_mid%d/gamma" % (i + 1)] for i in range(n)]
    stack.ln_beta = [views["ln_mid%d/beta" % (i + 1)] for i in range(n)]
    stack.lno_gamma, stack.lno_beta = views["ln_output/gamma"], views["ln_output/beta"]
    stack.mark_weights_changed()
    self.gflat = torch.zeros_like(self.opt.flat)
    self.gviews, off = {}, 0
    for nm, sz in zip(self.names, self.opt.sizes):
      self.gviews[nm] = self.gflat[off:off + sz]
      off += sz
    self.iteration = 0

  def step(self, emb, labels, input_lengths, label_lengths, dropout_masks=None):
    """Returns the summed CTC loss of this rank's utterances (device scalar)."""
    import torch.distributed as dist
    self.iteration += 1
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
    self.gflat.zero_()
    works = []

    def after_layer(i):
      # this layer's gradient slice is complete on the compute stream: its sum-all-reduce runs on the
      # collective's own stream while the layers below are still in their backward
      # (trainer_sr.py:70-71 all-reduces inside apply_gradients, after the whole tape)
      if distributed:
        lo, hi = self.layer_slices[i]
        works.append(dist.all_reduce(self.gflat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    # the loss is pre-scaled by 1/global_batch (trainer_sr.py:58,67-68) inside the CTC kernel
    loss, _, _ = self.stack.ctc_train_step_grads(emb, labels, input_lengths, label_lengths,
                                                 dropout_masks=dropout_masks,
                                                 grad_scale=1.0 / self.global_batch, out=self.gviews,
                                                 after_layer=after_layer)
    for w in works:
      w.wait()
    # Keras evaluates the schedule at optimizer.iterations BEFORE the increment (0 on the first
    # apply_gradients: the first learning rate is 0); only Adam's bias correction uses iterations+1
    self.opt.step(self.gflat, warmup_lr(self.iteration - 1, self.k, self.d_model, self.warmup_steps))
    self.stack.mark_weights_changed()
    return loss
