"""torch.autograd bridge of the routing stack: the CUDA library's forward_train / backward pair as
one differentiable op, so that a torch front-end (and a torch optimiser) can train through it the
way `tf.GradientTape` does in tfsr/trainer_sr.py:62-71.

    logits = route_stack(stack, emb)          # emb may require grad; stack.parameters() do
    loss.backward()                           # fills emb.grad and the .grad of every parameter

One forward per stack may be outstanding at a time (the stack keeps the tensors its backward
needs)."""
from __future__ import annotations

from typing import Optional, Sequence

import torch


class _RouteStackFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, emb, stack, masks, *params):   # params: the stack's tensors, for the graph only
    stack.mark_weights_changed()                  # an optimiser may have updated them in place
    logits = stack.forward_train(emb.detach(), dropout_masks=masks)
    ctx.stack = stack
    ctx.need_d_emb = emb.requires_grad
    return logits

  @staticmethod
  def backward(ctx, d_logits):
    stack = ctx.stack
    grads, d_emb = stack.backward(d_logits.contiguous().float(), need_d_emb=ctx.need_d_emb)
    names = [n for n, _ in stack.named_parameters()]
    return (d_emb if ctx.need_d_emb else None, None, None) + tuple(grads.get(n) for n in names)


def route_stack(stack, emb: torch.Tensor,
                dropout_masks: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
  """Training-mode forward of `stack` (inner dropout on unless its rate is 0 or masks are given)
  that torch.autograd can differentiate."""
  params = [t for _, t in stack.named_parameters()]
  if dropout_masks is None and stack.inn_dropout > 0:
    dropout_masks = stack.make_dropout_masks(emb.shape[0], emb.shape[1])
  return _RouteStackFn.apply(emb, stack, dropout_masks, *params)
