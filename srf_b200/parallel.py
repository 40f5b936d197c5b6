"""Multi-GPU plumbing for the routing path: one process per GPU, utterances sharded across
ranks (utterances are independent, forward needs no collective).  The reference
does the same split inside tf.distribute.MirroredStrategy (tfsr/trainer_sr.py:139,168-172);
here it is explicit and torch.distributed (NCCL on GPUs, gloo in the CPU tests) only moves the
per-rank results."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_utterances(lengths: Sequence[int], world_size: int) -> List[List[int]]:
  """Assign utterance indices to ranks, balancing the summed lengths (longest first, to the
  lightest rank; ties broken by rank index so that every rank computes the same plan)."""
  order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
  loads = [0] * world_size
  shards: List[List[int]] = [[] for _ in range(world_size)]
  for i in order:
    r = min(range(world_size), key=lambda k: (loads[k], k))
    shards[r].append(i)
    loads[r] += int(lengths[i])
  return [sorted(s) for s in shards]


def route_sharded(forward_fn, emb: torch.Tensor, lengths: Sequence[int], group=None) -> torch.Tensor:
  """Every rank holds the full batch `emb` [B,S,...]; each routes its own shard with
  `forward_fn(emb_shard) -> logits_shard` and the logits are all-gathered so that every rank
  returns the full [B,S,class_n] in the original utterance order."""
  world = dist.get_world_size(group)
  rank = dist.get_rank(group)
  shards = shard_utterances(lengths, world)
  mine = shards[rank]
  local = forward_fn(emb[mine].contiguous()) if mine else None
  n_max = max(len(s) for s in shards)
  probe = local if local is not None else forward_fn(emb[:1].contiguous())
  tail = tuple(probe.shape[1:])
  buf = torch.zeros((n_max,) + tail, dtype=probe.dtype, device=probe.device)
  if local is not None:
    buf[:len(mine)] = local
  gathered = [torch.empty_like(buf) for _ in range(world)]
  dist.all_gather(gathered, buf, group=group)
  out = torch.empty((emb.shape[0],) + tail, dtype=probe.dtype, device=probe.device)
  for r, idx in enumerate(shards):
    if idx:
      out[idx] = gathered[r][:len(idx)]
  return out


def allreduce_flat_grads(grads: Sequence[torch.Tensor], group=None) -> None:
  """Sum-all-reduce a list of gradient tensors as ONE flat fp32 buffer (the reference's
  implicit MirroredStrategy all-reduce at apply_gradients, tfsr/trainer_sr.py:71); results are
  copied back in place."""
  flat = torch.cat([g.reshape(-1) for g in grads])
  dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
  off = 0
  for g in grads:
    n = g.numel()
    g.copy_(flat[off:off + n].view_as(g))
    off += n
