"""World-size-2 gloo tests of the sharding logic (the per-rank compute is replaced by the CPU
oracle here; on the GPU box the same functions wrap RoutingStack.forward)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import srf_oracle as o
from srf_b200 import parallel


def test_shard_plan_is_balanced_and_complete():
  lens = [300, 120, 280, 90, 150, 310, 60]
  shards = parallel.shard_utterances(lens, 3)
  assert sorted(i for s in shards for i in s) == list(range(len(lens)))
  loads = [sum(lens[i] for i in s) for s in shards]
  assert max(loads) - min(loads) <= max(lens)
  assert parallel.shard_utterances(lens, 1) == [list(range(len(lens)))]
  assert parallel.shard_utterances([5], 4) == [[0], [], [], []]


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  torch.manual_seed(0)
  shapes = o.layer_shapes(2, 6, 5, 7, 4, 4, 4, 3)
  p = o.init_params(shapes, 7, seed=3)
  emb = torch.randn(5, 6, 6, 4, generator=torch.Generator().manual_seed(1))
  lens = [24, 10, 20, 24, 8]
  fwd = lambda e: o.route_stack(e, p, 1, 1, 1, True)
  full = parallel.route_sharded(fwd, emb, lens)
  ref = fwd(emb)
  ok_fwd = torch.allclose(full, ref, atol=1e-6)
  # gradient all-reduce of a flat buffer == sum of the per-rank gradients
  g = [torch.full((3, 2), float(rank + 1)), torch.arange(4, dtype=torch.float32) * (rank + 1)]
  parallel.allreduce_flat_grads(g)
  tot = sum(range(1, world + 1))
  ok_grad = torch.equal(g[0], torch.full((3, 2), float(tot))) and \
      torch.equal(g[1], torch.arange(4, dtype=torch.float32) * tot)
  q.put((rank, bool(ok_fwd), bool(ok_grad)))
  dist.destroy_process_group()


def test_two_rank_sharded_forward_and_grad_allreduce():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=120) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  assert sorted(r[0] for r in res) == [0, 1]
  assert all(r[1] and r[2] for r in res), res
