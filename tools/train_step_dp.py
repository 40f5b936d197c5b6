#!/usr/bin/env python
"""cfg-4 (BASELINE.json configs[3]) functional check: data-parallel SRF-SDR training step.
Launch with torchrun on N GPUs:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29533 tools/train_step_dp.py

Every rank routes its own utterance shard (fwd + CTC loss + bwd in the CUDA library, loss scaled
by 1/global_batch as tfsr/trainer_sr.py:58,67-68), the flat routing-weight gradient is
all-reduced with NCCL, and the result is compared with the single-GPU gradient of the whole
batch computed on rank 0.  Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from srf_b200 import RoutingStack, parallel  # noqa: E402


def main():
  rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
  local = int(os.environ.get("LOCAL_RANK", 0))
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  # TIMIT-shaped SDR stack (L7 PH60 CH30 DIM8 w3 cls63), 8 utterances x 40 routing frames
  L, PH, CH, cls, DIM, lpad, rpad, B, S = 7, 60, 30, 63, 8, 1, 1, 8, 40
  stack = RoutingStack(L, PH, CH, cls, DIM, DIM, DIM, lpad, rpad, 1, True, device=dev, seed=0,
                       inn_dropout=0.0)
  g = torch.Generator().manual_seed(1)
  emb = torch.randn(B, S, PH, DIM, generator=g).to(dev)
  labels = torch.randint(1, cls - 1, (B, S // 3), generator=g).to(dev)
  in_len = torch.full((B,), S, dtype=torch.long, device=dev)
  lab_len = torch.full((B,), S // 3, dtype=torch.long, device=dev)
  lens = [S] * B
  shard = parallel.shard_utterances(lens, world)[rank]
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  loss, grads, _ = stack.ctc_train_step_grads(emb[shard], labels[shard], in_len[shard], lab_len[shard])
  names = sorted(k for k, v in grads.items() if v is not None)
  flat = [grads[k].mul_(1.0 / B) for k in names]
  if world > 1:
    parallel.allreduce_flat_grads(flat)
  torch.cuda.synchronize()
  dt = time.perf_counter() - t0
  ok, worst = True, 0.0
  if rank == 0:
    _, ref, _ = stack.ctc_train_step_grads(emb, labels, in_len, lab_len)
    for k, t in zip(names, flat):
      r = ref[k] / B
      err = ((t - r).abs().max() / r.abs().max().clamp_min(1e-30)).item()
      worst = max(worst, err)
    ok = worst < 1e-4
    n_par = sum(t.numel() for t in flat)
    print(json.dumps({"check": "dp_train_step_grads", "world": world, "ok": ok, "max_rel_err": worst,
                      "flat_grad_floats": n_par, "step_s_unoptimised_bwd": dt,
                      "config": "TIMIT-shaped SDR L7 PH60 CH30 DIM8 w3 cls63, 8 x 40 routing frames"}))
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()
  sys.exit(0 if ok else 1)


if __name__ == "__main__":
  main()
