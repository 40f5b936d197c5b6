#!/usr/bin/env python
"""cfg-4 (BASELINE.json configs[3]) timing: SRF-SDR WSJ-shaped training step, data parallel.

  python tools/train_bench.py [--gpus N --steps K --warmup W --batch 64 --frames 375 --uhat bf16]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29544 tools/train_bench.py --gpus N

One step = forward (tcgen05 u_hat + streaming routing, training mode with dropout masks) +
CTC loss + backward through the routing stack + NCCL all-reduce of the flat fp32 gradient
(routing W, bias, LayerNorm) + fused Adam with the reference's warm-up schedule
(tfsr/trainer_sr.py:56-71, train_helper.py:32-68).  The global batch is split over the ranks
(strong scaling: B/N utterances per GPU).  Timed with CUDA events between barriers, max over
ranks; prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bench  # noqa: E402
from srf_b200 import RoutingStack, training  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=5)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--batch", type=int, default=64, help="global batch (utterances)")
  ap.add_argument("--frames", type=int, default=375, help="routing frames per utterance")
  ap.add_argument("--uhat", default="bf16", choices=["fp32", "tf32", "f16", "bf16", "fp32x3"])
  ap.add_argument("--bwd-uhat", default=None, choices=["fp32", "tf32", "f16", "bf16", "fp32x3"],
                  help="mode the backward recomputes u_hat in (default: the forward's)")
  ap.add_argument("--workload", default="cfg3")
  args = ap.parse_args()
  rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
  local = int(os.environ.get("LOCAL_RANK", 0))
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  w = bench.WORKLOADS[args.workload]
  B, S = args.batch // world, args.frames
  stack = RoutingStack(w["L"], w["PH"], w["CH"], w["class_n"], w["DIM"], w["DIM"], w["DIM"], w["lpad"],
                       w["rpad"], w["iters"], w["sdr"], device=dev, seed=0, inn_dropout=0.1,
                       uhat_mode=args.uhat, bwd_uhat_mode=args.bwd_uhat)
  trainer = training.TrainStep(stack, args.batch)
  g = torch.Generator().manual_seed(1 + rank)
  emb = torch.randn(B, S, w["PH"], w["DIM"], generator=g).to(dev)
  Lab = max(1, S // 3)
  labels = torch.randint(1, w["class_n"] - 1, (B, Lab), generator=g).to(dev)
  in_len = torch.full((B,), S, device=dev)
  lab_len = torch.full((B,), Lab, device=dev)
  h = stack.handle
  losses = []

  def step(it):
    losses.append(trainer.step(emb, labels, in_len, lab_len))

  for it in range(args.warmup):
    step(it)
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  l0 = h.launches
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for it in range(args.steps):
    step(args.warmup + it)
  e1.record()
  torch.cuda.synchronize()
  ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
  if world > 1:
    dist.barrier()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  ms_step = ms.item() / args.steps
  if rank == 0:
    print(json.dumps({
        "metric": "SRF-SDR training-step routing frames/sec (fwd + CTC + bwd + grad all-reduce + Adam)",
        "value": args.batch * S / (ms_step / 1e3), "unit": "routing frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "dtype": {"fp32": "f32", "fp32x3": "f32"}.get(args.uhat, args.uhat), "data": "synthetic",
        "config": {"workload": "cfg4: " + w["desc"] + ", training step", "global_batch": args.batch,
                   "routing_frames_per_utterance": S, "uhat": args.uhat,
                   "parallelism": "dp%d (utterance shards, NCCL all-reduce of %d gradient floats)"
                                  % (world, trainer.opt.flat.numel())},
        "gpu_launches": h.launches - l0,
        "loss_first_last": [losses[0].item(), losses[-1].item()]}), flush=True)
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
