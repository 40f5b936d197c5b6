#!/usr/bin/env python
"""bench.py -- SRF routing-layer throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload cfg3|cfg2|cfg1] [--uhat tf32|fp32x3|bf16|fp32]
    python bench.py --impl reference ...        # the CPU arm (oracle port; TensorFlow is not installable)

A "step" is one forward pass of the routing stack (primary capsules -> CTC logits,
tfsr/model/sequence_router_naive.py:145-193) over one synthetic batch.  Default workload is
cfg-3 of BASELINE.json (SRF-SDR WSJ-shaped, 31 labels + blank, batch 64 x 1500 fbank frames
= 64 x 375 routing frames, ITER=1), the configuration the metric "frames/sec at 1/2/4/8
B200" is quoted on; default arithmetic is the fused tcgen05 kernel (TF32 operands, u_hat never
leaves the SM; 1e-2 tolerance class with identical greedy-CTC strings, see
tests/test_routing_gpu.py::test_bench_mode_full_size_cfg3_parity).

The JSON line (rank 0) carries
  value / e2e   : every GPU routes its own 64-utterance batch, no collective (weak scaling);
  strong        : the 64-utterance batch sharded over the N GPUs (64/N per GPU), as BASELINE.json words cfg-3;
  train_cfg4    : the training step (fwd + CTC + bwd + NCCL gradient all-reduce + Adam), global batch 64 sharded;
  roofline      : the dominant kernel against the bound of the fused design (tensor: derived TF32 peak);
  cpu_baseline  : the oracle port on all host threads over the sample `--impl reference` times (N = 1).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: enc_num, PH, CH, class_n, DIM, lpad, rpad, iters, sdr, B, T(fbank frames)
    "cfg1": dict(desc="SRF-SDR TIMIT-shaped, ITER=1, L7 PH60 CH30 DIM8 w3 cls63, B=8 x T=300",
                 L=7, PH=60, CH=30, class_n=63, DIM=8, lpad=1, rpad=1, iters=1, sdr=True, B=8, T=300),
    "cfg2": dict(desc="SRF-DR TIMIT-shaped, ITER=3, LPAD=RPAD=3, L7 PH60 CH30 DIM8 cls63, B=8 x T=300",
                 L=7, PH=60, CH=30, class_n=63, DIM=8, lpad=3, rpad=3, iters=3, sdr=False, B=8, T=300),
    "cfg3": dict(desc="SRF-SDR WSJ-shaped, ITER=1, L10 PH60 CH30 DIM20 w5 cls32, B=64 x T=1500",
                 L=10, PH=60, CH=30, class_n=32, DIM=20, lpad=2, rpad=2, iters=1, sdr=True, B=64, T=1500),
}


def shapes_of(w):
  from srf_b200 import layer_shapes
  return layer_shapes(w["L"], w["PH"], w["CH"], w["class_n"], w["DIM"], w["DIM"], w["DIM"],
                      w["lpad"] + w["rpad"] + 1)


def algorithmic_work(w):
  """Per routing frame: F_uhat = 2 I O D d, F_route = 4 ITER I O D (SDR) /
  (4 ITER - 2) I O D (DR), bytes = 4 (H d + O D) per layer; weights once per launch."""
  window = w["lpad"] + w["rpad"] + 1
  f_uhat = f_route = bytes_frame = weights = 0
  for (I, O, D, d) in shapes_of(w):
    f_uhat += 2 * I * O * D * d
    f_route += (4 * w["iters"] if w["sdr"] else 4 * w["iters"] - 2) * I * O * D
    bytes_frame += 4 * (I // window * d + O * D)
    weights += 4 * (I * O * D * d + I * O * D)
  return f_uhat, f_route, bytes_frame, weights


class ClockSampler(threading.Thread):
  """Samples SM clock and throttle reasons of one GPU during the timed region."""
  REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
             0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
    self._stop_evt = threading.Event()
    self.ok = False
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
      self.ok = True
    except Exception:  # pylint: disable=broad-except
      self.ok = False

  def run(self):
    if not self.ok:
      return
    while not self._stop_evt.is_set():
      try:
        self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        try:
          mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:  # pylint: disable=broad-except
          mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, name in self.REASONS.items():
          if mask & bit:
            self.reasons.add(name)
      except Exception:  # pylint: disable=broad-except
        pass
      self._stop_evt.wait(0.05)

  def stop(self):
    self._stop_evt.set()
    self.join(timeout=2)
    s = sorted(self.samples)
    return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_reference_run(w, n_utts, n_frames, repeats=1, seed=0):
  """The reference's CPU path: TensorFlow is not installable here, so this
  times the oracle port (torch CPU fp32 restatement of naive:145-193) on all host threads."""
  from oracle import srf_oracle as o
  torch.set_num_threads(os.cpu_count() or 1)
  shapes = shapes_of(w)
  p = o.init_params(shapes, w["class_n"], seed=seed)
  emb = torch.randn(n_utts, n_frames, w["PH"], w["DIM"], generator=torch.Generator().manual_seed(seed))
  best = None
  for _ in range(repeats):
    t0 = time.perf_counter()
    o.route_stack(emb, p, w["lpad"], w["rpad"], w["iters"], w["sdr"])
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
  return n_utts * n_frames / best, best


def cpu_sample_shape(w):
  """Bounded sample of the workload for BOTH CPU legs (`cpu_baseline` of the GPU arm and every step
  of `--impl reference`), so that the two report the same thing.  Per-frame cost of the CPU path is
  independent of S and nearly independent of B; the sample is about 4 s of work on 16 cores."""
  if w["DIM"] >= 16:
    return (16, 375)
  return (w["B"], 75)


def run_reference_arm(args, w, rank, world):
  if rank != 0:
    return
  n_utts, n_frames = cpu_sample_shape(w)
  times = []
  for i in range(args.warmup + args.steps):
    fps, dt = cpu_reference_run(w, n_utts, n_frames)
    if i >= args.warmup:
      times.append(dt)
  ms = 1e3 * sum(times) / len(times)
  value = n_utts * n_frames / (ms / 1e3)
  sample = "%d utterances x %d routing frames of the %s model per step" % (n_utts, n_frames, args.workload)
  line = {
      "impl": "reference", "metric": "routing_frames_per_sec", "value": value, "unit": "routing frames/s",
      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
      "data": "synthetic",
      "config": {"workload": w["desc"], "sample": sample,
                 "note": "reference needs TensorFlow (not installable offline): oracle port timed"},
      "cpu_baseline": {"value": value, "unit": "routing frames/s", "cores": os.cpu_count(),
                       "kind": "port", "sample": sample},
      "e2e": {"value": value, "unit": "routing frames/s", "h2d_bytes_per_step": 0,
              "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
  }
  print(json.dumps(line), flush=True)


def build_stack(w, dev, uhat, inn_dropout=0.1, bwd_uhat=None):
  from srf_b200 import RoutingStack
  return RoutingStack(w["L"], w["PH"], w["CH"], w["class_n"], w["DIM"], w["DIM"], w["DIM"], w["lpad"],
                      w["rpad"], w["iters"], w["sdr"], device=dev, seed=0, uhat_mode=uhat,
                      inn_dropout=inn_dropout, bwd_uhat_mode=bwd_uhat)


UHAT_DESC = {
    "fp32": "FP32 CUDA cores, fused FP32 kernel",
    "tf32": "fused routing kernel: tcgen05 TF32 MMA into TMEM, u_hat consumed in place (never in HBM), "
            "SDR stack as one layer-wavefront launch (1e-2 tolerance class, identical greedy CTC)",
    "f16": "fused routing kernel: tcgen05 kind::f16 MMA on FP16 operand images (the 11-bit significand of TF32, "
           "2/3 of its operand bytes at d=20) into TMEM, u_hat consumed in place (never in HBM), SDR stack as one "
           "layer-wavefront launch (1e-2 tolerance class, identical greedy CTC)",
    "bf16": "two kernels: tcgen05 TF32 MMA, bf16 u_hat materialised in HBM, streaming routing kernel",
    "fp32x3": "3 x TF32 split (fp32-class u_hat, 1e-4 tolerance class); fused kernel or two-kernel path "
              "by the library's policy",
}


def packed_weight_bytes(w, uhat):
  """Bytes of the fused kernel's packed operand images (one pass over all layers): per input capsule
  32 * ceil(O/32) * 4 * ceil(D/4) rows of KC 16-byte chunks (csrc/capi.cu fused_geometry)."""
  total = 0
  for (I, O, D, d) in shapes_of(w):
    t4 = (D + 3) // 4
    t4 = 2 if t4 <= 2 else (4 if t4 <= 4 else 5)
    rows = 32 * ((O + 31) // 32) * 4 * t4
    kc = 2 * ((d + 1 + 15) // 16) if uhat == "f16" else 2 * ((d + 1 + 7) // 8)
    total += I * rows * kc * 16 * (2 if uhat == "fp32x3" else 1)
  return total


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-also", action="store_true", help="skip the strong-scaling, training and other-config legs")
  ap.add_argument("--uhat", default="f16", choices=["fp32", "tf32", "f16", "bf16", "fp32x3"],
                  help="u_hat arithmetic: f16 = fused tcgen05 kernel on FP16 operand images (default), tf32 = the "
                       "same kernel on TF32 operands, fp32x3 = the 1e-4 class, "
                       "bf16 = round-1 two-kernel path with bf16 u_hat in HBM, fp32 = CUDA-core kernel")
  args = ap.parse_args()
  w = dict(WORKLOADS[args.workload])

  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))

  if args.impl == "reference":
    run_reference_arm(args, w, rank, world)
    return

  if args.warmup < 3:
    args.warmup = 3
  if not torch.cuda.is_available():
    raise SystemExit("bench.py: no CUDA device; the routing path has no CPU fallback "
                     "(use --impl reference for the CPU arm)")
  import torch.distributed as dist
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)

  B, S = w["B"], (w["T"] + 3) // 4
  stack = build_stack(w, dev, args.uhat)
  g = torch.Generator().manual_seed(1000 + rank)
  n_bufs = 2
  host_emb = [torch.randn(B, S, w["PH"], w["DIM"], generator=g).pin_memory() for _ in range(n_bufs)]
  dev_emb = [h.to(dev) for h in host_emb]
  logits = torch.empty(B, S, w["class_n"], device=dev)
  host_logits = torch.empty(B, S, w["class_n"]).pin_memory()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(fn, steps):
    """`steps` calls between barriers, CUDA events on the launch stream, max over ranks (ms)."""
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for i in range(steps):
      fn(i)
    end.record()
    barrier()
    ms = start.elapsed_time(end)
    if world > 1:
      t = torch.tensor([ms], device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = t.item()
    return ms

  def step_resident(i):
    stack.forward(dev_emb[i % n_bufs], out_logits=logits)

  from srf_b200 import HostPipeline
  pipe = HostPipeline(stack, B, S)
  checks = []

  def step_e2e(i):
    # public host-buffer API: pinned host capsules in, pinned host logits out; every step's H2D
    # and D2H copies are issued inside the timed region (overlapped with the previous/next step's
    # routing by the pipeline's copy streams)
    pipe.submit(host_emb[i % n_bufs])
    if pipe.n_sub - pipe.n_res > 1:
      checks.append(float(pipe.result()[0, 0, 1]))

  for i in range(args.warmup):
    step_resident(i)
  l0 = stack.handle.launches
  sampler = ClockSampler(local_rank)
  sampler.start()
  ms_total = timed(step_resident, args.steps)
  launches = stack.handle.launches - l0

  def e2e_region(steps):
    for i in range(steps):
      step_e2e(i)
    while pipe.n_res < pipe.n_sub:                                # drain: last results reach the host
      checks.append(float(pipe.result()[0, 0, 1]))

  e2e_region(2)
  ms_e2e = timed(lambda i: e2e_region(args.steps) if i == 0 else None, 1)   # returns once the last logits are on the host
  # same steps once more with every kernel bracketed by CUDA events on its launch stream
  stack.handle.profile_begin()
  timed(step_resident, args.steps)
  kprof = stack.handle.profile_end()
  clocks = sampler.stop()
  kernel_name = stack.handle.last_kernel

  frames_step = B * S * world
  ms_step = ms_total / args.steps
  value = frames_step / (ms_step / 1e3)
  e2e_value = frames_step / (ms_e2e / args.steps / 1e3)

  peaks = {}
  try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
      peaks = json.load(f)
  except Exception:  # pylint: disable=broad-except
    pass
  p_bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
  p_tf32 = p_bf16 / 2.0          # TF32 is not in MEASURED_PEAKS.json: half the measured bf16 rate, "derived"
  p_hbm = peaks.get("hbm_gbs", 6650.0)
  peak_src = "measured" if peaks else "fallback"
  f_uhat, f_route, bytes_frame, weights = algorithmic_work(w)
  n_layers = w["L"]
  frames_rank = B * S
  sm_mhz = clocks.get("sm_mhz") or 1965
  fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
  # the fused roofline: t_roof = max(tensor time of the u_hat contraction, HBM time
  # of the algorithmic bytes); the dense contraction binds for every config of BASELINE.json
  # tensor peak of the mode's MMA kind: kind::f16 runs at the measured bf16/fp16 rate, kind::tf32 at half
  p_tensor = p_bf16 if args.uhat == "f16" else p_tf32
  t_tensor = f_uhat * frames_rank / (p_tensor * 1e12)
  t_hbm = (bytes_frame * frames_rank + weights) / (p_hbm * 1e9)
  t_roof = max(t_tensor, t_hbm)
  kernel_ms = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps}
               for k, v in kprof.items()}
  dom = max(("uhat_gemm", "routing"), key=lambda k: kprof[k][0])
  dom_launches = max(1, kprof[dom][1])
  dom_ms_launch = kprof[dom][0] / dom_launches
  fused = "route_fused_kernel" in kernel_name
  traffic = None
  try:
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
      traffic = json.load(f).get("%s/%s/%s" % (args.workload, args.uhat, "route_fused" if fused else dom))
  except Exception:  # pylint: disable=broad-except
    pass
  # the dominant kernel against the bound of the fused design: algorithmic u_hat FLOPs of the
  # launches' share of the step / measured launch time, peak = derived TF32
  share = dom_launches / args.steps                      # launches of the dominant kernel per step
  achieved = f_uhat * frames_rank / share / (dom_ms_launch / 1e3) / 1e12
  uhat_elems = sum(I * O * D for (I, O, D, d) in shapes_of(w))          # per routing frame, all layers
  roofline = {
      "bound": "tensor" if t_tensor >= t_hbm else "hbm", "achieved": achieved, "peak": p_tensor, "unit": "TFLOP/s",
      "frac": achieved / p_tensor, "traffic": traffic,
      "frac_vs_bf16_peak": achieved / p_bf16, "frac_vs_derived_tf32_peak": achieved / p_tf32,
      "note": "dominant kernel %s (%d launch(es) per step); achieved = algorithmic u_hat FLOPs (2 I O D d per "
              "frame-layer, SURVEY.md 8d) per launch / CUDA-event launch time; peak = %s"
              % ("route_fused_kernel" if fused else dom, round(share),
                 "the measured sustained bf16 GEMM rate (kind::f16 MMA)" if args.uhat == "f16" else
                 "TF32 derived as half the measured sustained bf16 GEMM rate"),
      "peak_source": peak_src + (" (bf16 sustained: the kind::f16 rate)" if args.uhat == "f16" else
                                 " (bf16 sustained / 2: derived TF32)"), "kernel": kernel_name, "dominant": dom,
      "launch_ms": dom_ms_launch, "kernel_ms": kernel_ms,
      "fused_roofline_frac": t_roof / (ms_step / 1e3),
      "hbm_frac_fused_min": (bytes_frame * frames_rank + weights) / (ms_step / 1e3) / 1e9 / p_hbm,
      "fp32_route_frac": f_route * frames_rank / (ms_step / 1e3) / 1e12 / fp32_peak,
  }
  if fused:
    # what actually bounds the fused kernel: every time step re-streams the layer's weights from L2
    # into shared memory (measured bulk-copy ingest 60-70 B/clk/SM = 17-19 TB/s, profiles/r2_ubench.txt)
    groups = (B + 31) // 32
    w_bytes_step = groups * S * packed_weight_bytes(w, args.uhat)
    roofline["l2_weight_stream"] = {"bytes_per_step": w_bytes_step, "achieved_gbs": w_bytes_step / (ms_step / 1e3) / 1e9,
                                    "measured_l2_to_smem_gbs": 17400.0}
  else:
    esize = {"fp32": 0, "tf32": 4, "f16": 4, "bf16": 2, "fp32x3": 4}[args.uhat]
    roofline["materialised_uhat_gbs"] = (uhat_elems * esize * frames_rank / n_layers + bytes_frame * frames_rank / n_layers) / (dom_ms_launch / 1e3) / 1e9

  line = {
      "metric": "routing_frames_per_sec", "value": value, "unit": "routing frames/s",
      "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": {"fp32": "f32", "tf32": "tf32", "f16": "f16", "bf16": "bf16", "fp32x3": "f32"}[args.uhat],
      "data": "synthetic",
      "config": {"workload": w["desc"], "global_batch": B * world, "routing_frames_per_utt": S,
                 "fbank_frames_per_sec": value * 4, "uhat": UHAT_DESC[args.uhat],
                 "parallelism": "dp%d: every GPU routes its own 64 utterances, no collective (weak); the sharded "
                                "64-utterance batch is under `strong`, the training step with the gradient "
                                "all-reduce under `train_cfg4`" % world,
                 "l2": "working set per step (inputs %d MB x2 rotating + packed weights %d MB + inter-layer capsules "
                       "%d MB) > 126 MB L2" % (host_emb[0].numel() * 4 >> 20, weights >> 20,
                                               (B * S * 600 * 4 * (n_layers - 1)) >> 20)},
      "e2e": {"value": e2e_value, "unit": "routing frames/s",
              "h2d_bytes_per_step": host_emb[0].numel() * 4 * world,
              "d2h_bytes_per_step": host_logits.numel() * 4 * world},
      "gpu_launches": launches,
      "clocks": clocks,
      "roofline": roofline,
  }

  if not args.no_also:
    # ---- cfg-3 as BASELINE.json words it: the 64-utterance batch SHARDED over the N GPUs ----
    Bs = max(1, B // world)
    st_s = stack if world == 1 else build_stack(w, dev, args.uhat)
    e_s = dev_emb[0][:Bs].contiguous()
    o_s = torch.empty(Bs, S, w["class_n"], device=dev)
    for _ in range(3):
      st_s.forward(e_s, out_logits=o_s)
    ms_s = timed(lambda i: st_s.forward(e_s, out_logits=o_s), 10) / 10
    line["strong"] = {"scaling": "strong", "workload": w["desc"] + ", sharded: %d utterances per GPU" % Bs,
                      "global_batch": Bs * world, "ms_per_step": ms_s, "value": Bs * world * S / (ms_s / 1e3),
                      "unit": "routing frames/s", "kernel": st_s.handle.last_kernel.split(" ")[0]}
    # ---- cfg-4: training step of the WSJ-shaped stack, global batch 64 sharded, NCCL all-reduce ----
    from srf_b200 import training
    w4 = WORKLOADS["cfg3"]
    B4, S4 = max(1, 64 // world), (w4["T"] + 3) // 4
    # reduced-precision classes: forward in the bench mode (f16: the fused wavefront kernel), the backward
    # recomputes u_hat with bf16 storage (the BPTT sweep streams it: half the bytes of fp32 u_hat)
    fwd4 = args.uhat if args.uhat in ("f16", "tf32") else ("bf16" if args.uhat == "bf16" else args.uhat)
    bwd4 = "bf16" if args.uhat in ("tf32", "f16", "bf16") else args.uhat
    train_mode = fwd4 if fwd4 == bwd4 else "%s forward / %s backward recompute" % (fwd4, bwd4)
    st4 = build_stack(w4, dev, fwd4, bwd_uhat=bwd4)
    tr4 = training.TrainStep(st4, B4 * world)
    g4 = torch.Generator().manual_seed(4 + rank)
    e4 = torch.randn(B4, S4, w4["PH"], w4["DIM"], generator=g4).to(dev)
    lab4 = torch.randint(1, w4["class_n"] - 1, (B4, S4 // 3), generator=g4).to(dev)
    il4 = torch.full((B4,), S4, device=dev)
    ll4 = torch.full((B4,), S4 // 3, device=dev)
    for _ in range(2):
      tr4.step(e4, lab4, il4, ll4)
    ms4 = timed(lambda i: tr4.step(e4, lab4, il4, ll4), 3) / 3
    line["train_cfg4"] = {
        "scaling": "strong",
        "workload": "SRF-SDR WSJ-shaped training step (fwd + CTC loss + bwd + NCCL all-reduce of %d gradient "
                    "floats + Adam), global batch %d x 375 routing frames, %d utterances per GPU"
                    % (tr4.opt.flat.numel(), B4 * world, B4),
        "uhat": train_mode, "ms_per_step": ms4, "value": B4 * world * S4 / (ms4 / 1e3), "unit": "routing frames/s"}
    del st4, tr4
  if world == 1 and not args.no_also:
    # the other single-GPU configurations of BASELINE.json and the other tolerance class on the
    # headline config, same build, short runs (not the headline)
    line["also"] = {}
    for name, mode in [(n, args.uhat) for n in sorted(WORKLOADS) if n != args.workload] + \
                      [(args.workload, "fp32x3" if args.uhat != "fp32x3" else "tf32")]:
      ww = WORKLOADS[name]
      Bw, Sw = ww["B"], (ww["T"] + 3) // 4
      st2 = build_stack(ww, dev, mode)
      e2 = torch.randn(Bw, Sw, ww["PH"], ww["DIM"], device=dev)
      o2 = torch.empty(Bw, Sw, ww["class_n"], device=dev)
      for _ in range(3):
        st2.forward(e2, out_logits=o2)
      ms2 = timed(lambda i: st2.forward(e2, out_logits=o2), 10) / 10
      line["also"]["%s_%s" % (name, mode)] = {
          "workload": ww["desc"], "uhat": mode, "ms_per_step": ms2, "value": Bw * Sw / (ms2 / 1e3),
          "unit": "routing frames/s", "kernel": st2.handle.last_kernel.split(" ")[0],
          "note": "inputs resident" + ("; 1e-4 tolerance class" if mode == "fp32x3" else "")}
      del st2
  if world == 1 and not args.no_also:
    # ---- next-1 row: the native capsulation front-end feeding the hot path (fbank -> logits) ----
    import types
    from srf_b200 import SequenceRouter
    cfg = types.SimpleNamespace(
        model_initializer="fan_avg", model_conv_layer_num=2, feat_dim=123, model_conv_filter_num=64,
        model_encoder_num=w["L"], model_caps_iter=w["iters"], model_caps_window_lpad=w["lpad"],
        model_caps_window_rpad=w["rpad"], model_caps_context=w["sdr"], model_caps_primary_num=w["PH"],
        model_caps_primary_dim=w["DIM"], model_caps_convolution_num=w["CH"], model_caps_convolution_dim=w["DIM"],
        model_caps_class_dim=w["DIM"], train_inp_dropout=0.1, train_inn_dropout=0.1)
    model = SequenceRouter(cfg, None, w["class_n"], device=dev, seed=0, uhat_mode=args.uhat)
    gfe = torch.Generator().manual_seed(7)
    fbank = torch.randn(B, w["T"], 123, generator=gfe).to(dev)
    flen = torch.cat([torch.tensor([w["T"]]), (w["T"] * (0.6 + 0.4 * torch.rand(B - 1, generator=gfe))).long()]).int().to(dev)
    for _ in range(2):
      model(fbank, input_lengths=flen)
    ms_fe = timed(lambda i: model.capsulate(fbank, flen), 5) / 5
    ms_all = timed(lambda i: model(fbank, input_lengths=flen), 5) / 5
    conv_flops = 2.0 * B * (((w["T"] + 1) // 2) * 62 * 9 * 1 * 128 + S * 31 * 9 * 64 * 128) + 2.0 * B * S * 31 * 64 * w["PH"]
    line["also"]["frontend_%s" % args.workload] = {
        "workload": "capsulation front-end (srf_capsulate_fwd: 2 CNN-FE stages with 64 filters, Dense(%d), encaps "
                    "maxout, squash, ln_input) on fbank [%d,%d,123], then the routing stack" % (w["PH"], B, w["T"]),
        "frontend_ms": ms_fe, "fbank_to_logits_ms": ms_all,
        "fbank_frames_per_sec": B * w["T"] / (ms_all / 1e3),
        "frontend_fp32_tflops": conv_flops / (ms_fe / 1e3) / 1e12,
        "note": "FP32 CUDA-core kernels (1e-4 class); inputs resident"}
    del model
  if rank == 0 and not args.no_cpu_baseline and world == 1:
    n_utts, n_frames = cpu_sample_shape(w)
    fps, dt = cpu_reference_run(w, n_utts, n_frames, repeats=2)
    line["cpu_baseline"] = {"value": fps, "unit": "routing frames/s", "cores": os.cpu_count(),
                            "kind": "port",
                            "sample": "%d utterances x %d routing frames of the same model, best of 2 x %.1f s "
                                      "(the sample `--impl reference` times per step)" % (n_utts, n_frames, dt)}
  if rank == 0:
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
