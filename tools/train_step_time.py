"""Time one native training step (fwd + CTC + bwd + Adam) of a BASELINE workload on one GPU."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from srf_b200 import RoutingStack, training
name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
w = bench.WORKLOADS[name]
B = int(sys.argv[2]) if len(sys.argv) > 2 else w["B"]
S = int(sys.argv[3]) if len(sys.argv) > 3 else (w["T"] + 3) // 4
stack = RoutingStack(w["L"], w["PH"], w["CH"], w["class_n"], w["DIM"], w["DIM"], w["DIM"], w["lpad"], w["rpad"],
                     w["iters"], w["sdr"], seed=0, inn_dropout=0.1, uhat_mode=os.environ.get("SRF_UHAT", "bf16"))
names = [n for n, _ in stack.named_parameters()]
opt = training.FlatAdam([t for _, t in stack.named_parameters()])
g = torch.Generator().manual_seed(1)
emb = torch.randn(B, S, w["PH"], w["DIM"], generator=g).cuda()
L = max(1, S // 3)
labels = torch.randint(1, w["class_n"] - 1, (B, L), generator=g).cuda()
in_len, lab_len = torch.full((B,), S).cuda(), torch.full((B,), L).cuda()
for it in range(3):
  torch.cuda.synchronize(); t0 = time.perf_counter()
  loss, grads, _ = stack.ctc_train_step_grads(emb, labels, in_len, lab_len)
  torch.cuda.synchronize(); t1 = time.perf_counter()
  flat = torch.cat([(grads[k] / B).reshape(-1) for k in names])
  opt.step(flat, training.warmup_lr(it + 1, 0.5, 256, 1200))
  torch.cuda.synchronize(); t2 = time.perf_counter()
  print("%s B=%d S=%d: fwd+ctc+bwd %.1f ms, adam %.2f ms, loss %.2f" % (name, B, S, (t1 - t0) * 1e3, (t2 - t1) * 1e3, loss.item()), flush=True)
