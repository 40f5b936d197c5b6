"""Build the sm_100a shared library in-tree: srf_b200/libsrf_b200.so.

    python -m srf_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsrf_b200.so")
STAMP = os.path.join(HERE, ".libsrf_b200.stamp")
SOURCES = ["capi.cu", "routing_fwd.cu", "uhat_gemm.cu", "routing_stream.cu", "routing_fused.cu",
           "routing_bwd.cu", "ctc_adam.cu", "frontend.cu"]
HEADERS = ["routing_kernels.h", "sm100_ptx.cuh", os.path.join("..", "..", "include", "srf_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
OBJDIR = os.path.join(HERE, "_build")
# developer switches, e.g. SRF_NVCC_EXTRA="-DSRF_STREAM_PHASE_TIMERS -DSRF_BWD_PHASE_TIMERS" for the
# clock64 phase timers read by tools/dev_phase_timers.py / tools/dev_bwd_phase_timers.py
NVCC_FLAGS += os.environ.get("SRF_NVCC_EXTRA", "").split()


def _nvcc() -> str:
  for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
    if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
      return cand
  return "nvcc"


def _digest() -> str:
  h = hashlib.sha256()
  for name in SOURCES + HEADERS:
    with open(os.path.join(CSRC, name), "rb") as f:
      h.update(f.read())
  h.update(" ".join(NVCC_FLAGS).encode())
  return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
  digest = _digest()
  if not force and os.path.exists(LIB) and os.path.exists(STAMP):
    with open(STAMP) as f:
      if f.read().strip() == digest:
        return LIB
  # one nvcc process per translation unit, in parallel, then one link
  from concurrent.futures import ThreadPoolExecutor
  os.makedirs(OBJDIR, exist_ok=True)
  extra = ["-Xptxas", "-v"] if verbose else []

  def compile_one(src):
    obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-c", "-o", obj, os.path.join(CSRC, src)]
    return src, obj, subprocess.run(cmd, capture_output=True, text=True)

  with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
    results = list(pool.map(compile_one, SOURCES))
  for src, _, res in results:
    if verbose or res.returncode != 0:
      sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
      raise RuntimeError("nvcc failed compiling %s (exit %d)" % (src, res.returncode))
  res = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] +
                       [obj for _, obj, _ in results], capture_output=True, text=True)
  if verbose or res.returncode != 0:
    sys.stderr.write(res.stdout + res.stderr)
  if res.returncode != 0:
    raise RuntimeError("nvcc failed linking libsrf_b200.so (exit %d)" % res.returncode)
  with open(STAMP, "w") as f:
    f.write(digest)
  return LIB


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
