"""ctypes binding of include/srf_b200.h.  There is no fallback: a missing or unloadable
libsrf_b200.so raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_uint64,
                    c_void_p)

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsrf_b200.so")
# developer override: another build of the SAME library (e.g. with the clock64 phase timers compiled in)
LIB_PATH = os.environ.get("SRF_B200_LIB", LIB_PATH)

SRF_UHAT_FP32, SRF_UHAT_TF32, SRF_UHAT_BF16, SRF_UHAT_FP32X3, SRF_UHAT_F16 = 0, 1, 2, 3, 4
UHAT_MODES = {"fp32": SRF_UHAT_FP32, "tf32": SRF_UHAT_TF32, "bf16": SRF_UHAT_BF16,
              "fp32x3": SRF_UHAT_FP32X3, "f16": SRF_UHAT_F16}

# every symbol include/srf_b200.h declares (tests check the .so exports all of them)
EXPORTS = ("srf_version", "srf_create", "srf_destroy", "srf_last_error", "srf_route_layer_fwd",
           "srf_route_stack_fwd", "srf_route_layer_bwd", "srf_route_stack_bwd", "srf_ctc_greedy_decode", "srf_ctc_loss",
           "srf_adam_step", "srf_uhat_fwd", "srf_capsulate_fwd", "srf_profile_begin", "srf_profile_end", "srf_launch_count",
           "srf_last_kernel")


class LayerDesc(Structure):
  """struct srf_layer_desc (include/srf_b200.h)."""
  _fields_ = [
      ("emb", c_void_p), ("W", c_void_p), ("bias", c_void_p),
      ("ln_gamma", c_void_p), ("ln_beta", c_void_p), ("dropout_mask", c_void_p),
      ("head_gamma", c_void_p), ("head_beta", c_void_p),
      ("out_caps", c_void_p), ("out_logits", c_void_p), ("out_raw", c_void_p),
      ("B", c_int32), ("S", c_int32), ("H", c_int32), ("d", c_int32),
      ("O", c_int32), ("D", c_int32), ("lpad", c_int32), ("rpad", c_int32),
      ("iters", c_int32), ("sdr", c_int32), ("mask_class0", c_int32), ("uhat_mode", c_int32),
      ("ln_eps", c_float), ("length_eps", c_float), ("weights_version", c_uint64),
  ]


class LayerGrads(Structure):
  """struct srf_layer_grads (include/srf_b200.h)."""
  _fields_ = [(n, c_void_p) for n in ("v_raw", "d_out", "d_logits", "d_raw", "dW", "dbias", "dgamma",
                                      "dbeta", "dhead_gamma", "dhead_beta", "d_emb")]


class FrontendDesc(Structure):
  """struct srf_frontend_desc (include/srf_b200.h)."""
  _fields_ = [
      ("feats", c_void_p), ("lengths", c_void_p),
      ("cnn_kernel", (c_void_p * 2) * 2), ("cnn_bias", (c_void_p * 2) * 2),
      ("bn_gamma", c_void_p * 2), ("bn_beta", c_void_p * 2), ("bn_mean", c_void_p * 2),
      ("bn_var", c_void_p * 2),
      ("dense_kernel", c_void_p), ("dense_bias", c_void_p),
      ("encaps_kernel", c_void_p * 2), ("encaps_bias", c_void_p * 2),
      ("ln_gamma", c_void_p), ("ln_beta", c_void_p),
      ("cnn_dropout", (c_void_p * 2) * 2), ("encaps_dropout", c_void_p * 2), ("inp_dropout", c_void_p),
      ("out_emb", c_void_p),
      ("B", c_int32), ("T", c_int32), ("F", c_int32), ("C", c_int32), ("PH", c_int32), ("PD", c_int32),
      ("training", c_int32), ("pos_enc", c_int32),
      ("bn_eps", c_float), ("bn_momentum", c_float), ("ln_eps", c_float), ("squash_eps", c_float),
  ]


_lib = None


def load() -> ctypes.CDLL:
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise RuntimeError(
        "srf_b200: %s is missing. Build it with `python -m srf_b200.build` (needs nvcc); "
        "there is no CPU or framework fallback for the routing path." % LIB_PATH)
  lib = ctypes.CDLL(LIB_PATH)
  lib.srf_version.restype = c_int
  lib.srf_create.argtypes = [c_int, POINTER(c_void_p)]
  lib.srf_create.restype = c_int
  lib.srf_destroy.argtypes = [c_void_p]
  lib.srf_destroy.restype = c_int
  lib.srf_last_error.argtypes = [c_void_p]
  lib.srf_last_error.restype = c_char_p
  lib.srf_route_layer_fwd.argtypes = [c_void_p, POINTER(LayerDesc), c_void_p]
  lib.srf_route_layer_fwd.restype = c_int
  lib.srf_route_stack_fwd.argtypes = [c_void_p, POINTER(LayerDesc), c_int32, c_void_p]
  lib.srf_route_stack_fwd.restype = c_int
  lib.srf_route_layer_bwd.argtypes = [c_void_p, POINTER(LayerDesc), POINTER(LayerGrads), c_void_p]
  lib.srf_route_layer_bwd.restype = c_int
  lib.srf_route_stack_bwd.argtypes = [c_void_p, POINTER(LayerDesc), POINTER(LayerGrads), c_int32, c_void_p]
  lib.srf_route_stack_bwd.restype = c_int
  lib.srf_ctc_greedy_decode.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                        c_void_p, c_void_p, c_void_p]
  lib.srf_ctc_greedy_decode.restype = c_int
  lib.srf_ctc_loss.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                               c_int32, c_int32, c_float, c_void_p, c_void_p, c_void_p]
  lib.srf_ctc_loss.restype = c_int
  lib.srf_adam_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float,
                                c_float, c_float, c_int64, c_void_p]
  lib.srf_adam_step.restype = c_int
  lib.srf_uhat_fwd.argtypes = [c_void_p, POINTER(LayerDesc), c_void_p, c_void_p]
  lib.srf_uhat_fwd.restype = c_int
  lib.srf_capsulate_fwd.argtypes = [c_void_p, POINTER(FrontendDesc), c_void_p]
  lib.srf_capsulate_fwd.restype = c_int
  lib.srf_profile_begin.argtypes = [c_void_p]
  lib.srf_profile_begin.restype = c_int
  lib.srf_profile_end.argtypes = [c_void_p, POINTER(c_float), POINTER(c_int32)]
  lib.srf_profile_end.restype = c_int
  lib.srf_launch_count.argtypes = [c_void_p]
  lib.srf_launch_count.restype = c_int64
  lib.srf_last_kernel.argtypes = [c_void_p]
  lib.srf_last_kernel.restype = c_char_p
  _lib = lib
  return lib


def check(lib, handle, rc: int, what: str) -> None:
  """Error convention of the C-ABI: negative = invalid argument -> ValueError,
  positive = CUDA/NCCL error -> RuntimeError."""
  if rc == 0:
    return
  msg = lib.srf_last_error(handle)
  msg = msg.decode() if msg else ""
  if rc < 0:
    raise ValueError("%s: %s (code %d)" % (what, msg, rc))
  raise RuntimeError("%s: %s (code %d)" % (what, msg, rc))
