"""Host side of the routing hot path: tensors in (torch or anything speaking DLPack), C-ABI
call, tensors out.  torch is used only for device memory and streams.

Reference semantics: tfsr/model/sequence_router_naive.py:145-193 (see include/srf_b200.h).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch

from . import _lib

LN_EPS = 1e-3       # Keras LayerNormalization default (naive:104-107)
LENGTH_EPS = 1e-7   # naive:256


def as_device_tensor(x, device: Optional[torch.device] = None) -> torch.Tensor:
  """Borrow `x` as a contiguous float32 CUDA tensor.  Accepts torch tensors and any object
  implementing the DLPack protocol (e.g. tf.experimental.dlpack / EagerTensor.__dlpack__)."""
  if not isinstance(x, torch.Tensor):
    if hasattr(x, "__dlpack__"):
      x = torch.from_dlpack(x)
    else:
      raise TypeError("expected a torch.Tensor or a DLPack-capable tensor, got %r" % type(x))
  if not x.is_cuda:
    if device is None:
      raise ValueError("srf_b200 routing needs CUDA tensors (there is no CPU path)")
    x = x.to(device, non_blocking=True)
  if x.dtype != torch.float32:
    raise ValueError("srf_b200 routing needs float32 tensors, got %s" % x.dtype)
  return x.contiguous()


def _ptr(t: Optional[torch.Tensor]):
  return None if t is None else ctypes.c_void_p(t.data_ptr())


class Handle:
  """Per-device workspace handle (srf_create / srf_destroy)."""

  def __init__(self, device=None):
    self.lib = _lib.load()
    if not torch.cuda.is_available():
      raise RuntimeError("srf_b200: no CUDA device visible; the routing path has no CPU fallback")
    self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
        else torch.device(device)
    h = ctypes.c_void_p()
    rc = self.lib.srf_create(self.device.index or 0, ctypes.byref(h))
    _lib.check(self.lib, None, rc, "srf_create")
    self._h = h

  def close(self):
    if getattr(self, "_h", None):
      self.lib.srf_destroy(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass

  def profile_begin(self):
    _lib.check(self.lib, self._h, self.lib.srf_profile_begin(self._h), "srf_profile_begin")

  def profile_end(self):
    """-> {"pack": (ms, launches), "uhat_gemm": (...), "routing": (...)} since profile_begin."""
    ms = (ctypes.c_float * 3)()
    n = (ctypes.c_int32 * 3)()
    _lib.check(self.lib, self._h, self.lib.srf_profile_end(self._h, ms, n), "srf_profile_end")
    return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(("pack", "uhat_gemm", "routing"))}

  @property
  def launches(self) -> int:
    return int(self.lib.srf_launch_count(self._h))

  @property
  def last_kernel(self) -> str:
    return self.lib.srf_last_kernel(self._h).decode()


_handles = {}


def default_handle(device=None) -> Handle:
  dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
  # torch.device("cuda") carries no index: it means the CURRENT device, not device 0
  key = dev.index if dev.index is not None else torch.cuda.current_device()
  if key not in _handles:
    _handles[key] = Handle(torch.device("cuda", key))
  return _handles[key]


@dataclass
class LayerArgs:
  """Python mirror of srf_layer_desc; tensors are kept alive until the call returns."""
  W: torch.Tensor
  bias: torch.Tensor
  lpad: int
  rpad: int
  iters: int
  sdr: bool
  mask_class0: bool
  ln_gamma: Optional[torch.Tensor] = None
  ln_beta: Optional[torch.Tensor] = None
  dropout_mask: Optional[torch.Tensor] = None
  head_gamma: Optional[torch.Tensor] = None
  head_beta: Optional[torch.Tensor] = None
  uhat_mode: str = "fp32"
  ln_eps: float = LN_EPS
  length_eps: float = LENGTH_EPS
  weights_version: int = 0
  keep: list = field(default_factory=list)


def exact_mode(d: int, O: int, D: int, emb_aligned: bool = True) -> str:
  """The fastest u_hat mode of the 1e-4 parity class for a layer shape: "fp32x3" (3 x TF32 split on
  the tensor cores) where its kernels are instantiated, else "fp32" (fused CUDA-core kernel).
  Mirrors uhat_geometry / route_layer_impl of csrc/capi.cu."""
  if d % 4 != 0 or not emb_aligned or O > 128 or D > 32 or d > 32:
    return "fp32"
  T = 8 if D <= 8 else (16 if D <= 16 else (20 if D <= 20 else 32))
  OPL = 1 if O <= 32 else (2 if O <= 64 else 4)   # the kernels' output capsules per lane
  if (T == 16 and OPL > 2) or (T == 20 and OPL > 2) or (T == 32 and OPL > 1):
    return "fp32"
  return "fp32x3"   # the GEMM keeps as many M tiles of W[i] resident as fit, so every shape fits


def _fill_desc(desc: _lib.LayerDesc, a: LayerArgs, emb, B, S, H, d, out_caps, out_logits):
  I, O, D, d_w = a.W.shape
  window = a.lpad + a.rpad + 1
  if I != window * H or d_w != d:
    raise ValueError("W shape %s does not match window*H=%d, d=%d" % (tuple(a.W.shape), window * H, d))
  if tuple(a.bias.shape) != (I, O, D):
    raise ValueError("bias shape %s != %s" % (tuple(a.bias.shape), (I, O, D)))
  for name, t, n in (("ln_gamma", a.ln_gamma, O * D), ("ln_beta", a.ln_beta, O * D),
                     ("head_gamma", a.head_gamma, O), ("head_beta", a.head_beta, O)):
    if t is not None and t.numel() != n:
      raise ValueError("%s has %d elements, expected %d" % (name, t.numel(), n))
  if a.dropout_mask is not None and a.dropout_mask.numel() != B * S * O * D:
    raise ValueError("dropout_mask has %d elements, expected %d" % (a.dropout_mask.numel(), B * S * O * D))
  mode = a.uhat_mode
  if mode == "exact":   # 1e-4 class, fastest available for this shape
    mode = exact_mode(d, O, D, emb is None or emb.data_ptr() % 16 == 0)
  if mode not in _lib.UHAT_MODES:
    raise ValueError("unknown uhat_mode %r" % (a.uhat_mode,))
  desc.emb = _ptr(emb)
  desc.W, desc.bias = _ptr(a.W), _ptr(a.bias)
  desc.ln_gamma, desc.ln_beta = _ptr(a.ln_gamma), _ptr(a.ln_beta)
  desc.dropout_mask = _ptr(a.dropout_mask)
  desc.head_gamma, desc.head_beta = _ptr(a.head_gamma), _ptr(a.head_beta)
  desc.out_caps, desc.out_logits = _ptr(out_caps), _ptr(out_logits)
  desc.B, desc.S, desc.H, desc.d, desc.O, desc.D = B, S, H, d, O, D
  desc.lpad, desc.rpad, desc.iters = a.lpad, a.rpad, a.iters
  desc.sdr, desc.mask_class0 = int(bool(a.sdr)), int(bool(a.mask_class0))
  desc.uhat_mode = _lib.UHAT_MODES[mode]
  desc.ln_eps, desc.length_eps = a.ln_eps, a.length_eps
  desc.weights_version = a.weights_version
  return O, D


def _prep(a: LayerArgs, dev):
  a.W = as_device_tensor(a.W, dev)
  a.bias = as_device_tensor(a.bias, dev)
  for n in ("ln_gamma", "ln_beta", "dropout_mask", "head_gamma", "head_beta"):
    t = getattr(a, n)
    if t is not None:
      setattr(a, n, as_device_tensor(t, dev))


def route_layer_fwd(emb, args: LayerArgs, handle: Optional[Handle] = None):
  """One routing layer (srf_route_layer_fwd).  emb [B,S,H,d] -> (capsules [B,S,O,D],
  logits [B,S,O] or None)."""
  emb = as_device_tensor(emb)
  if emb.dim() != 4:
    raise ValueError("emb must be [B,S,H,d], got %s" % (tuple(emb.shape),))
  h = handle or default_handle(emb.device)
  _prep(args, emb.device)
  B, S, H, d = emb.shape
  O, D = args.W.shape[1], args.W.shape[2]
  out_caps = torch.empty((B, S, O, D), dtype=torch.float32, device=emb.device)
  out_logits = torch.empty((B, S, O), dtype=torch.float32, device=emb.device) \
      if args.head_gamma is not None else None
  desc = _lib.LayerDesc()
  _fill_desc(desc, args, emb, B, S, H, d, out_caps, out_logits)
  stream = ctypes.c_void_p(torch.cuda.current_stream(emb.device).cuda_stream)
  rc = h.lib.srf_route_layer_fwd(h._h, ctypes.byref(desc), stream)
  _lib.check(h.lib, h._h, rc, "srf_route_layer_fwd")
  return out_caps, out_logits


def route_stack_fwd(emb, layers: Sequence[LayerArgs], handle: Optional[Handle] = None,
                    return_capsules: bool = False, out_logits: Optional[torch.Tensor] = None):
  """The whole routing stack (srf_route_stack_fwd): emb [B,S,PH,PD] -> logits [B,S,class_n].
  The last layer must carry head_gamma/head_beta.  With return_capsules every layer's
  output is also returned (used by the parity tests)."""
  emb = as_device_tensor(emb)
  if emb.dim() != 4:
    raise ValueError("emb must be [B,S,H,d], got %s" % (tuple(emb.shape),))
  if not layers:
    raise ValueError("no layers")
  h = handle or default_handle(emb.device)
  B, S, H, d = emb.shape
  n = len(layers)
  descs = (_lib.LayerDesc * n)()
  caps: List[torch.Tensor] = []
  if layers[-1].head_gamma is None:
    raise ValueError("the last layer must carry the head (head_gamma/head_beta)")
  O_last = layers[-1].W.shape[1]
  if out_logits is None:
    out_logits = torch.empty((B, S, O_last), dtype=torch.float32, device=emb.device)
  for i, a in enumerate(layers):
    _prep(a, emb.device)
    O, D = a.W.shape[1], a.W.shape[2]
    oc = torch.empty((B, S, O, D), dtype=torch.float32, device=emb.device) if return_capsules else None
    if oc is not None:
      caps.append(oc)
    _fill_desc(descs[i], a, emb if i == 0 else None, B, S, H, d, oc,
               out_logits if i == n - 1 else None)
    H, d = O, D
  stream = ctypes.c_void_p(torch.cuda.current_stream(emb.device).cuda_stream)
  rc = h.lib.srf_route_stack_fwd(h._h, descs, n, stream)
  _lib.check(h.lib, h._h, rc, "srf_route_stack_fwd")
  if return_capsules:
    return out_logits, caps
  return out_logits


def route_stack_fwd_train(emb, layers: Sequence[LayerArgs], handle: Optional[Handle] = None):
  """Training forward of the whole stack in ONE library call (srf_route_stack_fwd with out_caps and
  out_raw for every layer: the fused layer-wavefront kernel when the shape and mode allow it).
  -> (logits, [layer outputs], [pre-LayerNorm capsules]) -- what srf_route_stack_bwd needs saved."""
  emb = as_device_tensor(emb)
  if emb.dim() != 4:
    raise ValueError("emb must be [B,S,H,d], got %s" % (tuple(emb.shape),))
  if not layers or layers[-1].head_gamma is None:
    raise ValueError("the last layer must carry the head (head_gamma/head_beta)")
  h = handle or default_handle(emb.device)
  B, S, H, d = emb.shape
  n = len(layers)
  descs = (_lib.LayerDesc * n)()
  caps, raws = [], []
  logits = torch.empty((B, S, layers[-1].W.shape[1]), dtype=torch.float32, device=emb.device)
  for i, a in enumerate(layers):
    _prep(a, emb.device)
    O, D = a.W.shape[1], a.W.shape[2]
    caps.append(torch.empty((B, S, O, D), dtype=torch.float32, device=emb.device))
    raws.append(torch.empty((B, S, O, D), dtype=torch.float32, device=emb.device))
    _fill_desc(descs[i], a, emb if i == 0 else None, B, S, H, d, caps[i], logits if i == n - 1 else None)
    descs[i].out_raw = _ptr(raws[i])
    H, d = O, D
  stream = ctypes.c_void_p(torch.cuda.current_stream(emb.device).cuda_stream)
  rc = h.lib.srf_route_stack_fwd(h._h, descs, n, stream)
  _lib.check(h.lib, h._h, rc, "srf_route_stack_fwd")
  return logits, caps, raws


def route_stack_bwd(emb, layers: Sequence[LayerArgs], caps, raws, d_logits, need_d_emb: bool = True,
                    handle: Optional[Handle] = None):
  """Backward of the whole stack in ONE library call (srf_route_stack_bwd).  `caps` / `raws` are what
  route_stack_fwd_train (or the per-layer training forwards) returned.  -> list of per-layer gradient
  dicts (keys as route_layer_bwd; "d_emb" of layer 0 is the gradient w.r.t. `emb`)."""
  emb = as_device_tensor(emb)
  h = handle or default_handle(emb.device)
  dev = emb.device
  B, S, H, d = emb.shape
  n = len(layers)
  descs = (_lib.LayerDesc * n)()
  grs = (_lib.LayerGrads * n)()
  z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
  out, keep = [], []
  d_logits = as_device_tensor(d_logits, dev)
  for i, a in enumerate(layers):
    _prep(a, dev)
    I, O, D, _ = a.W.shape
    last = i == n - 1
    _fill_desc(descs[i], a, emb if i == 0 else None, B, S, H, d, caps[i], None)
    g = {"dW": z(I, O, D, d), "dbias": z(I, O, D),
         "dgamma": z(O * D) if a.ln_gamma is not None else None,
         "dbeta": z(O * D) if a.ln_gamma is not None else None,
         "dhead_gamma": z(O) if last else None, "dhead_beta": z(O) if last else None,
         "d_emb": z(B, S, H, d) if (need_d_emb or i > 0) else None}
    d_raw = torch.empty((B, S, O, D), dtype=torch.float32, device=dev)
    keep.append(d_raw)
    gr = grs[i]
    gr.v_raw, gr.d_raw = _ptr(as_device_tensor(raws[i], dev)), _ptr(d_raw)
    gr.d_logits = _ptr(d_logits) if last else None
    gr.dW, gr.dbias, gr.dgamma, gr.dbeta = _ptr(g["dW"]), _ptr(g["dbias"]), _ptr(g["dgamma"]), _ptr(g["dbeta"])
    gr.dhead_gamma, gr.dhead_beta, gr.d_emb = _ptr(g["dhead_gamma"]), _ptr(g["dhead_beta"]), _ptr(g["d_emb"])
    out.append(g)
    H, d = O, D
  stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
  rc = h.lib.srf_route_stack_bwd(h._h, descs, grs, n, stream)
  _lib.check(h.lib, h._h, rc, "srf_route_stack_bwd")
  return out


def uhat_fwd(emb, W, bias, lpad: int, rpad: int, uhat_mode: str = "tf32",
             handle: Optional[Handle] = None) -> torch.Tensor:
  """Prediction vectors alone (srf_uhat_fwd, naive:150-159): [B,S,H,d] -> [B,S,I,O,D] fp32,
  computed by the tcgen05 GEMM (tf32: fp32 u_hat storage, bf16: bf16 storage)."""
  emb = as_device_tensor(emb)
  h = handle or default_handle(emb.device)
  a = LayerArgs(W=W, bias=bias, lpad=lpad, rpad=rpad, iters=1, sdr=False, mask_class0=False,
                uhat_mode=uhat_mode)
  _prep(a, emb.device)
  B, S, H, d = emb.shape
  I, O, D, _ = a.W.shape
  out = torch.empty((B, S, I, O, D), dtype=torch.float32, device=emb.device)
  desc = _lib.LayerDesc()
  _fill_desc(desc, a, emb, B, S, H, d, None, None)
  stream = ctypes.c_void_p(torch.cuda.current_stream(emb.device).cuda_stream)
  rc = h.lib.srf_uhat_fwd(h._h, ctypes.byref(desc), ctypes.c_void_p(out.data_ptr()), stream)
  _lib.check(h.lib, h._h, rc, "srf_uhat_fwd")
  return out


def route_layer_fwd_train(emb, args: LayerArgs, handle: Optional[Handle] = None):
  """Forward of one layer that also returns the pre-LayerNorm capsules the backward needs:
  -> (capsules [B,S,O,D], logits or None, v_raw [B,S,O,D])."""
  emb = as_device_tensor(emb)
  h = handle or default_handle(emb.device)
  _prep(args, emb.device)
  B, S, H, d = emb.shape
  O, D = args.W.shape[1], args.W.shape[2]
  out_caps = torch.empty((B, S, O, D), dtype=torch.float32, device=emb.device)
  out_raw = torch.empty((B, S, O, D), dtype=torch.float32, device=emb.device)
  out_logits = torch.empty((B, S, O), dtype=torch.float32, device=emb.device) \
      if args.head_gamma is not None else None
  desc = _lib.LayerDesc()
  _fill_desc(desc, args, emb, B, S, H, d, out_caps, out_logits)
  desc.out_raw = _ptr(out_raw)
  stream = ctypes.c_void_p(torch.cuda.current_stream(emb.device).cuda_stream)
  rc = h.lib.srf_route_layer_fwd(h._h, ctypes.byref(desc), stream)
  _lib.check(h.lib, h._h, rc, "srf_route_layer_fwd")
  return out_caps, out_logits, out_raw


def route_layer_bwd(emb, args: LayerArgs, v_raw, d_out=None, d_logits=None, need_d_emb=True,
                    handle: Optional[Handle] = None, out: Optional[dict] = None):
  """Backward of one layer (srf_route_layer_bwd).  Returns a dict with dW, dbias, dgamma, dbeta,
  dhead_gamma, dhead_beta (None where the layer has no such parameter) and d_emb.  `out`: already
  ZEROED float32 device tensors to accumulate into instead of fresh ones (any subset of the keys;
  e.g. views into a flat gradient buffer)."""
  emb = as_device_tensor(emb)
  h = handle or default_handle(emb.device)
  _prep(args, emb.device)
  dev = emb.device
  B, S, H, d = emb.shape
  I, O, D, _ = args.W.shape
  z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
  g = {"dW": z(I, O, D, d), "dbias": z(I, O, D),
       "dgamma": z(O * D) if args.ln_gamma is not None else None,
       "dbeta": z(O * D) if args.ln_gamma is not None else None,
       "dhead_gamma": z(O) if d_logits is not None else None,
       "dhead_beta": z(O) if d_logits is not None else None,
       "d_emb": z(B, S, H, d) if need_d_emb else None} if out is None else None
  if out is not None:
    want = {"dW": ((I, O, D, d), True), "dbias": ((I, O, D), True),
            "dgamma": ((O * D,), args.ln_gamma is not None), "dbeta": ((O * D,), args.ln_gamma is not None),
            "dhead_gamma": ((O,), d_logits is not None), "dhead_beta": ((O,), d_logits is not None),
            "d_emb": ((B, S, H, d), need_d_emb)}
    g = {}
    for k, (shape, needed) in want.items():
      t = out.get(k)
      if not needed:
        g[k] = None
        continue
      if t is None:
        t = z(*shape)
      n = 1
      for v in shape:
        n *= v
      if t.numel() != n or t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
        raise ValueError("out[%r] must be a contiguous float32 tensor of %d elements on %s" % (k, n, dev))
      g[k] = t
  d_raw = torch.empty((B, S, O, D), dtype=torch.float32, device=dev)
  v_raw = as_device_tensor(v_raw, dev)
  d_out = None if d_out is None else as_device_tensor(d_out, dev)
  d_logits = None if d_logits is None else as_device_tensor(d_logits, dev)
  desc = _lib.LayerDesc()
  _fill_desc(desc, args, emb, B, S, H, d, None, None)
  gr = _lib.LayerGrads()
  gr.v_raw, gr.d_out, gr.d_logits, gr.d_raw = _ptr(v_raw), _ptr(d_out), _ptr(d_logits), _ptr(d_raw)
  gr.dW, gr.dbias, gr.dgamma, gr.dbeta = _ptr(g["dW"]), _ptr(g["dbias"]), _ptr(g["dgamma"]), _ptr(g["dbeta"])
  gr.dhead_gamma, gr.dhead_beta, gr.d_emb = _ptr(g["dhead_gamma"]), _ptr(g["dhead_beta"]), _ptr(g["d_emb"])
  stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
  rc = h.lib.srf_route_layer_bwd(h._h, ctypes.byref(desc), ctypes.byref(gr), stream)
  _lib.check(h.lib, h._h, rc, "srf_route_layer_bwd")
  return g


def capsulate_fwd(feats, lengths, fe: dict, C: int, PH: int, PD: int, training: bool = False,
                  dropout: Optional[dict] = None, pos_enc: bool = False,
                  handle: Optional[Handle] = None) -> torch.Tensor:
  """Capsulation front-end, forward (srf_capsulate_fwd; naive:129-142, sequence_router.py:44-82):
  fbank [B,T,F] + lengths [B] -> primary capsules emb [B,S,PH,PD].  `fe`: front-end parameters by
  the names of tests/golden (TF layouts, device tensors).  training=True: BatchNormalization with
  batch statistics (fe["bn*_mean"/"bn*_var"] are updated in place) and the keep masks of
  `dropout` ({"cnn<path>_<stage>", "encaps<path>", "inp"} -> already scaled mask) are applied."""
  feats = as_device_tensor(feats)
  if feats.dim() != 3:
    raise ValueError("feats must be [B,T,F], got %s" % (tuple(feats.shape),))
  dev = feats.device
  h = handle or default_handle(dev)
  B, T, F = feats.shape
  S = -(-(-(-T // 2)) // 2)
  lens = torch.as_tensor(lengths).to(device=dev, dtype=torch.int32).contiguous()
  if lens.numel() != B:
    raise ValueError("input_lengths has %d entries for %d utterances" % (lens.numel(), B))
  keep = [feats, lens]

  def par(name, numel=None):
    if name not in fe:
      raise ValueError("front-end parameter %r is missing" % name)
    t = as_device_tensor(fe[name], dev)
    if numel is not None and t.numel() != numel:
      raise ValueError("front-end parameter %s has %d elements, expected %d" % (name, t.numel(), numel))
    keep.append(t)
    return t.data_ptr()

  F1 = -(-F // 2)
  Fq = -(-F1 // 2)
  d = _lib.FrontendDesc()
  d.feats, d.lengths = feats.data_ptr(), lens.data_ptr()
  for p_ in range(2):
    for st in range(2):
      cin = 1 if st == 0 else C
      d.cnn_kernel[p_][st] = par("cnn%d_%d_kernel" % (p_, st), 9 * cin * C)
      d.cnn_bias[p_][st] = par("cnn%d_%d_bias" % (p_, st), C)
    d.encaps_kernel[p_] = par("encaps%d_kernel" % p_, 9 * PD)
    d.encaps_bias[p_] = par("encaps%d_bias" % p_, PD)
  for st in range(2):
    d.bn_gamma[st], d.bn_beta[st] = par("bn%d_gamma" % st, C), par("bn%d_beta" % st, C)
    if training and not fe["bn%d_mean" % st].is_contiguous():
      raise ValueError("BatchNormalization moving statistics must be contiguous (updated in place)")
    d.bn_mean[st], d.bn_var[st] = par("bn%d_mean" % st, C), par("bn%d_var" % st, C)
  d.dense_kernel, d.dense_bias = par("dense_kernel", Fq * C * PH), par("dense_bias", PH)
  d.ln_gamma, d.ln_beta = par("ln_input_gamma", PH * PD), par("ln_input_beta", PH * PD)
  if training and dropout:
    T1 = -(-T // 2)
    shp = {0: B * T1 * F1 * C, 1: B * S * Fq * C}
    for key, m in dropout.items():
      m = as_device_tensor(m, dev)
      keep.append(m)
      if key == "inp":
        want = B * S * PH * PD
        d.inp_dropout = m.data_ptr()
      elif key.startswith("encaps"):
        want = B * S * PH * PD
        d.encaps_dropout[int(key[6:])] = m.data_ptr()
      elif key.startswith("cnn"):
        p_, st = int(key[3]), int(key[5])
        want = shp[st]
        d.cnn_dropout[p_][st] = m.data_ptr()
      else:
        raise ValueError("unknown dropout mask %r" % key)
      if m.numel() != want:
        raise ValueError("dropout mask %s has %d elements, expected %d" % (key, m.numel(), want))
  emb = torch.empty((B, S, PH, PD), dtype=torch.float32, device=dev)
  d.out_emb = emb.data_ptr()
  d.B, d.T, d.F, d.C, d.PH, d.PD = B, T, F, C, PH, PD
  d.training, d.pos_enc = int(bool(training)), int(bool(pos_enc))
  d.bn_eps, d.bn_momentum, d.ln_eps, d.squash_eps = 1e-3, 0.99, LN_EPS, 1e-7
  stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
  rc = h.lib.srf_capsulate_fwd(h._h, ctypes.byref(d), stream)
  _lib.check(h.lib, h._h, rc, "srf_capsulate_fwd")
  return emb
