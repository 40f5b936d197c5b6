"""Weights interchange with the reference's variable layouts, checkpoints and checkpoint averaging
(SURVEY.md 8f next-4).  Host-side only: numpy dictionaries `name -> array`; the device tensors of
a RoutingStack / SequenceRouter are filled through `named_parameters()`.

Reference behaviour mirrored here:
  * routing variables of the three model variants are reshapes of one canonical layout
    W [I,O,D,d], bias [I,O,D] (names `W%d` / `b%d`):
      naive      W (1,1,I,O,D,d)  bias (1,1,I,O,D,1)   sequence_router_naive.py:88-103
      lowmemory  W (1,I,O,D,d)    bias (1,I,O,D,1)     sequence_router_lowmemory.py:85-101
      einsum     W (I,O,D,d)      bias (1,1,I,O,D)     sequence_router_einsum.py:82-97
  * `load_checkpoint` (tfsr/helper/misc_helper.py:140-163): restore `ckpt-<epoch>` when
    `path_ckpt_epoch > 0`, else the latest one, return the epoch offset (0 when nothing was loaded);
    `max_to_keep < 0` keeps everything.
  * checkpoint averaging (tfsr/utils/average_ckpt_sr.py:100-179): element-wise mean of the
    weights of the last `model_average_num` checkpoints, written to `<path_ckpt>/avg`.
Two containers: `.npz` files with the same variable names (`ckpt-<epoch>.npz`), and TensorFlow's own
tensor-bundle container (`ckpt-<epoch>.index` + `.data-00000-of-00001`, object-based variable names
`model/wgt/0/.ATTRIBUTES/VARIABLE_VALUE` ...) read and written without TensorFlow by
`srf_b200.tf_bundle` -- so checkpoints trained with the reference load here, and checkpoints saved
with `fmt="tf"` restore in the reference (`ckpt.restore(path).expect_partial()`).
"""
from __future__ import annotations

import glob
import os
import re
import shutil
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

VARIANTS = ("naive", "lowmemory", "einsum")


def _canonical_dims(W: np.ndarray) -> Tuple[int, int, int, int]:
  shape = tuple(int(s) for s in W.shape)
  while len(shape) > 4 and shape[0] == 1:
    shape = shape[1:]
  if len(shape) != 4:
    raise ValueError("routing weight of shape %r is not a reshape of [I,O,D,d]" % (tuple(W.shape),))
  return shape  # type: ignore[return-value]


def to_canonical(W, bias) -> Tuple[np.ndarray, np.ndarray]:
  """Any variant's (W, bias) -> canonical ([I,O,D,d], [I,O,D]); pure reshapes, no transposition."""
  W = np.asarray(W, dtype=np.float32)
  bias = np.asarray(bias, dtype=np.float32)
  I, O, D, d = _canonical_dims(W)
  if bias.size != I * O * D:
    raise ValueError("bias of shape %r does not match weight %r" % (tuple(bias.shape), tuple(W.shape)))
  return W.reshape(I, O, D, d), bias.reshape(I, O, D)


def from_canonical(W, bias, variant: str) -> Tuple[np.ndarray, np.ndarray]:
  """Canonical ([I,O,D,d], [I,O,D]) -> the variable shapes of a reference model variant."""
  W = np.asarray(W, dtype=np.float32)
  bias = np.asarray(bias, dtype=np.float32)
  if W.ndim != 4 or bias.shape != W.shape[:3]:
    raise ValueError("expected canonical W [I,O,D,d] and bias [I,O,D]")
  I, O, D, d = W.shape
  if variant == "naive":
    return W.reshape(1, 1, I, O, D, d), bias.reshape(1, 1, I, O, D, 1)
  if variant == "lowmemory":
    return W.reshape(1, I, O, D, d), bias.reshape(1, I, O, D, 1)
  if variant == "einsum":
    return W.copy(), bias.reshape(1, 1, I, O, D)
  raise ValueError("unknown variant %r (expected one of %s)" % (variant, ", ".join(VARIANTS)))


def state_dict(model) -> Dict[str, np.ndarray]:
  """name -> numpy copy of every parameter of a RoutingStack / SequenceRouter."""
  return {name: t.detach().cpu().numpy().copy() for name, t in model.named_parameters()}


def load_state_dict(model, state: Dict[str, np.ndarray], variant: Optional[str] = None,
                    strict: bool = True) -> List[str]:
  """Copy `state` into the model's device tensors.  Routing weights (`W%d`, `b%d`, possibly with a
  `.../` prefix) may come in any reference variant layout; they are reshaped to canonical.
  Returns the names that were loaded; `strict` raises on missing names (the reference's
  `expect_partial()` corresponds to strict=False)."""
  import torch
  loaded = []
  # a freshly built SequenceRouter creates its front-end parameters lazily (first call, like the
  # Keras layers of the reference); tf.train.Checkpoint defers the restore until then
  # (misc_helper.py:140-163) -- here the checkpoint's front-end arrays create them right away
  fe_state = {k[len("frontend/"):]: v for k, v in state.items() if k.startswith("frontend/")}
  if fe_state and hasattr(model, "load_frontend") and not getattr(model, "fe", None):
    model.load_frontend(fe_state)
    loaded.extend("frontend/" + k for k in fe_state)
  named = dict(model.named_parameters())
  pending_w, pending_b = {}, {}
  for name, arr in state.items():
    base = name.rsplit("/", 1)[-1]
    m = re.fullmatch(r"([Wb])(\d+)", base)
    if m and base in named:
      (pending_w if m.group(1) == "W" else pending_b)[int(m.group(2))] = np.asarray(arr)
  for idx, W in pending_w.items():
    if idx not in pending_b:
      raise ValueError("checkpoint has W%d but no b%d" % (idx, idx))
    Wc, bc = to_canonical(W, pending_b[idx])
    for nm, val in (("W%d" % idx, Wc), ("b%d" % idx, bc)):
      t = named[nm]
      if tuple(t.shape) != val.shape:
        raise ValueError("%s: checkpoint shape %r, model shape %r" % (nm, val.shape, tuple(t.shape)))
      with torch.no_grad():
        t.copy_(torch.from_numpy(val))
      loaded.append(nm)
  for name, arr in state.items():
    if name in named and name not in loaded:
      t = named[name]
      val = np.asarray(arr, dtype=np.float32)
      if val.size != t.numel():
        raise ValueError("%s: checkpoint has %d elements, model %d" % (name, val.size, t.numel()))
      with torch.no_grad():
        t.copy_(torch.from_numpy(val.reshape(tuple(t.shape))))
      loaded.append(name)
  if strict:
    missing = sorted(set(named) - set(loaded))
    if missing:
      raise KeyError("checkpoint is missing %s" % ", ".join(missing))
    consumed = set(loaded)
    for name in state:
      base = name.rsplit("/", 1)[-1]
      if name not in consumed and base not in consumed:
        raise KeyError("checkpoint entry %s has no counterpart in the model" % name)
  if hasattr(model, "mark_weights_changed"):
    model.mark_weights_changed()
  elif hasattr(model, "stack"):
    model.stack.mark_weights_changed()
  return loaded


# -- checkpoint files (tf.train.CheckpointManager semantics on .npz) ---------------------------
def _ckpt_path(path_ckpt: str, epoch: int) -> str:
  return os.path.join(path_ckpt, "ckpt-%d.npz" % epoch)


def list_checkpoints(path_ckpt: str) -> List[Tuple[int, str]]:
  """(epoch, path) sorted by epoch; `.npz` files and TensorFlow bundles (`ckpt-<n>.index`, returned
  as the bundle prefix `.../ckpt-<n>`).  When both exist for an epoch the `.npz` wins."""
  found = {}
  for p in glob.glob(os.path.join(path_ckpt, "ckpt-*.index")):
    m = re.fullmatch(r"ckpt-(\d+)\.index", os.path.basename(p))
    if m:
      found[int(m.group(1))] = p[:-len(".index")]
  for p in glob.glob(os.path.join(path_ckpt, "ckpt-*.npz")):
    m = re.fullmatch(r"ckpt-(\d+)\.npz", os.path.basename(p))
    if m:
      found[int(m.group(1))] = p
  return sorted(found.items())


def _remove_checkpoint(path: str) -> None:
  if path.endswith(".npz"):
    os.remove(path)
    return
  for p in glob.glob(path + ".index") + glob.glob(path + ".data-*"):
    os.remove(p)


def latest_checkpoint(path_ckpt: str) -> Optional[str]:
  c = list_checkpoints(path_ckpt)
  return c[-1][1] if c else None


def save_checkpoint(state: Dict[str, np.ndarray], path_ckpt: str, epoch: int,
                    max_to_keep: int = -1, variant: Optional[str] = None, fmt: str = "npz") -> str:
  """Write `ckpt-<epoch>.npz` (fmt="npz") or a TensorFlow tensor bundle `ckpt-<epoch>.index/.data-*`
  with the reference's object-based variable names (fmt="tf"; `variant` defaults to "naive" there);
  keep the newest `max_to_keep` checkpoints (< 0: keep all, misc_helper.py:143-145).  With `variant`,
  routing weights are stored in that reference layout."""
  if fmt not in ("npz", "tf"):
    raise ValueError("fmt must be 'npz' or 'tf'")
  if fmt == "tf" and variant is None:
    variant = "naive"
  os.makedirs(path_ckpt, exist_ok=True)
  out = {}
  for name, arr in state.items():
    out[name] = np.asarray(arr, dtype=np.float32)
  if variant is not None:
    idx = sorted(int(n[1:]) for n in out if re.fullmatch(r"W\d+", n))
    for i in idx:
      out["W%d" % i], out["b%d" % i] = from_canonical(*to_canonical(out["W%d" % i], out["b%d" % i]), variant)
  if fmt == "tf":
    from . import tf_bundle
    path = tf_bundle.write_reference_checkpoint(path_ckpt, epoch, out)
  else:
    path = _ckpt_path(path_ckpt, epoch)
    tmp = path + ".tmp.npz"
    np.savez(tmp, **out)
    os.replace(tmp, path)
  if max_to_keep is not None and max_to_keep >= 0:
    for _, p in list_checkpoints(path_ckpt)[:-max_to_keep or None]:
      if p != path:
        _remove_checkpoint(p)
  return path


def read_checkpoint(path: str) -> Dict[str, np.ndarray]:
  """`.npz` file or TensorFlow bundle prefix (`.../ckpt-7`) -> name -> array, srf_b200 names."""
  if path.endswith(".npz"):
    with np.load(path) as z:
      return {k: z[k] for k in z.files}
  from . import tf_bundle
  if path.endswith(".index"):
    path = path[:-len(".index")]
  state, _ = tf_bundle.read_reference_checkpoint(path)
  if not state:
    raise ValueError("%s holds no variable of a SequenceRouter model" % path)
  return state


def load_checkpoint(model, path_ckpt: str, path_ckpt_epoch: Optional[int] = None, logger=None,
                    strict: bool = False) -> int:
  """misc_helper.py:140-163: restore the requested epoch (> 0) or the latest checkpoint into
  `model`; returns the epoch offset (0 and nothing loaded when there is no checkpoint)."""
  loaded = None
  if path_ckpt_epoch is not None and path_ckpt_epoch > 0:
    loaded = dict(list_checkpoints(path_ckpt)).get(int(path_ckpt_epoch))
    if loaded is None:
      raise FileNotFoundError(os.path.join(path_ckpt, "ckpt-%d" % path_ckpt_epoch))
  else:
    loaded = latest_checkpoint(path_ckpt)
  epoch_offset = 0
  if loaded is not None:
    epoch_offset = int(re.search(r"ckpt-(\d+)(\.npz)?$", loaded).group(1))
    load_state_dict(model, read_checkpoint(loaded), strict=strict)
  if logger is not None:
    logger.info("Loaded ckpt: %s", loaded)
  return epoch_offset


def canonicalize_state(state: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
  """Routing variables (`W%d` / `b%d`, any variant layout) -> canonical shapes; the rest unchanged."""
  out = {k: np.asarray(v) for k, v in state.items()}
  for name in list(out):
    m = re.fullmatch(r"(.*/)?W(\d+)", name)
    if m:
      bname = (m.group(1) or "") + "b" + m.group(2)
      if bname in out:
        out[name], out[bname] = to_canonical(out[name], out[bname])
  return out


def average_states(states: Sequence[Dict[str, np.ndarray]]) -> Dict[str, np.ndarray]:
  """Element-wise mean of the same-named arrays (average_ckpt_sr.py:137-146); routing variables
  are brought to the canonical layout first, so checkpoints of different variants can be mixed."""
  if not states:
    raise ValueError("no checkpoints to average")
  states = [canonicalize_state(s) for s in states]
  names = list(states[0])
  for s in states[1:]:
    if set(s) != set(names):
      raise ValueError("checkpoints hold different variables")
  out = {}
  for n in names:
    stack = np.stack([np.asarray(s[n], dtype=np.float64) for s in states])
    out[n] = stack.mean(axis=0).astype(np.float32)
  return out


def average_checkpoints(path_ckpt: str, model_average_num: int, logger=None, fmt: str = "npz") -> str:
  """average_ckpt_sr.py:100-179: mean of the last `model_average_num` checkpoints, saved as the
  only checkpoint of `<path_ckpt>/avg` (that directory is recreated)."""
  ckpts = list_checkpoints(path_ckpt)[-model_average_num:]
  if not ckpts:
    raise FileNotFoundError("no ckpt-* checkpoint under %s" % path_ckpt)
  for _, p in ckpts:
    if logger is not None:
      logger.info(p)
  avg = average_states([read_checkpoint(p) for _, p in ckpts])
  if logger is not None:
    logger.info("Total %d models were loaded.", len(ckpts))
  avg_dir = os.path.join(path_ckpt, "avg")
  if os.path.exists(avg_dir):
    shutil.rmtree(avg_dir)
  out = save_checkpoint(avg, avg_dir, 1, max_to_keep=1, fmt=fmt)
  if logger is not None:
    logger.info("Saved to %s", out)
  return out
