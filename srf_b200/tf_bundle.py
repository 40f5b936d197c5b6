"""TensorFlow-free reader (and writer) for the reference's checkpoint container.

The reference saves `tf.train.Checkpoint(optimizer=optimizer, model=model)` through a
`tf.train.CheckpointManager` (tfsr/helper/misc_helper.py:140-163, tfsr/utils/average_ckpt_sr.py:
137-179, tfsr/trainer_sr.py saves once per epoch).  On disk that is TensorFlow's *tensor bundle*:

    <dir>/checkpoint                        text: model_checkpoint_path: "ckpt-7"
    <dir>/ckpt-7.index                      an SSTable (LevelDB table format): key -> proto
    <dir>/ckpt-7.data-00000-of-00001        the tensors' raw little-endian bytes

  * table: data blocks of prefix-compressed entries (varint shared / non-shared / value length,
    key suffix, value; restart offsets + count at the block end), each block followed by a 1-byte
    compression type and a masked CRC32C; a 48-byte footer (metaindex + index block handles as
    varints, padding, magic 0xdb4775248b80fb57); the index block maps separator keys to block handles.
    BundleWriter writes the index uncompressed (tensor_bundle.cc: options.compression = kNoCompression).
  * key "" -> BundleHeaderProto {1: num_shards, 2: endianness, 3: version};
    key <tensor name> -> BundleEntryProto {1: dtype, 2: TensorShapeProto, 3: shard_id, 4: offset,
    5: size, 6: crc32c (fixed32, masked), 7: slices}.
  * object-based checkpoints name a variable by its attribute path from the root object:
    `model/wgt/0/.ATTRIBUTES/VARIABLE_VALUE`, optimizer slots
    `model/wgt/0/.OPTIMIZER_SLOT/optimizer/m/.ATTRIBUTES/VARIABLE_VALUE`; the key
    `_CHECKPOINTABLE_OBJECT_GRAPH` holds the serialized object graph (a DT_STRING tensor, skipped here:
    the attribute paths of the SequenceRouter classes are fixed by their source, see
    `reference_key_map`).

Only what the routing path's checkpoints contain is supported: unsliced, uncompressed, little-endian
tensors of dtype float32 / float64 / int32 / int64 / bool (strings are listed but not decoded).  The
writer exists so that tests and tools can produce the container without TensorFlow
(tests/golden/make_tf_bundle.py); what it writes follows the same format description.
"""
from __future__ import annotations

import os
import re
import struct
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
FOOTER_LEN = 48
BLOCK_TRAILER = 5            # compression type byte + masked crc32c
HEADER_KEY = b""
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
VAR_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"

# tensorflow/core/framework/types.proto
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_STRING, DT_INT64, DT_BOOL = 1, 2, 3, 7, 9, 10
_NP_OF_DT = {DT_FLOAT: np.dtype("<f4"), DT_DOUBLE: np.dtype("<f8"), DT_INT32: np.dtype("<i4"),
             DT_INT64: np.dtype("<i8"), DT_BOOL: np.dtype("bool")}
_DT_OF_NP = {np.dtype("float32"): DT_FLOAT, np.dtype("float64"): DT_DOUBLE, np.dtype("int32"): DT_INT32,
             np.dtype("int64"): DT_INT64, np.dtype("bool"): DT_BOOL}


# ---------------------------------------------------------------------------------------------
# CRC32C (Castagnoli), masked the way LevelDB / TensorFlow store it
# ---------------------------------------------------------------------------------------------
def _make_crc_table():
  tab = []
  for n in range(256):
    c = n
    for _ in range(8):
      c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
    tab.append(c)
  return np.array(tab, dtype=np.uint32)


_CRC_TABLE = _make_crc_table()


def crc32c(data: bytes, crc: int = 0) -> int:
  c = (~crc) & 0xFFFFFFFF
  tab = _CRC_TABLE
  for byte in data:
    c = int(tab[(c ^ byte) & 0xFF]) ^ (c >> 8)
  return (~c) & 0xFFFFFFFF


def mask_crc(crc: int) -> int:
  return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------
# varints and the few protobuf messages involved
# ---------------------------------------------------------------------------------------------
def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
  shift, val = 0, 0
  while True:
    if pos >= len(buf):
      raise ValueError("truncated varint")
    b = buf[pos]
    pos += 1
    val |= (b & 0x7F) << shift
    if not b & 0x80:
      return val, pos
    shift += 7
    if shift > 70:
      raise ValueError("malformed varint")


def _put_varint(v: int) -> bytes:
  if v < 0:
    v += 1 << 64
  out = bytearray()
  while True:
    b = v & 0x7F
    v >>= 7
    if v:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _parse_proto(buf: bytes) -> Dict[int, list]:
  """field number -> list of raw values (int for varint / fixed, bytes for length-delimited)."""
  out: Dict[int, list] = {}
  pos = 0
  while pos < len(buf):
    tag, pos = _get_varint(buf, pos)
    field, wire = tag >> 3, tag & 7
    if wire == 0:
      val, pos = _get_varint(buf, pos)
    elif wire == 1:
      val = struct.unpack_from("<Q", buf, pos)[0]
      pos += 8
    elif wire == 2:
      n, pos = _get_varint(buf, pos)
      val = bytes(buf[pos:pos + n])
      if len(val) != n:
        raise ValueError("truncated protobuf field")
      pos += n
    elif wire == 5:
      val = struct.unpack_from("<I", buf, pos)[0]
      pos += 4
    else:
      raise ValueError("unsupported protobuf wire type %d" % wire)
    out.setdefault(field, []).append(val)
  return out


def _signed64(v: int) -> int:
  return v - (1 << 64) if v >= 1 << 63 else v


def _parse_shape(buf: bytes) -> Optional[Tuple[int, ...]]:
  msg = _parse_proto(buf)
  if msg.get(3, [0])[0]:
    return None                       # unknown rank
  dims = []
  for d in msg.get(2, []):
    dims.append(_signed64(_parse_proto(d).get(1, [0])[0]))
  return tuple(dims)


def _field(num: int, wire: int, payload: bytes) -> bytes:
  return _put_varint((num << 3) | wire) + payload


def _shape_proto(shape: Iterable[int]) -> bytes:
  out = b""
  for s in shape:
    dim = _field(1, 0, _put_varint(int(s)))
    out += _field(2, 2, _put_varint(len(dim)) + dim)
  return out


# ---------------------------------------------------------------------------------------------
# table (SSTable) reading
# ---------------------------------------------------------------------------------------------
def _read_block(buf: bytes, offset: int, size: int, what: str, verify: bool) -> bytes:
  if offset + size + BLOCK_TRAILER > len(buf):
    raise ValueError("%s block handle (%d, %d) outside the index file" % (what, offset, size))
  data = buf[offset:offset + size]
  ctype = buf[offset + size]
  if ctype != 0:
    raise ValueError("%s block is compressed (type %d); only uncompressed bundle indexes are supported"
                     % (what, ctype))
  if verify:
    want = struct.unpack_from("<I", buf, offset + size + 1)[0]
    got = mask_crc(crc32c(buf[offset:offset + size + 1]))
    if want != got:
      raise ValueError("%s block checksum mismatch (file %08x, computed %08x)" % (what, want, got))
  return data


def _block_entries(block: bytes) -> List[Tuple[bytes, bytes]]:
  if len(block) < 4:
    raise ValueError("table block too small")
  n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
  end = len(block) - 4 - 4 * n_restarts
  if end < 0:
    raise ValueError("bad restart array")
  out, pos, key = [], 0, b""
  while pos < end:
    shared, pos = _get_varint(block, pos)
    non_shared, pos = _get_varint(block, pos)
    vlen, pos = _get_varint(block, pos)
    if shared > len(key) or pos + non_shared + vlen > end:
      raise ValueError("corrupt table entry")
    key = key[:shared] + block[pos:pos + non_shared]
    pos += non_shared
    out.append((key, block[pos:pos + vlen]))
    pos += vlen
  return out


def read_table(path: str, verify: bool = True) -> List[Tuple[bytes, bytes]]:
  """All (key, value) pairs of an SSTable file, in key order."""
  with open(path, "rb") as f:
    buf = f.read()
  if len(buf) < FOOTER_LEN:
    raise ValueError("%s: too small to be a table" % path)
  footer = buf[-FOOTER_LEN:]
  if struct.unpack_from("<Q", footer, FOOTER_LEN - 8)[0] != TABLE_MAGIC:
    raise ValueError("%s: not a TensorFlow checkpoint index (bad table magic)" % path)
  pos = 0
  _, pos = _get_varint(footer, pos)       # metaindex handle (unused by tensor bundles)
  _, pos = _get_varint(footer, pos)
  idx_off, pos = _get_varint(footer, pos)
  idx_size, pos = _get_varint(footer, pos)
  out = []
  for _, handle in _block_entries(_read_block(buf, idx_off, idx_size, "index", verify)):
    off, p2 = _get_varint(handle, 0)
    size, _ = _get_varint(handle, p2)
    out.extend(_block_entries(_read_block(buf, off, size, "data", verify)))
  return out


# ---------------------------------------------------------------------------------------------
# bundle reading
# ---------------------------------------------------------------------------------------------
class BundleEntry:
  __slots__ = ("name", "dtype", "shape", "shard", "offset", "size", "crc", "sliced")

  def __init__(self, name, msg):
    self.name = name
    self.dtype = msg.get(1, [0])[0]
    self.shape = _parse_shape(msg.get(2, [b""])[0])
    self.shard = msg.get(3, [0])[0]
    self.offset = msg.get(4, [0])[0]
    self.size = msg.get(5, [0])[0]
    self.crc = msg.get(6, [None])[0]
    self.sliced = bool(msg.get(7))


class BundleReader:
  """`BundleReader(prefix)` with prefix = ".../ckpt-7".  `keys()`, `entry(name)`, `tensor(name)`."""

  def __init__(self, prefix: str, verify: bool = True):
    self.prefix = prefix
    self.verify = verify
    pairs = read_table(prefix + ".index", verify)
    if not pairs or pairs[0][0] != HEADER_KEY:
      raise ValueError("%s.index: bundle header entry is missing" % prefix)
    hdr = _parse_proto(pairs[0][1])
    self.num_shards = hdr.get(1, [1])[0]
    if hdr.get(2, [0])[0] != 0:
      raise ValueError("big-endian bundles are not supported")
    self.entries: Dict[str, BundleEntry] = {}
    for key, val in pairs[1:]:
      name = key.decode("utf-8")
      self.entries[name] = BundleEntry(name, _parse_proto(val))

  def keys(self) -> List[str]:
    return list(self.entries)

  def entry(self, name: str) -> BundleEntry:
    return self.entries[name]

  def _shard_path(self, shard: int) -> str:
    return "%s.data-%05d-of-%05d" % (self.prefix, shard, self.num_shards)

  def tensor(self, name: str) -> np.ndarray:
    e = self.entries[name]
    if e.sliced:
      raise ValueError("%s: partitioned (sliced) variables are not supported" % name)
    if e.dtype not in _NP_OF_DT:
      raise ValueError("%s: dtype %d is not supported" % (name, e.dtype))
    if e.shape is None:
      raise ValueError("%s: unknown-rank tensor" % name)
    dt = _NP_OF_DT[e.dtype]
    n = int(np.prod(e.shape, dtype=np.int64)) if e.shape else 1
    if n * dt.itemsize != e.size:
      raise ValueError("%s: %d bytes on disk, shape %r needs %d" % (name, e.size, e.shape, n * dt.itemsize))
    with open(self._shard_path(e.shard), "rb") as f:
      f.seek(e.offset)
      raw = f.read(e.size)
    if len(raw) != e.size:
      raise ValueError("%s: data file is truncated" % name)
    if self.verify and e.crc is not None and mask_crc(crc32c(raw)) != e.crc:
      raise ValueError("%s: tensor checksum mismatch" % name)
    return np.frombuffer(raw, dtype=dt).reshape(e.shape).copy()

  def variables(self) -> Dict[str, np.ndarray]:
    """Attribute path -> array for every numeric variable value (`.../.ATTRIBUTES/VARIABLE_VALUE`
    suffix stripped); string tensors (the object graph, save counters' metadata) are skipped."""
    out = {}
    for name, e in self.entries.items():
      if not name.endswith(VAR_SUFFIX) or e.dtype not in _NP_OF_DT:
        continue
      out[name[:-len(VAR_SUFFIX)]] = self.tensor(name)
    return out


def latest_checkpoint(ckpt_dir: str) -> Optional[str]:
  """tf.train.latest_checkpoint: the prefix named by `<dir>/checkpoint` (CheckpointState text proto)."""
  state = os.path.join(ckpt_dir, "checkpoint")
  if not os.path.exists(state):
    return None
  with open(state) as f:
    for line in f:
      m = re.match(r'\s*model_checkpoint_path:\s*"(.*)"', line)
      if m:
        p = m.group(1)
        return p if os.path.isabs(p) else os.path.join(ckpt_dir, p)
  return None


# ---------------------------------------------------------------------------------------------
# attribute paths of the reference's SequenceRouter classes -> this package's parameter names
# ---------------------------------------------------------------------------------------------
def reference_key_map(path: str) -> Optional[str]:
  """`model/...` attribute path of a variable of tfsr.model.sequence_router_{naive,lowmemory,einsum}.
  SequenceRouter (naive:72-107; the three variants use the same attribute names) -> the name used
  by srf_b200 (`named_parameters()`), or None for anything else (optimizer state, save counter)."""
  m = re.fullmatch(r"model/wgt/(\d+)", path)
  if m:
    return "W%s" % m.group(1)
  m = re.fullmatch(r"model/bias/(\d+)", path)
  if m:
    return "b%s" % m.group(1)
  m = re.fullmatch(r"model/ln_m/(\d+)/(gamma|beta)", path)
  if m:
    return "ln_mid%d/%s" % (int(m.group(1)) + 1, m.group(2))      # Keras names them ln_mid1..N
  m = re.fullmatch(r"model/ln_o/(gamma|beta)", path)
  if m:
    return "ln_output/%s" % m.group(1)
  m = re.fullmatch(r"model/ln_i/(gamma|beta)", path)
  if m:
    return "frontend/ln_input_%s" % m.group(1)
  m = re.fullmatch(r"model/conv/conv_layers/(\d+)/(\d+)/(kernel|bias)", path)
  if m:                                                           # conv_layers[p][s], sequence_router.py:52-61
    return "frontend/cnn%s_%s_%s" % m.groups()
  m = re.fullmatch(r"model/conv/bn_layers/(\d+)/(gamma|beta|moving_mean|moving_variance)", path)
  if m:
    return "frontend/bn%s_%s" % (m.group(1), {"moving_mean": "mean", "moving_variance": "var"}.get(m.group(2), m.group(2)))
  m = re.fullmatch(r"model/proj_pe/(kernel|bias)", path)
  if m:
    return "frontend/dense_%s" % m.group(1)
  m = re.fullmatch(r"model/ecs/(\d+)/(kernel|bias)", path)
  if m:
    return "frontend/encaps%s_%s" % m.groups()
  return None


def inverse_key_map(name: str) -> Optional[str]:
  """srf_b200 parameter name -> the reference's attribute path (for writing bundles)."""
  m = re.fullmatch(r"W(\d+)", name)
  if m:
    return "model/wgt/%s" % m.group(1)
  m = re.fullmatch(r"b(\d+)", name)
  if m:
    return "model/bias/%s" % m.group(1)
  m = re.fullmatch(r"ln_mid(\d+)/(gamma|beta)", name)
  if m:
    return "model/ln_m/%d/%s" % (int(m.group(1)) - 1, m.group(2))
  m = re.fullmatch(r"ln_output/(gamma|beta)", name)
  if m:
    return "model/ln_o/%s" % m.group(1)
  m = re.fullmatch(r"frontend/ln_input_(gamma|beta)", name)
  if m:
    return "model/ln_i/%s" % m.group(1)
  m = re.fullmatch(r"frontend/cnn(\d+)_(\d+)_(kernel|bias)", name)
  if m:
    return "model/conv/conv_layers/%s/%s/%s" % m.groups()
  m = re.fullmatch(r"frontend/bn(\d+)_(gamma|beta|mean|var)", name)
  if m:
    return "model/conv/bn_layers/%s/%s" % (m.group(1), {"mean": "moving_mean", "var": "moving_variance"}.get(m.group(2), m.group(2)))
  m = re.fullmatch(r"frontend/dense_(kernel|bias)", name)
  if m:
    return "model/proj_pe/%s" % m.group(1)
  m = re.fullmatch(r"frontend/encaps(\d+)_(kernel|bias)", name)
  if m:
    return "model/ecs/%s/%s" % m.groups()
  return None


def read_reference_checkpoint(prefix: str, verify: bool = True) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
  """-> (state, other): `state` maps srf_b200 parameter names to the arrays of the reference
  checkpoint at `prefix` (variant layouts as stored: checkpoint.load_state_dict reshapes them);
  `other` holds every remaining numeric variable by its attribute path (optimizer/iter, Adam slots
  `model/wgt/0/.OPTIMIZER_SLOT/optimizer/m`, save_counter ...)."""
  r = BundleReader(prefix, verify)
  state, other = {}, {}
  for path, arr in r.variables().items():
    name = reference_key_map(path)
    if name is None:
      other[path] = arr
    else:
      state[name] = arr
  return state, other


# ---------------------------------------------------------------------------------------------
# writer (tests / tools): same container, one data shard, uncompressed index
# ---------------------------------------------------------------------------------------------
def _build_block(entries: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
  out = bytearray()
  restarts = []
  prev = b""
  for n, (key, val) in enumerate(entries):
    shared = 0
    if n % restart_interval == 0:
      restarts.append(len(out))
    else:
      lim = min(len(prev), len(key))
      while shared < lim and prev[shared] == key[shared]:
        shared += 1
    out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(val))
    out += key[shared:] + val
    prev = key
  if not restarts:
    restarts.append(0)
  for r in restarts:
    out += struct.pack("<I", r)
  out += struct.pack("<I", len(restarts))
  return bytes(out)


def write_table(path: str, pairs: List[Tuple[bytes, bytes]], block_size: int = 4096) -> None:
  pairs = sorted(pairs)
  blocks: List[List[Tuple[bytes, bytes]]] = [[]]
  size = 0
  for key, val in pairs:
    if blocks[-1] and size >= block_size:
      blocks.append([])
      size = 0
    blocks[-1].append((key, val))
    size += len(key) + len(val) + 3
  out = bytearray()

  def emit(block: bytes) -> bytes:
    off = len(out)
    out.extend(block)
    out.append(0)                                                  # kNoCompression
    out.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
    return _put_varint(off) + _put_varint(len(block))

  index = []
  for blk in blocks:
    handle = emit(_build_block(blk))
    index.append((blk[-1][0] if blk else b"", handle))             # separator = the block's last key
  meta_handle = emit(_build_block([]))
  index_handle = emit(_build_block(index, restart_interval=1))
  footer = meta_handle + index_handle
  footer += b"\x00" * (FOOTER_LEN - 8 - len(footer))
  footer += struct.pack("<Q", TABLE_MAGIC)
  out.extend(footer)
  with open(path, "wb") as f:
    f.write(bytes(out))


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray], object_graph: bytes = b"") -> None:
  """Write `tensors` (full key -> array) as a one-shard tensor bundle at `prefix`; a non-empty
  `object_graph` is stored under _CHECKPOINTABLE_OBJECT_GRAPH as a scalar DT_STRING tensor the way
  TensorFlow lays strings out (varint length, masked crc32c of the lengths, bytes)."""
  os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
  data = bytearray()
  pairs = [(HEADER_KEY, _field(1, 0, _put_varint(1)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1))))]
  items = sorted(tensors.items())
  if object_graph:
    items.append((OBJECT_GRAPH_KEY, None))
    items.sort(key=lambda kv: kv[0])
  for name, arr in items:
    off = len(data)
    if arr is None:
      lens = _put_varint(len(object_graph))
      raw = lens + struct.pack("<I", mask_crc(crc32c(lens))) + object_graph
      dtype, shape = DT_STRING, ()
      crc = mask_crc(crc32c(object_graph, crc32c(struct.pack("<I", mask_crc(crc32c(lens))), crc32c(lens))))
    else:
      arr = np.asarray(arr)
      if arr.ndim and not arr.flags.c_contiguous:
        arr = np.ascontiguousarray(arr)
      if arr.dtype not in _DT_OF_NP:
        raise ValueError("%s: dtype %s cannot be written" % (name, arr.dtype))
      raw = arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes()
      dtype, shape = _DT_OF_NP[arr.dtype], arr.shape
      crc = mask_crc(crc32c(raw))
    data += raw
    shp = _shape_proto(shape)
    entry = _field(1, 0, _put_varint(dtype)) + _field(2, 2, _put_varint(len(shp)) + shp)
    if off:
      entry += _field(4, 0, _put_varint(off))
    entry += _field(5, 0, _put_varint(len(raw))) + _field(6, 5, struct.pack("<I", crc))
    pairs.append((name.encode("utf-8"), entry))
  with open(prefix + ".data-00000-of-00001", "wb") as f:
    f.write(bytes(data))
  write_table(prefix + ".index", pairs)


def write_reference_checkpoint(ckpt_dir: str, epoch: int, state: Dict[str, np.ndarray],
                               extra: Optional[Dict[str, np.ndarray]] = None) -> str:
  """Write `state` (srf_b200 parameter names; arrays in whatever variant layout they should have on
  disk) as `<ckpt_dir>/ckpt-<epoch>` with the reference's attribute paths, and point the
  `checkpoint` state file at it (what CheckpointManager.save does, misc_helper.py:146-147)."""
  tensors = {}
  for name, arr in state.items():
    path = inverse_key_map(name)
    if path is None:
      raise ValueError("no reference attribute path for parameter %r" % name)
    tensors[path + VAR_SUFFIX] = np.asarray(arr)
  for path, arr in (extra or {}).items():
    tensors[path + VAR_SUFFIX] = np.asarray(arr)
  tensors["save_counter" + VAR_SUFFIX] = np.asarray(epoch, dtype=np.int64)
  prefix = os.path.join(ckpt_dir, "ckpt-%d" % epoch)
  write_bundle(prefix, tensors, object_graph=b"\x0a\x00")
  names = sorted((int(m.group(1)) for m in (re.fullmatch(r"ckpt-(\d+)\.index", f) for f in os.listdir(ckpt_dir)) if m))
  with open(os.path.join(ckpt_dir, "checkpoint"), "w") as f:
    f.write('model_checkpoint_path: "ckpt-%d"\n' % epoch)
    for n in names:
      f.write('all_model_checkpoint_paths: "ckpt-%d"\n' % n)
  return prefix
