"""The C-ABI library builds for sm_100a, loads, and exports every symbol the header
declares.  No compute calls (there is no GPU on the CPU tier)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
  with open(os.path.join(ROOT, "include", "srf_b200.h")) as f:
    text = f.read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(srf_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(built_lib):
  lib = ctypes.CDLL(built_lib)
  names = _declared_symbols()
  assert "srf_route_layer_fwd" in names and "srf_route_stack_fwd" in names
  for n in names:
    assert hasattr(lib, n), "libsrf_b200.so does not export %s" % n


def test_python_binding_lists_every_symbol(built_lib):
  from srf_b200 import _lib
  assert sorted(_lib.EXPORTS) == _declared_symbols()
  lib = _lib.load()
  assert lib.srf_version() == 210


def test_layer_desc_matches_header_layout():
  """ctypes mirror and the C struct must agree field by field."""
  from srf_b200 import _lib
  with open(os.path.join(ROOT, "include", "srf_b200.h")) as f:
    text = f.read()
  body = re.search(r"typedef struct srf_layer_desc \{(.*?)\} srf_layer_desc;", text, re.S).group(1)
  body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
  fields = []
  for decl in body.split(";"):
    decl = decl.strip()
    if not decl:
      continue
    names = decl.replace("*", " ").split()
    first = decl.split(",")[0].replace("*", " ").split()[-1]
    fields.append(first)
    for extra in decl.split(",")[1:]:
      fields.append(extra.replace("*", " ").strip())
  assert fields == [n for n, _ in _lib.LayerDesc._fields_]
  # 11 pointers + 12 int32 + 2 float + 1 uint64
  assert ctypes.sizeof(_lib.LayerDesc) == 11 * 8 + 12 * 4 + 2 * 4 + 8


def test_frontend_desc_matches_header_layout():
  from srf_b200 import _lib
  with open(os.path.join(ROOT, "include", "srf_b200.h")) as f:
    text = f.read()
  body = re.search(r"typedef struct srf_frontend_desc \{(.*?)\} srf_frontend_desc;", text, re.S).group(1)
  body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
  fields = []
  for decl in body.split(";"):
    for part in decl.split(","):
      part = re.sub(r"\[\d+\]", "", part.replace("*", " ")).strip()
      if part:
        fields.append(part.split()[-1])
  assert fields == [n for n, _ in _lib.FrontendDesc._fields_]
  # 34 pointers + 8 int32 + 4 float
  assert ctypes.sizeof(_lib.FrontendDesc) == 34 * 8 + 8 * 4 + 4 * 4


def test_no_gpu_means_loud_failure(built_lib):
  import pytest
  import torch
  from srf_b200 import routing
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  with pytest.raises(RuntimeError):
    routing.Handle()
  lib = ctypes.CDLL(built_lib)
  h = ctypes.c_void_p()
  lib.srf_create.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
  assert lib.srf_create(0, ctypes.byref(h)) != 0
  lib.srf_last_error.restype = ctypes.c_char_p
  lib.srf_last_error.argtypes = [ctypes.c_void_p]
  assert len(lib.srf_last_error(None)) > 0


def test_product_never_imports_oracle():
  pkg = os.path.join(ROOT, "srf_b200")
  for dirpath, _, files in os.walk(pkg):
    for fn in files:
      if fn.endswith((".py", ".cu", ".h", ".cuh")):
        with open(os.path.join(dirpath, fn)) as f:
          src = f.read()
        assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn


def test_exact_mode_picks_the_tensor_split_where_it_fits():
  """host rule of uhat_mode="exact" (the default of RoutingStack / SequenceRouter): fp32x3 when the
  3 x TF32 path is instantiated for the shape."""
  from srf_b200 import routing
  assert routing.exact_mode(20, 30, 20) == "fp32x3"      # WSJ-shaped
  assert routing.exact_mode(8, 63, 8) == "fp32x3"        # TIMIT-shaped last layer
  assert routing.exact_mode(6, 30, 8) == "fp32"          # d % 4 != 0
  assert routing.exact_mode(32, 6, 32) == "fp32x3"       # W[i] resident in groups of M tiles
  assert routing.exact_mode(20, 60, 20) == "fp32x3"
  assert routing.exact_mode(20, 100, 20) == "fp32"       # O > 64 with D > 8 is not instantiated
  assert routing.exact_mode(8, 100, 8) == "fp32x3"
  assert routing.exact_mode(20, 30, 20, emb_aligned=False) == "fp32"
