"""The C-ABI library loads (no GPU needed) and exports every symbol include/vargp_sm100.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'vargp_sm100.h')
LIB = os.path.join(ROOT, 'vargp_b200', 'libvargp_sm100.so')


def declared_symbols():
  src = open(HEADER).read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'\b(vargp_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_entry_points():
  syms = declared_symbols()
  for must in ('vargp_init', 'vargp_gemm', 'vargp_gemm_tc', 'vargp_chol', 'vargp_trtri', 'vargp_scale_rows',
               'vargp_softmax_nll', 'vargp_softmax_predict', 'vargp_kl_fwd', 'vargp_marginal_reduce'):
    assert must in syms


def test_library_exports_every_declared_symbol():
  assert os.path.exists(LIB), 'build the library first: python -c "import __graft_entry__ as g; g.build()"'
  lib = ctypes.CDLL(LIB)
  missing = [s for s in declared_symbols() if not hasattr(lib, s)]
  assert not missing, missing
  lib.vargp_version.restype = ctypes.c_char_p
  assert b'sm_100a' in lib.vargp_version()
  lib.vargp_strerror.restype = ctypes.c_char_p
  assert lib.vargp_strerror(-1) == b'invalid argument'


def test_gemm_descriptor_layout_matches_header():
  """ctypes mirror and the C struct must agree on size (field order is checked by the GPU tests)."""
  from vargp_b200.ops import GemmDesc
  # 3 ptr + 3 + 6 + 3 + 9 int64, 2 float, 4 int32, 2 ptr, 6 int64, 1 ptr, 4 int64, 1 int64 (sm_limit)
  expect = 8 * (3 + 3 + 6 + 3 + 9) + 4 * 2 + 4 * 4 + 8 * 2 + 8 * 6 + 8 + 8 * 4 + 8
  assert ctypes.sizeof(GemmDesc) == expect


def test_product_path_fails_loudly_without_gpu():
  import torch
  if torch.cuda.is_available():
    return
  import pytest
  from vargp_b200 import ops
  old = ops._OPS
  ops.set_ops(None)
  try:
    with pytest.raises(ops.VargpError):
      ops.get_ops()
  finally:
    ops.set_ops(old)
