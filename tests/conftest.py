import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu on the GPU box')


def pytest_collection_modifyitems(config, items):
  if torch.cuda.is_available():
    return
  skip = pytest.mark.skip(reason='no CUDA device')
  for item in items:
    if 'gpu' in item.keywords:
      item.add_marker(skip)


@pytest.fixture
def emu_ops():
  """Install the torch emulation of the kernel interface (CPU tests of the host schedule only)."""
  from tests.emu_ops import EmuOps
  from vargp_b200 import ops
  old = ops._OPS
  ops.set_ops(EmuOps())
  yield ops._OPS
  ops.set_ops(old)


@pytest.fixture
def cuda_ops():
  from vargp_b200 import ops
  old = ops._OPS
  ops.set_ops(None)
  yield ops.get_ops()
  ops.set_ops(old)
