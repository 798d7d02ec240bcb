"""Kernel interface of the hot path (placeholder until the CUDA binding lands)."""
_OPS = None


def set_ops(o):
  global _OPS
  _OPS = o


def get_ops():
  if _OPS is None:
    raise RuntimeError('libvargp_sm100.so is not loaded')
  return _OPS
