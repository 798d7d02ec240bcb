"""Multi-GPU parity under NCCL (one process per GPU, torchrun-spawned): skipped on boxes with fewer than two GPUs.
See tests/dist_worker_gpu.py for what is asserted (N-rank gradients == 1-rank gradients, replicated and with the
factor stage sharded; every data-parallel tail -- graph + eager NCCL tail (default), NCCL captured into the graph, the fused
peer-memory all-reduce + Yogi kernels -- follows the same trajectory as the fully eager data-parallel step)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs at least two GPUs')
def test_nccl_data_parallel_step_matches_single_gpu():
  n = 2
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr', '127.0.0.1',
         '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'dist_worker_gpu.py')]
  out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
  assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
  rec = json.loads([l for l in out.stdout.splitlines() if l.startswith('{')][-1])
  assert rec['ok'] and rec['world'] == n, rec
