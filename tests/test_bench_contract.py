"""bench.py's reference arm runs on CPU (the unmodified reference from oracle/_ref on the host cores, else the oracle
port): check the JSON line it prints against the contract keys the driver reads.  (The GPU arm needs a B200; its line
is checked by the driver itself.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
  env = dict(os.environ, OMP_NUM_THREADS='4')
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup', '1',
                        '--task', '0'], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
  assert out.returncode == 0, out.stderr[-2000:]
  line = json.loads(out.stdout.strip().splitlines()[-1])
  assert line['impl'] == 'reference' and line['unit'] == 'steps/s' and line['higher_is_better'] is True
  assert line['value'] > 0 and line['steps'] >= 2 and line['gpu_launches'] == 0
  for k in ('metric', 'n_gpus', 'warmup', 'ms_per_step', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'e2e',
            'cpu_baseline'):
    assert k in line, k
  from oracle import ref_runner
  assert line['cpu_baseline']['kind'] == ('reference' if ref_runner.available() else 'port')
  assert line['cpu_baseline']['cores'] >= 1
  assert line['e2e'] == {'value': line['value'], 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
  assert 'workload' in line['config']
  # both arms print the SAME config dict for the same workload / world size
  sys.path.insert(0, ROOT)
  import bench
  assert line['config'] == bench.bench_config('split_mnist', 0, 1)


def test_port_arm_still_available():
  env = dict(os.environ, OMP_NUM_THREADS='4', VARGP_BENCH_PORT='1')
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup', '1',
                        '--task', '0'], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
  assert out.returncode == 0, out.stderr[-2000:]
  assert json.loads(out.stdout.strip().splitlines()[-1])['cpu_baseline']['kind'] == 'port'


def test_reference_arm_other_ranks_exit_quietly():
  env = dict(os.environ, RANK='1', WORLD_SIZE='2')
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                       capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
  assert out.returncode == 0 and out.stdout.strip() == ''
