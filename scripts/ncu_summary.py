#!/usr/bin/env python
"""Summarise ncu output for profiles/: `launches <csv>` (per-kernel share of a launch list) or
`rep <file.ncu-rep>` (key counters of every captured launch, one column per launch)."""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']


def launches(path):
  rows = [r for r in csv.reader(l for l in open(path) if not l.startswith('=='))]
  hdr = rows[0]
  ik, im, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
  agg = collections.OrderedDict()
  tot = 0.0
  for r in rows[1:]:
    if len(r) <= iv or r[im] != 'gpu__time_duration.sum':
      continue
    v = float(r[iv].replace(',', ''))
    v = v / 1e3 if r[iu] in ('ns', 'nsecond') else v
    d = agg.setdefault(r[ik][:96], [0.0, 0])
    d[0] += v
    d[1] += 1
    tot += v
  n = sum(d[1] for d in agg.values())
  print(f'# total {tot:.1f} us over {n} launches')
  for k, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'{v:10.1f} us {c:4d} launches {100 * v / tot:5.1f}%  {k}')


def rep(path):
  out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  hdr, units, data = rows[0], rows[1], rows[2:]
  ik = hdr.index('Kernel Name')
  print('# kernels: ' + ' | '.join(f'{i}:{r[ik][:60]}' for i, r in enumerate(data)))
  for k in KEYS:
    if k in hdr:
      j = hdr.index(k)
      print(f'{k:72s} [{units[j]:>16s}] ' + ' '.join(f'{r[j]:>16s}' for r in data))


if __name__ == '__main__':
  {'launches': launches, 'rep': rep}[sys.argv[1]](sys.argv[2])
