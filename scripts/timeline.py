#!/usr/bin/env python
"""Kernel timeline of graph-replayed ELBO steps (CUPTI through torch.profiler; nsys is not in the image).

    python scripts/timeline.py [--workload split_mnist] [--task T] [--out gpurun_out/timeline.json]

Writes, for one replayed step, every kernel's (name, stream, start us relative to the step, duration us) and prints
the step span, the summed kernel time and the idle gaps on the union of streams.  Profiler overhead inflates
absolute numbers a little; the picture (what overlaps, where the device idles) is what this is for."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--workload', default='split_mnist')
  ap.add_argument('--task', type=int, default=None)
  ap.add_argument('--out', default='gpurun_out/timeline.json')
  ap.add_argument('--no-graph', action='store_true')
  a = ap.parse_args()
  task = a.task if a.task is not None else bench.WORKLOADS[a.workload]['tasks'] - 1
  dev = torch.device('cuda', 0)
  cfg, params, prev = bench.make_problem(a.workload, task, dev)
  gp = bench.build_gpu_model(params, prev, dev)
  from vargp_b200.train import ElboStepper
  st = ElboStepper(gp, n_data=cfg['N'], batch_size=cfg['B'], beta=cfg['beta'], lr=3e-3, use_graph=not a.no_graph)
  xs, ys = bench.synth_batches(8, cfg['B'], cfg['D'], cfg['C'], task, dev, seed=0)
  torch.manual_seed(1234)
  for i in range(10):
    st.step(xs[i % 8], ys[i % 8])
  torch.cuda.synchronize()
  from torch.profiler import profile, ProfilerActivity
  with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(6):
      st.step(xs[i % 8], ys[i % 8])
    torch.cuda.synchronize()
  ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and 'emcpy' not in e.name and 'emset' not in e.name]
  ev.sort(key=lambda e: e.time_range.start)
  # split into steps at the yogi_step kernel (last kernel of a step)
  steps, cur = [], []
  for e in ev:
    cur.append(e)
    if 'yogi_step' in e.name:
      steps.append(cur); cur = []
  s = steps[3]
  t0 = s[0].time_range.start
  rows = [dict(name=e.name[:70], stream=getattr(e, 'device_resource_id', getattr(e, 'stream', -1)) if hasattr(e, 'device_resource_id') else -1,
               start=round(e.time_range.start - t0, 2), dur=round(e.time_range.end - e.time_range.start, 2)) for e in s]
  span = max(r['start'] + r['dur'] for r in rows)
  busy = sum(r['dur'] for r in rows)
  # idle time on the union of all streams
  iv = sorted((r['start'], r['start'] + r['dur']) for r in rows)
  idle, end = 0.0, 0.0
  gaps = []
  for b, e_ in iv:
    if b > end:
      idle += b - end; gaps.append((round(end, 1), round(b - end, 2)))
    end = max(end, e_)
  os.makedirs(os.path.dirname(a.out) or '.', exist_ok=True)
  json.dump(dict(workload=a.workload, task=task, span_us=span, kernel_sum_us=busy, idle_us=idle, kernels=rows), open(a.out, 'w'))
  print(f'step span {span:.1f} us, kernel sum {busy:.1f} us, idle (no kernel on any stream) {idle:.1f} us, {len(rows)} kernels')
  for r in rows:
    print(f"{r['start']:9.1f} {r['dur']:8.1f}  s{r['stream']}  {r['name']}")


if __name__ == '__main__':
  main()
