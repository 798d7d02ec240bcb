#!/usr/bin/env python
"""Parity report for profiles/: the product model on the B200 kernels (fp32) against the fp64 outputs of the LIVE
reference stored in tests/golden/*.pt, next to the reference's OWN fp32-vs-fp64 error where the fixture carries it.

    python scripts/parity_report.py > gpurun_out/r2_parity_report.txt

Columns: norm-relative error of every ELBO term and gradient (max |dp| for the predictive probabilities)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import util


def relerr(a, b, floor=0.0):
  a, b = a.detach().double().cpu(), b.detach().double().cpu()
  return ((a - b).norm() / max(b.norm().item(), floor, 1e-300)).item()


def main():
  print(f'{"fixture":24s} {"quantity":12s} {"B200 fp32 vs ref fp64":>22s} {"ref fp32 vs ref fp64":>22s}')
  for name in util.golden_names() + util.large_names():
    rec = util.load_golden(name)
    params, prev, x, y, noise, n_v, F, flags = util.case_tensors(rec['case'], torch.float32)
    gp = util.build_model(params, prev, n_v, F, flags, 'cuda', torch.float32)
    r64, r32 = rec['f64'], rec.get('f32')
    beta, Ntot = (rec['beta'], rec['Ntot']) if 'beta' in rec else (r64['beta'], r64['Ntot'])
    terms, grads = util.run_model(gp, x, y, noise, beta, Ntot)
    for k in ('kl_hypers', 'kl_u', 'nll', 'total'):
      if float(r64[k].abs()) > 0:
        own = f'{relerr(r32[k], r64[k]):22.2e}' if r32 else f'{"-":>22s}'
        print(f'{name:24s} {k:12s} {relerr(terms[k], r64[k]):22.2e} {own}')
    gmax = max((v['norm'] if isinstance(v, dict) else v.norm().item()) for v in r64['grads'].values())
    for k in util.GRAD_KEYS:
      ref = r64['grads'][k]
      err = util.compressed_err(grads[k], ref) if isinstance(ref, dict) else relerr(grads[k], ref, 1e-6 * gmax)
      own = f'{relerr(r32["grads"][k], ref, 1e-6 * gmax):22.2e}' if r32 else f'{"-":>22s}'
      print(f'{name:24s} {"grad " + k:12s} {err:22.2e} {own}')
    with torch.no_grad():
      probs = gp.predict(x.cuda(), noise={k: v.cuda() for k, v in noise.items()})
    own = f'{(r32["probs"].double() - r64["probs"]).abs().max().item():22.2e}' if r32 else f'{"-":>22s}'
    print(f'{name:24s} {"max |dprobs|":12s} {(probs.double().cpu() - r64["probs"]).abs().max().item():22.2e} {own}')
  # VARGPRetrain fixtures
  from oracle import vargp_oracle as orc
  from vargp_b200.synthetic import make_retrain_case
  for name in util.retrain_names():
    rec = util.load_golden(name)
    kw = rec['case']
    params, retrain, prev, x, y, noise = make_retrain_case(dtype=torch.float32, **kw)
    gp = util.build_retrain_model(params, retrain, prev, kw.get('H', 3), kw.get('F', 10), 'cuda', torch.float32)
    r64, r32 = rec['f64'], rec['f32']
    terms, grads = util.run_retrain_model(gp, x, y, noise, r64['beta'], r64['Ntot'])
    for k in ('kl_u', 'nll', 'total'):
      print(f'{name:24s} {k:12s} {relerr(terms[k], r64[k]):22.2e} {relerr(r32[k], r64[k]):22.2e}')
    gmax = max(v.norm().item() for v in r64['grads'].values())
    for k, ref in r64['grads'].items():
      print(f'{name:24s} {"grad " + k[:18]:12s} {relerr(grads[k], ref, 1e-6 * gmax):22.2e} {relerr(r32["grads"][k], ref, 1e-6 * gmax):22.2e}')


if __name__ == '__main__':
  main()
