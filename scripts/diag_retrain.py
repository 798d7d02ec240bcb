import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import util
from vargp_b200.synthetic import make_retrain_case
name = sys.argv[1] if len(sys.argv) > 1 else 'retrain_toy_t1'
rec = util.load_golden(name); kw = rec['case']
params, retrain, prev, x, y, noise = make_retrain_case(dtype=torch.float32, **kw)
gp = util.build_retrain_model(params, retrain, prev, kw.get('H', 3), kw.get('F', 10), 'cuda', torch.float32)
r64, r32 = rec['f64'], rec['f32']
terms, grads = util.run_retrain_model(gp, x, y, noise, r64['beta'], r64['Ntot'])
for k in ('log_logvar', 'log_mean', 'z', 'u_mean'):
  g = grads[k].double().cpu(); ref = r64['grads'][k]
  print(os.environ.get('TAGX', ''), k, 'err', ((g - ref).norm() / ref.norm()).item(), 'ref32 err', ((r32['grads'][k].double() - ref).norm() / ref.norm()).item())
