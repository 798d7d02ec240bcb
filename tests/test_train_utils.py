"""Host-side helpers of the training driver (var_gp/train_utils.py:59-98 semantics), CPU only."""
import torch

from vargp_b200.train import EarlyStopper, compute_bwt, TensorTask
from vargp_b200.dist import shard_coef, shard_loss


def test_early_stopper_counts_non_improving_evals():
  st = EarlyStopper(patience=2, delta=1e-4)
  st(0.5, 'a')
  st(0.50005, 'b')          # below best + delta: does not count as an improvement
  assert st.info() == 'a' and not st.is_done()
  st(0.6, 'c')              # improvement resets the counter
  assert st.info() == 'c'
  st(0.6, 'd')
  st(0.59, 'e')
  assert st.is_done() and st.info() == 'c'
  assert not EarlyStopper(patience=-1).is_done()


def test_compute_bwt_matches_definition():
  acc = torch.tensor([[0.9, 0.0, 0.0], [0.8, 0.95, 0.0], [0.7, 0.9, 0.99]])
  assert torch.isclose(compute_bwt(acc), torch.tensor(((0.7 - 0.9) + (0.9 - 0.95)) / 2))


def test_tensor_task_indexing_like_a_dataset():
  d = TensorTask(torch.arange(12.).view(6, 2), torch.arange(6))
  x, y = d[torch.tensor([4, 1])]
  assert len(d) == 6 and x.tolist() == [[8., 9.], [2., 3.]] and y.tolist() == [4, 1]
  t = TensorTask.for_classes(torch.arange(12.).view(6, 2), torch.tensor([0, 1, 2, 0, 1, 2]), (1, 2))
  assert len(t) == 4 and t.y.tolist() == [1, 2, 1, 2] and t[0][0].tolist() == [2., 3.]
  assert torch.unique(t.targets).numel() == 3          # the label vector stays unfiltered, like SplitMNIST


def test_shard_loss_coefficients_sum_to_full_batch_loss():
  kl_h, kl_u = torch.tensor(2.0), torch.tensor(3.0)
  nll_r = [torch.tensor(5.0), torch.tensor(7.0)]            # two ranks, half the minibatch each
  full = 1.7 * kl_h + kl_u + (240. / 24.) * (nll_r[0] + nll_r[1])
  parts = sum(shard_loss(kl_h, kl_u, n, 1.7, 240., 24, 2) for n in nll_r)
  assert torch.isclose(parts, full)
  assert shard_coef(1.7, 240., 24, 2) == (0.85, 0.5, 10.0)


def test_toy_data_reproduces_the_reference_dataset():
  """synthetic.toy_data issues the reference ToyDataset's draws (var_gp/datasets.py:21-51): same seed, same data
  (fixture recorded from the live reference by tests/golden/make_golden.py)."""
  import os
  from tests import util
  from vargp_b200.synthetic import toy_data
  rec = torch.load(os.path.join(util.GOLDEN_DIR, 'data_toy_seed3.pt'))
  torch.manual_seed(rec['seed'])
  X, Y = toy_data(N_K=rec['N_K'])
  assert torch.equal(Y, rec['targets'])
  assert (X - rec['data']).abs().max().item() < 1e-6      # same draws; allow for BLAS rounding of the 2 x 2 transform
  task = TensorTask.for_classes(X, Y, (0, 1))
  assert len(task) == 2 * rec['N_K'] and torch.unique(task.targets).numel() == 4
