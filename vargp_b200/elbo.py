"""Fused sparse-variational marginal + KL(u) for VAR-GP, with a hand-derived backward.

This is the host-side schedule of the hot path: it owns no arithmetic.  Every number is produced by a
kernel of ``libvargp_sm100.so`` reached through ``vargp_b200.ops`` (C-ABI, raw pointers + strides).

Formulation.  For each hyper sample h and class c, with
Z = [z_0; ...; z_t] (P = (t+1) M rows) and K = K_h(Z, Z) + eps I:

    W   = chol(K)^-1                         one Cholesky + one triangular inverse   (replaces the t+2
                                             Choleskys of var_gp/vargp.py:61-80,108,155)
    T_s = W_ss Lu_s ,  nu_s = W_ss m_s       whitened variational factor / mean: the autoregressive joint
                                             of var_gp/gp_utils.py:101-147 is block-diagonal here
    V   = W Kzx ,  N = blockdiag(T_s T_s^T) + eps W W^T   (P x P, symmetric, once per step)
    f_mean_b = nu . V_b
    f_var_b  = gamma^2 - |V_b|^2 + V_b^T N V_b                                (gp_utils.py:150-191)
             = gamma^2 - |V_b|^2 + sum_s |T_s^T V_sb|^2 + eps |W^T V_b|^2
    KL_hc    = -sum_{i in t} log W_ii - sum_i log Lu_t,ii + (|T_t|_F^2 + |nu_t|^2 - M) / 2
    kl_u     = (1/H) sum_hc KL_hc                                             (vargp.py:182-190)

(the KL line holds for ``ep_var_mean=True``; the block-diagonal ablation goes through
``vargp_b200.gp_utils``).  All O(P^2 B) and O(P^3) work is expressed as batched GEMMs so it can run on
the tensor cores; nothing of size B x B is ever formed.
"""
import math
import os

import torch

from . import ops as _ops_mod

JITTER = 1e-4   # var_gp/gp_utils.py:5


def _ops():
  return _ops_mod.get_ops()


class _Ctx:
  """Plain container for tensors saved between forward and backward."""
  pass


_SIDE = {}
USE_SIDE_STREAM = os.environ.get('VARGP_STREAMS', '1') != '0'
KZZ_FIRST = os.environ.get('VARGP_KZZ_FIRST', '1') != '0'
# Schedule knobs, measured on B200 at the Split-MNIST shape in round 2 (profiles/r2a_ab.txt): both on = 786 vs 756 steps/s.
#   VARGP_STACK_CLASSES    x is shared by all classes, so Kzx and Gz1 = (Kxbar . Kzx) xs are ONE (C P) x B / (C P) x D
#                          product per hyper sample instead of C products with P rows each: at P = 300 that is 24
#                          instead of 30 row tiles of 128 (the per-class tiling pads 300 rows to 384).
#   VARGP_V_SIDE           V = W Kzx on the side stream right behind Kzx (waiting for an event recorded after the
#                          factorisation), so that it overlaps the T / nu / KL / N chain instead of following it.
STACK_CLASSES = os.environ.get('VARGP_STACK_CLASSES', '1') != '0'
V_SIDE = os.environ.get('VARGP_V_SIDE', '1') != '0'
# VARGP_WHITEN=0: the per-task-block products through the batched GEMMs even when M fits the shared-memory kernels
USE_WHITEN = os.environ.get('VARGP_WHITEN', '1') != '0'
KZZ_LOWER = os.environ.get('VARGP_KZZ_LOWER', '1') != '0'
SIDE_AFTER_KZZ = os.environ.get('VARGP_SIDE_AFTER_KZZ', 'auto')           # '0' / '1' / 'auto': see marginal_forward
# SMs the persistent Kzx GEMM may occupy while it runs beside Kzz -> Cholesky (148 - H*C - a margin at the benched shape)
SIDE_SM_LIMIT = int(os.environ.get('VARGP_SIDE_SM_LIMIT', '112'))
G_PARALLEL = os.environ.get('VARGP_G_PARALLEL', '1') != '0'               # see marginal_backward: 1212 -> 1242 steps/s
GZ1_SM_LIMIT = int(os.environ.get('VARGP_GZ1_SM_LIMIT', '64'))           # same for Gz1 beside the adjoint chain (0 = all); measured
                                                                         # 1168 (no cap) / 1198 (112) / 1207 (74) / 1208 (48) steps/s


class _Fork:
  """`with fork:` queues the enclosed launches on a side stream, ordered after everything already queued on the
  current stream; `fork.join()` makes the current stream wait for them.  The 30-matrix Cholesky and the chain of
  P x P adjoint GEMMs leave most SMs idle, so the minibatch-sized Kzx branch runs beside them (both branches are
  captured into the step's CUDA graph as parallel paths).  Rules that keep the caching allocator safe: every tensor
  the side branch touches is allocated on the main stream BEFORE the fork and stays referenced until after join()."""

  def __init__(self, dev):
    self.side = None
    if USE_SIDE_STREAM and dev.type == 'cuda':
      key = dev.index if dev.index is not None else torch.cuda.current_device()
      if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
      self.side = _SIDE[key]
    self._ctx = None
    self._ev = None

  def mark(self):
    """The side branch depends on what is queued on the current stream UP TO HERE (not up to the point where it
    is queued): lets the branch be launched after the head of the critical chain without waiting for it."""
    if self.side is not None:
      self._ev = torch.cuda.Event()
      self._ev.record(torch.cuda.current_stream())

  def __enter__(self):
    if self.side is not None:
      if self._ev is not None:
        self.side.wait_event(self._ev)
      else:
        self.side.wait_stream(torch.cuda.current_stream())
      self._ctx = torch.cuda.stream(self.side)
      self._ctx.__enter__()
    return self

  def __exit__(self, *exc):
    if self._ctx is not None:
      self._ctx.__exit__(*exc)
      self._ctx = None
    return False

  def join(self):
    if self.side is not None:
      torch.cuda.current_stream().wait_stream(self.side)

  def side_event(self):
    """Event marking what has been queued on the side stream so far (None without a side stream)."""
    if self.side is None:
      return None
    ev = torch.cuda.Event()
    ev.record(self.side)
    return ev

  @staticmethod
  def main_wait(ev):
    if ev is not None:
      torch.cuda.current_stream().wait_event(ev)

  def after_main(self):
    """Make what is queued on the side stream from now on also wait for everything queued on the current stream."""
    if self.side is not None:
      ev = torch.cuda.Event()
      ev.record(torch.cuda.current_stream())
      self.side.wait_event(ev)


_SIDE2 = {}


class _Fork2:
  """A second helper stream (high priority, like the capture stream of the step): `with par:` queues launches that only
  depend on what is already queued on the current stream; `par.join()` makes the current stream wait for them."""

  def __init__(self, dev):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _SIDE2:
      _SIDE2[key] = torch.cuda.Stream(device=dev, priority=-1)
    self.side = _SIDE2[key]
    self._ctx = None

  def __enter__(self):
    self.side.wait_stream(torch.cuda.current_stream())
    self._ctx = torch.cuda.stream(self.side)
    self._ctx.__enter__()
    return self

  def __exit__(self, *exc):
    self._ctx.__exit__(*exc)
    self._ctx = None
    return False

  def join(self):
    torch.cuda.current_stream().wait_stream(self.side)


def _zeros_many(dev, dt, *shapes):
  """Zero-filled tensors of the given shapes carved out of ONE buffer (one fill launch; 128 B aligned segments)."""
  sizes = [math.prod(sh) for sh in shapes]
  offs, tot = [], 0
  for n in sizes:
    offs.append(tot)
    tot += (n + 31) // 32 * 32
  flat = torch.zeros(tot, device=dev, dtype=dt)
  return [flat[o:o + n].view(sh) for o, n, sh in zip(offs, sizes, shapes)]


def _blocks(mat, S, M):
  """(H, C, P, P) -> view (H, C, S, M, M) of the S diagonal M x M blocks (no copy)."""
  H, C, P, _ = mat.shape
  sH, sC, sR, sC2 = mat.stride()
  return mat.as_strided((H, C, S, M, M), (sH, sC, M * sR + M * sC2, sR, sC2))


def _rows(mat, S, M):
  """(H, C, P, B) -> view (H, C, S, M, B) of the S row blocks."""
  H, C, P, B = mat.shape
  sH, sC, sR, sB = mat.stride()
  return mat.as_strided((H, C, S, M, B), (sH, sC, M * sR, sR, sB))


class FactorShard:
  """Sharding of the replicated O(P^3) "factor" work (Kzz, Cholesky + inverse, whitening, KL, N and their adjoints)
  over the ranks of a data-parallel group.  The H*C (hyper sample, class) pairs are dealt out in contiguous ranges
  of k = ceil(H*C / R); every rank factors only its pairs, then W, N and nu are all-gathered (forward) and the
  minibatch-partial sums Wbar, G, nubar are reduce-scattered to the owners (backward).  At the scaled config
  (P = 2048, 8 ranks) this replaces ~21 ms of replicated work per rank by ~3 ms of work plus ~1.5 GB of NVLink
  traffic.  Everything downstream (parameter gradients) is a per-rank partial sum, summed by the flat-bucket
  all-reduce that data parallelism performs anyway; kl_u comes back as the rank's share (sum over ranks = kl_u)."""

  def __init__(self, group=None):
    import torch.distributed as dist
    self.dist, self.group = dist, group
    self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

  def split(self, G):
    k = -(-G // self.world)
    return k, min(self.rank * k, G), min((self.rank + 1) * k, G)

  def all_gather(self, full, k):
    """full (world*k, ...) with this rank's k slots already in place -> every slot filled."""
    mine = full[self.rank * k:(self.rank + 1) * k]
    try:
      self.dist.all_gather_into_tensor(full, mine, group=self.group)
    except (RuntimeError, NotImplementedError):        # backends without the flat variant (gloo in the CPU tests)
      self.dist.all_gather(list(full.chunk(self.world)), mine.clone(), group=self.group)

  def reduce_scatter(self, full, k):
    """full (world*k, ...) of per-rank partial sums -> this rank's k slots hold the total (other slots undefined)."""
    mine = full[self.rank * k:(self.rank + 1) * k]
    try:
      out = torch.empty_like(mine)
      self.dist.reduce_scatter_tensor(out, full, group=self.group)
      mine.copy_(out)
    except (RuntimeError, NotImplementedError):
      self.dist.all_reduce(full, group=self.group)


def _rects(H, C, g0, g1):
  """Pair range [g0, g1) of the flattened (h, c) grid as rectangles (h0, h1, c0, c1); the full grid is one."""
  if g0 == 0 and g1 == H * C:
    return [(0, H, 0, C)]
  out = []
  g = g0
  while g < g1:
    h, c0 = divmod(g, C)
    c1 = min(C, c0 + (g1 - g))
    out.append((h, h + 1, c0, c1))
    g += c1 - c0
  return out


def _pairs(G, P, *tail, dev, dt, slots, zero=False):
  """(H*C padded to `slots`, ...) buffer; returns (flat buffer, view of the first G slots)."""
  mk = torch.zeros if zero else torch.empty
  full = mk(slots, *tail, device=dev, dtype=dt)
  return full, full[:G]


def marginal_forward(theta, Zcat, x, m_all, Lu_all, M, want_kl, ctx=None, shard=None, zeroed=None):
  """theta (H, D+1); Zcat (C, P, D); x (B, D); m_all (S, C, M); Lu_all (S, C, M, M) lower.

  Returns f_mean, f_var (H, C, B), kl_u (0-d tensor or None) and fills ``ctx`` for the backward.
  With `shard` (a FactorShard) the factor stage only runs for this rank's (h, c) pairs, see FactorShard.
  `zeroed` = (info int32 (H*C,), kl 0-d, whiten workspace (1 + H*C,)): caller-provided, already zero-filled status
  words / KL accumulator / ticket workspace (the fused training step carves them out of its per-step arena instead of
  paying fill launches here).
  """
  ops = _ops()
  H, D1 = theta.shape
  D = D1 - 1
  C, P, _ = Zcat.shape
  B = x.shape[0]
  S = P // M
  G = H * C
  dev, dt = x.device, x.dtype
  new = lambda *s: torch.empty(*s, device=dev, dtype=dt)
  if shard is not None:
    k, g0, g1 = shard.split(G)
    slots = k * shard.world
  else:
    k, g0, g1, slots = G, 0, G, G
  rects = _rects(H, C, g0, g1)

  # (1) scaled operands and their squared norms                               [kernels.py:41-44,50]
  zs, zn = new(H, C * P, D), new(H, C * P)
  xs, xn = new(H, B, D), new(H, B)
  Kzz, Kzx = new(H, C, P, P), new(H, C, P, B)
  ops.scale_rows(Zcat.reshape(C * P, D), theta, zs, zn)
  zs4, zn3 = zs.view(H, C, P, D), zn.view(H, C, P)

  # (2) Gram matrices                                                          [kernels.py:45-56]
  #     the x side and Kzx (minibatch-sized) run on the side stream next to Kzz -> Cholesky -> whitening -> KL -> N
  fork = _Fork(dev)
  if KZZ_FIRST:
    fork.mark()                  # the Kzx branch needs theta and the scaled z only
  side_queued = False

  def queue_side():
    nonlocal side_queued
    if not side_queued:
      side_queued = True
      with fork:
        if not xs_queued:
          ops.scale_rows(x, theta, xs, xn)
        if STACK_CLASSES:
          # (beside the factorisation chain, whose shared-memory kernels need H*C SMs: the persistent GEMM leaves them free)
          ops.rbf_gram(zs.view(H, 1, C * P, D), zn.view(H, 1, C * P), xs.view(H, 1, B, D), xn.view(H, 1, B), theta,
                       Kzx.view(H, 1, C * P, B), False, tag='Kzx', sm_limit=SIDE_SM_LIMIT)
        else:
          ops.rbf_gram(zs4, zn3, xs.view(H, 1, B, D), xn.view(H, 1, B), theta, Kzx, False, tag='Kzx')

  # SIDE_AFTER_KZZ: the persistent Kzx GEMM must not grab its SMs BEFORE the cluster-cooperative factorisation (120 of 148 SMs,
  # potrf_cluster.cu) is resident -- its CTAs never yield, and the factorisation would wait for the whole Gram.  The x-side
  # scaling still starts early; Kzx itself becomes ready together with the factorisation (event after Kzz), is launched after
  # it, fills the 28 SMs the clusters leave free and takes the rest when they retire.
  # Measured (B200, Split-MNIST shape, steps/s): t=4 (P=300) 1133 -> 1166, t=1 (P=120) 2496 -> 2543, t=0 (P=60, 2-CTA clusters
  # on 60 SMs) 3288 -> 3225, Permuted t=9 (blocked factorisation) 245 -> 245: 'auto' = only with 4-CTA clusters.
  xs_queued = False
  after_kzz = SIDE_AFTER_KZZ == '1' or (SIDE_AFTER_KZZ == 'auto' and 96 < P <= 320 and hasattr(ops, 'chol_cluster_wants')
                                        and ops.chol_cluster_wants(P))
  if KZZ_FIRST and after_kzz and fork.side is not None and shard is None:
    with fork:
      ops.scale_rows(x, theta, xs, xn)
    xs_queued = True
  if not KZZ_FIRST:
    queue_side()

  # (3)-(6a) factor stage on this rank's (h, c) rectangles
  L = new(H, C, P, P)
  Wf, W = _pairs(G, P, P, P, dev=dev, dt=dt, slots=slots)
  Nf, N = _pairs(G, P, P, P, dev=dev, dt=dt, slots=slots)
  nuf, nu = _pairs(G, P, P, dev=dev, dt=dt, slots=slots)
  W, N, nu = W.view(H, C, P, P), N.view(H, C, P, P), nu.view(H, C, P)
  T = new(H, C, S, M, M)
  # V_SIDE: only when one rectangle covers every pair (unsharded), so that W is complete after its chol_inv
  v_early = V_SIDE and shard is None and len(rects) == 1
  V_early = new(H, C, P, B) if v_early else None
  wwork = None
  if zeroed is not None:
    info, kl0, wwork = zeroed
  elif dt == torch.float32:     # Cholesky status words and the KL accumulator share one zero-filled buffer
    zb = torch.zeros(G + 1, device=dev, dtype=dt)
    info, kl0 = zb[:G].view(torch.int32), zb[G]
  else:
    info, kl0 = torch.zeros(G, device=dev, dtype=torch.int32), torch.zeros((), device=dev, dtype=dt)
  kl = kl0 if want_kl else None
  LuB = Lu_all.permute(1, 0, 2, 3).unsqueeze(0)                 # (1, C, S, M, M), broadcast over h
  mB = m_all.permute(1, 0, 2).unsqueeze(0).unsqueeze(-1)        # (1, C, S, M, 1)
  use_wf = USE_WHITEN and hasattr(ops, 'whiten_fwd') and M <= ops.whiten_max_m(False)
  for (h0, h1, c0, c1) in rects:
    r = (slice(h0, h1), slice(c0, c1))
    th, Hs, Cs = theta[h0:h1], h1 - h0, c1 - c0
    # the factorisation reads the lower triangle only: a third fewer tiles on the head of the critical chain; the backward
    # pass, which needs the full symmetric Gram, mirrors it on its side branch (marginal_backward)
    ops.rbf_gram(zs4[r], zn3[r], zs4[r], zn3[r], th, Kzz[r], True, tag='Kzz', c_tri='lower' if KZZ_LOWER else None)
    if xs_queued and not side_queued:
      fork.mark()                              # Kzx waits for Kzz (see SIDE_AFTER_KZZ above)
    # W = chol(Kzz + eps I)^-1                                                 [gp_utils.py:5-11]
    ops.chol_inv(Kzz[r], L[r], W[r], JITTER, info.view(H, C)[r].reshape(-1) if Hs * Cs == G else
                 info[h0 * C + c0:h0 * C + c1])
    # the side branch is queued AFTER the head of the critical chain (Kzz -> Cholesky), so that launch order --
    # like the stream priorities under graph replay -- hands the SMs to the chain first
    queue_side()
    if v_early:
      fork.after_main()                        # W is complete
      with fork:
        ops.gemm(W, Kzx, V_early, a_tri='lower', tag='V=W*Kzx', zeroed=True)
    if use_wf:
      # small task blocks: N = eps W W^T by GEMM, then ONE shared-memory kernel for T_s = W_ss Lu_s, nu_s = W_ss m_s,
      # N_ss += T_s T_s^T and the KL (whiten.cu) instead of four launches on 128 x 128 tiles that are 95 % padding
      ops.gemm(W[r], W[r].transpose(-1, -2), N[r], alpha=JITTER, a_tri='lower', b_tri='upper', tag='N=eps*W*Wt',
               zeroed=True)
      ops.whiten_fwd(W, Lu_all, m_all, T, nu, N, kl if want_kl else None, rect=(h0, h1, c0, c1), work=wwork)
      continue
    # whitened variational parameters (block diagonal)
    Wd = _blocks(W[r], S, M)
    ops.gemm(Wd, LuB[:, c0:c1], T[r], a_tri='lower', b_tri='lower', tag='T=Wss*Lu', zeroed=True)
    ops.gemm(Wd, mB[:, c0:c1], nu[r].reshape(Hs, Cs, S, M, 1), a_tri='lower', tag='nu=Wss*m', zeroed=True)
    # KL(q(u_t | u_<t) || p(u_t | u_<t))                                       [vargp.py:182-190]
    if want_kl:
      if Hs == H:
        ops.kl_fwd(W[r], T[r], nu[r], Lu_all[S - 1][c0:c1], M, kl)
      else:                      # the kernel averages over ITS hyper samples: rescale a single-h rectangle by 1/H
        part = torch.zeros((), device=dev, dtype=dt)
        ops.kl_fwd(W[r], T[r], nu[r], Lu_all[S - 1][c0:c1].contiguous(), M, part)
        kl.add_(part, alpha=float(Hs) / H)
    # N = blockdiag(T_s T_s^T) + eps W W^T collects everything quadratic in V, so the minibatch-sized work is
    # two GEMMs (V, N V) and one streaming reduction instead of three GEMMs here and four more in the backward
    ops.gemm(W[r], W[r].transpose(-1, -2), N[r], alpha=JITTER, a_tri='lower', b_tri='upper', tag='N=eps*W*Wt',
             zeroed=True)
    ops.gemm(T[r], T[r].transpose(-1, -2), _blocks(N[r], S, M), beta=1., a_tri='lower', b_tri='upper',
             tag='N+=T*Tt', zeroed=True)
  queue_side()
  if shard is not None:
    shard.all_gather(Wf, k)
    shard.all_gather(Nf, k)
    shard.all_gather(nuf, k)

  # (6b) predictive marginal                                                   [gp_utils.py:150-191]
  V, NV = (V_early if v_early else new(H, C, P, B)), new(H, C, P, B)
  fork.join()
  if not v_early:
    ops.gemm(W, Kzx, V, a_tri='lower', tag='V=W*Kzx', zeroed=True)
  ops.gemm(N, V, NV, tag='NV=N*V')
  f_mean, f_var = new(H, C, B), new(H, C, B)
  ops.marginal_reduce(V, NV, nu, theta, f_mean, f_var)

  if ctx is not None:
    ctx.dims = (H, C, P, B, D, S, M)
    ctx.shard, ctx.part = shard, (k, g0, g1, slots)
    ctx.saved = dict(theta=theta, zs=zs4, xs=xs, Kzz=Kzz, Kzx=Kzx, W=W, T=T, nu=nu, V=V, NV=NV,
                     m_all=m_all, Lu_all=Lu_all)
    ctx.kzz_lower = KZZ_LOWER
  return f_mean, f_var, kl, info, L


def marginal_backward(ctx, g_mean, g_var, g_kl, need_x_grad=False, last_raw=False):
  """Adjoint of `marginal_forward`.  g_mean, g_var (H, C, B) or None; g_kl 0-d tensor or None.

  Returns grads (theta (H, D+1), Zcat (C, P, D), x (B, D) or None, m_all (S, C, M), Lu_all (S, C, M, M)); with a
  FactorShard these are the rank's partial sums.
  `last_raw` (fused training step: the previous tasks' variational parameters are constants): only the current task's
  block of the whitening adjoint is formed and the last two results are the per-hyper-sample terms mbar (H, C, M, 1),
  Lubar (H, C, M, M) of that block, WITHOUT the sum over h and without the -g_kl / Lu_ii term of the KL (both are
  folded into ops.step_grad_finish).
  """
  ops = _ops()
  H, C, P, B, D, S, M = ctx.dims
  small = P * B <= (1 << 20)      # Split / Permuted-MNIST sized steps: the side products fit beside the chain (G_PARALLEL, GZ1_SM_LIMIT);
                                  # at the scaled shape every product fills the chip and the caps cost 5 % (402 vs 383 ms / step)
  G = H * C
  shard = ctx.shard
  k, g0, g1, slots = ctx.part
  rects = _rects(H, C, g0, g1)
  sv = ctx.saved
  theta, zs, xs, Kzz, Kzx, W, T, nu = (sv[k_] for k_ in ('theta', 'zs', 'xs', 'Kzz', 'Kzx', 'W', 'T', 'nu'))
  V, NV, m_all, Lu_all = (sv[k_] for k_ in ('V', 'NV', 'm_all', 'Lu_all'))
  dev, dt = V.device, V.dtype
  new = lambda *s: torch.empty(*s, device=dev, dtype=dt)
  LuB = Lu_all.permute(1, 0, 2, 3).unsqueeze(0)                 # (1, C, S, M, M)
  mB = m_all.permute(1, 0, 2).unsqueeze(0)                      # (1, C, S, M)

  Wbarf, Gf, nubarf, Tbar, theta_bar, r1z, csumz = _zeros_many(dev, dt, (slots, P, P), (slots, P, P), (slots, P),
                                                               (H, C, S, M, M), (H, D + 1), (H, C, P), (H, B))
  Wbar, Gm, nubar = Wbarf[:G].view(H, C, P, P), Gf[:G].view(H, C, P, P), nubarf[:G].view(H, C, P)
  Kxbar = Gz1 = Gx = r1 = csum = kzz_ready = None
  fork = _Fork(dev)
  have_data = g_mean is not None or g_var is not None
  if have_data:
    if g_mean is None:
      g_mean = torch.zeros_like(g_var)
    if g_var is None:
      g_var = torch.zeros_like(g_mean)
    g_mean, g_var = g_mean.contiguous(), g_var.contiguous()
    # Vbar = nu gm^T + 2 gv (N V - V)  (overwrites NV) ;  Vg = gv V ;  theta_bar[:, D] += 2 gamma^2 sum_cb gv
    Vbar, Vg = NV, new(H, C, P, B)
    ops.marginal_bwd_prep(V, NV, nu, g_mean, g_var, theta, Vbar, Vg, theta_bar)
    # Kzx side of the adjoint on the side stream: Kzx_bar = W^T Vbar -> (.) Kzx, row / column sums -> Gz1 = Wk1 xs
    Kxbar, Gz1 = new(H, C, P, B), new(H, C, P, D)
    Gx = new(H, C, B, D) if need_x_grad else None
    r1, csum = r1z, csumz
    with fork:
      if getattr(ctx, 'kzz_lower', False):       # complete the symmetric Gram (its forward pass only formed the lower triangle)
        for (h0, h1, c0, c1) in rects:
          ops.sym_phi(Kzz[h0:h1, c0:c1], mirror=True)
        kzz_ready, ctx.kzz_lower = fork.side_event(), False
      ops.gemm(W.transpose(-1, -2), Vbar, Kxbar, a_tri='upper', tag='Kxbar=Wt*Vbar', zeroed=True,
               sm_limit=-1 if small else 0)     # beside the chain: one tile per CTA, so that the chain's launches can cut in
      ops.rbf_bwd_prep(Kxbar, Kzx, r1, csum)                    # Kxbar <- Kxbar * Kzx ; col sums over (c, i)
      if STACK_CLASSES:
        ops.gemm(Kxbar.view(H, 1, C * P, B), xs.view(H, 1, B, D), Gz1.view(H, 1, C * P, D), tag='Gz1=Wk1*xs',
                 sm_limit=GZ1_SM_LIMIT if small else 0)
      else:
        ops.gemm(Kxbar, xs.view(H, 1, B, D), Gz1, tag='Gz1=Wk1*xs')
      if need_x_grad:
        ops.gemm(Kxbar.transpose(-1, -2), zs, Gx, tag='Gx=Wk1t*zs')
    # minibatch sums for every pair:  Wbar = tril(Vbar Kzx^T),  G = Nbar = sum_b gv_b V_b V_b^T (lower),  nubar = V gm
    # Wbar and G are independent lower-triangular products of 180 tiles each (1.2 waves of 148 SMs: the second wave of
    # each runs on 32 SMs); queued on two streams the tiles of both pack into 2.4 waves (G_PARALLEL, a second helper
    # stream of the same high priority as the chain)
    par = _Fork2(dev) if (G_PARALLEL and small and fork.side is not None and shard is None) else None
    if par is not None:
      with par:
        ops.gemm(Vg, V.transpose(-1, -2), Gm, c_tri='lower', tag='G=Vg*Vt')
        ops.gemm(V, g_mean.unsqueeze(-1), nubar.unsqueeze(-1), tag='nubar=V*gm')
    ops.gemm(Vbar, Kzx.transpose(-1, -2), Wbar, c_tri='lower', tag='Wbar=Vbar*Kzxt')
    if par is not None:
      par.join()
    else:
      ops.gemm(Vg, V.transpose(-1, -2), Gm, c_tri='lower', tag='G=Vg*Vt')
      ops.gemm(V, g_mean.unsqueeze(-1), nubar.unsqueeze(-1), tag='nubar=V*gm')
    if shard is not None:        # ... summed over the ranks' minibatch slices, delivered to the owner of each pair
      shard.reduce_scatter(Wbarf, k)
      shard.reduce_scatter(Gf, k)
      shard.reduce_scatter(nubarf, k)

  # factor-stage adjoint on this rank's (h, c) rectangles; whatever is not owned stays zero
  sharded = shard is not None
  mk = torch.zeros if sharded else torch.empty
  Sg = 1 if last_raw else S                                     # task blocks whose parameter gradients are formed
  Lubar_h = mk(H, Sg, C, M, M, device=dev, dtype=dt)            # task-major: the sum over h is in parameter layout
  mbar_h = mk(H, Sg, C, M, 1, device=dev, dtype=dt)
  Gz2, r2, dg = mk(H, C, P, D, device=dev, dtype=dt), mk(H, C, P, device=dev, dtype=dt), mk(H, C, P, device=dev, dtype=dt)
  X, Y = new(H, C, P, P), new(H, C, P, P)
  g_kl_r = None
  use_wb = USE_WHITEN and hasattr(ops, 'whiten_bwd') and M <= ops.whiten_max_m(True)
  for (h0, h1, c0, c1) in rects:
    r = (slice(h0, h1), slice(c0, c1))
    Hs, Cs = h1 - h0, c1 - c0
    if have_data:
      ops.sym_phi(Gm[r], mirror=True)
    s0 = S - Sg
    if use_wb:
      # N = blockdiag(T_s T_s^T) + eps W W^T  =>  Wbar += tril(2 eps G W) by GEMM; everything that lives on the M x M task
      # blocks (Tbar_s = tril(2 G_ss T_s), the KL adjoint, the whitening adjoint into Wbar_ss, Lubar, mbar) in ONE
      # shared-memory kernel (whiten.cu) instead of a GEMM, the KL kernel and four more products
      if have_data:
        ops.gemm(Gm[r], W[r], Wbar[r], alpha=2. * JITTER, beta=1., b_tri='lower', c_tri='lower', tag='Wbar+=2eps*G*W',
                 zeroed=True)
      ops.whiten_bwd(W, T, nu, Lu_all, m_all, Gm, nubar, g_kl, Wbar, Lubar_h, mbar_h, s_grad0=s0, rect=(h0, h1, c0, c1))
    else:
      if have_data:
        # N = blockdiag(T_s T_s^T) + eps W W^T  =>  Tbar_s = tril(2 G_ss T_s),  Wbar += tril(2 eps G W)
        ops.gemm(_blocks(Gm[r], S, M), T[r], Tbar[r], alpha=2., b_tri='lower', c_tri='lower', tag='Tbar=2*Gss*T',
                 zeroed=True)
        ops.gemm(Gm[r], W[r], Wbar[r], alpha=2. * JITTER, beta=1., b_tri='lower', c_tri='lower', tag='Wbar+=2eps*G*W',
                 zeroed=True)
      if g_kl is not None:
        # adds (g_kl/H) T_t, (g_kl/H) nu_t and -(g_kl/H)/W_ii on the last block (the kernel divides by ITS H)
        if Hs == H:
          g_here = g_kl
        else:
          if g_kl_r is None:
            g_kl_r = g_kl * (1.0 / H)
          g_here = g_kl_r
        ops.kl_bwd(W[r], T[r], nu[r], M, g_here, Wbar[r], Tbar[r], nubar[r])
      # whitening adjoint: Wbar_ss += tril(Tbar_s Lu_s^T + nubar_s m_s^T); Lu_bar_s = sum_h W_ss^T Tbar_s ; m_bar_s = sum_h W_ss^T nubar_s
      Wd, Wbd = _blocks(W[r], S, M), _blocks(Wbar[r], S, M)
      nub5 = nubar[r].reshape(Hs, Cs, S, M, 1)
      ops.gemm(Tbar[r], LuB[:, c0:c1].transpose(-1, -2), Wbd, beta=1., a_tri='lower', b_tri='upper', c_tri='lower',
               tag='whiten_adj', zeroed=True)
      ops.gemm(nub5, mB[:, c0:c1].unsqueeze(-2), Wbd, beta=1., c_tri='lower', tag='whiten_adj')
      ops.gemm(Wd[:, :, s0:].transpose(-1, -2), Tbar[r][:, :, s0:], Lubar_h[h0:h1, :, c0:c1].permute(0, 2, 1, 3, 4),
               a_tri='upper', b_tri='lower', c_tri='lower', tag='whiten_adj', zeroed=True)
      ops.gemm(Wd[:, :, s0:].transpose(-1, -2), nub5[:, :, s0:], mbar_h[h0:h1, :, c0:c1].permute(0, 2, 1, 3, 4),
               a_tri='upper', tag='whiten_adj', zeroed=True)
    # Cholesky-inverse adjoint:  Kbar = -W^T Xi W,  Xi = (Phi(X) + Phi(X)^T)/2,  X = tril(Wbar W^T)
    ops.gemm(Wbar[r], W[r].transpose(-1, -2), X[r], a_tri='lower', b_tri='upper', c_tri='lower', tag='X=Wbar*Wt',
             zeroed=True)
    ops.sym_phi(X[r])                                            # in place -> Xi (full symmetric)
    ops.gemm(X[r], W[r], Y[r], b_tri='lower', tag='Y=Xi*W', zeroed=True)
    Kzzbar = X[r]                                                # Xi is dead after Y: reuse its storage
    ops.gemm(W[r].transpose(-1, -2), Y[r], Kzzbar, alpha=-1., a_tri='upper', tag='Kzzbar=-Wt*Y', zeroed=True)
    # RBF adjoint of the Kzz side                                                
    if getattr(ctx, 'kzz_lower', False):         # no side branch ran (no data term): mirror here
      ops.sym_phi(Kzz[r], mirror=True)
    _Fork.main_wait(kzz_ready)
    ops.rbf_bwd_prep(Kzzbar, Kzz[r], r2[r], None, dg[r])          # Kzzbar <- Kzzbar * Kzz (diag -> dg) ; row sums
    ops.gemm(Kzzbar, zs[r], Gz2[r], tag='Gz2=Wk2*zs')
  if last_raw:
    m_bar, Lu_bar = mbar_h[:, 0], Lubar_h[:, 0]                  # (H, C, M, 1), (H, C, M, M)
  else:
    Lu_bar = Lubar_h.sum(0)                                      # (S, C, M, M)
    m_bar = mbar_h.sum(0).squeeze(-1)                            # (S, C, M)
    if g_kl is not None and (not sharded or shard.rank == 0):
      # d/dLu_t of -sum_i log Lu_t,ii  (mean over h of H identical terms; counted once across the ranks)
      ops.kl_bwd_lu(Lu_all[S - 1], g_kl, Lu_bar[S - 1])

  Z_bar = new(C, P, D)
  fork.join()
  ops.rbf_bwd_finish(zs, Gz1, Gz2, r1, r2, theta, Z_bar, theta_bar, dg)
  x_bar = None
  if have_data:
    x_bar = new(B, D) if need_x_grad else None
    ops.rbf_bwd_xside(xs, csum, Gx, theta, theta_bar, x_bar)
  return theta_bar, Z_bar, x_bar, m_bar, Lu_bar
