"""ctypes binding of ``libvargp_sm100.so`` (include/vargp_sm100.h): the ONLY compute backend.

``get_ops()`` returns the kernel interface used by the host schedule (``elbo.py``, ``gp_utils.py`` ...).
Every method takes torch CUDA fp32 tensors (views allowed where the C entry point takes strides), checks
them, and launches the kernel asynchronously on torch's current stream.  There is no CPU fallback: if the
shared library is missing or the tensors are not on a B200, this raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libvargp_sm100.so')

_TRI = {None: 0, 'lower': 1, 'upper': 2}
EPI_NONE, EPI_RBF, EPI_RBF_SYM = 0, 1, 2
KL_CHUNKS = 16    # VARGP_KL_CHUNKS (include/vargp_sm100.h)

i64 = ctypes.c_int64
f32p = ctypes.c_void_p
vp = ctypes.c_void_p


class GemmDesc(ctypes.Structure):
  """Mirror of vargp_gemm_t (include/vargp_sm100.h)."""
  _fields_ = [
    ('A', vp), ('B', vp), ('C', vp),
    ('M', i64), ('N', i64), ('K', i64),
    ('a_rs', i64), ('a_cs', i64), ('b_rs', i64), ('b_cs', i64), ('c_rs', i64), ('c_cs', i64),
    ('nb', i64 * 3),
    ('a_bs', i64 * 3), ('b_bs', i64 * 3), ('c_bs', i64 * 3),
    ('alpha', ctypes.c_float), ('beta', ctypes.c_float),
    ('tri_a', ctypes.c_int32), ('tri_b', ctypes.c_int32), ('tri_c', ctypes.c_int32),
    ('epi', ctypes.c_int32),
    ('e_row', vp), ('e_col', vp),
    ('e_row_bs', i64 * 3), ('e_col_bs', i64 * 3),
    ('e_theta', vp),
    ('e_theta_bs', i64 * 3), ('e_D', i64),
    ('sm_limit', i64),
  ]


class VargpError(RuntimeError):
  pass


def _load():
  if not os.path.exists(_LIB_PATH):
    raise VargpError(
      f'{_LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
      f'(or vargp_b200/csrc/build.sh). vargp_b200 has no CPU / eager fallback.')
  lib = ctypes.CDLL(_LIB_PATH)
  lib.vargp_version.restype = ctypes.c_char_p
  lib.vargp_strerror.restype = ctypes.c_char_p
  lib.vargp_strerror.argtypes = [ctypes.c_int]
  lib.vargp_launch_count.restype = i64
  lib.vargp_init.argtypes = [ctypes.c_int]
  lib.vargp_set_pdl.argtypes = [ctypes.c_int]
  lib.vargp_gemm.argtypes = [ctypes.POINTER(GemmDesc), vp]
  lib.vargp_gemm_tc.argtypes = [ctypes.POINTER(GemmDesc), vp]
  lib.vargp_tc2_config.argtypes = [i64]
  lib.vargp_tc2_config.restype = i64
  lib.vargp_tc2_launch_count.restype = i64
  lib.vargp_tc_persist_config.argtypes = [i64]
  lib.vargp_tc_persist_config.restype = i64
  lib.vargp_scale_rows.argtypes = [vp, i64, i64, i64, vp, i64, i64, vp, vp, vp]
  lib.vargp_chol.argtypes = [vp, i64, i64, vp, i64, i64, i64, i64, ctypes.c_float, vp, vp]
  lib.vargp_trtri.argtypes = [vp, i64, i64, vp, i64, i64, i64, i64, vp]
  lib.vargp_chol_inv.argtypes = [vp, i64, i64, vp, i64, i64, vp, i64, i64, i64, i64, ctypes.c_float, vp, vp]
  lib.vargp_chol_config.argtypes = [i64, i64]
  lib.vargp_chol_config.restype = i64
  lib.vargp_chol_cluster_config.argtypes = [i64, i64]
  lib.vargp_chol_cluster_config.restype = i64
  lib.vargp_chol_cluster_wants.argtypes = [i64]
  lib.vargp_chol_inv_cluster.argtypes = [vp, i64, i64, vp, i64, i64, vp, i64, i64, i64, i64, ctypes.c_float, vp, vp]
  lib.vargp_tril_unpack.argtypes = [vp, i64, i64, vp, vp]
  lib.vargp_tril_unpack_bwd.argtypes = [vp, vp, i64, i64, vp, vp]
  lib.vargp_kl_fwd.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, vp, vp, vp]
  lib.vargp_kl_bwd.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, vp, vp, vp, vp]
  lib.vargp_kl_bwd_lu.argtypes = [vp, vp, i64, i64, vp, vp]
  lib.vargp_marginal_reduce.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, vp, vp, vp]
  lib.vargp_marginal_bwd_prep.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, vp, vp, vp, vp]
  lib.vargp_sym_phi.argtypes = [vp, i64, i64, ctypes.c_int, vp]
  lib.vargp_rbf_bwd_prep.argtypes = [vp, vp, i64, i64, i64, i64, vp, vp, vp, vp]
  lib.vargp_rbf_bwd_finish.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, vp, vp, vp]
  lib.vargp_rbf_bwd_xside.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, i64, vp, vp, vp]
  lib.vargp_softmax_nll.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, vp, vp, vp, ctypes.c_float, vp, vp]
  lib.vargp_whiten_fwd_work.argtypes = [i64, i64]
  lib.vargp_whiten_fwd_work.restype = i64
  lib.vargp_whiten_max_m.argtypes = [ctypes.c_int]
  lib.vargp_whiten_max_m.restype = i64
  lib.vargp_whiten_fwd.argtypes = [vp, vp, vp] + [i64] * 9 + [vp] * 6
  lib.vargp_whiten_bwd.argtypes = [vp] * 8 + [i64] * 10 + [vp] * 4
  lib.vargp_step_assemble.argtypes = [vp, vp, vp, i64, i64, i64, i64, vp, vp, vp, vp]
  lib.vargp_step_grad_finish.argtypes = [vp, vp, i64, vp, i64] + [vp] * 10 + [i64] * 5 + [vp] * 6
  lib.vargp_softmax_nll_work.argtypes = [i64, i64]
  lib.vargp_softmax_nll_work.restype = i64
  lib.vargp_softmax_predict.argtypes = [vp, vp, vp, i64, i64, i64, i64, vp, vp]
  lib.vargp_yogi_step.argtypes = [vp, vp, vp, vp, i64] + [ctypes.c_float] * 4 + [vp, vp]
  lib.vargp_peer_buffer_floats.argtypes = [i64]
  lib.vargp_peer_buffer_floats.restype = i64
  lib.vargp_peer_allreduce_yogi.argtypes = [vp, ctypes.c_int, ctypes.c_int, i64, vp, vp, vp, vp] + [ctypes.c_float] * 4 + [vp, vp, vp]
  lib.vargp_peer_allreduce_yogi_nvls.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, i64, vp, vp, vp, vp] + [ctypes.c_float] * 4 + [vp, vp, vp]
  lib.vargp_hyper_fwd.argtypes = [vp, vp, vp, vp, vp, i64, i64, vp, vp, vp]
  lib.vargp_hyper_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i64, vp, vp, vp]
  return lib


def _f32(t, name, contiguous=True):
  if not isinstance(t, torch.Tensor) or t.dtype != torch.float32 or not t.is_cuda:
    raise VargpError(f'{name}: expected a CUDA float32 tensor, got '
                     f'{getattr(t, "dtype", type(t))} on {getattr(t, "device", "?")}')
  if contiguous and not t.is_contiguous():
    raise VargpError(f'{name}: expected a contiguous tensor, got strides {t.stride()} for shape {tuple(t.shape)}')
  return t.data_ptr()


def _bstrides(t, nbatch):
  """Batch strides of `t` (last two dims are the matrix), left-padded to 3, 0 for broadcast dims."""
  sizes, strides = list(t.shape[:-2]), list(t.stride()[:-2])
  pad = 3 - len(sizes)
  if pad < 0:
    raise VargpError('at most 3 batch dimensions are supported')
  sizes, strides = [1] * pad + sizes, [0] * pad + strides
  out = []
  for sz, st, nb in zip(sizes, strides, nbatch):
    if sz == nb:
      out.append(st if sz > 1 else 0)
    elif sz == 1:
      out.append(0)
    else:
      raise VargpError(f'batch dims {sizes} do not broadcast to {list(nbatch)}')
  return out


class CudaOps:
  """Kernel interface backed by libvargp_sm100.so.  Contracts: see tests/emu_ops.py (same method names)
  and include/vargp_sm100.h."""
  name = 'sm100'

  def __init__(self):
    self.lib = _load()
    if not torch.cuda.is_available():
      raise VargpError('vargp_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
    self._inited = set()
    self.use_tc = os.environ.get('VARGP_TC', '1') != '0'
    self.prof = None              # list of (tag, kernel, flops, bytes, start_event, end_event) when profiling
    self.tc_calls = self.simt_calls = 0

  # -- per-launch device timing (bench.py roofline pass) ------------------------------------------
  _STREAM_OPS = ('scale_rows', 'rbf_bwd_prep', 'rbf_bwd_finish', 'rbf_bwd_xside', 'chol', 'trtri', 'chol_inv', 'tril_unpack',
                 'tril_unpack_bwd', 'kl_fwd', 'kl_bwd', 'kl_bwd_lu', 'whiten_fwd', 'whiten_bwd', 'step_assemble', 'step_grad_finish', 'marginal_reduce', 'marginal_bwd_prep',
                 'sym_phi', 'nll_fwd_bwd', 'predict', 'yogi_step', 'hyper_fwd', 'hyper_bwd')

  # positional tensor arguments that a kernel both reads and writes (Kbar of rbf_bwd_prep, X of sym_phi, p / m / v of Yogi)
  _RMW_ARGS = {'rbf_bwd_prep': (0,), 'sym_phi': (0,), 'yogi_step': (0, 2, 3)}

  def profile_start(self):
    """Bracket every launch with CUDA events (slows the host side; never on during a timed region)."""
    self.prof = []
    for name in self._STREAM_OPS:
      fn = getattr(type(self), name)

      def timed(*a, _fn=fn, _name=name, **kw):
        # algorithmic bytes of a streaming kernel: every tensor argument is touched once, read-modify-write ones twice
        nbytes = sum(t.numel() * t.element_size() for t in a if isinstance(t, torch.Tensor))
        nbytes += sum(a[i].numel() * a[i].element_size() for i in self._RMW_ARGS.get(_name, ()) if i < len(a))
        flops = 0.0
        if _name in ('chol', 'trtri', 'chol_inv'):
          n = a[0].shape[-1]
          flops = (a[0].numel() // (n * n)) * n ** 3 / 3.0 * (2 if _name == 'chol_inv' else 1)
        return self._timed(_name, _name, flops, nbytes, lambda: _fn(self, *a, **kw))
      setattr(self, name, timed)

  def profile_stop(self):
    """-> {tag: dict(kernel, calls, ms, flops, bytes)} aggregated over the recorded launches."""
    torch.cuda.synchronize()
    out = {}
    for tag, kern, fl, by, e0, e1 in self.prof:
      d = out.setdefault(tag, dict(kernel=kern, calls=0, ms=0.0, flops=0.0, bytes=0.0))
      d['calls'] += 1
      d['ms'] += e0.elapsed_time(e1)
      d['flops'] += fl
      d['bytes'] += by
    self.prof = None
    for name in self._STREAM_OPS:
      if name in self.__dict__:
        delattr(self, name)
    return out

  def _timed(self, tag, kern, flops, nbytes, fn):
    if self.prof is None:
      return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    self.prof.append((tag, kern, float(flops), float(nbytes), e0, e1))
    return r

  # -- plumbing -------------------------------------------------------------------------------
  def _stream(self, t):
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if dev not in self._inited:
      self._check(self.lib.vargp_init(dev), 'vargp_init')
      self._inited.add(dev)
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

  def _check(self, rc, what):
    if rc != 0:
      raise VargpError(f'{what} failed: {self.lib.vargp_strerror(rc).decode()} (code {rc})')

  def launch_count(self):
    return int(self.lib.vargp_launch_count())

  def set_pdl(self, mode):
    """Programmatic dependent launch between the library's kernels: 0 off, 1 every launch, 2 light kernels only
    (default); returns the previous setting."""
    return int(self.lib.vargp_set_pdl(int(mode)))

  def tc2_config(self, min_tiles=None):
    """Set (or with None query) the tile-count threshold above which GEMMs take the 2-CTA kernel; < 0 disables."""
    return int(self.lib.vargp_tc2_config(-2 ** 63 if min_tiles is None else int(min_tiles)))

  def tc2_launch_count(self):
    return int(self.lib.vargp_tc2_launch_count())

  # -- GEMM -----------------------------------------------------------------------------------
  def _desc(self, A, B, C, alpha, beta, a_tri, b_tri, c_tri):
    for t, nm in ((A, 'A'), (B, 'B'), (C, 'C')):
      _f32(t, 'gemm ' + nm, contiguous=False)
    M, K = A.shape[-2:]
    K2, N = B.shape[-2:]
    if K != K2 or tuple(C.shape[-2:]) != (M, N):
      raise VargpError(f'gemm shape mismatch: A {tuple(A.shape)} B {tuple(B.shape)} C {tuple(C.shape)}')
    cb = list(C.shape[:-2])
    nb = [1] * (3 - len(cb)) + cb
    d = GemmDesc()
    d.A, d.B, d.C = A.data_ptr(), B.data_ptr(), C.data_ptr()
    d.M, d.N, d.K = M, N, K
    d.a_rs, d.a_cs = A.stride(-2), A.stride(-1)
    d.b_rs, d.b_cs = B.stride(-2), B.stride(-1)
    d.c_rs, d.c_cs = C.stride(-2), C.stride(-1)
    d.nb = (i64 * 3)(*nb)
    d.a_bs = (i64 * 3)(*_bstrides(A, nb))
    d.b_bs = (i64 * 3)(*_bstrides(B, nb))
    d.c_bs = (i64 * 3)(*_bstrides(C, nb))
    d.alpha, d.beta = float(alpha), float(beta)
    d.tri_a, d.tri_b, d.tri_c = _TRI[a_tri], _TRI[b_tri], _TRI[c_tri]
    d.epi = EPI_NONE
    return d, nb

  def _run_gemm(self, d, ref, tag, zeroed=False):
    """tcgen05 path when the problem qualifies (K- or M/N-contiguous operands, TMA-alignable strides) and any
    declared operand triangle is physically zero (`zeroed`); otherwise the general SIMT kernel."""
    s = self._stream(ref)
    if self.prof is not None:
      nbat = d.nb[0] * d.nb[1] * d.nb[2]
      frac = (0.5 if (d.tri_a or d.tri_b) else 1.0) * (0.5 if d.tri_c else 1.0)
      flops = 2.0 * d.M * d.N * d.K * nbat * frac
      nbytes = 4.0 * nbat * (d.M * d.K + d.K * d.N + d.M * d.N)
    else:
      flops = nbytes = 0.0
    if self.use_tc and (zeroed or not (d.tri_a or d.tri_b)):
      if self.prof is not None:
        c2 = self.lib.vargp_tc2_launch_count()
      rc = self._timed(tag, 'gemm_tc', flops, nbytes, lambda: self.lib.vargp_gemm_tc(ctypes.byref(d), s))
      if rc == 0:
        self.tc_calls += 1
        if self.prof is not None:          # name the kernel that actually ran (1-CTA or persistent 2-CTA)
          kern = 'gemm_tc2' if self.lib.vargp_tc2_launch_count() != c2 else 'gemm_tc'
          self.prof[-1] = (self.prof[-1][0], kern) + self.prof[-1][2:]
        return
      if self.prof is not None:
        self.prof.pop()
      if rc != -2:
        self._check(rc, 'vargp_gemm_tc')
    self.simt_calls += 1
    self._check(self._timed(tag, 'gemm_simt', flops, nbytes, lambda: self.lib.vargp_gemm(ctypes.byref(d), s)),
                'vargp_gemm')

  def gemm(self, A, B, C, alpha=1., beta=0., a_tri=None, b_tri=None, c_tri=None, tag='gemm', zeroed=False, sm_limit=0):
    """C = alpha tri_a(A) tri_b(B) [tri_c] + beta C.  `zeroed=True` promises that the other triangle of a
    declared-triangular operand holds actual zeros (lets the TMA-fed tensor-core kernel take the call).
    `sm_limit`: at most that many SMs for a product that runs beside a critical chain (persistent 2-CTA kernel)."""
    d, _ = self._desc(A, B, C, alpha, beta, a_tri, b_tri, c_tri)
    d.sm_limit = int(sm_limit)
    self._run_gemm(d, C, tag, zeroed)

  def rbf_gram(self, a, an, b, bn, theta, out, sym, tag='rbf_gram', sm_limit=0, c_tri=None):
    """out[h,c] = gamma2[h] exp(a b^T - |a|^2/2 - |b|^2/2); a (H,C,Pa,D), b (H,Cb,Pb,D), Cb in {1, C}.
    c_tri='lower': only the lower triangle of a symmetric Gram is formed (a third fewer tiles); complete it with sym_phi(mirror)."""
    _f32(theta, 'theta', contiguous=False)
    d, nb = self._desc(a, b.transpose(-1, -2), out, 1., 0., None, None, None)
    _f32(an, 'an', contiguous=False)
    _f32(bn, 'bn', contiguous=False)
    if an.stride(-1) != 1 or bn.stride(-1) != 1 or theta.stride(-1) != 1:
      raise VargpError('rbf_gram: norm vectors / theta must be contiguous in their last dim')
    d.epi = EPI_RBF_SYM if sym else EPI_RBF
    d.e_row, d.e_col = an.data_ptr(), bn.data_ptr()
    d.e_row_bs = (i64 * 3)(*_bstrides(an.unsqueeze(-1), nb))
    d.e_col_bs = (i64 * 3)(*_bstrides(bn.unsqueeze(-1), nb))
    d.e_theta = theta.data_ptr()
    H = theta.shape[0]
    if out.dim() < 3 or out.shape[0] != H:
      raise VargpError('rbf_gram: the leading batch dimension of `out` must be the hyper-sample index')
    th = theta.as_strided((H,) + (1,) * (out.dim() - 3) + (1, 1), (theta.stride(0),) + (0,) * (out.dim() - 3) + (1, 1))
    d.e_theta_bs = (i64 * 3)(*_bstrides(th, nb))
    d.e_D = theta.shape[1] - 1
    d.sm_limit = int(sm_limit)
    d.tri_c = _TRI[c_tri]               # symmetric Gram: only that triangle is computed (the other is zero-filled)
    self._run_gemm(d, out, tag)

  # -- RBF operand prep / adjoint -------------------------------------------------------------
  def scale_rows(self, src, theta, dst, norms):
    _f32(src, 'src', contiguous=False)
    if src.dim() != 2 or src.stride(1) != 1:
      raise VargpError('scale_rows: src must be (R, D) with unit inner stride')
    R, D = src.shape
    H = theta.shape[0]
    _f32(theta, 'theta', contiguous=False)
    if theta.stride(1) != 1 or theta.shape[1] != D + 1:
      raise VargpError('scale_rows: theta must be (H, D+1) with unit inner stride')
    if tuple(dst.shape) != (H, R, D) or tuple(norms.shape) != (H, R):
      raise VargpError('scale_rows: bad output shapes')
    self._check(self.lib.vargp_scale_rows(src.data_ptr(), R, D, src.stride(0), theta.data_ptr(), H, theta.stride(0),
                                          _f32(dst, 'dst'), _f32(norms, 'norms'), self._stream(dst)), 'scale_rows')

  def rbf_bwd_prep(self, Kbar, K, rsum, csum, dsum=None):
    H, C, Pa, Pb = K.shape
    self._check(self.lib.vargp_rbf_bwd_prep(_f32(Kbar, 'Kbar'), _f32(K, 'K'), H, C, Pa, Pb, _f32(rsum, 'rsum'),
                                            None if csum is None else _f32(csum, 'csum'),
                                            None if dsum is None else _f32(dsum, 'dsum'), self._stream(K)),
                'rbf_bwd_prep')

  def rbf_bwd_finish(self, zs, Gz1, Gz2, r1, r2, theta, Z_bar, theta_bar, dg=None):
    H, C, P, D = zs.shape
    _f32(theta, 'theta', contiguous=False)
    self._check(self.lib.vargp_rbf_bwd_finish(
      _f32(zs, 'zs'), None if Gz1 is None else _f32(Gz1, 'Gz1'), None if Gz2 is None else _f32(Gz2, 'Gz2'),
      None if Gz1 is None else _f32(r1, 'r1'), None if Gz2 is None else _f32(r2, 'r2'),
      None if dg is None else _f32(dg, 'dg'),
      theta.data_ptr(), theta.stride(0), H, C, P, D, _f32(Z_bar, 'Z_bar'), _f32(theta_bar, 'theta_bar'),
      self._stream(zs)), 'rbf_bwd_finish')

  def rbf_bwd_xside(self, xs, csum, Gx, theta, theta_bar, x_bar):
    H, B, D = xs.shape
    C = 1 if Gx is None else Gx.shape[1]
    _f32(theta, 'theta', contiguous=False)
    self._check(self.lib.vargp_rbf_bwd_xside(
      _f32(xs, 'xs'), _f32(csum, 'csum'), None if Gx is None else _f32(Gx, 'Gx'), theta.data_ptr(), theta.stride(0),
      H, C, B, D, _f32(theta_bar, 'theta_bar'), None if x_bar is None else _f32(x_bar, 'x_bar'),
      self._stream(xs)), 'rbf_bwd_xside')

  # -- factorisations -------------------------------------------------------------------------
  @staticmethod
  def _mat_batch(t, name):
    """(.., n, n) tensor whose batch dims collapse to one stride; returns ptr, ld, batch stride, n, batch."""
    _f32(t, name, contiguous=False)
    n = t.shape[-1]
    if t.shape[-2] != n or t.stride(-1) != 1:
      raise VargpError(f'{name}: expected (..., n, n) with unit inner stride')
    bshape, bstr = list(t.shape[:-2]), list(t.stride()[:-2])
    batch = 1
    for s in bshape:
      batch *= s
    bs = 0
    if batch > 1:
      # collapse: require stride[i] == stride[i+1] * size[i+1]
      dims = [(s, st) for s, st in zip(bshape, bstr) if s > 1]
      for (s0, st0), (s1, st1) in zip(dims[:-1], dims[1:]):
        if st0 != st1 * s1:
          raise VargpError(f'{name}: batch dims are not collapsible: shape {bshape} strides {bstr}')
      bs = dims[-1][1]
    return t.data_ptr(), t.stride(-2), bs, n, batch

  def chol(self, K, L, jitter, info):
    ap, ald, abs_, n, batch = self._mat_batch(K, 'K')
    lp, lld, lbs, n2, batch2 = self._mat_batch(L, 'L')
    if (n, batch) != (n2, batch2) or info.numel() != batch or info.dtype != torch.int32:
      raise VargpError('chol: shape mismatch')
    self._check(self.lib.vargp_chol(ap, ald, abs_, lp, lld, lbs, n, batch, float(jitter), info.data_ptr(),
                                    self._stream(L)), 'chol')

  def trtri(self, L, W):
    lp, lld, lbs, n, batch = self._mat_batch(L, 'L')
    wp, wld, wbs, n2, batch2 = self._mat_batch(W, 'W')
    if (n, batch) != (n2, batch2):
      raise VargpError('trtri: shape mismatch')
    self._check(self.lib.vargp_trtri(lp, lld, lbs, wp, wld, wbs, n, batch, self._stream(W)), 'trtri')

  def chol_inv(self, K, L, W, jitter, info):
    """L = chol(K + jitter I), W = L^-1 (blocked, tensor-core driven for large n; see potrf_blocked.cu)."""
    ap, ald, abs_, n, batch = self._mat_batch(K, 'K')
    lp, lld, lbs, n2, batch2 = self._mat_batch(L, 'L')
    wp, wld, wbs, n3, batch3 = self._mat_batch(W, 'W')
    if not ((n, batch) == (n2, batch2) == (n3, batch3)) or info.numel() != batch or info.dtype != torch.int32:
      raise VargpError('chol_inv: shape mismatch')
    self._check(self.lib.vargp_chol_inv(ap, ald, abs_, lp, lld, lbs, wp, wld, wbs, n, batch, float(jitter),
                                        info.data_ptr(), self._stream(L)), 'chol_inv')

  def tc_persist_config(self, mode=-1):
    """Routing of the persistent form of the 1-CTA tensor-core GEMM (gemm_tcp.cu): 0 off, 1 above one wave of tiles (default),
    2 always, 3 from four waves; negative only queries.  Returns the previous mode."""
    return int(self.lib.vargp_tc_persist_config(int(mode)))

  def chol_cluster_config(self, min_n=-1, max_n=-1):
    """Routing window [min_n, max_n] of the cluster-cooperative kernel (potrf_cluster.cu; 0, 0 disables, negative only
    queries); returns the previous (min_n, max_n)."""
    r = int(self.lib.vargp_chol_cluster_config(int(min_n), int(max_n)))
    return r >> 32, r & 0xffffffff

  def chol_inv_cluster(self, K, L, W, jitter, info):
    """The cluster-cooperative kernel directly (32 < n <= 320); K may alias L or W."""
    ap, ald, abs_, n, batch = self._mat_batch(K, 'K')
    lp, lld, lbs, n2, batch2 = self._mat_batch(L, 'L')
    wp, wld, wbs, n3, batch3 = self._mat_batch(W, 'W')
    if not ((n, batch) == (n2, batch2) == (n3, batch3)) or info.numel() != batch or info.dtype != torch.int32:
      raise VargpError('chol_inv_cluster: shape mismatch')
    self._check(self.lib.vargp_chol_inv_cluster(ap, ald, abs_, lp, lld, lbs, wp, wld, wbs, n, batch, float(jitter),
                                                info.data_ptr(), self._stream(L)), 'chol_inv_cluster')

  def chol_cluster_wants(self, n):
    return bool(self.lib.vargp_chol_cluster_wants(int(n)))

  def chol_config(self, block=0, min_n=-1):
    """Set block size (0 keeps it, 1 = automatic: 256 with cluster-factored diagonal blocks, else 128) / minimum n of the
    blocked factorisation; returns (block, min_n) in effect (block 1 while automatic)."""
    r = int(self.lib.vargp_chol_config(int(block), int(min_n)))
    return r & 0xffffffff, r >> 32

  def tril_unpack(self, vec, out):
    C, M = out.shape[0], out.shape[-1]
    if tuple(vec.shape) != (C, M * (M + 1) // 2):
      raise VargpError('tril_unpack: shape mismatch')
    self._check(self.lib.vargp_tril_unpack(_f32(vec, 'vec'), C, M, _f32(out, 'out'), self._stream(out)),
                'tril_unpack')

  def tril_unpack_bwd(self, Lbar, vec, vec_bar):
    C, M = Lbar.shape[0], Lbar.shape[-1]
    self._check(self.lib.vargp_tril_unpack_bwd(_f32(Lbar, 'Lbar'), _f32(vec, 'vec'), C, M, _f32(vec_bar, 'vec_bar'),
                                               self._stream(vec)), 'tril_unpack_bwd')

  # -- KL(u) ----------------------------------------------------------------------------------
  def kl_fwd(self, W, T, nu, Lu_t, M, kl):
    H, C, P, _ = W.shape
    work = torch.empty(H * C * KL_CHUNKS, device=W.device, dtype=W.dtype)
    self._check(self.lib.vargp_kl_fwd(_f32(W, 'W'), _f32(T, 'T'), _f32(nu, 'nu'), _f32(Lu_t, 'Lu_t'), H, C, P, M,
                                      _f32(kl, 'kl'), work.data_ptr(), self._stream(W)), 'kl_fwd')

  def kl_bwd(self, W, T, nu, M, g_kl, Wbar, Tbar, nubar):
    H, C, P, _ = W.shape
    self._check(self.lib.vargp_kl_bwd(_f32(W, 'W'), _f32(T, 'T'), _f32(nu, 'nu'), _f32(g_kl, 'g_kl'), H, C, P, M,
                                      _f32(Wbar, 'Wbar'), _f32(Tbar, 'Tbar'), _f32(nubar, 'nubar'),
                                      self._stream(W)), 'kl_bwd')

  def kl_bwd_lu(self, Lu_t, g_kl, Lu_bar_t):
    C, M = Lu_t.shape[0], Lu_t.shape[-1]
    self._check(self.lib.vargp_kl_bwd_lu(_f32(Lu_t, 'Lu_t'), _f32(g_kl, 'g_kl'), C, M, _f32(Lu_bar_t, 'Lu_bar_t'),
                                         self._stream(Lu_t)), 'kl_bwd_lu')

  # -- whitening on the diagonal task blocks (shared-memory kernels, small M) --------------------
  def whiten_max_m(self, adjoint):
    return int(self.lib.vargp_whiten_max_m(int(bool(adjoint))))

  def whiten_work(self, H, C):
    return int(self.lib.vargp_whiten_fwd_work(H, C))

  def whiten_fwd(self, W, Lu_all, m_all, T, nu, N, kl, rect=None, work=None):
    """T_s = W_ss Lu_s, nu_s = W_ss m_s, N_ss += T_s T_s^T, kl += (1/H) sum KL_hc on the (h, c) rectangle `rect`."""
    H, C, P, _ = W.shape
    S, M = Lu_all.shape[0], Lu_all.shape[-1]
    h0, h1, c0, c1 = rect or (0, H, 0, C)
    if kl is not None and work is None:
      work = torch.zeros(self.whiten_work(H, C), device=W.device, dtype=W.dtype)
    self._check(self.lib.vargp_whiten_fwd(_f32(W, 'W'), _f32(Lu_all, 'Lu_all'), _f32(m_all, 'm_all'), H, C, S, M, P, h0, h1, c0, c1,
                                          _f32(T, 'T'), _f32(nu, 'nu'), _f32(N, 'N'), None if kl is None else _f32(kl, 'kl'),
                                          None if kl is None else _f32(work, 'work'), self._stream(W)), 'whiten_fwd')

  def whiten_bwd(self, W, T, nu, Lu_all, m_all, G, nubar, g_kl, Wbar, Lubar, mbar, s_grad0=0, rect=None):
    H, C, P, _ = W.shape
    S, M = Lu_all.shape[0], Lu_all.shape[-1]
    h0, h1, c0, c1 = rect or (0, H, 0, C)
    if tuple(Lubar.shape) != (H, S - s_grad0, C, M, M) or mbar.numel() != H * (S - s_grad0) * C * M:
      raise VargpError('whiten_bwd: Lubar / mbar must be (H, S - s_grad0, C, M, M) / (H, S - s_grad0, C, M[, 1])')
    self._check(self.lib.vargp_whiten_bwd(_f32(W, 'W'), _f32(T, 'T'), _f32(nu, 'nu'), _f32(Lu_all, 'Lu_all'), _f32(m_all, 'm_all'),
                                          _f32(G, 'G'), _f32(nubar, 'nubar'), None if g_kl is None else _f32(g_kl, 'g_kl'),
                                          H, C, S, M, P, h0, h1, c0, c1, s_grad0, _f32(Wbar, 'Wbar'), _f32(Lubar, 'Lubar'),
                                          _f32(mbar, 'mbar'), self._stream(W)), 'whiten_bwd')

  # -- predictive marginal --------------------------------------------------------------------
  def marginal_reduce(self, V, NV, nu, theta, f_mean, f_var):
    H, C, P, B = V.shape
    _f32(theta, 'theta', contiguous=False)
    self._check(self.lib.vargp_marginal_reduce(
      _f32(V, 'V'), _f32(NV, 'NV'), _f32(nu, 'nu'), theta.data_ptr(), theta.stride(0),
      theta.shape[1] - 1, H, C, P, B, _f32(f_mean, 'f_mean'), _f32(f_var, 'f_var'),
      self._stream(V)), 'marginal_reduce')

  def marginal_bwd_prep(self, V, NV, nu, g_mean, g_var, theta, Vbar, Vg, theta_bar):
    H, C, P, B = V.shape
    _f32(theta, 'theta', contiguous=False)
    self._check(self.lib.vargp_marginal_bwd_prep(
      _f32(V, 'V'), _f32(NV, 'NV'), _f32(nu, 'nu'), _f32(g_mean, 'g_mean'), _f32(g_var, 'g_var'),
      theta.data_ptr(), theta.stride(0), theta.shape[1] - 1, H, C, P, B, _f32(Vbar, 'Vbar'), _f32(Vg, 'Vg'),
      _f32(theta_bar, 'theta_bar'), self._stream(V)), 'marginal_bwd_prep')

  def sym_phi(self, X, mirror=False):
    n = X.shape[-1]
    self._check(self.lib.vargp_sym_phi(_f32(X, 'X'), n, X.numel() // (n * n), int(bool(mirror)), self._stream(X)),
                'sym_phi')

  # -- likelihood -----------------------------------------------------------------------------
  def nll_work(self, H, B):
    """floats of zero-initialised workspace `nll_fwd_bwd` needs (ticket counter + one partial sum per CTA)."""
    return int(self.lib.vargp_softmax_nll_work(H, B))

  def nll_fwd_bwd(self, f_mean, f_var, eps_f, y, nll, g_mean, g_var, work=None, gscale=1.0):
    H, F, C, B = eps_f.shape
    if y.dtype != torch.int64 or not y.is_cuda or not y.is_contiguous():
      raise VargpError('nll: y must be a contiguous CUDA int64 tensor')
    if work is None:
      work = torch.zeros(self.nll_work(H, B), device=f_mean.device, dtype=f_mean.dtype)
    elif work.numel() < self.nll_work(H, B):
      raise VargpError('nll: workspace too small')
    self._check(self.lib.vargp_softmax_nll(
      _f32(f_mean, 'f_mean'), _f32(f_var, 'f_var'), _f32(eps_f, 'eps_f'), y.data_ptr(), H, F, C, B,
      _f32(nll, 'nll'), _f32(g_mean, 'g_mean'), _f32(g_var, 'g_var'), float(gscale), _f32(work, 'work'),
      self._stream(f_mean)),
      'softmax_nll')

  def predict(self, f_mean, f_var, eps_f, probs):
    H, F, C, B = eps_f.shape
    self._check(self.lib.vargp_softmax_predict(
      _f32(f_mean, 'f_mean'), _f32(f_var, 'f_var'), _f32(eps_f, 'eps_f'), H, F, C, B, _f32(probs, 'probs'),
      self._stream(f_mean)), 'softmax_predict')


  # -- kernel hyper-parameters ----------------------------------------------------------------
  def hyper_fwd(self, log_mean, log_logvar, prior_log_mean, prior_log_logvar, eps, theta, kl):
    H, D1 = eps.shape
    self._check(self.lib.vargp_hyper_fwd(_f32(log_mean, 'log_mean'), _f32(log_logvar, 'log_logvar'),
                                         _f32(prior_log_mean, 'prior_log_mean'),
                                         _f32(prior_log_logvar, 'prior_log_logvar'), _f32(eps, 'eps'), H, D1,
                                         _f32(theta, 'theta'), None if kl is None else _f32(kl, 'kl'),
                                         self._stream(theta)), 'hyper_fwd')

  def hyper_bwd(self, log_mean, log_logvar, prior_log_mean, prior_log_logvar, eps, theta_bar, g_kl, m_bar, lv_bar):
    H, D1 = eps.shape
    self._check(self.lib.vargp_hyper_bwd(_f32(log_mean, 'log_mean'), _f32(log_logvar, 'log_logvar'),
                                         _f32(prior_log_mean, 'prior_log_mean'),
                                         _f32(prior_log_logvar, 'prior_log_logvar'), _f32(eps, 'eps'),
                                         None if theta_bar is None else _f32(theta_bar, 'theta_bar'),
                                         None if g_kl is None else _f32(g_kl, 'g_kl'), H, D1,
                                         _f32(m_bar, 'm_bar'), _f32(lv_bar, 'lv_bar'), self._stream(m_bar)), 'hyper_bwd')

  # -- fused training step: parameter plumbing ------------------------------------------------
  def step_assemble(self, z, u_mean, u_tril_vec, Zcat, m_last, Lu_last):
    C, M, D = z.shape
    P = Zcat.shape[1]
    self._check(self.lib.vargp_step_assemble(_f32(z, 'z'), _f32(u_mean, 'u_mean'), _f32(u_tril_vec, 'u_tril_vec'), C, M, D, P,
                                             _f32(Zcat, 'Zcat'), _f32(m_last, 'm_last'), _f32(Lu_last, 'Lu_last'),
                                             self._stream(z)), 'step_assemble')

  def step_grad_finish(self, Zbar, mbar, Lubar, Lu, u_tril_vec, g_kl_u, log_mean, log_logvar, prior_log_mean, prior_log_logvar,
                       eps, theta_bar, g_kl_h, z_g, um_g, ut_g, lm_g, llv_g):
    """mbar (H, C, M[, 1]) / Lubar (H, C, M, M): views whose trailing dims are contiguous (any stride along h)."""
    C, P, D = Zbar.shape
    H, M = Lubar.shape[0], Lubar.shape[-1]
    for t, nm in ((mbar, 'mbar'), (Lubar, 'Lubar')):
      _f32(t, nm, contiguous=False)
      if not t[0].is_contiguous():
        raise VargpError(f'step_grad_finish: {nm}[h] must be contiguous')
    self._check(self.lib.vargp_step_grad_finish(
      _f32(Zbar, 'Zbar'), mbar.data_ptr(), mbar.stride(0), Lubar.data_ptr(), Lubar.stride(0), _f32(Lu, 'Lu'),
      _f32(u_tril_vec, 'u_tril_vec'), None if g_kl_u is None else _f32(g_kl_u, 'g_kl_u'), _f32(log_mean, 'log_mean'),
      _f32(log_logvar, 'log_logvar'), _f32(prior_log_mean, 'prior_log_mean'), _f32(prior_log_logvar, 'prior_log_logvar'),
      _f32(eps, 'eps'), _f32(theta_bar, 'theta_bar'), _f32(g_kl_h, 'g_kl_h'), H, C, M, D, P,
      _f32(z_g, 'z_g'), _f32(um_g, 'um_g'), _f32(ut_g, 'ut_g'), _f32(lm_g, 'lm_g'), _f32(llv_g, 'llv_g'),
      self._stream(Zbar)), 'step_grad_finish')

  # -- optimizer ------------------------------------------------------------------------------
  def yogi_step(self, p, g, m, v, lr, b1, b2, eps, pows):
    self._check(self.lib.vargp_yogi_step(_f32(p, 'p'), _f32(g, 'g'), _f32(m, 'm'), _f32(v, 'v'), p.numel(),
                                         float(lr), float(b1), float(b2), float(eps), _f32(pows, 'pows'),
                                         self._stream(p)), 'yogi_step')


  def peer_buffer_floats(self, n):
    return int(self.lib.vargp_peer_buffer_floats(n))

  def peer_allreduce_yogi(self, peer_ptrs, rank, flat_g, p, m, v, lr, b1, b2, eps, pows, ctr, multicast=0):
    """flat_g <- sum over ranks (through the ranks' symmetric-memory staging buffers `peer_ptrs`; with `multicast`, the
    multicast mapping of those buffers, by in-switch reduction), then the Yogi update."""
    arr = (vp * len(peer_ptrs))(*[int(q) for q in peer_ptrs])
    if multicast:
      if ctr.dtype != torch.int32 or ctr.numel() < 4:
        raise VargpError('peer_allreduce_yogi: ctr must be an int32 tensor of 4 words')
      self._check(self.lib.vargp_peer_allreduce_yogi_nvls(arr, int(multicast), len(peer_ptrs), int(rank), p.numel(),
                                                          _f32(flat_g, 'flat_g'), _f32(p, 'p'), _f32(m, 'm'), _f32(v, 'v'),
                                                          float(lr), float(b1), float(b2), float(eps), _f32(pows, 'pows'),
                                                          ctr.data_ptr(), self._stream(p)), 'peer_allreduce_yogi_nvls')
      return
    if ctr.dtype != torch.int32 or ctr.numel() < 4:
      raise VargpError('peer_allreduce_yogi: ctr must be an int32 tensor of 4 words')
    self._check(self.lib.vargp_peer_allreduce_yogi(arr, len(peer_ptrs), int(rank), p.numel(), _f32(flat_g, 'flat_g'), _f32(p, 'p'),
                                                   _f32(m, 'm'), _f32(v, 'v'), float(lr), float(b1), float(b2), float(eps),
                                                   _f32(pows, 'pows'), ctr.data_ptr(), self._stream(p)), 'peer_allreduce_yogi')


_OPS = None


def set_ops(o):
  """Install a kernel interface (used by the CPU tests to inject tests/emu_ops.EmuOps)."""
  global _OPS
  _OPS = o


def get_ops():
  global _OPS
  if _OPS is None:
    _OPS = CudaOps()
  return _OPS
