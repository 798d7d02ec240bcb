"""Torch emulation of the ``vargp_b200.ops`` kernel interface  --  TEST INFRASTRUCTURE ONLY.

Lets the CPU test-suite (``-m "not gpu"``) exercise the *host schedule* of the hot path
(``vargp_b200/elbo.py``: which kernel is called on which view with which flags, and the hand-derived
backward) without a GPU, in fp32 or fp64.  Each method states the contract the CUDA kernel of the same
name implements; the GPU tests check the kernels against these same contracts.  Never imported by
``vargp_b200``: the product path fails loudly when the CUDA library is missing.
"""
import torch


def _tri(x, kind):
  if kind is None:
    return x
  return torch.tril(x) if kind == 'lower' else torch.triu(x)


class EmuOps:
  name = 'emu'

  def scale_rows(self, src, theta, dst, norms):
    D = src.shape[-1]
    dst.copy_(src.unsqueeze(0) * torch.exp(-theta[:, :D]).unsqueeze(1))
    norms.copy_((dst * dst).sum(-1))

  def rbf_gram(self, a, an, b, bn, theta, out, sym, tag=None, sm_limit=0, c_tri=None):
    g2 = torch.exp(2. * theta[:, -1]).view(-1, 1, 1, 1)
    dot = a @ b.transpose(-1, -2)
    val = g2 * torch.exp(dot - 0.5 * an.unsqueeze(-1) - 0.5 * bn.unsqueeze(-2))
    if sym:
      eye = torch.eye(a.shape[-2], dtype=torch.bool, device=a.device)
      val = torch.where(eye, g2.expand_as(val), val)
    if c_tri == 'lower':                 # contract: the other triangle is zero-filled
      val = torch.tril(val)
    elif c_tri == 'upper':
      val = torch.triu(val)
    out.copy_(val)

  def gemm(self, A, B, C, alpha=1., beta=0., a_tri=None, b_tri=None, c_tri=None, tag=None, zeroed=False, sm_limit=0):
    prod = alpha * (_tri(A, a_tri) @ _tri(B, b_tri))
    if c_tri is not None:
      prod = _tri(prod, c_tri)
    if beta == 0.:
      C.copy_(prod)
    else:
      C.copy_(prod + beta * C)

  def chol(self, K, L, jitter, info):
    eye = torch.eye(K.shape[-1], dtype=K.dtype, device=K.device)
    Lc, inf = torch.linalg.cholesky_ex(K + jitter * eye)
    L.copy_(Lc)
    info.copy_(inf.reshape(-1))

  def trtri(self, L, W):
    eye = torch.eye(L.shape[-1], dtype=L.dtype, device=L.device).expand_as(L)
    W.copy_(torch.linalg.solve_triangular(L, eye, upper=False))

  def chol_inv(self, K, L, W, jitter, info):
    self.chol(K, L, jitter, info)
    self.trtri(L, W)

  def kl_fwd(self, W, T, nu, Lu_t, M, kl):
    H = W.shape[0]
    P = W.shape[-1]
    wd = W.diagonal(dim1=-2, dim2=-1)[..., P - M:]
    val = -wd.log().sum() - H * Lu_t.diagonal(dim1=-2, dim2=-1).log().sum() \
          + 0.5 * ((T[:, :, -1] ** 2).sum() + (nu[..., P - M:] ** 2).sum() - M * W.shape[0] * W.shape[1])
    kl.add_(val / H)

  def kl_bwd(self, W, T, nu, M, g_kl, Wbar, Tbar, nubar):
    H = W.shape[0]
    P = W.shape[-1]
    s = g_kl / H
    Tbar[:, :, -1] += s * T[:, :, -1]
    nubar[..., P - M:] += s * nu[..., P - M:]
    wd = W.diagonal(dim1=-2, dim2=-1)[..., P - M:]
    Wbar.diagonal(dim1=-2, dim2=-1)[..., P - M:] -= s / wd

  def kl_bwd_lu(self, Lu_t, g_kl, Lu_bar_t):
    Lu_bar_t.diagonal(dim1=-2, dim2=-1).sub_(g_kl / Lu_t.diagonal(dim1=-2, dim2=-1))

  def marginal_reduce(self, V, NV, nu, theta, f_mean, f_var):
    g2 = torch.exp(2. * theta[:, -1]).view(-1, 1, 1)
    f_mean.copy_((V * nu.unsqueeze(-1)).sum(-2))
    f_var.copy_(g2 + (V * (NV - V)).sum(-2))

  def marginal_bwd_prep(self, V, NV, nu, g_mean, g_var, theta, Vbar, Vg, theta_bar):
    """Vbar may alias NV."""
    gv = g_var.unsqueeze(-2)
    theta_bar[:, -1] += 2. * torch.exp(2. * theta[:, -1]) * g_var.sum((1, 2))
    Vbar.copy_(nu.unsqueeze(-1) * g_mean.unsqueeze(-2) + 2. * gv * (NV - V))
    Vg.copy_(V * gv)

  def sym_phi(self, X, mirror=False):
    low = torch.tril(X, -1)
    d = torch.diag_embed(X.diagonal(dim1=-2, dim2=-1))
    X.copy_((1.0 if mirror else 0.5) * (low + low.transpose(-1, -2) + d))

  def rbf_bwd_prep(self, Kbar, K, rsum, csum, dsum=None):
    Kbar.mul_(K)
    if dsum is not None:
      d = Kbar.diagonal(dim1=-2, dim2=-1)
      dsum.copy_(d)
      d.zero_()
    rsum.copy_(Kbar.sum(-1))
    if csum is not None:
      csum.add_(Kbar.sum(-2).sum(1))

  def rbf_bwd_finish(self, zs, Gz1, Gz2, r1, r2, theta, Z_bar, theta_bar, dg=None):
    D = zs.shape[-1]
    z0 = torch.zeros_like(zs)
    r1 = torch.zeros_like(zs[..., 0]) if Gz1 is None else r1
    r2 = torch.zeros_like(zs[..., 0]) if Gz2 is None else r2
    g1 = z0 if Gz1 is None else Gz1
    g2 = z0 if Gz2 is None else Gz2
    zsb = -(r1 + 2. * r2).unsqueeze(-1) * zs + g1 + 2. * g2
    Z_bar.copy_((zsb * torch.exp(-theta[:, :D]).view(-1, 1, 1, D)).sum(0))
    theta_bar[:, :D] += (-zs * zsb - zs * g1).sum((1, 2))
    theta_bar[:, D] += 2. * (r1 + r2).sum((1, 2))
    if dg is not None:
      theta_bar[:, D] += 2. * dg.sum((1, 2))

  def rbf_bwd_xside(self, xs, csum, Gx, theta, theta_bar, x_bar):
    D = xs.shape[-1]
    theta_bar[:, :D] += (csum.unsqueeze(-1) * xs * xs).sum(1)
    if x_bar is not None:
      xsb = -csum.unsqueeze(-1) * xs + Gx.sum(1)
      x_bar.copy_((xsb * torch.exp(-theta[:, :D]).unsqueeze(1)).sum(0))

  def tril_unpack(self, vec, out):
    M = out.shape[-1]
    r, c = torch.tril_indices(M, M)
    out.zero_()
    out[..., r, c] = vec
    d = out.diagonal(dim1=-2, dim2=-1)
    d.copy_(torch.nn.functional.softplus(d))

  def tril_unpack_bwd(self, Lbar, vec, vec_bar):
    M = Lbar.shape[-1]
    r, c = torch.tril_indices(M, M)
    g = Lbar[..., r, c].clone()
    on = (r == c)
    g[..., on] = g[..., on] * torch.sigmoid(vec[..., on])
    vec_bar.copy_(g)

  def nll_fwd_bwd(self, f_mean, f_var, eps_f, y, nll, g_mean, g_var):
    H, F, C, B = eps_f.shape
    sd = f_var.sqrt().unsqueeze(1)
    f = f_mean.unsqueeze(1) + sd * eps_f
    logp = torch.log_softmax(f, dim=-2)
    idx = y.view(1, 1, 1, B).expand(H, F, 1, B)
    nll.add_(-logp.gather(2, idx).sum() / (H * F))
    gf = logp.exp()
    gf.scatter_add_(2, idx, -torch.ones_like(gf[:, :, :1]))
    gf = gf / (H * F)
    g_mean.copy_(gf.sum(1))
    g_var.copy_((gf * eps_f).sum(1) / (2. * sd.squeeze(1)))

  def predict(self, f_mean, f_var, eps_f, probs):
    H, F, C, B = eps_f.shape
    f = f_mean.unsqueeze(1) + f_var.sqrt().unsqueeze(1) * eps_f
    probs.copy_(torch.softmax(f, dim=-2).sum((0, 1)).T / (H * F))

  def hyper_fwd(self, log_mean, log_logvar, prior_log_mean, prior_log_logvar, eps, theta, kl):
    theta.copy_(log_mean + torch.exp(0.5 * log_logvar) * eps)
    if kl is not None:
      dl, dm = log_logvar - prior_log_logvar, log_mean - prior_log_mean
      kl.copy_((0.5 * (dl.exp() + dm * dm * torch.exp(-prior_log_logvar) - 1. - dl)).sum())

  def hyper_bwd(self, log_mean, log_logvar, prior_log_mean, prior_log_logvar, eps, theta_bar, g_kl, m_bar, lv_bar):
    m_bar.zero_()
    lv_bar.zero_()
    if theta_bar is not None:
      m_bar += theta_bar.sum(0)
      lv_bar += 0.5 * torch.exp(0.5 * log_logvar) * (theta_bar * eps).sum(0)
    if g_kl is not None:
      m_bar += g_kl * (log_mean - prior_log_mean) * torch.exp(-prior_log_logvar)
      lv_bar += 0.5 * g_kl * (torch.exp(log_logvar - prior_log_logvar) - 1.)

  def yogi_step(self, p, g, m, v, lr, b1, b2, eps, pows):
    pows[0] *= b1
    pows[1] *= b2
    bc1, bc2 = 1. - pows[0], 1. - pows[1]
    g2 = g * g
    m.mul_(b1).add_(g, alpha=1. - b1)
    v.sub_((1. - b2) * torch.sign(v - g2) * g2)
    p.sub_((lr / bc1) * m / (v.sqrt() / bc2.sqrt() + eps))

  def step_assemble(self, z, u_mean, u_tril_vec, Zcat, m_last, Lu_last):
    M = z.shape[1]
    Zcat[:, Zcat.shape[1] - M:] = z
    m_last.copy_(u_mean.reshape(m_last.shape))
    self.tril_unpack(u_tril_vec, Lu_last)

  def step_grad_finish(self, Zbar, mbar, Lubar, Lu, u_tril_vec, g_kl_u, log_mean, log_logvar, prior_log_mean, prior_log_logvar,
                       eps, theta_bar, g_kl_h, z_g, um_g, ut_g, lm_g, llv_g):
    M = Lu.shape[-1]
    z_g.copy_(Zbar[:, Zbar.shape[1] - M:])
    um_g.copy_(mbar.sum(0).reshape(um_g.shape))
    Lb = Lubar.sum(0)
    if g_kl_u is not None:
      Lb = Lb.clone()
      self.kl_bwd_lu(Lu, g_kl_u[0], Lb)
    self.tril_unpack_bwd(Lb, u_tril_vec, ut_g)
    self.hyper_bwd(log_mean, log_logvar, prior_log_mean, prior_log_logvar, eps, theta_bar, g_kl_h[0], lm_g, llv_g)

  # -- whitening on the diagonal task blocks (contracts of whiten.cu) --
  def whiten_max_m(self, adjoint):
    return 96 if adjoint else 128

  def whiten_fwd(self, W, Lu_all, m_all, T, nu, N, kl, rect=None, work=None):
    H, C, P, _ = W.shape
    S, M = Lu_all.shape[0], Lu_all.shape[-1]
    h0, h1, c0, c1 = rect or (0, H, 0, C)
    for s in range(S):
      sl = slice(s * M, (s + 1) * M)
      Wss = W[h0:h1, c0:c1, sl, sl]
      Ts = Wss @ Lu_all[s, c0:c1]
      T[h0:h1, c0:c1, s] = Ts
      nu[h0:h1, c0:c1, sl] = (Wss @ m_all[s, c0:c1].unsqueeze(-1)).squeeze(-1)
      N[h0:h1, c0:c1, sl, sl] += Ts @ Ts.transpose(-1, -2)
    if kl is not None:
      wd = W[h0:h1, c0:c1].diagonal(dim1=-2, dim2=-1)[..., P - M:]
      val = -wd.log().sum() - (h1 - h0) * Lu_all[S - 1, c0:c1].diagonal(dim1=-2, dim2=-1).log().sum() \
            + 0.5 * ((T[h0:h1, c0:c1, -1] ** 2).sum() + (nu[h0:h1, c0:c1, P - M:] ** 2).sum() - M * (h1 - h0) * (c1 - c0))
      kl.add_(val / H)

  def whiten_bwd(self, W, T, nu, Lu_all, m_all, G, nubar, g_kl, Wbar, Lubar, mbar, s_grad0=0, rect=None):
    H, C, P, _ = W.shape
    S, M = Lu_all.shape[0], Lu_all.shape[-1]
    h0, h1, c0, c1 = rect or (0, H, 0, C)
    mb = mbar.view(H, S - s_grad0, C, M)
    for s in range(S):
      sl = slice(s * M, (s + 1) * M)
      k = (g_kl.reshape(-1)[0] / H) if (g_kl is not None and s == S - 1) else 0.
      Ts, Wss = T[h0:h1, c0:c1, s], W[h0:h1, c0:c1, sl, sl]
      Tb = torch.tril(2. * G[h0:h1, c0:c1, sl, sl] @ Ts) + k * Ts
      nb = nubar[h0:h1, c0:c1, sl] + k * nu[h0:h1, c0:c1, sl]
      upd = torch.tril(Tb @ Lu_all[s, c0:c1].transpose(-1, -2) + nb.unsqueeze(-1) * m_all[s, c0:c1].unsqueeze(-2))
      if g_kl is not None and s == S - 1:
        upd = upd - torch.diag_embed(k / Wss.diagonal(dim1=-2, dim2=-1))
      Wbar[h0:h1, c0:c1, sl, sl] += upd
      if s >= s_grad0:
        Lubar[h0:h1, s - s_grad0, c0:c1] = torch.tril(Wss.transpose(-1, -2) @ Tb)
        mb[h0:h1, s - s_grad0, c0:c1] = (Wss.transpose(-1, -2) @ nb.unsqueeze(-1)).squeeze(-1)
