"""KL-divergence upper bound of the minibatch ELBO: drop-in for
elbo_functions.minibatch_KLD_upper_bound (elbo_functions.py:118-193) and
minibatch_KLD_upper_bound_iter (:196-285): same arguments, same returns
(kld_total attached to autograd, grad_m, grad_H).

Data flow (DESIGN.md has the derivation):
  1. M x M pre-stage, replicated, float64:  K0zz + eps I, Cholesky, iK, iH, w = iK m,
     G = iK H iK - iK.
  2. streaming stage, CUDA kernels behind the C ABI (hlvae_kl_subject, hlvae_kl_panel):
     per-latent accumulators S, p, dJ/dw, the scalars A, B + sum(iB*K0_st), C, F, and the
     gradients of mu, log_v, Z and kernel hyper-parameters, all in one pass.
  3. (data parallel) one all-reduce of the accumulator buffer.
  4. M x M post-stage: D, E, kld_qu_pu, kld_total, natural-gradient pieces.
"""
from __future__ import annotations

import torch

from . import _lib, config
from .kernels import compile_spec, constrained_pair
from .subjects import SubjectLayout

N_SM = 148
PANEL_WAVES = 4        # waves of kl_panel CTAs per launch (see _subjects_per_chunk)

_side_streams = {}
last_terms = None      # see config.keep_terms


def _side_stream(device):
    """A stream other than the current one, on which hlvae_kl_subject (T x T work, needs nothing from
    the M x M pre-stage) runs next to hlvae_mxm_pre (2 L CTAs only)."""
    cur = torch.cuda.current_stream(device)
    pool = _side_streams.setdefault(str(device), [torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)])
    return pool[0] if pool[0] != cur else pool[1]


def _noise_vector(likelihood, L, device):
    nz = likelihood.noise_covar.noise if hasattr(likelihood, "noise_covar") else likelihood.noise
    return nz.detach().to(device=device, dtype=torch.float64).reshape(-1).expand(L).contiguous()


def _raise_status(status, info):
    st = status.tolist()
    if st[0] == _lib.STATUS_NOT_PD:
        what = {-1: "K0zz + eps I", -2: "H", -3: "H", -4: "H^-1 + lr (grad_H + grad_H^T)"}.get(
            st[2], f"B_s of subject {st[2]}")
        raise RuntimeError(f"hlvae_b200: cholesky: {what} (latent dim {st[1]}) is not positive-definite")
    if st[0] == _lib.STATUS_T_TOO_LARGE:
        raise RuntimeError(f"hlvae_b200: subject {st[2]} has more than {_lib.TMAX} rows")
    if info is not None and bool((info != 0).any()):
        raise RuntimeError("hlvae_b200: cholesky: K0zz + eps I or H is not positive-definite")


class _KLD(torch.autograd.Function):
    """(mu, log_v, z, m, H, constrained hyper-parameters) -> kld_total, grad_m, grad_H.

    Every gradient is produced in the forward pass (the streaming kernels need only quantities
    of the M x M pre-stage, see DESIGN.md "single-pass gradient"), so backward is a scaling."""

    @staticmethod
    def forward(ctx, mu, log_v, z, m, H, os0, ls0, os1, ls1, noise, x, fs0, fs1, layout, scale, const, eps,
                natural_gradient, out_shape):
        dev = x.device
        N, L = mu.shape
        M, Q = z.shape[-2], z.shape[-1]
        if mu.dtype != log_v.dtype:
            raise TypeError("mu and log_v must share a dtype")
        dcode = _lib.dtype_code(mu)
        f64 = dict(dtype=torch.float64, device=dev)
        mu_c, lv_c = mu.detach().contiguous(), log_v.detach().contiguous()
        x_c = x.detach().to(torch.float64).contiguous()
        z_c = z.detach().to(torch.float64).contiguous()
        m_c = m.detach().to(torch.float64).reshape(L, M).contiguous()
        H_c = H.detach().to(torch.float64).contiguous()
        os0c, ls0c = os0.detach().contiguous(), ls0.detach().contiguous()
        os1c, ls1c = os1.detach().contiguous(), ls1.detach().contiguous()
        st = _lib.stream_ptr()
        ws = _lib.workspace(L, M, dev)
        status = torch.zeros(4, dtype=torch.int32, device=dev)

        # ---- buffers
        mats = torch.empty(4, L, M, M, **f64)          # iK, iH, G, (later) dkld/dK0zz / c0
        iK, iH, G, gK = mats[0], mats[1], mats[2], mats[3]
        vecs = torch.empty(2, L, M, **f64)             # w, dkld/dm
        w, gm = vecs[0], vecs[1]
        pre = torch.empty(L, 4, **f64)
        off = _lib.acc_layout(L, M, Q)
        acc = torch.zeros(off["total"] + 2 + L, **f64) # tail: kld_total, arrival counter, per-latent kld terms
        full = layout.n_rows == N
        g_mu = torch.empty_like(mu_c) if full else torch.zeros_like(mu_c)
        g_lv = torch.empty_like(lv_c) if full else torch.zeros_like(lv_c)
        binv = torch.empty(L, max(layout.tt_total, 1), **f64)

        # ---- 1. M x M pre-stage (replicated): iK, iH, w = iK m, G = iK H iK - iK.  2 L CTAs of 166 KB shared
        # memory each: launched FIRST, so that they are resident before the per-subject stage floods the block
        # scheduler (launched second, its small CTAs never leave an SM enough shared memory for one of these,
        # and the two kernels serialise)
        cur = torch.cuda.current_stream(dev)
        side = _side_stream(dev) if (config.overlap and layout.n_subj > 0) else None
        if side is not None:
            side.wait_stream(cur)
        _lib.call("hlvae_mxm_pre", fs0.cspec, _lib.ptr(os0c), _lib.ptr(ls0c), L, Q, M, _lib.ptr(z_c), eps,
                  _lib.ptr(m_c), _lib.ptr(H_c), _lib.ptr(iK), _lib.ptr(iH), _lib.ptr(w), _lib.ptr(G), _lib.ptr(pre),
                  _lib.ptr(ws), _lib.ptr(status), st)
        # ---- 2a. per-subject T x T stage on a side stream (independent of the M x M pre-stage)
        if layout.n_subj > 0:
            with torch.cuda.stream(side if side is not None else cur):
                _lib.call("hlvae_kl_subject", fs0.cspec, _lib.ptr(os0c), _lib.ptr(ls0c), fs1.cspec, _lib.ptr(os1c),
                          _lib.ptr(ls1c), _lib.ptr(noise), L, Q, _lib.ptr(x_c), Q, _lib.ptr(layout.row_idx),
                          _lib.ptr(layout.subj_ptr), _lib.ptr(layout.tt_ptr), layout.n_subj, max(layout.t_max, 1),
                          _lib.ptr(lv_c), L, dcode, _lib.ptr(binv), binv.shape[1], _lib.ptr(acc), M, _lib.ptr(g_lv),
                          scale, _lib.ptr(status), _lib.stream_ptr())
        if side is not None:
            cur.wait_stream(side)
        # ---- 2b. streaming stage over the minibatch rows
        if layout.n_subj > 0:
            rp = _row_panel(layout, M)
            spc = _subjects_per_chunk(layout.n_subj, layout.t_max, L, M, rp=rp)
            _lib.call("hlvae_kl_panel", fs0.cspec, _lib.ptr(os0c), _lib.ptr(ls0c), fs1.cspec, _lib.ptr(os1c),
                      _lib.ptr(ls1c), L, Q, M, _lib.ptr(x_c), Q, _lib.ptr(z_c), _lib.ptr(layout.row_idx),
                      _lib.ptr(layout.subj_ptr), _lib.ptr(layout.tt_ptr), layout.n_subj, spc, _lib.ptr(mu_c), L,
                      dcode, _lib.ptr(w), _lib.ptr(G), _lib.ptr(binv), binv.shape[1], _lib.ptr(acc), _lib.ptr(g_mu),
                      None, scale, _lib.ptr(status), rp, st)
        # ---- 3. data parallel: one all-reduce of every accumulator (S, p, scalars, replicated-parameter grads)
        if config.process_group is not None:
            torch.distributed.all_reduce(acc, group=config.process_group)

        def view(name, *shape):
            n = 1
            for s_ in shape:
                n *= s_
            return acc[off[name]:off[name] + n].view(*shape)

        S, p, gw, scal = view("S", L, M, M), view("p", L, M), view("gw", L, M), view("scal", L, _lib.NSCAL)
        kld = acc[off["total"]:]
        # ---- 4. M x M post-stage: kld_total, natural-gradient pieces, d kld / d{K0zz, m, H}
        gH = torch.empty(L, M, M, **f64)
        if natural_gradient:
            ng_H_c = torch.empty(L, M, M, **f64)
            ng_m_c = torch.empty(L, M, 1, **f64)
        _lib.call("hlvae_mxm_post", L, M, scale, const, _lib.ptr(iK), _lib.ptr(iH), _lib.ptr(H_c), _lib.ptr(m_c),
                  _lib.ptr(w), _lib.ptr(G), _lib.ptr(pre), _lib.ptr(S), _lib.ptr(p), _lib.ptr(gw), _lib.ptr(scal),
                  _lib.ptr(kld), _lib.ptr(gK), _lib.ptr(gH), _lib.ptr(gm),
                  _lib.ptr(ng_m_c) if natural_gradient else None, _lib.ptr(ng_H_c) if natural_gradient else None,
                  _lib.ptr(ws), st)
        # d kld / dK0zz -> Z and the K0 hyper-parameters, accumulated (over c0) onto the streaming gradients
        gZ = view("gZ", L, M, Q)
        _lib.call("hlvae_kernel_eval_bwd", fs0.cspec, _lib.ptr(os0c), _lib.ptr(ls0c), L, Q, _lib.ptr(z_c), M, Q,
                  M * Q, _lib.ptr(z_c), M, Q, M * Q, _lib.ptr(gK), _lib.ptr(view("gos0", _lib.MAX_COMPS, L)),
                  _lib.ptr(view("gls0", _lib.MAX_COMPS, L)), _lib.ptr(gZ), _lib.ptr(gZ), st)
        if config.check_errors:
            _raise_status(status, None)
        if config.keep_terms:
            global last_terms
            last_terms = dict(scal=scal.clone(), pre=pre.clone(), S=S.clone(), p=p.clone(), iK=iK.clone(), G=G.clone(),
                              H=H_c.clone(), iH=iH.clone(), kld_per_latent=kld[2:2 + L].clone(), scale=scale)
        nc0, nc1 = fs0.ncomp, fs1.ncomp
        grads = acc[off["gZ"]:off["total"]] * scale          # one launch for every replicated-parameter gradient
        o0 = off["gZ"]

        def gview(name, *shape):
            n = 1
            for s_ in shape:
                n *= s_
            return grads[off[name] - o0:off[name] - o0 + n].view(*shape)

        ctx.save_for_backward(g_mu, g_lv, gview("gZ", L, M, Q).to(z.dtype), gm.reshape(m.shape).to(m.dtype),
                              gH.to(H.dtype), gview("gos0", _lib.MAX_COMPS, L)[:nc0],
                              gview("gls0", _lib.MAX_COMPS, L)[:nc0], gview("gos1", _lib.MAX_COMPS, L)[:nc1],
                              gview("gls1", _lib.MAX_COMPS, L)[:nc1])
        grad_m = ng_m_c if natural_gradient else None
        grad_H = ng_H_c if natural_gradient else None
        ctx.mark_non_differentiable(iH, *[t for t in (grad_m, grad_H) if t is not None])
        return kld[:1].clone().reshape(out_shape), grad_m, grad_H, iH

    @staticmethod
    def backward(ctx, g_kld, g_gm, g_gH, g_iH):
        g = g_kld.reshape(())
        out = []
        for t in ctx.saved_tensors:
            out.append((g * t).to(t.dtype) if t.numel() else t)
        return tuple(out) + (None,) * 10


def _kld(covar_module0, covar_module1, likelihood, latent_dim, m, H, train_xt, mu, log_v, z, layout, scale, const,
         natural_gradient, eps, out_shape):
    if not train_xt.is_cuda:
        raise RuntimeError("hlvae_b200: the KL upper bound runs on CUDA tensors only (no CPU fallback)")
    fs0, fs1 = compile_spec(covar_module0), compile_spec(covar_module1)
    L = latent_dim
    os0, ls0, os1, ls1 = constrained_pair(fs0, fs1, L, train_xt.device)
    noise = _noise_vector(likelihood, L, train_xt.device)
    kld, grad_m, grad_H, iH = _KLD.apply(mu, log_v, z, m, H, os0, ls0, os1, ls1, noise, train_xt, fs0, fs1, layout,
                                         float(scale), float(const), float(eps), bool(natural_gradient), out_shape)
    if grad_H is not None:
        # H^-1 of this step rides along with grad_H so that natural_gradient_update (training.py:131-132) need
        # not factorise the same H again; valid only while H is unchanged (checked there)
        grad_H._hlvae_iH = (iH, H.data_ptr(), H._version)
    return kld, grad_m, grad_H


def minibatch_KLD_upper_bound(covar_module0, covar_module1, likelihood, latent_dim, m, H, train_xt, mu, log_v, z,
                              P_tot, P_batch, T, natural_gradient, eps, layout=None):
    """elbo_functions.py:118-193 (fixed T).  Rows must be subject-contiguous, T per subject."""
    if layout is None:
        layout = SubjectLayout.fixed(train_xt.shape[0], T, train_xt.device)
    return _kld(covar_module0, covar_module1, likelihood, latent_dim, m, H, train_xt, mu, log_v, z, layout,
                P_tot / P_batch, latent_dim * P_tot * T / 2, natural_gradient, eps, ())


def minibatch_KLD_upper_bound_iter(covar_module0, covar_module1, likelihood, latent_dim, m, H, train_xt, mu, log_v, z,
                                   P, P_in_current_batch, N, natural_gradient, id_covariate, eps, layout=None):
    """elbo_functions.py:196-285 (varying T).  `layout` (optional) skips the id grouping when the
    caller already knows the subject lengths (SubjectLayout.from_lengths)."""
    if layout is None:
        layout = SubjectLayout.from_ids(train_xt[:, id_covariate])
    return _kld(covar_module0, covar_module1, likelihood, latent_dim, m, H, train_xt, mu, log_v, z, layout,
                P / P_in_current_batch, latent_dim * N / 2, natural_gradient, eps, (1,))


PANEL_COST_40 = 0.60      # device time of a 40-row panel (two CTAs per SM) relative to a 64-row one, measured (T = 20)
PANEL_COST_M128 = {32: 5.2, 48: 7.75, 64: 9.6}     # relative device time of a panel at 64 < M <= 128, measured (T = 20)


def _row_panel(layout, M):
    """Rows per panel for hlvae_kl_panel, chosen per minibatch from how its whole subjects pack (host arithmetic on
    the lengths): M <= 32 -> 64; 32 < M <= 64 -> 40 (the two-CTAs-per-SM shape) when the subjects need so few more
    40-row panels than 64-row ones that the cheaper panel wins (fixed T = 20: 2 subjects per 40 rows against 3 per
    64), else 64; 64 < M <= 128 -> the cheapest of 32 / 48 / 64."""
    import os
    forced = os.environ.get("HLVAE_PANEL_RP")
    if forced is not None:
        return int(forced)
    if M <= 32 or layout.n_subj == 0:
        return 64
    if M <= 64:
        if layout.t_max > 40:
            return 64
        return 40 if layout.panels(40) * PANEL_COST_40 < layout.panels(64) else 64
    fits = [rp for rp in (32, 48, 64) if rp >= layout.t_max] or [64]
    return min(fits, key=lambda rp: layout.panels(rp) * PANEL_COST_M128[rp])


def _two_ctas_per_sm(M, rp):
    return M <= 32 or (M <= 64 and rp == 40)


def _subjects_per_chunk(n_subj, t_max, L, M, waves=None, rp=0):
    """Subjects per hlvae_kl_panel CTA.  A CTA walks its chunk in row panels of `rp` rows (see _row_panel) that hold
    whole subjects, so a chunk should be a whole number of full panels (a
    trailing panel with one subject costs nearly as much as a full one), and L * n_chunks CTAs should fill a whole
    number of waves of resident CTAs."""
    import os
    rp = rp or 64
    two = _two_ctas_per_sm(M, rp)
    spp = max(1, rp // max(int(t_max), 1))                    # subjects per full panel
    # (measured: 8 waves for the two-CTA shape at M = 64, 4 waves everywhere else incl. the two-CTA shape at M <= 32)
    waves = int(os.environ.get("HLVAE_PANEL_WAVES", waves or (2 * PANEL_WAVES if (two and M > 32) else PANEL_WAVES)))
    n_panels = (n_subj + spp - 1) // spp
    # at least `waves` waves of CTAs, and no more than ~30 panels per CTA (large batches: more, shorter CTAs balance
    # the tail better - measured at 64 000 rows)
    n_chunks = max(1, min(max((N_SM * waves) // L, (n_panels + 29) // 30), n_panels))
    spc = (n_subj + n_chunks - 1) // n_chunks
    return ((spc + spp - 1) // spp) * spp


def natural_gradient_update(m, H, grad_m, grad_H, lr):
    """training.py:130-137 as one kernel (hlvae_natgrad_update): returns detached (m_new, H_new)."""
    L, M = H.shape[0], H.shape[-1]
    dev = H.device
    m_c = m.detach().to(torch.float64).reshape(L, M).contiguous()
    H_c = H.detach().to(torch.float64).contiguous()
    gm = grad_m.detach().to(torch.float64).reshape(L, M).contiguous()
    gH = grad_H.detach().to(torch.float64).contiguous()
    m_new, H_new = torch.empty_like(m_c), torch.empty_like(H_c)
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    ws = _lib.workspace(L, M, dev)
    iH = None
    stash = getattr(grad_H, "_hlvae_iH", None)
    if stash is not None and stash[1] == H.data_ptr() and stash[2] == H._version and H.dtype == torch.float64:
        iH = stash[0]
    _lib.call("hlvae_natgrad_update", L, M, float(lr), _lib.ptr(m_c), _lib.ptr(H_c), _lib.ptr(iH), _lib.ptr(gm), _lib.ptr(gH),
              _lib.ptr(m_new), _lib.ptr(H_new), _lib.ptr(ws), _lib.ptr(status), _lib.stream_ptr())
    if config.check_errors:
        _raise_status(status, None)
    return m_new.reshape(m.shape).to(m.dtype), H_new.to(H.dtype)
