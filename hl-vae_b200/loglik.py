"""Heterogeneous masked log-likelihoods with the reference's operator surface.

Two entry levels, both backed by the fused CUDA kernel (hlvae_loglik_fwd / _bwd):
  * `loglik_real / pos / cat / ordinal / count` - same signature and returned dict as
    HL_VAE/loglik.py:27,73,124,149,191, looked up by name from HLVAE.py:388;
  * `loglik_and_reconstruction` - a replacement for the method HLVAE.py:381-414 that
    evaluates every type group in ONE launch (no per-type gather/scatter) and also
    produces the monitoring transforms of read_functions.statistics (:268-302) and
    discrete_variables_transformation (:221-235).
There is no CPU path: tensors must be on a CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributions as td

from . import _lib

# Draw `samples` with the same torch.distributions calls as the reference so that the global
# RNG stream advances identically (training.py never reads them).  The fused method skips them.
PRODUCE_SAMPLES = True

SUPPORTED_TYPES = tuple(_lib.VAR_KINDS)        # variable types the fused kernels evaluate


class VarLayout:
    """Per-variable descriptors of the packed [N, E_x] data / [N, P_theta] parameter layout
    (read_functions.py:144-173, logvar_network=False)."""

    def __init__(self, types, device):
        kinds, ncls, dcol, pcol, gpos = [], [], [], [], []
        e = p = 0
        counters = {}
        for kind, C in types:
            C = int(C) if kind in ("cat", "ordinal") else 1
            if kind not in _lib.VAR_KINDS:
                raise NotImplementedError(f"likelihood type '{kind}' is not supported by the fused kernel")
            if C > _lib.MAX_CLASS:
                raise NotImplementedError(f"at most {_lib.MAX_CLASS} classes per variable")
            kinds.append(_lib.VAR_KINDS[kind]); ncls.append(C); dcol.append(e); pcol.append(p)
            gpos.append(counters.get(kind, 0))
            counters[kind] = gpos[-1] + 1
            e += C
            p += C
        self.types = [(k, int(c)) for k, c in types]
        self.D, self.E_x, self.P_theta = len(kinds), e, p
        self.max_class = max(ncls) if ncls else 1
        self.device = device
        i32 = dict(dtype=torch.int32, device=device)
        self.var_kind = torch.tensor(kinds, **i32)
        self.var_nclass = torch.tensor(ncls, **i32)
        self.var_dcol = torch.tensor(dcol, **i32)
        self.var_pcol = torch.tensor(pcol, **i32)
        kinds_np = np.array(kinds)
        gpos_np = np.array(gpos)
        self.idx = {k: torch.tensor(np.nonzero(kinds_np == v)[0], dtype=torch.long, device=device)
                    for k, v in _lib.VAR_KINDS.items()}
        self.gpos = {k: torch.tensor(gpos_np[kinds_np == v], dtype=torch.long, device=device)
                     for k, v in _lib.VAR_KINDS.items()}
        self.pcol_host, self.dcol_host, self.ncls_host = pcol, dcol, ncls

    @staticmethod
    def from_types_info(types_info, device):
        types = [(t['type'], int(t['nclass'])) for t in types_info['types_dict']]
        return VarLayout(types, device)

    def vparam(self, log_vy_real=None, log_vy_pos=None, norm_real=None, norm_pos=None, conv=False):
        """[4, D] float64: normalisation mean, clamped normalisation variance, raw log-variance
        parameter, data divisor.  Differentiable w.r.t. the log-variance parameters."""
        D, dev = self.D, self.device
        nm = torch.zeros(D, dtype=torch.float64, device=dev)
        nv = torch.ones(D, dtype=torch.float64, device=dev)
        lvy = torch.zeros(D, dtype=torch.float64, device=dev)
        div = torch.ones(D, dtype=torch.float64, device=dev)
        ir, ip = self.idx["real"], self.idx["pos"]
        if ir.numel():
            if norm_real is not None and len(norm_real) == 2:
                nm = nm.index_put((ir,), norm_real[0].to(torch.float64)[self.gpos["real"]])
                nv = nv.index_put((ir,), torch.clamp(norm_real[1].to(torch.float64), min=3e-4)[self.gpos["real"]])  # loglik.py:38
            if conv:
                div = div.index_put((ir,), torch.full((ir.numel(),), 255.0, dtype=torch.float64, device=dev))     # HLVAE.py:394
            if log_vy_real is None:
                raise NotImplementedError("real likelihood without a per-variable log-variance (logvar_network=True) "
                                          "is not supported")
            lvy = lvy.index_put((ir,), log_vy_real.to(torch.float64)[self.gpos["real"]])
        if ip.numel():
            if norm_pos is None or len(norm_pos) != 2:
                raise ValueError("positive likelihood needs normalisation parameters [mean_log, var_log]")
            nm = nm.index_put((ip,), norm_pos[0].to(torch.float64)[self.gpos["pos"]])
            nv = nv.index_put((ip,), torch.clamp(norm_pos[1].to(torch.float64), min=1e-3)[self.gpos["pos"]])      # loglik.py:80
            if log_vy_pos is None:
                raise NotImplementedError("positive likelihood without a per-variable log-variance is not supported")
            lvy = lvy.index_put((ip,), log_vy_pos.to(torch.float64)[self.gpos["pos"]])
        return torch.stack([nm, nv, lvy, div])


def _storage_code(t, base_dtype, what):
    """Storage code of `data` / `mask`: uint8 (bool) travels as is, anything else is brought to the
    storage dtype of theta."""
    if t.dtype in (torch.uint8, torch.bool):
        t = t.detach().contiguous()
        return (t.view(torch.uint8) if t.dtype == torch.bool else t), _lib.U8
    return t.detach().to(base_dtype).contiguous(), (_lib.F64 if base_dtype == torch.float64 else _lib.F32)


class _FusedLoglik(torch.autograd.Function):
    """(theta, vparam) -> log_p_x, log_p_x_missing, params, log_p_x_sum[, recon_mean, recon_mode, data_tr].

    log_p_x_sum (float64 scalar) is accumulated inside the forward kernel; when a loss uses only
    the sum (training.py:83,104: -sum(log_p_x)), backward never materialises an [N, D] gradient."""

    @staticmethod
    def forward(ctx, theta, vparam, data, mask, layout, monitor):
        if not theta.is_cuda:
            raise RuntimeError("hlvae_b200: likelihoods run on CUDA tensors only (no CPU fallback)")
        N = theta.shape[0]
        D = layout.D
        dt = theta.dtype
        dcode = _lib.dtype_code(theta)
        th = theta.detach().contiguous()
        da, da_code = _storage_code(data, dt, "data")
        mk, mk_code = _storage_code(mask, dt, "mask")
        vp = vparam.detach().to(torch.float64).contiguous()
        new = lambda *s: torch.empty(*s, dtype=dt, device=th.device)
        lpx, lpm, prm = new(N, D), new(N, D), new(N, layout.P_theta)
        rmean = new(N, D) if monitor else None
        rmode = new(N, D) if monitor else None
        dtr = new(N, D) if monitor else None
        total = torch.zeros((), dtype=torch.float64, device=th.device)
        if N > 0:
            _lib.call("hlvae_loglik_fwd", N, D, layout.E_x, layout.P_theta, _lib.ptr(layout.var_kind),
                      _lib.ptr(layout.var_nclass), _lib.ptr(layout.var_dcol), _lib.ptr(layout.var_pcol), _lib.ptr(vp),
                      _lib.ptr(da), _lib.ptr(th), _lib.ptr(mk), dcode, da_code, mk_code, layout.max_class, _lib.ptr(lpx), _lib.ptr(lpm),
                      _lib.ptr(prm), _lib.ptr(rmean), _lib.ptr(rmode), _lib.ptr(dtr), _lib.ptr(total),
                      _lib.stream_ptr())
        ctx.layout, ctx.codes = layout, (dcode, da_code, mk_code)
        ctx.save_for_backward(th, vp, da, mk)
        ctx.set_materialize_grads(False)
        outs = (lpx, lpm, prm, total) + ((rmean, rmode, dtr) if monitor else ())
        ctx.mark_non_differentiable(lpm, prm, *outs[4:])
        return outs

    @staticmethod
    def backward(ctx, g_lpx, g_lpm, g_prm, g_total, *unused):
        th, vp, da, mk = ctx.saved_tensors
        layout = ctx.layout
        dcode, da_code, mk_code = ctx.codes
        N, D = th.shape[0], layout.D
        g_theta = torch.empty_like(th)
        g_lvy = torch.zeros(D, dtype=torch.float64, device=th.device)
        if g_lpx is None and g_total is None:
            return torch.zeros_like(th), torch.zeros_like(vp), None, None, None, None
        g = g_lpx.to(th.dtype).contiguous() if g_lpx is not None else None
        gs = g_total.detach().to(torch.float64).reshape(1).contiguous() if g_total is not None else None
        if N > 0:
            _lib.call("hlvae_loglik_bwd", N, D, layout.E_x, layout.P_theta, _lib.ptr(layout.var_kind),
                      _lib.ptr(layout.var_nclass), _lib.ptr(layout.var_dcol), _lib.ptr(layout.var_pcol), _lib.ptr(vp),
                      _lib.ptr(da), _lib.ptr(th), _lib.ptr(mk), dcode, da_code, mk_code, layout.max_class, _lib.ptr(g),
                      _lib.ptr(gs),
                      _lib.ptr(g_theta), _lib.ptr(g_lvy), _lib.stream_ptr())
        g_vp = torch.zeros_like(vp)
        g_vp[2] = g_lvy
        return g_theta, g_vp, None, None, None, None


def fused_loglik(layout, data, mask, theta, vparam, monitor=True):
    """All type groups in one launch.  Returns a dict with log_p_x, log_p_x_missing [N,D], params
    [N,P_theta], log_p_x_sum (float64 scalar = log_p_x.sum(), differentiable) and (monitor=True)
    recon_mean, recon_mode, data_transformed [N,D].  `data` and `mask` may be uint8."""
    outs = _FusedLoglik.apply(theta, vparam, data, mask, layout, monitor)
    names = ("log_p_x", "log_p_x_missing", "params", "log_p_x_sum", "recon_mean", "recon_mode", "data_transformed")
    return dict(zip(names, outs))


# --------------------------------------------------------------------------------------
# per-type functions, HL_VAE/loglik.py signatures
# --------------------------------------------------------------------------------------
_group_cache = {}


def _group_layout(kind, C, Dg, device):
    key = (kind, C, Dg, str(device))
    if key not in _group_cache:
        _group_cache[key] = VarLayout([(kind, C)] * Dg, device)
    return _group_cache[key]


def _one_hot(idx, depth):
    """HL_VAE/utils.py:70-76."""
    return torch.zeros(idx.shape + (depth,), device=idx.device).scatter_(idx.dim(), idx.unsqueeze(-1), 1)


def _sequence_mask(lengths, maxlen, dtype):
    """HL_VAE/utils.py:79-85."""
    return (torch.ones((lengths.shape[0], lengths.shape[1], maxlen), device=lengths.device).cumsum(dim=2)
            <= lengths[:, :, None]).type(dtype)


def _run_group(kind, C, batch_data, theta, vp_builder):
    data, mask = batch_data
    mask = mask.float()                                   # loglik.py:33 (float32 mask, promoted on use)
    N, Dg = mask.shape
    layout = _group_layout(kind, C, Dg, theta.device)
    vparam = vp_builder(layout)
    out = fused_loglik(layout, data.reshape(N, -1), mask.to(theta.dtype), theta.reshape(N, -1), vparam, monitor=False)
    return out, N, Dg


class _AuxLoglik(torch.autograd.Function):
    """One type group through hlvae_loglik_aux_fwd / _bwd: (th_a, th_b | raw dispersion) -> log_p_x, log_p_x_missing,
    prm_a, prm_b, each [N, Dg].  th_a / th_b may be column slices of theta (row stride passed as is); th_a [N, 1]
    with col_stride_a = 0 is read by every variable of the group."""

    @staticmethod
    def forward(ctx, th_a, th_b, data, mask, vparam, mode, cs_a):
        if not th_a.is_cuda:
            raise RuntimeError("hlvae_b200: likelihoods run on CUDA tensors only (no CPU fallback)")
        N, Dg = data.shape
        dt = th_a.dtype
        dcode = _lib.dtype_code(th_a)
        ta = th_a.detach()
        ta = ta if ta.stride(-1) == 1 or ta.shape[1] == 1 else ta.contiguous()
        beta = mode == _lib.AUX_BETA
        if beta:
            tb = th_b.detach().to(torch.float64).reshape(-1)[:1].contiguous()      # extra_params[0], loglik.py:240
        else:
            tb = th_b.detach().to(dt)
            tb = tb if tb.stride(-1) == 1 else tb.contiguous()
        da, da_code = _storage_code(data, dt, "data")
        mk, mk_code = _storage_code(mask, dt, "mask")
        vp = vparam.detach().to(torch.float64).contiguous()
        new = lambda: torch.empty(N, Dg, dtype=dt, device=ta.device)
        lpx, lpm, pa, pb = new(), new(), new(), new()
        ctx.args = (mode, N, Dg, da_code, mk_code, cs_a, dcode)
        ctx.save_for_backward(ta, tb, da, mk, vp)
        ctx.shape_b, ctx.dtype_b = th_b.shape, th_b.dtype
        _lib.call("hlvae_loglik_aux_fwd", mode, N, Dg, _lib.ptr(da), da.stride(0), da_code, _lib.ptr(mk), mk.stride(0),
                  mk_code, _lib.ptr(ta), ta.stride(0), cs_a, None if beta else _lib.ptr(tb),
                  0 if beta else tb.stride(0), dcode, _lib.ptr(vp), _lib.ptr(tb) if beta else None, _lib.ptr(lpx),
                  _lib.ptr(lpm), _lib.ptr(pa), _lib.ptr(pb), _lib.stream_ptr())
        ctx.mark_non_differentiable(lpm, pa, pb)
        return lpx, lpm, pa, pb

    @staticmethod
    def backward(ctx, g_lpx, g_lpm, g_pa, g_pb):
        ta, tb, da, mk, vp = ctx.saved_tensors
        mode, N, Dg, da_code, mk_code, cs_a, dcode = ctx.args
        beta = mode == _lib.AUX_BETA
        g = g_lpx.to(ta.dtype).contiguous()
        g_a = torch.empty(N, Dg, dtype=ta.dtype, device=ta.device)
        g_b = None if beta else torch.empty(N, Dg, dtype=ta.dtype, device=ta.device)
        g_disp = torch.zeros(1, dtype=torch.float64, device=ta.device) if beta else None
        _lib.call("hlvae_loglik_aux_bwd", mode, N, Dg, _lib.ptr(da), da.stride(0), da_code, _lib.ptr(mk), mk.stride(0),
                  mk_code, _lib.ptr(ta), ta.stride(0), cs_a, None if beta else _lib.ptr(tb),
                  0 if beta else tb.stride(0), dcode, _lib.ptr(vp), _lib.ptr(tb) if beta else None, _lib.ptr(g), None,
                  _lib.ptr(g_a), _lib.ptr(g_b), _lib.ptr(g_disp), _lib.stream_ptr())
        if cs_a == 0:
            g_a = g_a.sum(1, keepdim=True)
        if beta:
            g_second = torch.zeros(ctx.shape_b, dtype=torch.float64, device=ta.device).reshape(-1)
            g_second[0] = g_disp[0]
            g_second = g_second.reshape(ctx.shape_b).to(ctx.dtype_b)
        else:
            g_second = g_b
        return g_a, g_second, None, None, None, None, None


def _aux_vparam(Dg, device, mean=None, var=None, div=None):
    vp = torch.zeros(4, Dg, dtype=torch.float64, device=device)
    vp[1] = 1.0
    vp[3] = 1.0
    if mean is not None:
        vp[0] = torch.as_tensor(mean, dtype=torch.float64, device=device).reshape(-1).expand(Dg)
    if var is not None:
        vp[1] = torch.as_tensor(var, dtype=torch.float64, device=device).reshape(-1).expand(Dg)
    if div is not None:
        vp[3] = div
    return vp


def loglik_real(batch_data, list_type, theta, normalization_params, extra_params=None):
    """HL_VAE/loglik.py:27-70.  `extra_params` = the per-variable log-variance parameter (logvar_network=False) or
    None (logvar_network=True, :45-48: theta holds the means followed by per-row raw log-variances)."""
    norm = None if (isinstance(normalization_params, list) and normalization_params == []) else normalization_params
    if extra_params is None:
        data, mask = batch_data
        N, Dg = data.shape
        if theta.shape[1] < 2 * Dg:
            raise IndexError("loglik_real without extra_params expects theta = [means, log-variances] (2 D columns)")
        vp = _aux_vparam(Dg, theta.device, None if norm is None else norm[0],
                         None if norm is None else torch.clamp(norm[1].to(torch.float64), min=3e-4))       # :36-41
        lpx, lpm, mean, var = _AuxLoglik.apply(theta[:, :Dg], theta[:, Dg:2 * Dg], data, mask.float(), vp,
                                               _lib.AUX_REAL, 1)
        res = {'log_p_x': lpx, 'log_p_x_missing': lpm, 'params': [mean, var]}                                # :64-65
        res['samples'] = td.Normal(mean, torch.sqrt(var)).rsample() if PRODUCE_SAMPLES else None
        return res
    out, N, Dg = _run_group("real", 1, batch_data, theta[:, :batch_data[0].shape[1]],
                            lambda lay: lay.vparam(log_vy_real=extra_params, norm_real=norm))
    res = {'log_p_x': out['log_p_x'], 'log_p_x_missing': out['log_p_x_missing'], 'params': out['params']}
    if PRODUCE_SAMPLES:
        dvar = torch.clamp(norm[1], 3e-4, np.inf) if norm is not None else torch.tensor(1., device=theta.device)
        est_var = dvar * torch.exp(-8.0 + torch.nn.functional.softplus(extra_params + 8.0))
        res['samples'] = td.Normal(out['params'], torch.sqrt(est_var)).rsample()          # :59,68
    else:
        res['samples'] = None
    return res


def loglik_pos(batch_data, list_type, theta, normalization_params, extra_params=None):
    """HL_VAE/loglik.py:73-121; extra_params None -> per-row log-variances from theta (:89,104-108)."""
    if extra_params is None:
        data, mask = batch_data
        N, Dg = data.shape
        if theta.shape[1] < 2 * Dg:
            raise UnboundLocalError("loglik_pos without extra_params expects theta = [means, log-variances]")
        vp = _aux_vparam(Dg, theta.device, normalization_params[0],
                         torch.clamp(normalization_params[1].to(torch.float64), min=1e-3))                  # :79-80
        lpx, lpm, mean, var = _AuxLoglik.apply(theta[:, :Dg], theta[:, Dg:2 * Dg], data, mask.float(), vp,
                                               _lib.AUX_POS, 1)
        res = {'log_p_x': lpx, 'log_p_x_missing': lpm, 'params': [mean, var]}                                # :114-115
        res['samples'] = torch.clamp(torch.exp(td.Normal(mean, torch.sqrt(var)).rsample()) - 1.0, 0, 1e20) \
            if PRODUCE_SAMPLES else None
        return res
    out, N, Dg = _run_group("pos", 1, batch_data, theta[:, :batch_data[0].shape[1]],
                            lambda lay: lay.vparam(log_vy_pos=extra_params, norm_pos=normalization_params))
    res = {'log_p_x': out['log_p_x'], 'log_p_x_missing': out['log_p_x_missing'], 'params': out['params']}
    if PRODUCE_SAMPLES:
        est_var = torch.clamp(normalization_params[1], 1e-3, np.inf) * torch.exp(extra_params)
        res['samples'] = torch.clamp(torch.exp(td.Normal(out['params'], torch.sqrt(est_var)).rsample()) - 1.0, 0, 1e20)  # :118-119
    else:
        res['samples'] = None
    return res


def loglik_beta(batch_data, list_type, theta, normalization_params, extra_params=None):
    """HL_VAE/loglik.py:216-256.  normalization_params: the concatenated [min, max + 1e-3] ranges of the group
    (HLVAE.py:400); extra_params: the raw dispersion `_disp_param`.  The reference reads theta[:, :D] and
    theta[:, D:2D] and, when theta is narrower (it is: beta variables own ONE parameter column each,
    read_functions.py:164-173), falls back to columns 0 and 1 as [N, 1] tensors (:232-235) - of which only the first
    is used afterwards (:238-245): every variable of the group then shares theta[:, 0].  Reproduced as is."""
    data, mask = batch_data
    N, Dg = data.shape
    rng = torch.as_tensor(np.asarray(normalization_params.detach().cpu() if torch.is_tensor(normalization_params)
                                     else normalization_params), dtype=torch.float64).reshape(Dg, -1)       # :222
    vp = _aux_vparam(Dg, theta.device)
    vp[0], vp[1] = rng[:, 0].to(theta.device), rng[:, 1].to(theta.device)
    if theta.shape[1] >= 2 * Dg:
        th_a, cs = theta[:, :Dg], 1
    else:
        if theta.shape[1] < 2:
            raise IndexError("index 1 is out of bounds for dimension 1 with size 1")       # what :235 raises
        th_a, cs = theta[:, 0:1], 0
    lpx, lpm, al, be = _AuxLoglik.apply(th_a, extra_params, data, mask.float(), vp, _lib.AUX_BETA, cs)
    if cs == 0:
        al, be = al[:, :1], be[:, :1]
    res = {'log_p_x': lpx, 'log_p_x_missing': lpm, 'params': [al, be]}                                       # :253
    if PRODUCE_SAMPLES:
        res['samples'] = td.Beta(al, be).sample() * (vp[1] - vp[0]).to(al.dtype) + vp[0].to(al.dtype)         # :254
    else:
        res['samples'] = None
    return res


def loglik_cat(batch_data, list_type, theta, normalization_params, extra_params=None):
    """HL_VAE/loglik.py:124-146."""
    C = int(list_type[1])
    out, N, Dg = _run_group("cat", C, batch_data, theta, lambda lay: lay.vparam())
    log_pi = out['params'].reshape(N, Dg, C)
    res = {'log_p_x': out['log_p_x'], 'log_p_x_missing': out['log_p_x_missing'], 'params': log_pi}
    if PRODUCE_SAMPLES:
        res['samples'] = _one_hot(td.Categorical(probs=torch.softmax(log_pi, 1)).sample(), C).to(torch.float64)  # :141-142
    else:
        res['samples'] = None
    return res


def loglik_ordinal(batch_data, list_type, theta, normalization_params, extra_params=None):
    """HL_VAE/loglik.py:149-188."""
    C = int(list_type[1])
    out, N, Dg = _run_group("ordinal", C, batch_data, theta, lambda lay: lay.vparam())
    probs = out['params'].reshape(N, Dg, C)
    res = {'log_p_x': out['log_p_x'], 'log_p_x_missing': out['log_p_x_missing'], 'params': probs}
    if PRODUCE_SAMPLES:
        draw = td.Categorical(logits=torch.log(torch.clamp(probs, 1e-6, 1e20))).sample()
        res['samples'] = _sequence_mask(1 + draw, C, dtype=torch.float64)                   # :184-186
    else:
        res['samples'] = None
    return res


def loglik_count(batch_data, list_type, theta, normalization_params, extra_params=None):
    """HL_VAE/loglik.py:191-213."""
    out, N, Dg = _run_group("count", 1, batch_data, theta, lambda lay: lay.vparam())
    res = {'log_p_x': out['log_p_x'], 'log_p_x_missing': out['log_p_x_missing'], 'params': out['params']}
    res['samples'] = td.Poisson(out['params']).sample() if PRODUCE_SAMPLES else None       # :204,211
    return res


# --------------------------------------------------------------------------------------
# HLVAE.loglik_and_reconstruction replacement
# --------------------------------------------------------------------------------------
def _model_layout(model, device):
    lay = getattr(model, "_hlvae_b200_layout", None)
    if lay is None or str(lay.device) != str(device):
        lay = VarLayout.from_types_info(model.types_info, device)
        groups = []
        for i, tpl in enumerate(model.types_info['set_of_types']):
            cols = torch.tensor(np.nonzero(np.asarray(model.types_info['param_indexes']) == i)[0], dtype=torch.long,
                                device=device)
            groups.append(cols)
        lay.group_param_cols = groups
        model._hlvae_b200_layout = lay
    return lay


def loglik_and_reconstruction(self, theta, batch_data_list, miss_list, param_miss_list, normalization_params, s=None):
    """Drop-in for HLVAE.loglik_and_reconstruction (HLVAE.py:381-414); bind with
    `HLVAE.loglik_and_reconstruction = hlvae_b200.loglik.loglik_and_reconstruction`.
    Returns (log_p_x, log_p_x_missing, samples_x, params_x) with params_x one tensor per type
    group, consumable by read_functions.p_params_concatenation_by_key (:206-218).  The fused
    monitoring outputs are left on `self.hlvae_b200_monitor`."""
    if getattr(self, "logvar_network", False) or any(t[0] == 'beta' for t in self.types_info['set_of_types']):
        return _loglik_by_groups(self, theta, batch_data_list, miss_list, normalization_params)
    lay = _model_layout(self, theta.device)
    nr = normalization_params[0] if len(normalization_params) > 0 else []
    npos = normalization_params[1] if len(normalization_params) > 1 else []
    vparam = lay.vparam(getattr(self, "_log_vy_real", None), getattr(self, "_log_vy_pos", None),
                        None if (isinstance(nr, list) and nr == []) else nr,
                        None if (isinstance(npos, list) and npos == []) else npos, conv=bool(self.conv))
    out = fused_loglik(lay, batch_data_list, miss_list, theta, vparam, monitor=True)
    self.hlvae_b200_monitor = out
    params_x = [out['params'][:, cols] for cols in lay.group_param_cols]
    samples_x = [None] * len(params_x)
    return out['log_p_x'], out['log_p_x_missing'], samples_x, params_x


def _loglik_by_groups(model, theta, batch_data_list, miss_list, normalization_params):
    """HLVAE.loglik_and_reconstruction (HLVAE.py:381-414) for the configurations the single fused launch does not
    cover (variance network, beta variables): the reference's own loop over the type groups, each group evaluated by
    this module's loglik_<type> kernels."""
    ti = model.types_info
    dev = theta.device
    log_p_x = torch.zeros(miss_list.shape, dtype=theta.dtype, device=dev)
    log_p_x_missing = torch.zeros_like(log_p_x)
    samples_x, params_x = [], []
    sel = lambda key, i: torch.as_tensor(np.nonzero(np.asarray(ti[key]) == i)[0], dtype=torch.long, device=dev)
    this = globals()
    for i, tpl in enumerate(ti['set_of_types']):
        dcols, vcols, pcols = sel('exp_types_indexes', i), sel('data_types_indexes', i), sel('param_indexes', i)
        data_g = batch_data_list[:, dcols]
        extra, norm = None, torch.tensor(0.)
        if tpl[0] == 'real':
            norm = normalization_params[0]
            if model.conv:
                data_g = data_g / 255                                                          # :393-394
            extra = model._log_vy_real
        elif tpl[0] == 'pos':
            norm, extra = normalization_params[1], model._log_vy_pos
        elif tpl[0] == 'beta':
            norm, extra = np.concatenate(ti['beta_ranges']), model._disp_param                   # :399-401
        out = this['loglik_' + tpl[0]]([data_g, miss_list[:, vcols]], tpl, theta[:, pcols], norm, extra)
        log_p_x = log_p_x.index_copy(1, vcols, out['log_p_x'].to(log_p_x.dtype))
        log_p_x_missing = log_p_x_missing.index_copy(1, vcols, out['log_p_x_missing'].to(log_p_x.dtype))
        samples_x.append(out['samples'])
        params_x.append(out['params'])
    return log_p_x, log_p_x_missing, samples_x, params_x


def statistics(layout, params, vparam):
    """read_functions.statistics (:268-302) on packed params -> (mean, mode) [N, D]."""
    N = params.shape[0]
    prm = params.detach().contiguous()
    mean, mode = torch.empty(N, layout.D, dtype=prm.dtype, device=prm.device), \
        torch.empty(N, layout.D, dtype=prm.dtype, device=prm.device)
    vp = vparam.detach().to(torch.float64).contiguous()
    _lib.call("hlvae_statistics", N, layout.D, layout.P_theta, _lib.ptr(layout.var_kind),
                                           _lib.ptr(layout.var_nclass), _lib.ptr(layout.var_pcol), _lib.ptr(vp),
                                           _lib.ptr(prm), _lib.dtype_code(prm), _lib.ptr(mean), _lib.ptr(mode),
                                           _lib.stream_ptr())
    return mean, mode


def statistics_general(loglik_params, types_info, conv=False, log_vy=None):
    """read_functions.statistics (:268-339) for any layout of `types_info`: the packed five-type layout goes to the
    kernel as is; with a variance network (real / positive groups own [means, variances] columns, :276-291) or beta
    groups (:303-337) the kernel evaluates the gathered mean columns and the few affected groups are finished with
    element-wise device ops."""
    ti = types_info
    dev = loglik_params.device
    dti, pidx = np.asarray(ti['data_types_indexes']), np.asarray(ti['param_indexes'])
    types = [(t['type'], int(t['nclass'])) for t in ti['types_dict']]
    base_types = [("real" if k == "beta" else k, c) for k, c in types]
    lay = VarLayout(base_types, dev)
    plain = len(pidx) == lay.P_theta and not any(k == "beta" for k, _ in types)
    lv_pos = log_vy[1] if (log_vy is not None and len(log_vy) > 1) else None
    vparam = torch.zeros(4, lay.D, dtype=torch.float64, device=dev)
    if lv_pos is not None and lay.idx["pos"].numel():
        vparam[2, lay.idx["pos"]] = lv_pos.detach().to(torch.float64)[lay.gpos["pos"]]
    if plain:
        return statistics(lay, loglik_params, vparam)
    gather = np.zeros(lay.P_theta, dtype=np.int64)
    fix = []
    for i, tpl in enumerate(ti['set_of_types']):
        cols, vars_g = np.nonzero(pidx == i)[0], np.nonzero(dti == i)[0]
        if tpl[0] in ("real", "pos", "beta"):
            n = len(vars_g)
            gather[[lay.pcol_host[d] for d in vars_g]] = cols[:n]
            if tpl[0] == "beta" or (tpl[0] == "pos" and len(cols) == 2 * n):
                fix.append((tpl[0], cols, vars_g))
        else:
            gather[np.concatenate([np.arange(lay.pcol_host[d], lay.pcol_host[d] + lay.ncls_host[d]) for d in vars_g])] = cols
    mean, mode = statistics(lay, loglik_params.index_select(1, torch.as_tensor(gather, device=dev)).contiguous(), vparam)
    for kind, cols, vars_g in fix:
        c = torch.as_tensor(cols, device=dev)
        vg = torch.as_tensor(vars_g, device=dev)
        prm = loglik_params.index_select(1, c)
        n = len(vars_g)
        if kind == "pos":                                                   # :283-291 with the per-row variance
            mu, var = prm[:, :n], prm[:, n:2 * n]
            mean = mean.index_copy(1, vg, (torch.exp(mu + 0.5 * var) - 1.0).to(mean.dtype))
            mode = mode.index_copy(1, vg, (torch.exp(mu - var) - 1.0).to(mode.dtype))
        else:                                                               # beta, :303-337
            sz = prm.shape[1] // 2
            al, be = prm[:, :sz], prm[:, sz:2 * sz]
            rng = torch.as_tensor(np.concatenate(ti['beta_ranges']).reshape(sz, -1), dtype=prm.dtype, device=dev)
            lo, hi = rng[:, 0], rng[:, 1]
            m01 = al / (al + be)
            md = torch.where((al > 1) & (be > 1), (al - 1) / (al + be - 2),
                             torch.where((al > 1) & (be <= 1), torch.ones_like(al), torch.zeros_like(al)))
            eq = (al == 1) & (be == 1)
            if bool(eq.any()):                                              # :333: uniform density, a random mode
                md = torch.where(eq, torch.rand(al.shape, device=dev).to(md.dtype), md)
            mean = mean.index_copy(1, vg, (m01 * (hi - lo) + lo).expand(-1, n).to(mean.dtype))
            mode = mode.index_copy(1, vg, (md * (hi - lo) + lo).expand(-1, n).to(mode.dtype))
    return mean, mode


def discrete_variables_transformation(layout, data):
    """read_functions.discrete_variables_transformation (:221-235) on packed data -> [N, D]."""
    N = data.shape[0]
    da = data.detach().contiguous()
    out = torch.empty(N, layout.D, dtype=da.dtype, device=da.device)
    _lib.call("hlvae_discrete_transform", N, layout.D, layout.E_x, _lib.ptr(layout.var_kind),
                                                   _lib.ptr(layout.var_nclass), _lib.ptr(layout.var_dcol),
                                                   _lib.ptr(da), _lib.dtype_code(da), _lib.ptr(out),
                                                   _lib.stream_ptr())
    return out
