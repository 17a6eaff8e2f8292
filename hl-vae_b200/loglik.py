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


def loglik_real(batch_data, list_type, theta, normalization_params, extra_params=None):
    """HL_VAE/loglik.py:27-70."""
    if extra_params is None:
        raise NotImplementedError("loglik_real with a variance network (extra_params=None) is not supported")
    norm = None if (isinstance(normalization_params, list) and normalization_params == []) else normalization_params
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
    """HL_VAE/loglik.py:73-121."""
    if extra_params is None:
        raise NotImplementedError("loglik_pos with a variance network (extra_params=None) is not supported")
    out, N, Dg = _run_group("pos", 1, batch_data, theta[:, :batch_data[0].shape[1]],
                            lambda lay: lay.vparam(log_vy_pos=extra_params, norm_pos=normalization_params))
    res = {'log_p_x': out['log_p_x'], 'log_p_x_missing': out['log_p_x_missing'], 'params': out['params']}
    if PRODUCE_SAMPLES:
        est_var = torch.clamp(normalization_params[1], 1e-3, np.inf) * torch.exp(extra_params)
        res['samples'] = torch.clamp(torch.exp(td.Normal(out['params'], torch.sqrt(est_var)).rsample()) - 1.0, 0, 1e20)  # :118-119
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
