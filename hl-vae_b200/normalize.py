"""Batch normalisation of the heterogeneous data batch with the reference's operator surface.

`batch_normalization(batch_data_list, miss_list, param_mask, types_info)` mirrors HL_VAE/utils.py:88-143: it
returns the encoder input X_list [N, E_x] and `[[mean, var] | [], [mean_log, var_log] | []]`, the normalisation
parameters that HLVAE.forward hands to the likelihoods (HLVAE.py:367-375).  Three streaming CUDA launches
(hlvae_batch_norm_stats pass 0 / pass 1, hlvae_batch_norm_apply) replace the per-type-group boolean gathers,
reductions and scatters.  No CPU path; no autograd (the reference's result does not require grad either: it is
a function of the data alone).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .loglik import VarLayout, _storage_code


class NormLayout:
    """Descriptors of the packed data layout for the normalisation kernels."""

    def __init__(self, types, conv, device):
        self.var = VarLayout(types, device)
        v = self.var
        self.conv = bool(conv)
        self.device = device
        dcol_var = np.zeros(v.E_x, dtype=np.int32)
        for d in range(v.D):
            dcol_var[v.dcol_host[d]:v.dcol_host[d] + v.ncls_host[d]] = d
        stat = [d for d, (k, _) in enumerate(v.types) if k == "pos" or (k == "real" and not self.conv)]
        i32 = dict(dtype=torch.int32, device=device)
        self.dcol_var = torch.tensor(dcol_var, **i32)
        self.stat_vars = torch.tensor(stat, **i32)
        self.n_stat = len(stat)

    @staticmethod
    def from_types_info(types_info, device):
        types = [(t['type'], int(t['nclass'])) for t in types_info['types_dict']]
        return NormLayout(types, bool(types_info.get('conv', False)), device)


_layout_cache = {}
_LAYOUT_CACHE_MAX = 16


def _layout_for(types_info, device):
    """Layout descriptors of `types_info`, cached by CONTENT ((type, nclass) per variable, conv flag, device): an
    id()-keyed cache would hand a stale layout to a dict that was mutated in place or whose id CPython reused."""
    key = (tuple((t['type'], int(t['nclass'])) for t in types_info['types_dict']), bool(types_info.get('conv', False)),
           str(device))
    lay = _layout_cache.get(key)
    if lay is None:
        if len(_layout_cache) >= _LAYOUT_CACHE_MAX:
            _layout_cache.pop(next(iter(_layout_cache)))
        lay = NormLayout.from_types_info(types_info, device)
        _layout_cache[key] = lay
    return lay


def normalize(layout: NormLayout, data, mask, out_dtype=None):
    """(X_list [N, E_x], mean [D], var [D]) on the packed layout; mean / var are float64 and meaningful for the
    real (non-convolutional) and positive variables only.  `data` and `mask` may be uint8."""
    if not data.is_cuda:
        raise RuntimeError("hlvae_b200: batch normalisation runs on CUDA tensors only (no CPU fallback)")
    v = layout.var
    N = data.shape[0]
    if data.dim() != 2 or data.shape[1] != v.E_x or mask.dim() != 2 or mask.shape[0] != N or mask.shape[1] != v.D:
        raise ValueError(f"hlvae_b200: batch normalisation expects data [N, {v.E_x}] and mask [N, {v.D}], got "
                         f"{tuple(data.shape)} and {tuple(mask.shape)}")
    if out_dtype is None:
        out_dtype = data.dtype if data.dtype in (torch.float32, torch.float64) else torch.float32
    dcode = _lib.F64 if out_dtype == torch.float64 else _lib.F32
    da, da_code = _storage_code(data, out_dtype, "data")
    mk, mk_code = _storage_code(mask, out_dtype, "mask")
    stats = torch.zeros(3, v.D, dtype=torch.float64, device=data.device)
    out = torch.empty(N, v.E_x, dtype=out_dtype, device=data.device)
    if N > 0 and layout.n_stat > 0:
        for p in (0, 1):
            _lib.call("hlvae_batch_norm_stats", N, v.D, v.E_x, _lib.ptr(v.var_kind), _lib.ptr(v.var_dcol),
                      _lib.ptr(layout.stat_vars), layout.n_stat, _lib.ptr(da), _lib.ptr(mk), dcode, da_code, mk_code, p,
                      _lib.ptr(stats), _lib.stream_ptr())
    mean = stats[1] / stats[0]                                                # utils.py:106 / :125
    var = stats[2] / stats[0]                                                 # :107 / :126
    ip = v.idx["pos"]
    if ip.numel():
        var = var.index_put((ip,), torch.clamp(var[ip], 1e-6, 1e20))          # :127
    meanvar = torch.stack([mean, var]).contiguous()
    if N > 0:
        _lib.call("hlvae_batch_norm_apply", N, v.D, v.E_x, _lib.ptr(v.var_kind), _lib.ptr(layout.dcol_var), _lib.ptr(da),
                  _lib.ptr(mk), dcode, da_code, mk_code, int(layout.conv), _lib.ptr(meanvar), _lib.ptr(out),
                  _lib.stream_ptr())
    return out, mean, var


def batch_normalization(batch_data_list, miss_list, param_mask, types_info):
    """Drop-in for HL_VAE.utils.batch_normalization (HL_VAE/utils.py:88-143): same arguments (`param_mask` is not
    read, as in the reference) and the same `(normalized_data, [[mean, var] | [], [mean_log, var_log] | []])`."""
    lay = _layout_for(types_info, batch_data_list.device)
    out, mean, var = normalize(lay, batch_data_list, miss_list)
    v = lay.var
    params = [[], []]
    ir, ip = v.idx["real"], v.idx["pos"]
    if ir.numel() and not lay.conv:
        params[0] = [mean[ir].to(out.dtype), var[ir].to(out.dtype)]
    if ip.numel():
        params[1] = [mean[ip].to(out.dtype), var[ip].to(out.dtype)]
    return out, params
