"""Per-variable observation heads: y (decoder output, [N, D, y_dim]) -> theta ([N, P_theta]).

Mirror of `HLVAE.theta_estimation` (HLVAE.py:416-453) over the `Observation_*` modules (HLVAE.py:11-89),
logvar_network=False, backed by ONE streaming CUDA kernel per direction (hlvae_theta_fwd / _bwd) instead of
two einsum passes per type group, mask products and boolean-index scatters.

For 0/1 masks the reference's observed / missing double evaluation is
    theta[n, p] = act_p(bias[p] + sum_k weight[p, k] * y[n, var(p), k])
in the forward direction (the "missing" pass only supplies the values where mask = 0, under no_grad), and the
gradient reaches y and the head parameters through observed entries only.  `pack_heads` gathers the modules'
parameters into the per-column [P_theta, y_dim] / [P_theta] form with differentiable torch ops, so autograd
carries the kernel's gradients back to the `nn.Parameter`s the optimiser holds.
There is no CPU path: tensors must be on a CUDA device.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib
from .loglik import VarLayout

MAX_TILE = 256      # theta columns (and variables) per CTA tile, csrc/theta.cu


class HeadLayout:
    """Column descriptors of the packed theta layout for the head kernel, plus the index maps that
    gather the per-type-group module parameters (HLVAE.py:276-299) into per-column form."""

    def __init__(self, types, conv, device, set_of_types=None, data_types_indexes=None):
        self.var = VarLayout(types, device)
        v = self.var
        self.conv = bool(conv)
        self.device = device
        D, P = v.D, v.P_theta
        col_var = np.zeros(P, dtype=np.int32)
        col_mode = np.full(P, _lib.HEAD_AFFINE, dtype=np.int32)
        for d, (kind, C) in enumerate(v.types):
            p0, nc = v.pcol_host[d], v.ncls_host[d]
            col_var[p0:p0 + nc] = d
            if kind == "cat":
                col_mode[p0] = _lib.HEAD_ZERO                      # HLVAE.py:66-67
            elif kind == "ordinal":
                col_mode[p0:p0 + nc - 1] = _lib.HEAD_BIAS           # thresholds, HLVAE.py:85
            elif kind == "real" and self.conv:
                col_mode[p0] = _lib.HEAD_SIGMOID                   # HLVAE.py:292-295, 431-433
        var_pcol = np.array(list(v.pcol_host) + [P], dtype=np.int32)
        tiles = [0]
        for d in range(D):                                         # greedy: whole variables, <= MAX_TILE columns each
            if var_pcol[d + 1] - var_pcol[tiles[-1]] > MAX_TILE or d - tiles[-1] >= MAX_TILE:
                tiles.append(d)
        tiles.append(D)
        i32 = dict(dtype=torch.int32, device=device)
        self.col_var = torch.tensor(col_var, **i32)
        self.col_mode = torch.tensor(col_mode, **i32)
        self.var_pcol = torch.tensor(var_pcol, **i32)
        self.tile_var = torch.tensor(tiles, **i32)
        self.n_tiles = len(tiles) - 1
        # widest variable: layouts with <= 5 columns per variable take the thread-per-variable backward kernel
        # (csrc/theta.cu); HLVAE_THETA_PER_COLUMN=1 keeps the thread-per-column one (A/B runs)
        self.max_cols = 0 if os.environ.get("HLVAE_THETA_PER_COLUMN") else int(np.diff(var_pcol).max())
        self.D, self.P = D, P
        # type groups in the reference's order: sorted set of (type, nclass-string) tuples (read_functions.py:176-180)
        if set_of_types is None:
            tt = [(k, str(c if k in ("cat", "ordinal") else 1)) for k, c in v.types]
            set_of_types = sorted(set(tt))
            data_types_indexes = np.array([set_of_types.index(t) for t in tt])
        self.set_of_types = [tuple(t) for t in set_of_types]
        self.groups = []
        pcol = np.array(v.pcol_host)
        for i, tpl in enumerate(self.set_of_types):
            idx = np.nonzero(np.asarray(data_types_indexes) == i)[0]
            lt = lambda a: torch.tensor(np.asarray(a).reshape(-1), dtype=torch.long, device=device)
            g = dict(kind=tpl[0], C=int(tpl[1]), n=len(idx))
            if tpl[0] == "cat":
                C = int(tpl[1])
                g["cols"] = lt(pcol[idx][:, None] + 1 + np.arange(C - 1)[None, :])
            elif tpl[0] == "ordinal":
                C = int(tpl[1])
                g["thr_cols"] = lt(pcol[idx][:, None] + np.arange(C - 1)[None, :])
                g["cols"] = lt(pcol[idx] + C - 1)
            else:
                g["cols"] = lt(pcol[idx])
            self.groups.append(g)

    @staticmethod
    def from_types_info(types_info, device):
        types = [(t['type'], int(t['nclass'])) for t in types_info['types_dict']]
        return HeadLayout(types, bool(types_info.get('conv', False)), device, types_info['set_of_types'],
                          types_info['data_types_indexes'])


def pack_heads(obs_layer, layout: HeadLayout, y_dim: int):
    """(weight [P_theta, y_dim], bias [P_theta]) float64 from the `obs_layer` ModuleList of an HLVAE model
    (HLVAE.py:276-299): one Observation_* module per type group, followed by an nn.Sigmoid for the real group of
    the convolutional model.  Differentiable w.r.t. the module parameters."""
    dev = layout.device
    W = torch.zeros(layout.P, y_dim, dtype=torch.float64, device=dev)
    b = torch.zeros(layout.P, dtype=torch.float64, device=dev)
    layer = 0
    for g in layout.groups:
        mod = obs_layer[layer]
        kind = g["kind"]
        if kind == "count":                                          # Observation_Count, HLVAE.py:11-23
            W = W.index_put((g["cols"],), mod.weight[:, :, 0].to(torch.float64))
            b = b.index_put((g["cols"],), mod.bias[:, 0].to(torch.float64))
        elif kind in ("real", "pos"):                                # Observation_Real_Pos_Beta, :26-52
            # (a variance network adds weight_logvar / bias_logvar, HLVAE.py:30-35: those heads are evaluated by
            #  _theta_estimation_logvar below; this function packs the mean heads)
            W = W.index_put((g["cols"],), mod.weight_mean[:, :, 0].to(torch.float64))
            b = b.index_put((g["cols"],), mod.bias_mean[:, 0].to(torch.float64))
            if kind == "real" and layout.conv:
                layer += 1                                           # the nn.Sigmoid entry, :292-295
        elif kind == "cat":                                          # Observation_Cat, :55-68
            W = W.index_put((g["cols"],), mod.weight.permute(0, 2, 1).reshape(-1, y_dim).to(torch.float64))
            b = b.index_put((g["cols"],), mod.bias.reshape(-1).to(torch.float64))
        elif kind == "ordinal":                                      # Observation_Ordinal, :70-89
            b = b.index_put((g["thr_cols"],), mod.weight_thresholds.reshape(-1).to(torch.float64))
            W = W.index_put((g["cols"],), mod.weight_region[:, :, 0].to(torch.float64))
            b = b.index_put((g["cols"],), mod.bias_region[:, 0].to(torch.float64))
        else:
            raise NotImplementedError(f"observation head for type '{kind}' is not supported")
        layer += 1
    return W, b


def _mask_arg(mask, dt):
    if mask.dtype in (torch.uint8, torch.bool):
        m = mask.detach().contiguous()
        return (m.view(torch.uint8) if m.dtype == torch.bool else m), _lib.U8
    return mask.detach().to(dt).contiguous(), (_lib.F64 if dt == torch.float64 else _lib.F32)


class _ThetaHeads(torch.autograd.Function):
    """(y [N, D, Y] any strides, weight [P, Y], bias [P]) -> theta [N, P]."""

    @staticmethod
    def forward(ctx, y, weight, bias, mask, layout):
        if not y.is_cuda:
            raise RuntimeError("hlvae_b200: observation heads run on CUDA tensors only (no CPU fallback)")
        N, D, Y = y.shape
        if D != layout.D:
            raise ValueError(f"y has {D} variables, layout has {layout.D}")
        if Y > _lib.MAX_Y:
            raise NotImplementedError(f"y_dim {Y} > {_lib.MAX_Y}")
        yd = y.detach()
        W = weight.detach().to(torch.float64).contiguous()
        b = bias.detach().to(torch.float64).contiguous()
        theta = torch.empty(N, layout.P, dtype=y.dtype, device=y.device)
        if N > 0:
            sn, sd, sk = yd.stride()
            _lib.call("hlvae_theta_fwd", N, D, layout.P, Y, layout.n_tiles, _lib.ptr(layout.col_var),
                      _lib.ptr(layout.col_mode), _lib.ptr(layout.var_pcol), _lib.ptr(layout.tile_var), _lib.ptr(W),
                      _lib.ptr(b), _lib.ptr(yd), sn, sd, sk, _lib.dtype_code(yd), _lib.ptr(theta), layout.P,
                      _lib.stream_ptr())
        mk, mk_code = _mask_arg(mask, y.dtype)
        ctx.layout, ctx.mk_code = layout, mk_code
        ctx.save_for_backward(yd, W, b, mk)
        return theta

    @staticmethod
    def backward(ctx, g_theta):
        yd, W, b, mk = ctx.saved_tensors
        layout = ctx.layout
        N, D, Y = yd.shape
        g_y = torch.empty_strided(yd.shape, yd.stride(), dtype=yd.dtype, device=yd.device)
        g_W = torch.zeros_like(W)
        g_b = torch.zeros_like(b)
        if N > 0:
            g = g_theta.to(yd.dtype).contiguous()
            sn, sd, sk = yd.stride()
            _lib.call("hlvae_theta_bwd", N, D, layout.P, Y, layout.n_tiles, layout.max_cols, _lib.ptr(layout.col_var),
                      _lib.ptr(layout.col_mode), _lib.ptr(layout.var_pcol), _lib.ptr(layout.tile_var), _lib.ptr(W),
                      _lib.ptr(b), _lib.ptr(yd), sn, sd, sk, _lib.dtype_code(yd), _lib.ptr(mk), ctx.mk_code, _lib.ptr(g),
                      layout.P, _lib.ptr(g_y), _lib.ptr(g_W), _lib.ptr(g_b), _lib.stream_ptr())
        else:
            g_y.zero_()
        return g_y, g_W, g_b, None, None


def theta_heads(layout: HeadLayout, y, mask, weight, bias):
    """theta [N, P_theta] from y [N, D, y_dim] (any strides; a tensor whose storage is not dense, i.e. with
    overlapping or gapped strides, is made contiguous first), the 0/1 observation mask [N, D] and the packed
    head parameters."""
    if y.dim() != 3:
        raise ValueError("y must be [N, D, y_dim]")
    if not _dense(y):
        y = y.contiguous()
    return _ThetaHeads.apply(y, weight, bias, mask, layout)


def _dense(t):
    """True when `t` is a permutation of a contiguous tensor (every element owned once, no gaps)."""
    sizes, strides = t.shape, t.stride()
    order = sorted(range(t.dim()), key=lambda i: (strides[i], sizes[i]))
    expect = 1
    for i in order:
        if sizes[i] == 1:
            continue
        if strides[i] != expect:
            return False
        expect *= sizes[i]
    return True


def _model_head_layout(model, device):
    lay = getattr(model, "_hlvae_b200_head_layout", None)
    if lay is None or str(lay.device) != str(device):
        ti = dict(model.types_info)
        ti['conv'] = bool(model.conv)
        lay = HeadLayout.from_types_info(ti, device)
        model._hlvae_b200_head_layout = lay
    return lay


def _logvar_plan(model, lay, device):
    """Index maps for logvar_network=True (HLVAE.py:30-51, read_functions.py:164-185): every real / positive variable
    owns TWO theta columns, and inside a type group the columns hold all means first, then all raw log-variances.
    Returns (variables with a log-variance head, in variable order; gather index into cat([theta_base, theta_lv], 1)
    that yields theta in the reference's column order; per-group positions of those variables)."""
    plan = getattr(model, "_hlvae_b200_logvar_plan", None)
    if plan is not None and plan["device"] == str(device):
        return plan
    ti = model.types_info
    dti, pidx = np.asarray(ti['data_types_indexes']), np.asarray(ti['param_indexes'])
    v = lay.var
    rp_vars = [d for d, (k, _) in enumerate(v.types) if k in ("real", "pos")]
    rp_pos = {d: j for j, d in enumerate(rp_vars)}
    P_base = v.P_theta
    gather = np.zeros(len(pidx), dtype=np.int64)
    for i, tpl in enumerate(ti['set_of_types']):
        cols = np.nonzero(pidx == i)[0]
        vars_g = np.nonzero(dti == i)[0]
        if tpl[0] in ("real", "pos"):
            n = len(vars_g)
            gather[cols[:n]] = [v.pcol_host[d] for d in vars_g]                    # means, HLVAE.py:51
            gather[cols[n:2 * n]] = [P_base + rp_pos[d] for d in vars_g]           # raw log-variances
        else:
            base_cols = np.concatenate([np.arange(v.pcol_host[d], v.pcol_host[d] + v.ncls_host[d]) for d in vars_g])
            gather[cols] = base_cols
    plan = dict(device=str(device), rp_vars=torch.tensor(rp_vars, dtype=torch.long, device=device),
                gather=torch.tensor(gather, dtype=torch.long, device=device),
                sub=HeadLayout([("count", 1)] * len(rp_vars), False, device), rp_list=rp_vars)
    model._hlvae_b200_logvar_plan = plan
    return plan


def _theta_estimation_logvar(model, y, miss_list):
    """theta_estimation with the variance network: the mean heads through the packed layout as usual, the
    log-variance heads (affine, no Sigmoid: HLVAE.py:431-433 applies it to the first cov_dim columns only) as a
    second launch over the real / positive variables, then one gather into the reference's column order."""
    lay = _model_head_layout(model, y.device)
    plan = _logvar_plan(model, lay, y.device)
    Y = y.shape[2]
    W, b = pack_heads(model.obs_layer, lay, Y)
    theta_base = theta_heads(lay, y, miss_list, W, b)
    n_rp = len(plan["rp_list"])
    Wl = torch.zeros(n_rp, Y, dtype=torch.float64, device=y.device)
    bl = torch.zeros(n_rp, dtype=torch.float64, device=y.device)
    layer = 0
    rp_pos = {d: j for j, d in enumerate(plan["rp_list"])}
    dti = np.asarray(model.types_info['data_types_indexes'])
    for i, tpl in enumerate(model.types_info['set_of_types']):
        mod = model.obs_layer[layer]
        if tpl[0] in ("real", "pos"):
            rows = torch.tensor([rp_pos[d] for d in np.nonzero(dti == i)[0]], dtype=torch.long, device=y.device)
            Wl = Wl.index_put((rows,), mod.weight_logvar[:, :, 0].to(torch.float64))
            bl = bl.index_put((rows,), mod.bias_logvar[:, 0].to(torch.float64))
            if tpl[0] == "real" and lay.conv:
                layer += 1
        layer += 1
    y_sub = y.index_select(1, plan["rp_vars"])
    theta_lv = theta_heads(plan["sub"], y_sub, miss_list.index_select(1, plan["rp_vars"]), Wl, bl)
    return torch.cat([theta_base, theta_lv], 1).index_select(1, plan["gather"])


def theta_estimation(self, y, miss_list, param_miss_list):
    """Drop-in for HLVAE.theta_estimation (HLVAE.py:416-453); bind with
    `HLVAE.theta_estimation = hlvae_b200.theta.theta_estimation`.  `param_miss_list` (the mask repeated per
    parameter column) is implied by `miss_list` and the layout and is not read.  Masks must be 0/1."""
    if getattr(self, "logvar_network", False):
        return _theta_estimation_logvar(self, y, miss_list)
    lay = _model_head_layout(self, y.device)
    W, b = pack_heads(self.obs_layer, lay, y.shape[2])
    return theta_heads(lay, y, miss_list, W, b)
