"""Additive covariance-kernel objects with the reference's operator surface.

Mirrors kernel_spec.py (BinKernel 9-23, CatKernel 26-32, RbfKernel 58-69) and the
gpytorch Scale/Product/Additive kernels that kernel_gen.generate_kernel_batched
(kernel_gen.py:199-310) assembles: same constructor arguments, same parameter names
(`raw_outputscale`, `raw_lengthscale`, `kernels.N...`, `base_kernel...`) and transforms
(softplus), so optimisers, `.double()`, `k0 + k1`, `state_dict()` and
`k(x1, x2).evaluate()` behave as the callers in HLVAE_main.py:215-236,
elbo_functions.py:147-151 and training.py:249-255 expect.  Evaluation runs in the
hand-written CUDA kernels behind the C ABI (include/hlvae_b200.h); there is no CPU path.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib

__all__ = ["Kernel", "RBFKernel", "RbfKernel", "CatKernel", "BinKernel", "ScaleKernel", "ProductKernel",
           "AdditiveKernel", "generate_kernel_batched", "compile_spec", "FlatSpec"]


def _inv_softplus(v: torch.Tensor) -> torch.Tensor:
    return v + torch.log(-torch.expm1(-v))


class _Positive(nn.Module):
    """Parameter constraint softplus(raw) (+ lower bound), with the buffers gpytorch keeps."""

    def __init__(self, lower_bound=0.0):
        super().__init__()
        self.lower_bound_value = float(lower_bound)      # host copy: the per-step path must not read the device
        self.register_buffer("lower_bound", torch.as_tensor(float(lower_bound)))
        self.register_buffer("upper_bound", torch.as_tensor(math.inf))

    def transform(self, raw):
        return F.softplus(raw) + self.lower_bound

    def inverse_transform(self, value):
        return _inv_softplus(value - self.lower_bound)


class _LazyKernelTensor:
    """Return value of `kernel(x1, x2)`; `.evaluate()` gives the dense matrix."""

    def __init__(self, kernel, x1, x2):
        self.kernel, self.x1, self.x2 = kernel, x1, x2

    def evaluate(self):
        return evaluate_dense(self.kernel, self.x1, self.x2)

    to_dense = evaluate


class Kernel(nn.Module):
    has_lengthscale = False

    def __init__(self, active_dims=None, batch_shape=torch.Size([]), has_lengthscale=None, **kwargs):
        super().__init__()
        self._batch_shape = torch.Size(batch_shape)
        if active_dims is not None and not torch.is_tensor(active_dims):
            active_dims = torch.tensor(active_dims, dtype=torch.long)
        self.register_buffer("active_dims", active_dims)
        # host copy of the covariate column: compile_spec runs every step and must not read the device; refreshed
        # whenever the `active_dims` buffer (the source of truth, as in gpytorch) is replaced or written in place
        self._columns, self._columns_tag = None, None
        self._refresh_columns()
        if has_lengthscale is None:
            has_lengthscale = type(self).has_lengthscale
        if has_lengthscale:
            self.register_parameter("raw_lengthscale", nn.Parameter(torch.zeros(*self._batch_shape, 1, 1)))
            self.raw_lengthscale_constraint = _Positive()

    @property
    def batch_shape(self):
        return self._batch_shape

    @property
    def lengthscale(self):
        return self.raw_lengthscale_constraint.transform(self.raw_lengthscale)

    def initialize(self, **kwargs):
        for name, val in kwargs.items():
            raw = getattr(self, "raw_" + name)
            con = getattr(self, "raw_" + name + "_constraint")
            with torch.no_grad():
                raw.copy_(con.inverse_transform(torch.as_tensor(val).to(raw)).expand_as(raw))
        return self

    def _refresh_columns(self):
        ad = self.active_dims
        tag = None if ad is None else (ad.data_ptr(), ad._version, ad.device)
        if tag != self._columns_tag:
            self._columns = None if ad is None else [int(v) for v in ad.reshape(-1).tolist()]
            self._columns_tag = tag

    def _column(self):
        if self.active_dims is None:
            raise ValueError(f"{type(self).__name__} needs active_dims (a covariate column)")
        self._refresh_columns()      # load_state_dict / reassignment of the buffer changes the tag (no device read otherwise)
        if len(self._columns) != 1:
            raise ValueError("only one covariate column per base kernel is supported")
        return self._columns[0]

    def __call__(self, x1, x2=None, **params):
        return _LazyKernelTensor(self, x1, x1 if x2 is None else x2)

    def __add__(self, other):
        ks = list(self.kernels) if isinstance(self, AdditiveKernel) else [self]
        ks += list(other.kernels) if isinstance(other, AdditiveKernel) else [other]
        return AdditiveKernel(*ks)

    def __mul__(self, other):
        ks = list(self.kernels) if isinstance(self, ProductKernel) else [self]
        ks += list(other.kernels) if isinstance(other, ProductKernel) else [other]
        return ProductKernel(*ks)


class BinKernel(Kernel):
    """kernel_spec.py:9-23: 1 where x1 + x2 == 2."""

    def __init__(self, value=1, **kwargs):
        super().__init__(has_lengthscale=False, **kwargs)
        self.value = value


class CatKernel(Kernel):
    """kernel_spec.py:26-32: 1 where x1 == x2."""


class RBFKernel(Kernel):
    """gpytorch RBFKernel as used by kernel_spec.py:58-69: exp(-(x1-x2)^2 / (2 l^2))."""
    has_lengthscale = True


def RbfKernel(active_dims, batch_shape=None):
    """kernel_spec.py:58-69 (lengthscale initialised to 2.5 in the parameter's dtype)."""
    k = RBFKernel(active_dims=active_dims) if batch_shape is None else RBFKernel(active_dims=active_dims,
                                                                               batch_shape=batch_shape)
    k.initialize(lengthscale=2.5)
    return k


class ScaleKernel(Kernel):
    def __init__(self, base_kernel, batch_shape=torch.Size([]), **kwargs):
        if base_kernel.active_dims is not None:
            kwargs["active_dims"] = base_kernel.active_dims
        super().__init__(batch_shape=batch_shape, **kwargs)
        self.base_kernel = base_kernel
        self.register_parameter("raw_outputscale", nn.Parameter(torch.zeros(self.batch_shape)))
        self.raw_outputscale_constraint = _Positive()

    @property
    def outputscale(self):
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)


class ProductKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = nn.ModuleList(kernels)


class AdditiveKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = nn.ModuleList(kernels)


# --------------------------------------------------------------------------------------
# Flattening the module tree into the C descriptor
# --------------------------------------------------------------------------------------
class FlatSpec:
    """One additive kernel as the C ABI sees it: `cspec` (hlvae_kspec_t) plus, per component,
    the modules that own its outputscale and lengthscale."""

    def __init__(self):
        self.cspec = _lib.KSpec()
        self.lb_cache = {}        # device copies of non-zero constraint lower bounds (owned by the kernel module)
        self.scale_mods = []      # ScaleKernel or None (unit scale)
        self.rbf_mods = []        # RBFKernel or None

    @property
    def ncomp(self):
        return self.cspec.ncomp

    def signature(self):
        s = self.cspec
        return tuple((s.comp[r].se_col, tuple((s.comp[r].disc_kind[f], s.comp[r].disc_col[f])
                                              for f in range(s.comp[r].ndisc))) for r in range(s.ncomp))

    def latent_dim(self):
        L = 1
        for m in self.scale_mods:
            if m is not None and m.raw_outputscale.dim() > 0:
                L = max(L, m.raw_outputscale.shape[0])
        for m in self.rbf_mods:
            if m is not None and m.raw_lengthscale.dim() > 2:
                L = max(L, m.raw_lengthscale.shape[0])
        return L

    def constrained(self, L, device, dtype=torch.float64):
        """(outputscale, lengthscale) each [ncomp, L], attached to autograd.  The raw parameters of all components
        go through ONE softplus per kind (stack -> softplus -> + lower bounds) instead of one per module: the same
        values bit for bit, a dozen fewer launches per step in front of the KL kernels."""
        if not self.scale_mods:
            z = torch.zeros(0, L, dtype=dtype, device=device)
            return z, z
        return (self._constrained_rows(self.scale_mods, "raw_outputscale", L, device, dtype),
                self._constrained_rows(self.rbf_mods, "raw_lengthscale", L, device, dtype))

    def _constrained_rows(self, mods, raw_name, L, device, dtype):
        live = [i for i, m in enumerate(mods) if m is not None]
        if not live:
            return torch.ones(len(mods), L, dtype=dtype, device=device)
        raws = torch.stack([getattr(mods[i], raw_name).to(dtype).reshape(-1).expand(L) for i in live])
        vals = F.softplus(raws)
        lbs = [getattr(mods[i], raw_name + "_constraint").lower_bound_value for i in live]   # fixed at construction
        if any(v != 0.0 for v in lbs):
            key = (raw_name, str(device), dtype, tuple(lbs))
            if key not in self.lb_cache:
                self.lb_cache[key] = torch.tensor(lbs, dtype=dtype, device=device).reshape(-1, 1)
            vals = vals + self.lb_cache[key]
        if len(live) == len(mods):
            return vals.contiguous()
        one = torch.ones(L, dtype=dtype, device=device)       # components without this parameter: unit rows
        pos = {i: j for j, i in enumerate(live)}              # (no device index tensors: the step is graph-captured)
        return torch.stack([vals[pos[i]] if i in pos else one for i in range(len(mods))]).contiguous()


class _ConstrainRows(torch.autograd.Function):
    """softplus(raw) + lower bound for every outputscale / lengthscale of both additive kernels in ONE launch
    (hlvae_hyper_constrain), and one launch for the chain rule on the way back.  `plan`: per output row the index
    into `raws` or -1 (no parameter: the row is 1) and the lower bound; `blocks`: the row ranges handed back as
    separate tensors (outputscale / lengthscale of kernel 0 / 1)."""

    @staticmethod
    def forward(ctx, plan, blocks, L, *raws):
        import ctypes as C
        n_rows = len(plan)
        dev = raws[0].device
        out = torch.empty(n_rows, L, dtype=torch.float64, device=dev)
        ptrs = (C.c_void_p * n_rows)(*[raws[j].data_ptr() if j >= 0 else None for j, _ in plan])
        bc = (C.c_int32 * n_rows)(*[1 if (j >= 0 and raws[j].numel() == 1 and L > 1) else 0 for j, _ in plan])
        lb = (C.c_double * n_rows)(*[float(b) for _, b in plan])
        ctx.hold = (ptrs, bc, lb, n_rows, L, plan, blocks)
        ctx.save_for_backward(*raws)
        _lib.call("hlvae_hyper_constrain", n_rows, L, ptrs, bc, lb, _lib.ptr(out), None, None, _lib.stream_ptr())
        return tuple(out[a:b] for a, b in blocks)

    @staticmethod
    def backward(ctx, *g_blocks):
        ptrs, bc, lb, n_rows, L, plan, blocks = ctx.hold
        raws = ctx.saved_tensors
        dev = raws[0].device
        parts = [g if g is not None else torch.zeros(b - a, L, dtype=torch.float64, device=dev)
                 for g, (a, b) in zip(g_blocks, blocks)]
        g_out = torch.cat(parts) if len(parts) > 1 else parts[0].contiguous()
        g_raw = torch.empty(n_rows, L, dtype=torch.float64, device=dev)
        _lib.call("hlvae_hyper_constrain", n_rows, L, ptrs, bc, lb, None, _lib.ptr(g_out), _lib.ptr(g_raw),
                  _lib.stream_ptr())
        grads = [None] * len(raws)
        for row, (j, _) in enumerate(plan):
            if j >= 0:
                r = raws[j]
                grads[j] = (g_raw[row, :1] if r.numel() == 1 and L > 1 else g_raw[row]).reshape(r.shape)
        return (None, None, None) + tuple(grads)


def constrained_pair(fs0, fs1, L, device):
    """(outputscale0, lengthscale0, outputscale1, lengthscale1), each [ncomp, L] float64 and attached to autograd:
    what FlatSpec.constrained returns for the two kernels of the KL bound, computed by one launch.  Falls back to the
    stock-PyTorch formulation when a raw parameter is not a contiguous float64 CUDA tensor of L (or 1) values."""
    plan, raws = [], []
    blocks = []
    for fs in (fs0, fs1):
        for mods, name in ((fs.scale_mods, "raw_outputscale"), (fs.rbf_mods, "raw_lengthscale")):
            start = len(plan)
            for m in mods:
                if m is None:
                    plan.append((-1, 0.0))
                else:
                    r = getattr(m, name)
                    if not (r.is_cuda and r.dtype == torch.float64 and r.is_contiguous() and r.numel() in (1, L)):
                        return fs0.constrained(L, device) + fs1.constrained(L, device)
                    plan.append((len(raws), getattr(m, name + "_constraint").lower_bound_value))
                    raws.append(r)
            blocks.append((start, len(plan)))
    if not raws or len(plan) > 4 * _lib.MAX_COMPS:
        return fs0.constrained(L, device) + fs1.constrained(L, device)
    return _ConstrainRows.apply(tuple(plan), tuple(blocks), L, *raws)


def _flatten_product(k, factors):
    if isinstance(k, ProductKernel):
        for c in k.kernels:
            _flatten_product(c, factors)
    elif isinstance(k, (RBFKernel, CatKernel, BinKernel)):
        factors.append(k)
    else:
        raise TypeError(f"unsupported kernel inside a product: {type(k).__name__}")


def compile_spec(kernel) -> FlatSpec:
    """Flatten AdditiveKernel(ScaleKernel(product of base kernels), ...) into a FlatSpec."""
    terms = list(kernel.kernels) if isinstance(kernel, AdditiveKernel) else [kernel]
    fs = FlatSpec()
    if len(terms) > _lib.MAX_COMPS:
        raise ValueError(f"at most {_lib.MAX_COMPS} additive components are supported")
    for r, t in enumerate(terms):
        scale = t if isinstance(t, ScaleKernel) else None
        factors = []
        _flatten_product(t.base_kernel if scale is not None else t, factors)
        comp = fs.cspec.comp[r]
        comp.se_col, comp.ndisc = -1, 0
        rbf = None
        for f in factors:
            if isinstance(f, RBFKernel):
                if rbf is not None:
                    raise ValueError("at most one squared-exponential factor per component is supported")
                rbf, comp.se_col = f, f._column()
            else:
                if comp.ndisc >= _lib.MAX_DISC:
                    raise ValueError(f"at most {_lib.MAX_DISC} categorical/binary factors per component")
                comp.disc_kind[comp.ndisc] = _lib.KIND_CAT if isinstance(f, CatKernel) else _lib.KIND_BIN
                comp.disc_col[comp.ndisc] = f._column()
                comp.ndisc += 1
        fs.scale_mods.append(scale)
        fs.rbf_mods.append(rbf)
    fs.cspec.ncomp = len(terms)
    fs.lb_cache = kernel.__dict__.setdefault("_hlvae_lb_cache", {})
    return fs


# --------------------------------------------------------------------------------------
# Dense evaluation (autograd)
# --------------------------------------------------------------------------------------
class _KernelEval(torch.autograd.Function):
    """out[l, i, j] = sum_r os[r, l] * comp_r(x1[l?, i], x2[l?, j]); x* are [n, Q] or [L, n, Q]."""

    @staticmethod
    def forward(ctx, fs, os_, ls_, x1, x2):
        L = os_.shape[1]
        x1c, x2c = x1.detach().contiguous(), x2.detach().contiguous()
        Q = x1c.shape[-1]
        n1, n2 = x1c.shape[-2], x2c.shape[-2]
        out = torch.empty(L, n1, n2, dtype=torch.float64, device=x1c.device)
        os_c, ls_c = os_.detach().contiguous(), ls_.detach().contiguous()
        bs1 = n1 * Q if x1c.dim() == 3 else 0
        bs2 = n2 * Q if x2c.dim() == 3 else 0
        ctx.fs = fs
        ctx.save_for_backward(os_c, ls_c, x1c, x2c)
        if n1 == 0 or n2 == 0:
            return out
        _lib.call("hlvae_kernel_eval_fwd", fs.cspec, _lib.ptr(os_c), _lib.ptr(ls_c), L, Q, _lib.ptr(x1c), n1,
                                                    Q, bs1, _lib.ptr(x2c), n2, Q, bs2, _lib.ptr(out),
                                                    _lib.stream_ptr())
        return out

    @staticmethod
    def backward(ctx, g):
        os_c, ls_c, x1c, x2c = ctx.saved_tensors
        fs = ctx.fs
        L = os_c.shape[1]
        Q = x1c.shape[-1]
        n1, n2 = x1c.shape[-2], x2c.shape[-2]
        g = g.contiguous()
        g_os, g_ls = torch.zeros_like(os_c), torch.zeros_like(ls_c)
        need1, need2 = ctx.needs_input_grad[3], ctx.needs_input_grad[4]
        g_x1 = torch.zeros(L, n1, Q, dtype=torch.float64, device=g.device) if need1 else None
        g_x2 = torch.zeros(L, n2, Q, dtype=torch.float64, device=g.device) if need2 else None
        bs1 = n1 * Q if x1c.dim() == 3 else 0
        bs2 = n2 * Q if x2c.dim() == 3 else 0
        if n1 > 0 and n2 > 0:
            _lib.call("hlvae_kernel_eval_bwd", fs.cspec, _lib.ptr(os_c), _lib.ptr(ls_c), L, Q, _lib.ptr(x1c),
                                                        n1, Q, bs1, _lib.ptr(x2c), n2, Q, bs2, _lib.ptr(g),
                                                        _lib.ptr(g_os), _lib.ptr(g_ls), _lib.ptr(g_x1),
                                                        _lib.ptr(g_x2), _lib.stream_ptr())
        if g_x1 is not None and x1c.dim() == 2:
            g_x1 = g_x1.sum(0)
        if g_x2 is not None and x2c.dim() == 2:
            g_x2 = g_x2.sum(0)
        return None, g_os, g_ls, g_x1, g_x2


def evaluate_dense(kernel, x1, x2):
    """`kernel(x1, x2).evaluate()` for x of shape [n,Q], [L,n,Q] or [P,L,n,Q]
    (elbo_functions.py:147-150, 248-249; utils.py:128-130)."""
    fs = compile_spec(kernel)
    if not x1.is_cuda or not x2.is_cuda:
        raise RuntimeError("hlvae_b200 kernels evaluate on CUDA tensors only (no CPU fallback)")
    x1 = x1.to(torch.float64)
    x2 = x2.to(torch.float64)
    L = fs.latent_dim()
    for xx in (x1, x2):
        if xx.dim() >= 3:
            L = max(L, xx.shape[-3])
    os_, ls_ = fs.constrained(L, x1.device)
    nd = max(x1.dim(), x2.dim())
    if nd <= 3:
        a = x1 if x1.dim() == 2 else x1.expand(L, *x1.shape[-2:])
        b = x2 if x2.dim() == 2 else x2.expand(L, *x2.shape[-2:])
        out = _KernelEval.apply(fs, os_, ls_, a, b)
        unbatched = fs.latent_dim() == 1 and x1.dim() == 2 and x2.dim() == 2 and \
            all(m is None or m.raw_outputscale.dim() == 0 for m in fs.scale_mods)
        return out[0] if unbatched else out
    if nd == 4:
        P = max(x1.shape[0] if x1.dim() == 4 else 1, x2.shape[0] if x2.dim() == 4 else 1)
        a = x1.expand(P, L, *x1.shape[-2:]).reshape(P * L, *x1.shape[-2:])
        b = x2.expand(P, L, *x2.shape[-2:]).reshape(P * L, *x2.shape[-2:])
        out = _KernelEval.apply(fs, os_.repeat(1, P), ls_.repeat(1, P), a, b)
        return out.reshape(P, L, out.shape[-2], out.shape[-1])
    raise ValueError("kernel inputs must have 2, 3 or 4 dimensions")


# --------------------------------------------------------------------------------------
# kernel_gen.generate_kernel_batched
# --------------------------------------------------------------------------------------
def _build_additive(bs, cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel, bin_int_kernel, covariate_missing_val,
                    id_covariate):
    """The two additive kernels every generator of kernel_gen.py assembles - same component order (categorical,
    squared-exponential, binary, categorical x continuous, binary x continuous), same masking by BinKernel factors
    of missing covariates, same module tree; `bs` is the batch shape of the hyper-parameters and `id_covariate`
    (None: everything goes to the first kernel) routes the terms that carry the id covariate to the second."""
    missing = {d['covariate']: d['mask'] for d in covariate_missing_val}
    k0, k1 = AdditiveKernel(), AdditiveKernel()
    kw = {} if bs is None else dict(batch_shape=bs)

    def with_mask(k, col):
        return k * BinKernel(active_dims=missing[col], value=1) if col in missing else k

    def rbf(col):
        return RbfKernel(active_dims=col, batch_shape=bs)

    def is_id(col):
        return id_covariate is not None and col == id_covariate

    for idx in cat_kernel:                                                   # :225-242 / :27-35 / :121-137
        (k1 if is_id(idx) else k0).kernels.append(ScaleKernel(with_mask(CatKernel(active_dims=idx), idx), **kw))
    for idx in sqexp_kernel:                                                 # :245-254 / :38-46 / :140-148
        k0.kernels.append(ScaleKernel(with_mask(rbf(idx), idx), **kw))
    for idx in bin_kernel:                                                   # :257-266 / :49-57 / :151-159
        k0.kernels.append(ScaleKernel(with_mask(BinKernel(active_dims=idx, value=1), idx), **kw))
    for d in cat_int_kernel:                                                 # :269-289 / :60-75 / :162-181
        a = with_mask(CatKernel(active_dims=d['cat_covariate']), d['cat_covariate'])
        b = with_mask(rbf(d['cont_covariate']), d['cont_covariate'])
        (k1 if is_id(d['cat_covariate']) else k0).kernels.append(ScaleKernel(ProductKernel(a, b), **kw))
    for d in bin_int_kernel:                                                 # :292-308 / :78-93 / :184-195
        a = with_mask(BinKernel(active_dims=d['bin_covariate'], value=1), d['bin_covariate'])
        b = with_mask(rbf(d['cont_covariate']), d['cont_covariate'])
        k0.kernels.append(ScaleKernel(ProductKernel(a, b), **kw))
    return k0, k1


def generate_kernel_batched(latent_dim, cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel, bin_int_kernel,
                            covariate_missing_val, id_covariate):
    """Same signature, component order and module structure as kernel_gen.py:199-310:
    returns (additive kernel without the id covariate, additive kernel with it), hyper-parameters batched over the
    latent dimensions."""
    k0, k1 = _build_additive(torch.Size([latent_dim]), cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel,
                             bin_int_kernel, covariate_missing_val, id_covariate)
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")    # :219, :310
    return k0.to(device), k1.to(device)


def generate_kernel_approx(cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel, bin_int_kernel, covariate_missing_val,
                           id_covariate):
    """kernel_gen.py:97-197: the un-batched pair (one latent dimension per kernel object), as
    elbo_functions.elbo / deviance_upper_bound take them."""
    return _build_additive(None, cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel, bin_int_kernel,
                           covariate_missing_val, id_covariate)


def generate_kernel(cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel, bin_int_kernel, covariate_missing_val):
    """kernel_gen.py:9-94: ONE un-batched additive kernel holding every term."""
    return _build_additive(None, cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel, bin_int_kernel,
                           covariate_missing_val, None)[0]
