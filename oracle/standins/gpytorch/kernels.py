import os

import torch
from .constraints import Positive


class _Lazy:
    """What Kernel.__call__ returns: something with .evaluate()."""
    def __init__(self, fn):
        self._fn = fn

    def evaluate(self):
        return self._fn()

    def to_dense(self):
        return self._fn()


class Kernel(torch.nn.Module):
    has_lengthscale = False

    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), active_dims=None,
                 lengthscale_prior=None, lengthscale_constraint=None, eps=1e-6, **kwargs):
        super().__init__()
        self._batch_shape = torch.Size(batch_shape)
        if active_dims is not None and not torch.is_tensor(active_dims):
            active_dims = torch.tensor(active_dims, dtype=torch.long)
        self.register_buffer("active_dims", active_dims)
        self.ard_num_dims = ard_num_dims
        has_ls = kwargs.get("has_lengthscale", self.has_lengthscale)
        if has_ls:
            n = 1 if ard_num_dims is None else ard_num_dims
            self.register_parameter("raw_lengthscale",
                                    torch.nn.Parameter(torch.zeros(*self._batch_shape, 1, n)))
            self.raw_lengthscale_constraint = lengthscale_constraint or Positive()

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
            val = torch.as_tensor(val).to(raw)
            with torch.no_grad():
                raw.copy_(con.inverse_transform(val).expand_as(raw))
        return self

    def _select(self, x):
        if x.dim() == 1:
            x = x.unsqueeze(1)
        if self.active_dims is not None:
            x = x.index_select(-1, self.active_dims.reshape(-1))
        return x

    def __call__(self, x1, x2=None, **params):
        x2 = x1 if x2 is None else x2
        x1_, x2_ = self._select(x1), self._select(x2)
        return _Lazy(lambda: self.forward(x1_, x2_, **params))

    def __add__(self, other):
        ks = list(self.kernels) if isinstance(self, AdditiveKernel) else [self]
        ks += list(other.kernels) if isinstance(other, AdditiveKernel) else [other]
        return AdditiveKernel(*ks)

    def __mul__(self, other):
        ks = list(self.kernels) if isinstance(self, ProductKernel) else [self]
        ks += list(other.kernels) if isinstance(other, ProductKernel) else [other]
        return ProductKernel(*ks)


def _sq_dist(x1, x2):
    # gpytorch's matmul form with mean-centring and clamp at zero
    adj = x1.mean(-2, keepdim=True)
    x1 = x1 - adj
    x2 = x2 - adj
    x1n = x1.pow(2).sum(-1, keepdim=True)
    x2n = x2.pow(2).sum(-1, keepdim=True)
    a = torch.cat([-2.0 * x1, x1n, torch.ones_like(x1n)], dim=-1)
    b = torch.cat([x2, torch.ones_like(x2n), x2n], dim=-1)
    return a.matmul(b.transpose(-2, -1)).clamp_min(0)


class RBFKernel(Kernel):
    has_lengthscale = True

    def forward(self, x1, x2, **params):
        x1_ = x1.div(self.lengthscale)
        x2_ = x2.div(self.lengthscale)
        if os.environ.get("STANDIN_RBF_DIRECT"):      # diagnostic: closed-form distance instead of the matmul form
            d = x1_.unsqueeze(-2) - x2_.unsqueeze(-3)
            return d.pow(2).sum(-1).div(-2).exp()
        return _sq_dist(x1_, x2_).div(-2).exp()


class ScaleKernel(Kernel):
    def __init__(self, base_kernel, outputscale_prior=None, outputscale_constraint=None, **kwargs):
        if base_kernel.active_dims is not None:
            kwargs["active_dims"] = base_kernel.active_dims
        super().__init__(**kwargs)
        self.base_kernel = base_kernel
        self.register_parameter("raw_outputscale", torch.nn.Parameter(torch.zeros(self.batch_shape)))
        self.raw_outputscale_constraint = outputscale_constraint or Positive()

    @property
    def outputscale(self):
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)

    def forward(self, x1, x2, **params):
        out = self.base_kernel.forward(x1, x2, **params)
        s = self.outputscale
        return out * s.view(*s.shape, 1, 1)


class AdditiveKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = torch.nn.ModuleList(kernels)

    def forward(self, x1, x2, **params):
        res = 0
        for k in self.kernels:
            res = res + k(x1, x2, **params).evaluate()
        return res


class ProductKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = torch.nn.ModuleList(kernels)

    def forward(self, x1, x2, **params):
        res = self.kernels[0](x1, x2, **params).evaluate()
        for k in self.kernels[1:]:
            res = res * k(x1, x2, **params).evaluate()
        return res
