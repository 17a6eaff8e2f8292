"""A/B probe of the sufficient-statistics contraction S = K^T V (elbo_functions.py:161 / :254,266) on the two tensor
pipes (csrc/contraction_probe.cu): FP64 mma.sync against tcgen05.mma kind::i8 fed with an error-free 8-bit splitting
of the float64 operands.  Both must reproduce the float64 product; the integer path's error falls by 2^-8 per slice
and reaches float64 round-off at 7 slices - the accuracy oracle/emulate_tensor_contraction.py shows the path needs."""
import math

import pytest
import torch

from hlvae_b200 import _lib

pytestmark = pytest.mark.gpu


def _probe(mode, nslice, K, V, rows_per_cta=2048):
    L, N, M = K.shape
    S = torch.zeros(L, M, M, dtype=torch.float64, device=K.device)
    status = torch.zeros(4, dtype=torch.int32, device=K.device)
    ks = 2.0 ** math.ceil(math.log2(float(K.abs().max()) * (1 + 1e-12) + 1e-300))
    vs = 2.0 ** math.ceil(math.log2(float(V.abs().max()) * (1 + 1e-12) + 1e-300))
    _lib.call("hlvae_contraction_probe", mode, nslice, L, N, M, _lib.ptr(K), _lib.ptr(V), ks, vs, rows_per_cta,
              _lib.ptr(S), _lib.ptr(status), _lib.stream_ptr())
    torch.cuda.synchronize()
    assert int(status[0]) == 0, f"probe status {status.tolist()}"
    return S


@pytest.mark.parametrize("N", [32, 1000, 5000])
def test_both_pipes_reproduce_the_float64_product(N, device):
    gen = torch.Generator(device=device).manual_seed(N)
    L, M = 3, 64
    K = torch.rand(L, N, M, generator=gen, device=device, dtype=torch.float64) * 2.0        # K0xz >= 0
    V = torch.randn(L, N, M, generator=gen, device=device, dtype=torch.float64)
    ref = K.transpose(1, 2) @ V
    scale = float((K.abs().transpose(1, 2) @ V.abs()).max())
    rel = lambda S: float((S - ref).abs().max()) / scale
    assert rel(_probe(0, 0, K, V)) < 1e-15
    errs = {ns: rel(_probe(1, ns, K, V)) for ns in (3, 4, 5, 6, 7)}
    print(N, {k: f"{v:.1e}" for k, v in errs.items()})
    for ns in (3, 4, 5, 6):
        assert errs[ns] < 2.0 ** (-8 * ns + 6)             # truncation: the slice pairs of weight 2^-8 ns are dropped
    assert errs[7] < 1e-15
