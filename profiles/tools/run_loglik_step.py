"""Three fused-likelihood forward + backward calls at the configs[1] shape (16 000 rows, D4: 324 real + 972
categorical x 5, float32 theta, uint8 data / mask): the short command ncu profiles for the likelihood kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import __graft_entry__ as g
g.build()
from hlvae_b200 import loglik, synth
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
N = int(os.environ.get("ROWS", "16000"))
if os.environ.get("LAYOUT") == "tabular":                  # BASELINE.json configs[3]: float32 data, uint8 mask
    import numpy as np
    lay = loglik.VarLayout(synth.TABULAR_TYPES, dev)
    d4k, m4k = synth.likelihood_batch(synth.TABULAR_TYPES, 4000, np.random.default_rng(9))
    data = d4k.float().repeat(N // 4000, 1).to(dev)
    mask = m4k.to(torch.uint8).repeat(N // 4000, 1).to(dev)
    N = data.shape[0]
    z32 = torch.zeros(32, dtype=torch.float64, device=dev)
    vparam = lay.vparam(z32.clone(), z32.clone(), [z32, torch.ones_like(z32)], [z32, torch.ones_like(z32)])
else:
    lay = loglik.VarLayout(synth.HEALTHMNIST_D4_TYPES, dev)
    data, mask = synth.device_likelihood_batch(lay, N, dev, gen, dtype=torch.uint8)
    vparam = lay.vparam(log_vy_real=torch.zeros(324, dtype=torch.float64, device=dev), conv=True)
theta = torch.randn(N, lay.P_theta, device=dev, generator=gen)
if os.environ.get("STORAGE") == "f64":                     # the reference's storage: every tensor float64
    theta, data, mask = theta.double(), data.double(), mask.double()
theta.requires_grad_(True)
for _ in range(3):
    theta.grad = None
    out = loglik.fused_loglik(lay, data, mask, theta, vparam, monitor=True)
    (-out["log_p_x_sum"]).backward()
torch.cuda.synchronize()
print("nll", float(-out["log_p_x_sum"]))
if os.environ.get("TIME_LL"):
    import numpy as np
    from hlvae_b200 import _lib
    def step():
        theta.grad = None
        out = loglik.fused_loglik(lay, data, mask, theta, vparam, monitor=True)
        (-out["log_p_x_sum"]).backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    _lib.PROFILE = []
    for _ in range(20): step()
    torch.cuda.synchronize()
    per = {}
    for name, a, b in _lib.PROFILE: per.setdefault(name, []).append(a.elapsed_time(b))
    _lib.PROFILE = None
    print({k[6:]: round(float(np.mean(v)), 4) for k, v in per.items()})
