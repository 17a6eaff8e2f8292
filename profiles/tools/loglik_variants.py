"""Timing probe: fused likelihood forward with and without the monitoring outputs (configs[1] shape)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
import __graft_entry__ as g
g.build()
from hlvae_b200 import _lib, loglik, synth
dev = torch.device("cuda:0")
lay = loglik.VarLayout(synth.HEALTHMNIST_D4_TYPES, dev)
gen = torch.Generator(device=dev).manual_seed(0)
N = 16000
data, mask = synth.device_likelihood_batch(lay, N, dev, gen, dtype=torch.uint8)
theta = torch.randn(N, lay.P_theta, device=dev, generator=gen)
lvr = torch.zeros(324, dtype=torch.float64, device=dev)
vparam = lay.vparam(log_vy_real=lvr, conv=True)
def run(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    _lib.PROFILE = []
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    per = {}
    for name, a, b in _lib.PROFILE: per.setdefault(name, []).append(a.elapsed_time(b))
    _lib.PROFILE = None
    return {k[6:]: round(float(np.mean(v)), 3) for k, v in per.items()}
print("monitor=True ", run(lambda: loglik.fused_loglik(lay, data, mask, theta, vparam, monitor=True)))
print("monitor=False", run(lambda: loglik.fused_loglik(lay, data, mask, theta, vparam, monitor=False)))
df = data.float()
print("f32 data     ", run(lambda: loglik.fused_loglik(lay, df, mask, theta, vparam, monitor=True)))
