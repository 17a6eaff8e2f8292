"""Observation-head forward + backward calls at the configs[1] shape (16 000 rows, D4, y_dim 5, conv layout, float32
storage, uint8 mask): the short command ncu profiles for hlvae_theta_fwd / _bwd; TIME_TH=1 prints event timings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
from hlvae_b200 import _lib, synth, theta as th
dev = torch.device("cuda:0")
tabular = os.environ.get("LAYOUT") == "tabular"
types = synth.TABULAR_TYPES if tabular else synth.HEALTHMNIST_D4_TYPES
N, Y, D = int(os.environ.get("ROWS", "16000")), int(os.environ.get("YDIM", "5")), len(types)
gen = torch.Generator(device=dev).manual_seed(0)
lay = th.HeadLayout(types, not tabular, dev)
P = lay.P
if tabular:
    y = torch.randn(N, D, Y, generator=gen, device=dev, dtype=torch.float32).requires_grad_(True)
else:
    y = torch.randn(N, Y, D, generator=gen, device=dev, dtype=torch.float32).permute(0, 2, 1).requires_grad_(True)
mask = (torch.rand(N, D, generator=gen, device=dev) < 0.75).to(torch.uint8)
W = (torch.randn(P, Y, generator=gen, device=dev, dtype=torch.float64) * 0.3).requires_grad_(True)
b = (torch.randn(P, generator=gen, device=dev, dtype=torch.float64) * 0.3).requires_grad_(True)
g_up = torch.randn(N, P, generator=gen, device=dev, dtype=torch.float32)


def step():
    y.grad = W.grad = b.grad = None
    th.theta_heads(lay, y, mask, W, b).backward(g_up)


for _ in range(3):
    step()
torch.cuda.synchronize()
print("checksum", float(y.grad.double().sum()), float(W.grad.sum()))
if os.environ.get("TIME_TH"):
    _lib.PROFILE = []
    for _ in range(20):
        step()
    torch.cuda.synchronize()
    per = {}
    for name, a, c in _lib.PROFILE:
        per.setdefault(name, []).append(a.elapsed_time(c))
    _lib.PROFILE = None
    print({k[6:]: round(float(np.mean(v)), 4) for k, v in per.items()})
    bf, bb = N * (4 * D * Y + 4 * P), N * (4 * D * Y + D + 4 * P + 4 * D * Y)
    print("GB/s fwd", round(bf / np.mean(per["hlvae_theta_fwd"]) / 1e6), "bwd", round(bb / np.mean(per["hlvae_theta_bwd"]) / 1e6))
