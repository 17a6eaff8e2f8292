"""Kernel launches of one KL forward + natural-gradient update + backward, in launch order (torch profiler): what is
left of stock-PyTorch glue around the custom kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
import __graft_entry__ as g
g.build()
from hlvae_b200 import config, elbo
config.check_errors = False
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
s = bench.build_gpu_state(dev, 800, 0)
def kl_only():
    for t_ in (s["mu"], s["lv"], s["z"], *s["k0"].parameters(), *s["k1"].parameters()): t_.grad = None
    kld, gm, gH = elbo.minibatch_KLD_upper_bound_iter(s["k0"], s["k1"], s["lik"], bench.L, s["m"], s["H"], s["x"], s["mu"], s["lv"], s["z"], bench.P_TOTAL, 800, bench.N_TOTAL, True, 2, bench.EPS, layout=s["layout"])
    m_new, H_new = elbo.natural_gradient_update(s["m"], s["H"], gm, gH, bench.NG_LR)
    kld.backward()
for _ in range(3): kl_only()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    kl_only(); torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
for e in evs:
    print(f"{e.time_range.start - evs[0].time_range.start:9.1f} us  {e.cuda_time if hasattr(e,'cuda_time') else e.device_time:7.1f} us  {e.name[:90]}")
print("launches", len(evs))
