"""Three KL forward + backward calls at the configs[1] shape (L=32, M=64, 800 subjects x T=20, float32 mu / log_v):
the short command that ncu profiles for the KL kernels (ncu -k regex:kl_subject|kl_panel ...)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
import __graft_entry__ as g
g.build()
from hlvae_b200 import config, elbo
config.check_errors = False
config.overlap = False
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
s = bench.build_gpu_state(dev, int(os.environ.get("SUBJECTS", "800")), 0)
for _ in range(3):
    for t_ in (s["mu"], s["lv"], s["z"], *s["k0"].parameters(), *s["k1"].parameters()):
        t_.grad = None
    kld, gm, gH = elbo.minibatch_KLD_upper_bound_iter(s["k0"], s["k1"], s["lik"], bench.L, s["m"], s["H"], s["x"], s["mu"],
                                                      s["lv"], s["z"], bench.P_TOTAL, s["n_subj"], bench.N_TOTAL, True, 2,
                                                      bench.EPS, layout=s["layout"])
    kld.backward()
torch.cuda.synchronize()
print("kld", float(kld))
if os.environ.get("WITH_NATGRAD"):
    for _ in range(2):
        elbo.natural_gradient_update(s["m"], s["H"], gm, gH, bench.NG_LR)
    torch.cuda.synchronize()
