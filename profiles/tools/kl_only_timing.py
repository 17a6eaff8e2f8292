import os, sys, torch
sys.path.insert(0, "/root/repo")
import bench
import __graft_entry__ as g
g.build()
from hlvae_b200 import config, elbo, loglik
from hlvae_b200.graph import StepGraph
config.check_errors = False
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
s = bench.build_gpu_state(dev, 800, 0)
s.update(side=None, side2=None)
def kl_only():
    for t_ in (s["mu"], s["lv"], s["z"], *s["k0"].parameters(), *s["k1"].parameters()): t_.grad = None
    kld, gm, gH = elbo.minibatch_KLD_upper_bound_iter(s["k0"], s["k1"], s["lik"], bench.L, s["m"], s["H"], s["x"], s["mu"], s["lv"], s["z"], bench.P_TOTAL, 800, bench.N_TOTAL, True, 2, bench.EPS, layout=s["layout"])
    m_new, H_new = elbo.natural_gradient_update(s["m"], s["H"], gm, gH, bench.NG_LR)
    kld.backward()
    return kld.detach()
def ll_only():
    s["theta"].grad = None; s["log_vy_real"].grad = None
    vparam = s["lay"].vparam(log_vy_real=s["log_vy_real"], conv=True)
    out = loglik.fused_loglik(s["lay"], s["data"], s["mask"], s["theta"], vparam, monitor=True)
    nll = -out["log_p_x_sum"]; nll.backward(); return nll.detach()
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, fn in (("kl_only", kl_only), ("ll_only", ll_only)):
    gr = StepGraph(fn, warmup=3)
    print(name, "graph ms", round(timeit(gr.replay), 4), "eager ms", round(timeit(fn), 4))
for ov in (True, False):
    config.overlap = ov
    gr = StepGraph(kl_only, warmup=3)
    print("kl_only overlap", ov, round(timeit(gr.replay), 4))
