import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import __graft_entry__ as g
g.build()
from hlvae_b200 import _lib, config, elbo, loglik
config.check_errors = False
dev = torch.device("cuda:0")
s = bench.build_gpu_state(dev, 40, 0)
L = bench.L

def piece_loglik():
    s["theta"].grad = None; s["log_vy_real"].grad = None
    vparam = s["lay"].vparam(log_vy_real=s["log_vy_real"], conv=True)
    out = loglik.fused_loglik(s["lay"], s["data"], s["mask"], s["theta"], vparam, monitor=True)
    (-out["log_p_x_sum"] * 2.0).backward()
    return out["log_p_x_sum"].detach()

def piece_vparam():
    return s["lay"].vparam(log_vy_real=s["log_vy_real"], conv=True).detach()

def piece_kld_fwd():
    kld, gm, gH = elbo.minibatch_KLD_upper_bound_iter(s["k0"], s["k1"], s["lik"], L, s["m"], s["H"], s["x"], s["mu"], s["lv"], s["z"], 5000, 40, 100000, True, 2, 1e-6, layout=s["layout"])
    return kld.detach()

def piece_kld():
    for t_ in (s["mu"], s["lv"], s["z"], *s["k0"].parameters(), *s["k1"].parameters()):
        t_.grad = None
    kld, gm, gH = elbo.minibatch_KLD_upper_bound_iter(s["k0"], s["k1"], s["lik"], L, s["m"], s["H"], s["x"], s["mu"], s["lv"], s["z"], 5000, 40, 100000, True, 2, 1e-6, layout=s["layout"])
    kld.sum().backward()
    return kld.detach()

def piece_constrained():
    from hlvae_b200.kernels import compile_spec
    fs0 = compile_spec(s["k0"])
    a, b = fs0.constrained(L, dev)
    return a.detach()

def piece_noise():
    return elbo._noise_vector(s["lik"], L, dev)

def piece_natgrad():
    gm = torch.zeros_like(s["m"]); gH = torch.zeros_like(s["H"])
    m2, H2 = elbo.natural_gradient_update(s["m"], s["H"], gm, gH, 0.01)
    return m2

for name in ("piece_vparam", "piece_constrained", "piece_noise", "piece_natgrad", "piece_loglik", "piece_kld_fwd", "piece_kld"):
    fn = globals()[name]
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn(); fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            out = fn()
        gr.replay(); torch.cuda.synchronize()
        print(name, "OK", flush=True)
    except Exception as e:
        print(name, "FAILED", repr(e)[:300], flush=True)
        traceback.print_exc(limit=6)
        torch.cuda.synchronize()
