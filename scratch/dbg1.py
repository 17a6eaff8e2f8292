import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, helpers as h
dev = torch.device("cuda:0")
for (L, M, n_subj, T, ragged) in [(3, 120, 12, 20, True), (2, 64, 1, 1, False), (2, 64, 2, 1, False), (2, 64, 1, 2, False)]:
    inp = h.make_kl_inputs(L, M, n_subj, T, 100 + M, ragged=ragged)
    got = h.run_kl_product(inp, dev)
    ref = h.oracle_kl(inp["kargs"], L, inp["x"], inp["mu"], inp["lv"], inp["z"], inp["m"], inp["H"], inp["ros0"], inp["rls0"], inp["ros1"], inp["rls1"], inp["noise"], 200, n_subj, 200 * T, 1e-6)
    print("CASE", L, M, n_subj, T)
    for k in ("kld", "grad_m", "grad_H", "d_mu", "d_logv", "d_z", "d_m", "d_H", "d_os0", "d_ls0", "d_os1", "d_ls1"):
        print("  ", k, "%.2e" % h.rel_err(got[k], ref[k]), "max|ref| %.3e" % float(torch.as_tensor(ref[k]).abs().max()))
    print("  d_os0 got", got["d_os0"].cpu().numpy().round(4).tolist())
    print("  d_os0 ref", ref["d_os0"].numpy().round(4).tolist())
    print("  d_os1 got", got["d_os1"].cpu().numpy().round(4).tolist())
    print("  d_os1 ref", ref["d_os1"].numpy().round(4).tolist())
