"""Multi-GPU (NCCL) parity of the subject-sharded KL upper bound, BASELINE.json configs[4] pattern:
each rank streams its own subjects, ONE all-reduce of the accumulator buffer, replicated M x M stage.
Needs >= 2 GPUs on the box (skipped otherwise): run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`.
Every rank must reproduce the single-GPU result of the whole minibatch: kld_total, grad_m, grad_H,
dZ and the kernel hyper-parameter gradients replicated; d mu / d log_v on the rank's own rows."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import helpers as h
    from hlvae_b200 import parallel, subjects
    try:
        L, M, P_b, T = 8, 64, 40, 20
        inp = h.make_kl_inputs(L, M, P_b, T, seed=5, ragged=True)
        full = h.run_kl_product(inp, dev)                     # unsharded, process-local (no group enabled yet)
        parallel.enable()
        lay = subjects.SubjectLayout.from_lengths(inp["lens"], dev).shard(rank, world)
        part = h.run_kl_product(inp, dev, layout=lay)         # P_batch stays the GLOBAL subject count
        parallel.disable()
        errs = {}
        for key in ("kld", "grad_m", "grad_H", "d_z", "d_m", "d_H", "d_os0", "d_ls0", "d_os1", "d_ls1"):
            errs[key] = h.rel_err(part[key], full[key])
        rows = lay.row_idx.long()
        errs["d_mu_own"] = h.rel_err(part["d_mu"][rows], full["d_mu"][rows])
        errs["d_logv_own"] = h.rel_err(part["d_logv"][rows], full["d_logv"][rows])
        other = torch.ones(inp["x"].shape[0], dtype=torch.bool, device=dev)
        other[rows] = False
        errs["d_mu_foreign"] = float(part["d_mu"][other].abs().max())
        # replicated outputs must be BIT-identical across ranks (they feed the replicated m, H update)
        flat = torch.cat([part["kld"].reshape(1), part["grad_m"].reshape(-1), part["grad_H"].reshape(-1)])
        ref = flat.clone()
        dist.broadcast(ref, 0)
        errs["replica_bits"] = 0.0 if torch.equal(flat, ref) else 1.0
        # only the summation order of the accumulators differs (per-rank partial sums, then NCCL);
        # cond(K0zz + eps I) ~ 1e7 amplifies that to ~1e-8 on the M x M outputs
        ok = all(v < 1e-6 for k, v in errs.items() if k not in ("d_os0", "d_ls0", "d_os1", "d_ls1")) and \
            all(h._hyper_ok(part[k], full[k], full["kld"], 1e-4) for k in ("d_os0", "d_ls0", "d_os1", "d_ls1"))
        bad = {k: float(v) for k, v in errs.items() if not v < 1e-6}      # hyper-gradients: scale-aware bound above
        q.put((rank, ok, {} if ok else (bad or {k: float(errs[k]) for k in ("d_os0", "d_ls0", "d_os1", "d_ls1")})))
    except Exception as e:  # surface the failure to the parent instead of hanging the queue
        q.put((rank, False, repr(e)))
    dist.destroy_process_group()


def test_sharded_kl_matches_single_gpu_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import __graft_entry__ as g
    g.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res), res
