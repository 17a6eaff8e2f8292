"""CPU, world_size 2 over gloo: the data-parallel decomposition of SURVEY.md section 8(e).
Each rank evaluates the oracle on its own subjects; one all-reduce(sum) of the packed
accumulators must reproduce the unsharded sufficient statistics and scalars."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import helpers as h
    from hlvae_b200 import parallel, subjects
    from oracle import hlvae_oracle as orc
    parallel.enable()
    inp = h.make_kl_inputs(L=3, M=10, n_subj=7, T=6, seed=21, ragged=True)
    spec0, spec1 = orc.compile_spec(**inp["kargs"])
    prm0, prm1 = orc.KernelParams(inp["ros0"], inp["rls0"]), orc.KernelParams(inp["ros1"], inp["rls1"])
    subj = orc.split_subjects_by_id(inp["x"], 2)
    full = orc.kld_terms(spec0, prm0, spec1, prm1, inp["noise"], inp["m"], inp["H"], inp["x"], inp["mu"], inp["lv"],
                         inp["z"], subj, 1e-6)
    lo, hi = parallel.shard_subjects(len(subj), rank, world)
    mine = subj[lo:hi]
    rows = torch.cat(mine)
    # a rank sees only its own rows; F is the only term summed over rows outside the subject loop
    part = orc.kld_terms(spec0, prm0, spec1, prm1, inp["noise"], inp["m"], inp["H"], inp["x"], inp["mu"], inp["lv"],
                         inp["z"], mine, 1e-6)
    part_F = inp["lv"][rows].sum()
    # D and E contain replicated pieces; shard only their S-linear parts: compare S, p, A, B, C and F
    buf = torch.cat([part["S"].reshape(-1), part["p"].reshape(-1),
                     torch.stack([part["A"].reshape(()), part["B"].reshape(()), part["C"].reshape(()), part_F])])
    parallel.allreduce_sum_(buf)
    L, M = 3, 10
    S = buf[:L * M * M].view(L, M, M)
    p = buf[L * M * M:L * M * M + L * M].view(L, M, 1)
    sc = buf[-4:]
    ok = (h.rel_err(S, full["S"]) < 1e-12 and h.rel_err(p, full["p"]) < 1e-12 and
          h.rel_err(sc[0], full["A"]) < 1e-12 and h.rel_err(sc[1], full["B"]) < 1e-12 and
          h.rel_err(sc[2], full["C"]) < 1e-12 and h.rel_err(sc[3], full["F"]) < 1e-12)
    # layout sharding agrees with the subject split
    lay = subjects.SubjectLayout.from_ids(inp["x"][:, 2]).shard(rank, world)
    ok = ok and sorted(lay.row_idx.tolist()) == sorted(rows.tolist())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sum_of_partitions_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok in res), res


def _glue_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    import hlvae_b200  # noqa: F401
    from hlvae_b200 import parallel
    parallel.enable()
    torch.manual_seed(0)
    ids = np.repeat(np.array([7, 3, 9, 1, 4]), [3, 2, 4, 1, 3])          # 5 subjects, rows adjacent
    N = len(ids)
    data = torch.randn(N, 6, dtype=torch.float64)
    net = torch.nn.Linear(6, 2).double()                                  # stands for the NN trunk
    ref = torch.nn.Linear(6, 2).double()
    ref.load_state_dict(net.state_dict())
    # single-process loss: a sum over ALL rows (as nll and the row part of the KL bound are)
    (ref(data) ** 2).sum().backward()
    rows = parallel.shard_dataset_rows(ids, rank, world)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    parallel.attach_gradient_sync(opt, net.parameters())
    (net(data[rows]) ** 2).sum().backward()                               # this rank's partial sum
    opt.step()                                                            # pre-hook: one all-reduce(sum) of the bucket
    ok = all(torch.allclose(p.grad, r.grad, rtol=1e-12, atol=1e-14) for p, r in zip(net.parameters(), ref.parameters()))
    # whole subjects, disjoint, covering
    parts = [None] * world
    dist.all_gather_object(parts, rows.tolist())
    allrows = sorted(sum(parts, []))
    ok = ok and allrows == list(range(N)) and all(len(set(ids[np.array(p_)]) & set(ids[np.array(q_)])) == 0
                                                   for i, p_ in enumerate(parts) for q_ in parts[i + 1:])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_training_glue_sums_partial_gradients_gloo():
    """each rank owns whole subjects; NN-weight gradients are SUMMED through an optimiser
    step pre-hook and then equal the single-process gradients."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_glue_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok in res), res
