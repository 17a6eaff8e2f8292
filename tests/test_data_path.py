"""CPU: the device data path (hl-vae_b200/data.py) composes minibatches exactly like the
reference's samplers (utils.py:36-97) - checked against the frozen outputs of the reference classes
(tests/golden/samplers.npz) and against the oracle restatement on fresh seeds - and gathers the right rows."""
import numpy as np
import torch

import helpers as h
from hlvae_b200 import data as D
from oracle import hlvae_oracle as orc


def _split(rows, ptr):
    return [rows[ptr[i]:ptr[i + 1]].tolist() for i in range(len(ptr) - 1)]


def _dataset(ids, E_x=7, Dv=3, Q=4):
    n = len(ids)
    rng = np.random.default_rng(0)
    cov = np.zeros((n, Q))
    cov[:, 0] = rng.random(n)
    cov[:, 2] = ids
    cov[0, 1] = np.nan                                         # the reference applies nan_to_num (dataset_def.py:85)
    return D.DeviceDataset(rng.integers(0, 256, (n, E_x)).astype(np.float64), rng.integers(0, 2, (n, Dv)), cov,
                           param_mask=rng.integers(0, 2, (n, E_x)), id_covariate=2, device="cpu")


def test_oracle_samplers_match_reference_goldens():
    g = h.load("samplers")
    np.random.seed(int(g["fixed_seed"]))
    assert orc.fixed_T_batches(int(g["fixed_P"]), int(g["fixed_T"]), int(g["fixed_batch"])) == _split(g["fixed_rows"], g["fixed_ptr"])
    np.random.seed(int(g["var_seed"]))
    assert orc.varying_T_batches(g["var_ids"].tolist(), int(g["var_batch"])) == _split(g["var_rows"], g["var_ptr"])


def test_loader_matches_reference_goldens():
    g = h.load("samplers")
    P, T = int(g["fixed_P"]), int(g["fixed_T"])
    ds = _dataset(np.repeat(np.arange(P), T))
    np.random.seed(int(g["fixed_seed"]))
    plan = D.DeviceSubjectLoader(ds, int(g["fixed_batch"]), varying_T=False, P=P, T=T).batches()
    assert [r.tolist() for r, _ in plan] == _split(g["fixed_rows"], g["fixed_ptr"])
    assert all(sum(lens) == len(r) for r, lens in plan)
    ds2 = _dataset(g["var_ids"])
    np.random.seed(int(g["var_seed"]))
    plan2 = D.DeviceSubjectLoader(ds2, int(g["var_batch"]), varying_T=True).batches()
    assert [r.tolist() for r, _ in plan2] == _split(g["var_rows"], g["var_ptr"])


def test_loader_vs_oracle_fresh_seeds_and_batch_contents():
    rng = np.random.default_rng(3)
    ids = np.repeat(rng.permutation(23) + 100, rng.integers(1, 9, 23))      # ragged subjects, arbitrary id values
    ds = _dataset(ids)
    assert ds.data.dtype == torch.uint8 and ds.mask.dtype == torch.uint8    # exact small integers travel as uint8
    assert not torch.isnan(ds.covariates).any()
    for seed, bs in ((1, 4), (2, 5), (3, 23), (4, 1)):
        np.random.seed(seed)
        ref = orc.varying_T_batches(ids.tolist(), bs)
        np.random.seed(seed)
        loader = D.DeviceSubjectLoader(ds, bs, varying_T=True)
        got = list(loader)
        assert len(got) == len(loader) == len(ref)
        for b, rows in zip(got, ref):
            assert b['idx'].tolist() == rows
            assert torch.equal(b['digit'], ds.data[rows]) and torch.equal(b['mask'], ds.mask[rows])
            assert torch.equal(b['label'], ds.covariates[rows]) and torch.equal(b['param_mask'], ds.param_mask[rows])
            lay = b['layout']
            assert lay.n_rows == len(rows) and lay.n_subj <= bs
            # every CSR segment is one subject
            sp = lay.subj_ptr.tolist()
            for a, c in zip(sp, sp[1:]):
                assert len(set(ids[rows[a:c]].tolist())) == 1
    # every row exactly once per epoch
    np.random.seed(9)
    seen = np.concatenate([r for r, _ in D.DeviceSubjectLoader(ds, 6).batches()])
    assert sorted(seen.tolist()) == list(range(len(ids)))


def test_fixed_T_batch_may_cut_a_subject_like_the_reference():
    P, T, bs = 5, 4, 6                                         # 6 rows per batch: subjects are cut (training.py:46-47)
    ds = _dataset(np.repeat(np.arange(P), T))
    np.random.seed(0)
    ref = orc.fixed_T_batches(P, T, bs)
    np.random.seed(0)
    plan = D.DeviceSubjectLoader(ds, bs, varying_T=False, P=P, T=T).batches()
    assert [r.tolist() for r, _ in plan] == ref
    for rows, lens in plan:
        k = 0
        for n in lens:                                          # each CSR segment stays inside one subject
            assert len({int(v) // T for v in rows[k:k + n]}) == 1
            k += n
