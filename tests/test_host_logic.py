"""CPU: host-side logic and the C-ABI surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import helpers as h
from hlvae_b200 import _lib, kernels, likelihoods, loglik, subjects, synth
from oracle import hlvae_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "hlvae_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t)\s+(hlvae_\w+)\s*\(", hdr, flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTED), declared ^ set(_lib.EXPORTED)
    for name in declared:
        assert getattr(built_lib, name) is not None
    assert built_lib.hlvae_version() == 1
    assert built_lib.hlvae_sizeof_kspec() == ctypes.sizeof(_lib.KSpec)


def test_binding_signatures_match_header():
    """Every prototype of include/hlvae_b200.h against the ctypes signature in _lib._SIGS: same number of parameters
    and the same kind per position (pointer / int / int64_t / double) - a parameter added on one side only would
    shift every later argument silently."""
    hdr = open(os.path.join(ROOT, "include", "hlvae_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"^(?:int|int64_t)\s+(hlvae_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.M | re.S)
    assert len(protos) == len(_lib.EXPORTED)

    def kind(param):
        param = " ".join(param.split())
        if "*" in param:
            return "ptr"
        if param.startswith("int64_t"):
            return "i64"
        if param.startswith("double"):
            return "f64"
        if param.startswith("int"):
            return "int"
        raise AssertionError(f"unparsed parameter '{param}'")

    ckind = {_lib._P: "ptr", _lib._I: "int", _lib._L: "i64", _lib._D: "f64"}
    for name, params in protos:
        params = [] if params.strip() in ("", "void") else [q for q in params.split(",")]
        args, _ = _lib._SIGS[name]
        got = [ckind.get(a, "ptr") for a in args]                      # POINTER(...) types are pointers
        assert [kind(q) for q in params] == got, name


def test_acc_layout(built_lib):
    off = _lib.acc_layout(4, 12, 6)
    assert off["S"] == 0 and off["p"] == 4 * 12 * 12 and off["gw"] == off["p"] + 48
    assert off["total"] == off["gls1"] + 8 * 4


@pytest.mark.parametrize("kargs", [synth.DEFAULT_KERNEL_ARGS, synth.SWEEP_KERNEL_ARGS, synth.MASKED_KERNEL_ARGS])
def test_spec_matches_oracle_compile(kargs):
    k0, k1 = kernels.generate_kernel_batched(3, **kargs)
    o0, o1 = orc.compile_spec(**kargs)
    for km, os_ in ((k0, o0), (k1, o1)):
        fs = kernels.compile_spec(km)
        assert fs.ncomp == len(os_.comps)
        for r, comp in enumerate(os_.comps):
            se = [f.col for f in comp.factors if f.kind == orc.SE]
            disc = [(_lib.KIND_CAT if f.kind == orc.CAT else _lib.KIND_BIN, f.col) for f in comp.factors if f.kind != orc.SE]
            c = fs.cspec.comp[r]
            assert c.se_col == (se[0] if se else -1)
            assert [(c.disc_kind[i], c.disc_col[i]) for i in range(c.ndisc)] == disc


def test_parameter_names_and_defaults():
    """gpytorch-compatible names (checkpoint row) and initial values
    (kernel_spec.py:68: lengthscale 2.5 set in float32, raw_outputscale 0)."""
    k0, k1 = kernels.generate_kernel_batched(4, **synth.DEFAULT_KERNEL_ARGS)
    names = [n for n, _ in k0.named_parameters()]
    assert names == ['kernels.0.raw_outputscale', 'kernels.0.base_kernel.raw_lengthscale', 'kernels.1.raw_outputscale',
                     'kernels.1.base_kernel.kernels.1.raw_lengthscale', 'kernels.2.raw_outputscale',
                     'kernels.2.base_kernel.kernels.1.raw_lengthscale']
    assert [n for n, _ in k1.named_parameters()] == ['kernels.0.raw_outputscale', 'kernels.1.raw_outputscale',
                                                    'kernels.1.base_kernel.kernels.1.raw_lengthscale']
    k0.double()
    p = dict(k0.named_parameters())
    assert p['kernels.0.raw_outputscale'].shape == (4,) and p['kernels.0.base_kernel.raw_lengthscale'].shape == (4, 1, 1)
    raw = float(np.float32(orc.softplus_inv(2.5)))
    assert abs(float(p['kernels.0.base_kernel.raw_lengthscale'][0, 0, 0]) - raw) < 1e-12
    both = k0 + k1
    assert len(both.kernels) == 5 and any(k.startswith("kernels.4.") for k in both.state_dict())
    sd = both.state_dict()
    both.load_state_dict(sd)
    lik = likelihoods.GaussianLikelihood(batch_shape=torch.Size([4]), noise_constraint=likelihoods.GreaterThan(1e-8))
    lik.noise = 1
    assert lik.noise_covar.noise.shape == (4, 1) and abs(float(lik.noise[0]) - 1.0) < 1e-6


def test_subject_layouts():
    lay = subjects.SubjectLayout.fixed(12, 4, "cpu")
    assert lay.n_subj == 3 and lay.subj_ptr.tolist() == [0, 4, 8, 12] and lay.tt_ptr.tolist() == [0, 16, 32, 48]
    ids = torch.tensor([5., 2., 5., 9., 2., 2.])
    lay = subjects.SubjectLayout.from_ids(ids)
    assert lay.subj_ptr.tolist() == [0, 3, 5, 6] and lay.row_idx.tolist() == [1, 4, 5, 0, 2, 3]
    assert lay.t_max == 3 and lay.tt_total == 9 + 4 + 1
    ref = orc.split_subjects_by_id(torch.stack([ids, ids], 1), 0)
    assert [r.tolist() for r in ref] == [[1, 4, 5], [0, 2], [3]]
    lay2 = subjects.SubjectLayout.from_lengths([2, 3, 1], "cpu")
    assert lay2.subj_ptr.tolist() == [0, 2, 5, 6]
    a, b = lay2.shard(0, 2), lay2.shard(1, 2)
    assert a.n_subj == 2 and b.n_subj == 1 and a.row_idx.tolist() == [0, 1, 2, 3, 4] and b.row_idx.tolist() == [5]
    empty = subjects.SubjectLayout.from_lengths([], "cpu")
    assert empty.n_subj == 0 and empty.tt_total == 0
    assert subjects.SubjectLayout.from_lengths([33, 64], "cpu").t_max == 64      # HLVAE_TMAX = 64
    with pytest.raises(RuntimeError):
        subjects.SubjectLayout.from_lengths([65], "cpu")


def test_var_layout_matches_oracle():
    rng = np.random.default_rng(0)
    types = synth.mixed_types(rng, 17)
    lay = loglik.VarLayout(types, "cpu")
    descs, E_x, P_th = orc.build_layout(types)
    assert (lay.D, lay.E_x, lay.P_theta) == (len(descs), E_x, P_th)
    assert lay.var_dcol.tolist() == [d.data_col for d in descs]
    assert lay.var_pcol.tolist() == [d.theta_col for d in descs]
    ti = orc.types_info_from_layout(types)
    lay2 = loglik.VarLayout.from_types_info(ti, "cpu")
    assert lay2.var_kind.tolist() == lay.var_kind.tolist() and lay2.var_nclass.tolist() == lay.var_nclass.tolist()


def test_cpu_tensors_fail_loudly():
    k0, k1 = kernels.generate_kernel_batched(2, **synth.DEFAULT_KERNEL_ARGS)
    x = torch.zeros(3, 6, dtype=torch.float64)
    with pytest.raises(RuntimeError, match="CUDA"):
        k0(x, x).evaluate()
    lay = loglik.VarLayout([("real", 1)], "cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        loglik.fused_loglik(lay, torch.zeros(2, 1), torch.ones(2, 1), torch.zeros(2, 1), torch.zeros(4, 1))


def test_panel_chunks_are_whole_panels():
    """hlvae_kl_panel chunk sizing (elbo._subjects_per_chunk): whole row panels per chunk, every subject covered, at
    least one chunk, and a bounded number of panels per CTA for large batches."""
    from hlvae_b200 import elbo
    for n_subj, t_max, L, M in ((1, 20, 32, 64), (7, 20, 32, 64), (200, 20, 32, 64), (800, 20, 32, 64), (3200, 20, 32, 64),
                                (800, 5, 8, 32), (800, 32, 32, 128), (513, 13, 4, 64)):
        spc = elbo._subjects_per_chunk(n_subj, t_max, L, M)
        rp = 64
        spp = max(1, rp // t_max)
        assert spc >= 1 and spc % spp == 0
        n_chunks = (n_subj + spc - 1) // spc
        assert n_chunks >= 1 and n_chunks * spc >= n_subj
        assert spc // spp <= max(30, -(-n_subj // spp) // max(1, (148 * elbo.PANEL_WAVES) // L) + 1) + 1
