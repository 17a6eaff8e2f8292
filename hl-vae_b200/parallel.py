"""Data-parallel sharding of the ELBO path: one process per GPU, subjects partitioned across
ranks, ONE all-reduce per step of the per-latent accumulator buffer (S, p, scalars and the
gradients of replicated parameters) - SURVEY.md section 8(e).  The reference has no parallelism
at all; this is new surface.

After `enable()`, `hlvae_b200.elbo.minibatch_KLD_upper_bound*` all-reduce their accumulators
over `config.process_group`; every rank then runs the identical replicated M x M stage, so
kld_total, grad_m, grad_H and the gradients of Z and the kernel hyper-parameters are the same
on every rank, while mu / log_v gradients stay local to the rank's rows.  Callers pass the
GLOBAL subject count as P_in_current_batch / P_batch."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import config


def enable(group=None):
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    config.process_group = group if group is not None else dist.group.WORLD
    return config.process_group


def disable():
    config.process_group = None


def shard_subjects(n_subj, rank, world):
    """Contiguous block of subjects owned by `rank` (whole subjects only: the rows of a subject
    are coupled through B_s, elbo_functions.py:150-161, 244-254)."""
    per = (n_subj + world - 1) // world
    return min(n_subj, rank * per), min(n_subj, (rank + 1) * per)


def allreduce_sum_(t, group=None):
    g = group if group is not None else config.process_group
    if g is not None:
        dist.all_reduce(t, group=g)
    return t
