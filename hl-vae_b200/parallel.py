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


def shard_dataset_rows(subject_ids, rank, world):
    """Row indices of a data set (host array of per-row subject ids, rows of a subject adjacent as
    dataset_def.py stores them) that belong to `rank`: a contiguous block of WHOLE subjects in order of first
    appearance.  `training.hensman_training` builds its DataLoader and P from the dataset object it is handed
    (training.py:37-47), so rank awareness has to come from handing each rank a dataset restricted to these rows
    (SURVEY.md Appendix D)."""
    import numpy as np
    ids = np.asarray(subject_ids)
    _, first = np.unique(ids, return_index=True)
    order = ids[np.sort(first)]                               # subjects in order of first appearance
    lo, hi = shard_subjects(len(order), rank, world)
    return np.nonzero(np.isin(ids, order[lo:hi]))[0]


def sync_gradients(params, group=None):
    """SUM the gradients of `params` over the ranks in one all-reduce of a flat bucket (missing gradients count as
    zero).  With subject-sharded ranks the reference's loss nll * P / P_b + kld (training.py:121-124) splits as
    follows: every rank's nll is a PARTIAL sum of the global one (callers pass the GLOBAL P_in_current_batch), and
    mu / log_v gradients of the KL bound are local to the rank's rows - so gradients of the NN weights (and of the
    likelihood log-variances) must be SUMMED, not averaged.  Gradients of the replicated GP parameters (Z, kernel
    hyper-parameters) are already global on every rank - they are built from the all-reduced accumulators - and
    must NOT be passed here."""
    g = group if group is not None else config.process_group
    if g is None:
        return
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float64)
                      for p in params])
    dist.all_reduce(flat, group=g)
    o = 0
    for p in params:
        n = p.numel()
        gsum = flat[o:o + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = gsum.clone()
        else:
            p.grad.copy_(gsum)
        o += n


def attach_gradient_sync(optimizer, local_sum_params, group=None):
    """Register `sync_gradients(local_sum_params)` as a step pre-hook of `optimizer` (the reference builds the
    optimiser outside the training loop, HLVAE_main.py:278, and `training.py` reaches the model by attribute, so a
    DistributedDataParallel wrapper would break the unchanged caller - SURVEY.md Appendix D).  Returns the hook
    handle."""
    params = list(local_sum_params)
    return optimizer.register_step_pre_hook(lambda opt, args, kwargs: sync_gradients(params, group))


def allreduce_sum_(t, group=None):
    g = group if group is not None else config.process_group
    if g is not None:
        dist.all_reduce(t, group=g)
    return t
