"""Device-resident data path (second part).

The reference feeds every minibatch through a pandas-backed `Dataset` (`dataset_def.py:57-92`: four `.iloc` row
reads per sample, 23 % of a CPU training step) and a `HensmanDataLoader` over subject samplers
(`utils.py:24-97`, `training.py:38-47`).  Here the whole data set lives in HBM once - data as uint8 when its
values are exact small integers (pixel values, one-hot / thermometer codes), the mask as uint8, covariates
float64 - and a loader yields minibatches of WHOLE subjects by gathering rows on the device:

  * same batch composition as the reference samplers (`SubjectSampler` + `BatchSampler`, `training.py:46-47`, and
    `VaryingLengthSubjectSampler` + `VaryingLengthBatchSampler`, `:41-45`), drawing the subject permutation from
    `np.random.shuffle` exactly like `utils.py:45,69`, so a seeded run visits the same rows in the same order;
  * same batch dict keys (`digit`, `label`, `idx`, `mask`, `param_mask`), values already on the device;
  * plus `layout`, the subject CSR the KL kernels take (built from the known subject lengths: no device sync, no
    `torch.unique` on the id column).
Only host-side index arithmetic and torch gathers: nothing here needs the CUDA library.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from .subjects import SubjectLayout


def _exact_uint8(t):
    t = torch.as_tensor(t)
    if t.dtype == torch.uint8:
        return True
    tf = t.to(torch.float64)
    return bool(((tf >= 0) & (tf <= 255) & (tf == tf.round())).all())


class DeviceDataset:
    """Rows of a heterogeneous longitudinal data set as device tensors.

    data [N, E_x], mask [N, D], covariates [N, Q] (the reference's `label`, after its column reorder and
    nan_to_num, dataset_def.py:47,83-85), optional param_mask [N, P_theta]."""

    def __init__(self, data, mask, covariates, param_mask=None, id_covariate=2, device="cuda", compress=True):
        data, mask = torch.as_tensor(data), torch.as_tensor(mask)
        self.device = torch.device(device)
        if compress and _exact_uint8(data):
            data = data.to(torch.uint8)
        self.data = data.to(self.device).contiguous()
        self.mask = mask.to(torch.uint8).to(self.device).contiguous()
        self.covariates = torch.nan_to_num(torch.as_tensor(covariates, dtype=torch.float64)).to(self.device).contiguous()
        self.param_mask = None if param_mask is None else torch.as_tensor(param_mask).to(torch.uint8).to(self.device).contiguous()
        self.id_covariate = int(id_covariate)
        ids = self.covariates[:, self.id_covariate].cpu().numpy().astype(np.int64).tolist()
        # first occurrence of every subject id, in order of appearance (utils.py:63-65)
        firsts = OrderedDict()
        for i, v in enumerate(ids):
            firsts.setdefault(v, i)
        self.start_indices = list(firsts.values())
        self.end_indices = self.start_indices[1:] + [len(ids)]
        self.P = len(self.start_indices)

    @classmethod
    def from_reference_dataset(cls, ds, id_covariate=2, device="cuda"):
        """From an instance of the reference's dataset classes (dataset_def.py:13-92): reads its pandas sources once."""
        label = np.nan_to_num(ds.label_source.to_numpy(dtype=np.float64))
        return cls(ds.data_source.to_numpy(), ds.mask_source.to_numpy(), label,
                   param_mask=ds.param_mask_source.to_numpy(), id_covariate=id_covariate, device=device)

    def __len__(self):
        return self.data.shape[0]

    def rows(self, idx):
        """Batch dict for row indices `idx` (host list / array or device tensor), keys as dataset_def.py:91."""
        host = idx if isinstance(idx, torch.Tensor) else torch.as_tensor(np.asarray(idx, dtype=np.int64))
        dev_idx = host.to(self.device, non_blocking=True)
        out = {'digit': self.data.index_select(0, dev_idx), 'label': self.covariates.index_select(0, dev_idx),
               'idx': dev_idx, 'mask': self.mask.index_select(0, dev_idx)}
        if self.param_mask is not None:
            out['param_mask'] = self.param_mask.index_select(0, dev_idx)
        return out


class DeviceSubjectLoader:
    """Minibatches of whole subjects, composed like the reference's loaders.

    varying_T=False: `BatchSampler(SubjectSampler(dataset, P, T), batch_size, drop_last=False)` (training.py:46-47):
      subjects own rows [T s, T (s+1)), `batch_size` ROWS per batch.
    varying_T=True: `VaryingLengthBatchSampler(VaryingLengthSubjectSampler(dataset, id_covariate), subjects_per_batch)`
      (training.py:41-45): subjects own the rows from their first occurrence to the next subject's,
      `batch_size` SUBJECTS per batch.
    One pass = one epoch; the subject permutation comes from `np.random.shuffle` as in utils.py:45,69."""

    def __init__(self, dataset: DeviceDataset, batch_size, varying_T=True, P=None, T=None, shuffle=True):
        self.ds, self.batch_size, self.varying_T, self.shuffle = dataset, int(batch_size), bool(varying_T), bool(shuffle)
        if self.varying_T:
            self.starts, self.ends = dataset.start_indices, dataset.end_indices
        else:
            if P is None or T is None:
                raise ValueError("fixed-T sampling needs P and T (utils.py:37-41)")
            # batch_size counts ROWS here; when it is not a multiple of T a batch boundary cuts a subject (the
            # reference's BatchSampler does the same and its minibatch_KLD_upper_bound then fails in
            # reshape([P_batch, T, Q]), elbo_functions.py:144): such batches carry no subject layout
            self.whole_subjects = self.batch_size % int(T) == 0
            self.starts = [T * s for s in range(P)]
            self.ends = [T * (s + 1) for s in range(P)]
        self.P = len(self.starts)

    def batches(self):
        """Host-side plan of one epoch: list of (row indices, subject lengths in batch order)."""
        r = np.arange(self.P)
        if self.shuffle:
            np.random.shuffle(r)
        plan = []
        if self.varying_T:
            for b in range(0, self.P, self.batch_size):
                subj = r[b:b + self.batch_size]
                lens = [self.ends[s] - self.starts[s] for s in subj]
                rows = np.concatenate([np.arange(self.starts[s], self.ends[s]) for s in subj]) if len(subj) else np.zeros(0, np.int64)
                plan.append((rows, lens))
        else:
            order = np.concatenate([np.arange(self.starts[s], self.ends[s]) for s in r]) if self.P else np.zeros(0, np.int64)
            T = self.ends[0] - self.starts[0] if self.P else 1
            for b in range(0, len(order), self.batch_size):
                rows = order[b:b + self.batch_size]
                # a batch boundary may cut a subject when batch_size is not a multiple of T (as in the reference)
                lens, k = [], 0
                while k < len(rows):
                    n = min(T - (rows[k] % T), len(rows) - k)
                    lens.append(int(n))
                    k += n
                plan.append((rows, lens))
        return plan

    def __len__(self):
        if self.varying_T:
            return (self.P + self.batch_size - 1) // self.batch_size
        n = sum(e - s for s, e in zip(self.starts, self.ends))
        return (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for rows, lens in self.batches():
            batch = self.ds.rows(rows)
            # fragments of a cut subject must not pass as subjects of their own: no layout -> the fixed-T KL bound
            # builds SubjectLayout.fixed(N, T), which raises when N is not a whole number of subjects
            whole = self.varying_T or self.whole_subjects
            batch['layout'] = SubjectLayout.from_lengths(lens, self.ds.device) if whole else None
            yield batch
