"""Subject structure of a minibatch in CSR form (what the streaming kernels consume).

The reference finds subjects with `torch.unique(train_xt[:, id_covariate])` plus one boolean
mask per subject (elbo_functions.py:242-244) or assumes T contiguous rows per subject
(elbo_functions.py:144,159).  Here that becomes three int32 device arrays:
row_idx (rows grouped by subject), subj_ptr (row offsets) and tt_ptr (offsets of the T_s x T_s
blocks)."""
from __future__ import annotations

import torch

from . import _lib


class SubjectLayout:
    def __init__(self, row_idx, subj_ptr, tt_ptr, n_subj, n_rows, t_max, tt_total, lens=None):
        self.row_idx, self.subj_ptr, self.tt_ptr = row_idx, subj_ptr, tt_ptr
        self.n_subj, self.n_rows, self.t_max, self.tt_total = int(n_subj), int(n_rows), int(t_max), int(tt_total)
        self._lens = lens               # host copy of the subject lengths (int64 tensor) for panel planning
        self._panels = {}

    def panels(self, rows_per_panel):
        """Number of row panels hlvae_kl_panel needs when it packs whole subjects, in order, into panels of
        `rows_per_panel` rows (and at most 16 subjects) - host arithmetic on the lengths, cached."""
        rp = int(rows_per_panel)
        if rp not in self._panels:
            if self._lens is None or self.n_subj == 0:
                per = max(1, min(16, rp // max(self.t_max, 1)))
                self._panels[rp] = (self.n_subj + per - 1) // per
            else:
                lens = self._lens.tolist()
                if len(set(lens)) == 1:
                    per = max(1, min(16, rp // max(lens[0], 1)))
                    self._panels[rp] = (len(lens) + per - 1) // per
                else:
                    n, rows, ns = 0, 0, 0
                    for t in lens:
                        if ns > 0 and (rows + t > rp or ns == 16):
                            n, rows, ns = n + 1, 0, 0
                        rows, ns = rows + t, ns + 1
                    self._panels[rp] = n + (1 if ns else 0)
        return self._panels[rp]

    @staticmethod
    def _finish(row_idx, lengths_host, device):
        lens = torch.as_tensor(lengths_host, dtype=torch.int64)
        n_subj = int(lens.numel())
        t_max = int(lens.max()) if n_subj else 0
        if t_max > _lib.TMAX:
            raise RuntimeError(f"hlvae_b200: a subject has {t_max} rows; the per-subject kernels support at most "
                               f"{_lib.TMAX} (HLVAE_TMAX)")
        sp = torch.zeros(n_subj + 1, dtype=torch.int64)
        tp = torch.zeros(n_subj + 1, dtype=torch.int64)
        if n_subj:
            sp[1:] = torch.cumsum(lens, 0)
            tp[1:] = torch.cumsum(lens * lens, 0)
        return SubjectLayout(row_idx.to(device=device, dtype=torch.int32).contiguous(),
                             sp.to(device=device, dtype=torch.int32), tp.to(device=device, dtype=torch.int32),
                             n_subj, int(sp[-1]), t_max, int(tp[-1]), lens=lens)

    @staticmethod
    def fixed(n_rows, T, device):
        """Rows are subject-contiguous, T per subject (elbo_functions.py:144)."""
        if T <= 0 or n_rows % T != 0:
            # the reference's reshape([P_batch, T, Q]) (elbo_functions.py:144) raises as well; dropping the trailing
            # rows silently would bias the loss
            raise ValueError(f"hlvae_b200: {n_rows} rows are not a whole number of subjects of T = {T} rows")
        n_subj = n_rows // T
        return SubjectLayout._finish(torch.arange(n_subj * T), [T] * n_subj, device)

    @staticmethod
    def from_lengths(lengths, device):
        """Subject-contiguous rows with known per-subject lengths (no device sync)."""
        n = int(sum(int(v) for v in lengths))
        return SubjectLayout._finish(torch.arange(n), [int(v) for v in lengths], device)

    @staticmethod
    def from_ids(ids):
        """Group rows by id value, sorted unique ids as torch.unique gives them
        (elbo_functions.py:242-244).  Needs one device->host copy of the counts, as the
        reference's own `.tolist()` does."""
        order = torch.argsort(ids, stable=True)
        _, counts = torch.unique_consecutive(ids[order], return_counts=True)
        return SubjectLayout._finish(order, counts.tolist(), ids.device)

    def shard(self, rank, world):
        """Subjects [rank::world]-style contiguous split for data-parallel runs: returns the layout
        holding only this rank's subjects (row indices still refer to the full batch)."""
        per = (self.n_subj + world - 1) // world
        lo, hi = min(self.n_subj, rank * per), min(self.n_subj, (rank + 1) * per)
        sp = self.subj_ptr.cpu().to(torch.int64)
        lens = (sp[lo + 1:hi + 1] - sp[lo:hi]).tolist()
        rows = self.row_idx[int(sp[lo]):int(sp[hi])]
        return SubjectLayout._finish(rows, lens, self.row_idx.device)
