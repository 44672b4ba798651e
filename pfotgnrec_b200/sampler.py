"""Candidate sampling and mean-variance efficient selection on the device (K5).

`CandidateSampler` mirrors reference utils/utils.py:65-114 (RandEdgeSampler);
`MVSelector` is the entry point the reference lacks: its MV block is inline script code
(reference main.py:197-304), see INTEGRATION.md for the two-line change that calls it.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import ptr


def _dev(a, dtype, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), device=device).to(dtype).contiguous()


class CandidateSampler:
    """Uniform candidates from (universe \\ held items) per interaction, Philox keyed by
    (seed, interaction id, item position)."""

    def __init__(self, items_sorted, device="cuda"):
        self.device = torch.device(device)
        items = np.unique(np.asarray(items_sorted.cpu() if isinstance(items_sorted, torch.Tensor) else items_sorted))
        self.items = torch.as_tensor(items.astype(np.int32), device=self.device)

    def sample(self, event_ids, port_ptr, port_items, size, seed=0):
        dev = self.device
        ev = _dev(event_ids, torch.int64, dev)
        pp = _dev(port_ptr, torch.int64, dev)
        pi = _dev(port_items, torch.int32, dev)
        if pi.numel() == 0:
            pi = torch.zeros(1, dtype=torch.int32, device=dev)
        B = ev.shape[0]
        out = torch.empty(B, size, dtype=torch.int32, device=dev)
        _lib.call("pfo_sample_candidates", ptr(ev), ptr(pp), ptr(pi), ptr(self.items), self.items.shape[0],
                  B, int(size), int(seed), ptr(out))
        return out


class MVSelector:
    """p_pos / p_neg of the PfoTGNRec training step.

    logret  float64 [n_days, n_stocks, T] daily log-returns (log(p[1:]/p[:-1]) of the
            reference's `time_feature` price rows, computed once on the host);
    universe  0-based stock indices that occur as destinations in the training split."""

    def __init__(self, logret, universe_stocks, n_users, gamma=2.0, lam=0.5, n_candidates=20,
                 n_pos=1, n_neg=3, seed=0, device="cuda"):
        self.device = torch.device(device)
        self.logret = _dev(logret, torch.float64, self.device)
        self.n_days, self.n_stocks, self.T = self.logret.shape
        u = np.unique(np.asarray(universe_stocks)).astype(np.int32)
        self.universe = torch.as_tensor(u, device=self.device)
        self.n_users = int(n_users)
        self.gamma, self.lam = float(gamma), float(lam)
        self.K, self.n_pos, self.n_neg, self.seed = int(n_candidates), int(n_pos), int(n_neg), int(seed)
        if not 1 <= self.K <= 31:
            raise ValueError(f"num_negatives={self.K}: the MV kernel ranks the true item + candidates in one warp (K <= 31)")
        if self.n_pos + self.n_neg > self.K + 1:
            raise ValueError("p_pos_num + p_neg_num exceeds the number of ranked items (num_negatives + 1)")

    def select(self, event_ids, day_idx, dst_items, port_ptr, port_items, cand=None, return_scores=False):
        """dst_items: item ids (U+1..U+I) of the true destinations; portfolio CSR over 0-based stocks.
        Returns (p_pos int32[B*n_pos], p_neg int32[B*n_neg]) as ITEM ids, interaction-major."""
        dev = self.device
        ev = _dev(event_ids, torch.int64, dev)
        di = _dev(day_idx, torch.int32, dev)
        pos = _dev(dst_items, torch.int32, dev)          # item ids; the kernel subtracts / adds the item offset
        pp = _dev(port_ptr, torch.int64, dev)
        pi = _dev(port_items, torch.int32, dev)
        if pi.numel() == 0:
            pi = torch.zeros(1, dtype=torch.int32, device=dev)
        B, C = ev.shape[0], self.K + 1
        sample = cand is None
        cand_t = torch.empty(B, C, dtype=torch.int32, device=dev) if sample else _dev(cand, torch.int32, dev)
        y = torch.empty(B, C, dtype=torch.float64, device=dev) if return_scores else None
        p_pos = torch.empty(B * self.n_pos, dtype=torch.int32, device=dev)
        p_neg = torch.empty(B * self.n_neg, dtype=torch.int32, device=dev)
        _lib.call("pfo_mv_select", ptr(ev), ptr(di), ptr(pos), ptr(pp), ptr(pi), ptr(self.universe),
                  self.universe.shape[0], ptr(self.logret), self.n_stocks, self.T, B, self.K, self.gamma, self.lam,
                  self.n_pos, self.n_neg, self.seed, int(sample), ptr(cand_t), ptr(y), ptr(p_pos), ptr(p_neg),
                  self.n_users + 1)
        if return_scores:
            return p_pos, p_neg, cand_t, y
        return p_pos, p_neg
