"""Scoring, ranking and the metric block of the evaluation loop, on the device.

Replaces reference evaluation.py:107-115 (scores), :134-138 (ranking) and :127-258 (Recall/NDCG@{1,3,5},
delta-return / delta-Sharpe@{1,3,5} of the held portfolio in and out of sample, means and fraction-positive),
which in the reference copy every score to the host and loop over interactions in Python.  Two C-ABI calls per
batch (`pfo_eval_score`, `pfo_eval_metrics`); 31 running sums stay on the device until `summary`.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import ptr

KS = (1, 3, 5)


class EvalMetricBlock:
    """logret_past / logret_future: float64 [n_days, n_stocks, T] daily log-returns log(p[1:]/p[:-1]) of the
    reference's `time_feature_past` / `time_feature_future` price rows (computed once on the host with numpy, like
    the reference's np.log); item_offset: item id of stock 0 (upper_u + 1)."""

    def __init__(self, logret_past, logret_future, item_offset, device="cuda"):
        self.device = torch.device(device)
        f64 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=self.device).contiguous()
        self.lr_past, self.lr_future = f64(logret_past), f64(logret_future)
        assert self.lr_past.shape == self.lr_future.shape and self.lr_past.dim() == 3
        self.n_stocks, self.T = int(self.lr_past.shape[1]), int(self.lr_past.shape[2])
        self.item_offset = int(item_offset)
        self.acc = torch.zeros(31, dtype=torch.float64, device=self.device)

    def reset(self):
        self.acc.zero_()

    def step(self, e_src, e_dst, e_cand, dst_items, cand_items, day_idx, port_ptr, port_items):
        """e_src/e_dst [B,d], e_cand [B*N,d] fp32; dst_items int32[B] and cand_items int32[B,N] item ids; day_idx
        int32[B] rows of the price tables; portfolio CSR (int64[B+1], int32[nnz]) over 0-based stocks.  Returns
        (pos_rank int32[B], top5 int32[B,5], scores fp32[B,1+N], per_event float64[B,18]) and adds to `acc`."""
        dev = self.device
        B, d = e_src.shape
        N = e_cand.shape[0] // B
        e_src, e_dst, e_cand = e_src.contiguous(), e_dst.contiguous(), e_cand.contiguous()
        scores = torch.empty(B, 1 + N, device=dev)
        pos_rank = torch.empty(B, dtype=torch.int32, device=dev)
        top = torch.empty(B, 5, dtype=torch.int32, device=dev)
        _lib.call("pfo_eval_score", ptr(e_src), ptr(e_dst), ptr(e_cand), B, N, d, 5, ptr(scores), ptr(pos_rank), ptr(top))
        per_event = torch.empty(B, 18, dtype=torch.float64, device=dev)
        _lib.call("pfo_eval_metrics", ptr(pos_rank), ptr(top), 5, ptr(dst_items), ptr(cand_items), N, self.item_offset,
                  ptr(day_idx), ptr(port_ptr), ptr(port_items), ptr(self.lr_past), ptr(self.lr_future),
                  self.n_stocks, self.T, B, ptr(per_event), ptr(self.acc))
        return pos_rank, top, scores, per_event

    def summary(self, EVAL="val", reduce=None):
        """The dictionary reference eval_recommendation returns (evaluation.py:209-258) from the running sums: one
        248-byte device->host copy.  `reduce` (optional) sums the 31 doubles over ranks first."""
        acc = self.acc.clone()
        if reduce is not None:
            acc = reduce(acc)
        acc = acc.cpu().numpy()
        n = max(acc[30], 1.0)
        out = {}
        for j, k in enumerate(KS):
            out[f"{EVAL}_recall_avg_{k}"] = acc[j] / n
            out[f"{EVAL}_ndcg_avg_{k}"] = acc[3 + j] / n
            for s, suf in enumerate(("", "_")):
                out[f"{EVAL}_return_avg_{k}{suf}"] = acc[6 + 6 * s + j] / n
                out[f"{EVAL}_return_percent_{k}{suf}"] = acc[18 + 6 * s + j] / n
                out[f"{EVAL}_sharpe_avg_{k}{suf}"] = acc[9 + 6 * s + j] / n
                out[f"{EVAL}_sharpe_percent_{k}{suf}"] = acc[21 + 6 * s + j] / n
        return out
