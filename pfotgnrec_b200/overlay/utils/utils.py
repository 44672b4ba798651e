"""Drop-in for reference utils/utils.py: same names, B200-native internals.

`get_neighbor_finder` / `NeighborFinder` run the K1 CUDA kernel over a device CSR;
`RandEdgeSampler` draws from the shared Philox stream in the K5 sampling kernel
(documented deviation from numpy's MT19937).
"""
import numpy as np
import torch

from pfotgnrec_b200 import _lib
from pfotgnrec_b200.graph import NeighborFinder, TemporalCSR, get_neighbor_finder as _gnf
from pfotgnrec_b200.sampler import CandidateSampler
from pfotgnrec_b200.containers import MergeLayer  # noqa: F401  (reference utils/utils.py:4-17)


class MLP(torch.nn.Module):
    """Importable name only (reference utils/utils.py:19-35 is imported by main.py:7 and never called)."""

    def __init__(self, dim, drop=0.3):
        super().__init__()
        self.fc_1, self.fc_2, self.fc_3 = torch.nn.Linear(dim, 80), torch.nn.Linear(80, 10), torch.nn.Linear(10, 1)
        self.act, self.dropout = torch.nn.ReLU(), torch.nn.Dropout(p=drop, inplace=False)

    def forward(self, x):
        raise NotImplementedError("MLP is not on the PfoTGNRec hot path")


class EarlyStopMonitor(object):
    """Patience counter with the interface of reference utils/utils.py:38-62 (imported by main.py:7, never called):
    `early_stop_check(value)` is True once `max_round` consecutive values failed to improve on the best one by
    more than `tolerance` (relative)."""

    def __init__(self, max_round=3, higher_better=True, tolerance=1e-10):
        self.max_round, self.higher_better, self.tolerance = max_round, higher_better, tolerance
        self.num_round = self.epoch_count = self.best_epoch = 0
        self.last_best = None

    def early_stop_check(self, curr_val):
        val = curr_val if self.higher_better else -curr_val
        improved = self.last_best is not None and (val - self.last_best) / np.abs(self.last_best) > self.tolerance
        if self.last_best is None:
            self.last_best = val
        elif improved:
            self.last_best, self.num_round, self.best_epoch = val, 0, self.epoch_count
        else:
            self.num_round += 1
        self.epoch_count += 1
        return self.num_round >= self.max_round


_UNIVERSE_CACHE = {}


def _universe_sampler(dst_list, device):
    """unique(dst_list) on the device, cached across constructions.  main.py:193-194 / :344-347 / evaluation.py:84-88
    build a new RandEdgeSampler for EVERY batch from the same `train_data.destinations` / `full_data.destinations`
    array, and the reference pays an O(E log E) np.unique each time (utils/utils.py:73).  The cache key is the array's
    identity (buffer address, length, dtype) guarded by a strided content fingerprint, so a different or modified array
    is never served a stale universe."""
    a = np.asarray(dst_list)
    step = max(1, a.shape[0] // 64) if a.ndim == 1 and a.shape[0] else 1
    key = (a.__array_interface__["data"][0], a.shape, str(a.dtype), str(device),
           a[::step].tobytes() if a.ndim == 1 else a.tobytes())
    hit = _UNIVERSE_CACHE.get(key)
    if hit is None:
        if len(_UNIVERSE_CACHE) > 8:
            _UNIVERSE_CACHE.clear()
        hit = _UNIVERSE_CACHE[key] = CandidateSampler(np.unique(a), device=device)
    return hit


class RandEdgeSampler(object):
    """Reference utils/utils.py:65-114.  `sample(size)` -> int64 [B, size] item ids drawn
    uniformly from unique(dst_list) minus each interaction's portfolio; without replacement
    when enough items are available.  Draws are keyed by (seed, interaction, item) in the
    Philox stream, so they do not depend on batch size or call order."""

    _event_counter = 0      # stands in for the global interaction index when the caller has none

    def __init__(self, src_list, dst_list, portfolio_list, upper_u, map_item_id, seed=None,
                 event_ids=None, device="cuda"):
        self.src_list = src_list
        self.seed = seed
        B = len(src_list)
        held = [[map_item_id[item] + upper_u + 1 for item in sub if item] for sub in portfolio_list]
        ptr_ = np.zeros(B + 1, dtype=np.int64)
        np.cumsum([len(h) for h in held], out=ptr_[1:])
        items = np.array([x for h in held for x in h], dtype=np.int32)
        if event_ids is None:
            if seed is not None:     # evaluation recreates RandomState(seed) per batch (evaluation.py:88)
                event_ids = np.arange(B, dtype=np.int64)
            else:
                event_ids = np.arange(B, dtype=np.int64) + RandEdgeSampler._event_counter
                RandEdgeSampler._event_counter += B
        self._sampler = _universe_sampler(dst_list, device)
        self._args = (np.asarray(event_ids, dtype=np.int64), ptr_, items)

    def sample(self, size):
        ev, ptr_, items = self._args
        out = self._sampler.sample(ev, ptr_, items, int(size), seed=0 if self.seed is None else int(self.seed))
        return out.cpu().numpy().astype(np.int64)


def get_neighbor_finder(data, uniform, max_node_idx=None):
    """Reference utils/utils.py:117-127."""
    return _gnf(data, uniform, max_node_idx=max_node_idx, device="cuda")
