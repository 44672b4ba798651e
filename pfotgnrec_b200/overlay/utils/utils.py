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


class MergeLayer(torch.nn.Module):
    """Parameter container of fc2(relu(fc1([x1 | x2]))) (reference utils/utils.py:4-17).
    The arithmetic runs inside the fused attention path (pfo_linear_*); same init order."""

    def __init__(self, dim1, dim2, dim3, dim4):
        super().__init__()
        self.fc1 = torch.nn.Linear(dim1 + dim2, dim3)
        self.fc2 = torch.nn.Linear(dim3, dim4)
        self.act = torch.nn.ReLU()
        torch.nn.init.xavier_normal_(self.fc1.weight)
        torch.nn.init.xavier_normal_(self.fc2.weight)

    def forward(self, x1, x2):
        raise NotImplementedError("MergeLayer is evaluated inside TGN.compute_temporal_embeddings*")


class MLP(torch.nn.Module):
    """Importable name only (reference utils/utils.py:19-35; never called by main.py)."""

    def __init__(self, dim, drop=0.3):
        super().__init__()
        self.fc_1 = torch.nn.Linear(dim, 80)
        self.fc_2 = torch.nn.Linear(80, 10)
        self.fc_3 = torch.nn.Linear(10, 1)
        self.act = torch.nn.ReLU()
        self.dropout = torch.nn.Dropout(p=drop, inplace=False)

    def forward(self, x):
        raise NotImplementedError("MLP is not on the PfoTGNRec hot path")


class EarlyStopMonitor(object):
    """Reference utils/utils.py:38-62 (host-side bookkeeping)."""

    def __init__(self, max_round=3, higher_better=True, tolerance=1e-10):
        self.max_round, self.num_round = max_round, 0
        self.epoch_count, self.best_epoch = 0, 0
        self.last_best, self.higher_better, self.tolerance = None, higher_better, tolerance

    def early_stop_check(self, curr_val):
        if not self.higher_better:
            curr_val *= -1
        if self.last_best is None:
            self.last_best = curr_val
        elif (curr_val - self.last_best) / np.abs(self.last_best) > self.tolerance:
            self.last_best, self.num_round, self.best_epoch = curr_val, 0, self.epoch_count
        else:
            self.num_round += 1
        self.epoch_count += 1
        return self.num_round >= self.max_round


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
        self._sampler = CandidateSampler(np.unique(dst_list), device=device)
        self._args = (np.asarray(event_ids, dtype=np.int64), ptr_, items)

    def sample(self, size):
        ev, ptr_, items = self._args
        out = self._sampler.sample(ev, ptr_, items, int(size), seed=0 if self.seed is None else int(self.seed))
        return out.cpu().numpy().astype(np.int64)


def get_neighbor_finder(data, uniform, max_node_idx=None):
    """Reference utils/utils.py:117-127."""
    return _gnf(data, uniform, max_node_idx=max_node_idx, device="cuda")
