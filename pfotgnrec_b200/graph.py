"""Temporal adjacency as a device-resident, time-sorted CSR and the neighbour finder on top.

Mirrors the interface of reference utils/utils.py:117-220 (`get_neighbor_finder`,
`NeighborFinder.get_temporal_neighbor / find_before`) but the lookup runs in the K1 CUDA
kernel (csrc/graph_kernels.cu) instead of a Python loop over queries.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


class TemporalCSR:
    """rowptr int64[N+1], nbr int32[2E], eidx int32[2E], ts float64[2E], sorted by
    (node, timestamp, stream order) -- the stable per-node sort of utils/utils.py:139."""

    def __init__(self, sources, destinations, edge_idxs, timestamps, n_nodes=None, device="cuda"):
        src = np.asarray(sources, dtype=np.int64)
        dst = np.asarray(destinations, dtype=np.int64)
        eid = np.asarray(edge_idxs, dtype=np.int64)
        ts = np.asarray(timestamps, dtype=np.float64)
        E = src.shape[0]
        if n_nodes is None:
            n_nodes = int(max(src.max(), dst.max())) + 1 if E else 1
        self.n_nodes = int(n_nodes)
        self.n_events = int(E)
        self.device = torch.device(device)
        if self.device.type == "cuda" and E > 0:
            # device build: one stable sort by timestamp, one stable sort by node
            s = torch.as_tensor(src, device=self.device)
            d = torch.as_tensor(dst, device=self.device)
            node = torch.stack([s, d], dim=1).reshape(-1)
            other = torch.stack([d, s], dim=1).reshape(-1)
            e2 = torch.as_tensor(eid, device=self.device).repeat_interleave(2)
            t2 = torch.as_tensor(ts, device=self.device).repeat_interleave(2)
            o1 = torch.sort(t2, stable=True).indices
            o2 = torch.sort(node[o1], stable=True).indices
            order = o1[o2]
            self.nbr = other[order].to(torch.int32).contiguous()
            self.eidx = e2[order].to(torch.int32).contiguous()
            self.ts = t2[order].contiguous()
            counts = torch.bincount(node, minlength=self.n_nodes)
            self.rowptr = torch.zeros(self.n_nodes + 1, dtype=torch.int64, device=self.device)
            torch.cumsum(counts, 0, out=self.rowptr[1:])
        else:
            node = np.stack([src, dst], axis=1).ravel()
            other = np.stack([dst, src], axis=1).ravel()
            order = np.lexsort((np.arange(2 * E), np.repeat(ts, 2), node))
            rowptr = np.zeros(self.n_nodes + 1, dtype=np.int64)
            if E:
                np.cumsum(np.bincount(node, minlength=self.n_nodes), out=rowptr[1:])
            self.rowptr = torch.as_tensor(rowptr, device=self.device)
            self.nbr = torch.as_tensor(other[order].astype(np.int32), device=self.device)
            self.eidx = torch.as_tensor(np.repeat(eid, 2)[order].astype(np.int32), device=self.device)
            self.ts = torch.as_tensor(np.repeat(ts, 2)[order], device=self.device)

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.rowptr, self.nbr, self.eidx, self.ts))


class NeighborFinder:
    """Drop-in for reference utils/utils.py:130-220 over a `TemporalCSR`."""

    def __init__(self, csr: TemporalCSR, uniform=False, seed=0):
        self.csr = csr
        self.uniform = bool(uniform)
        self.seed = int(seed)
        self.call_id = 0
        # device part of the uniform mode's call counter: a captured CUDA graph bakes `call_id` in, so the step bumps
        # this counter on the stream instead (engine.TGNStepFunction) and every replay draws a fresh Philox stream
        self.call_ctr = torch.zeros(1, dtype=torch.int32, device=csr.device) if self.uniform else None
        self.lanes_per_query = 0        # K1 search width: 0 = chosen by the kernel launcher from the query count

    # device API used by the engine -----------------------------------------------------
    def sample(self, q_nodes: torch.Tensor, q_ts: torch.Tensor, n_neighbors: int, out=None, q_ids=None, ld_out=0):
        """q_nodes int32[Q], q_ts float64[Q] on the device -> (nbr i32, eidx i32, etime f32, dt f32) [Q, n].
        q_ids / ld_out / out[2] = None: the node-sharded caller's query ids, reply-row stride and skipped edge times
        (see include/pfo_b200.h)."""
        Q = q_nodes.shape[0]
        n = max(int(n_neighbors), 1)
        dev = q_nodes.device
        if out is None:
            out = (torch.empty((Q, n), dtype=torch.int32, device=dev),
                   torch.empty((Q, n), dtype=torch.int32, device=dev),
                   torch.empty((Q, n), dtype=torch.float32, device=dev),
                   torch.empty((Q, n), dtype=torch.float32, device=dev))
        if n_neighbors <= 0:                       # utils/utils.py:175: one all-zero column
            for o in out[:3]:
                o.zero_()
            out[3].copy_(q_ts.to(torch.float32).unsqueeze(1))
            return out
        c = self.csr
        _lib.call("pfo_neighbor_sample", _lib.ptr(c.rowptr), _lib.ptr(c.nbr), _lib.ptr(c.eidx), _lib.ptr(c.ts),
                  _lib.ptr(q_nodes), _lib.ptr(q_ts), Q, n, int(self.uniform), self.seed, self.call_id,
                  _lib.ptr(self.call_ctr), int(self.lanes_per_query), _lib.ptr(q_ids), int(ld_out), _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.ptr(out[2]), _lib.ptr(out[3]))
        self.call_id += 1
        return out

    # numpy API of the reference --------------------------------------------------------
    def get_temporal_neighbor(self, source_nodes, timestamps, n_neighbors=20):
        assert len(source_nodes) == len(timestamps)
        dev = self.csr.device
        qn = torch.as_tensor(np.asarray(source_nodes).astype(np.int32), device=dev)
        qt = torch.as_tensor(np.asarray(timestamps, dtype=np.float64), device=dev)
        nbr, eidx, etime, _ = self.sample(qn, qt, n_neighbors)
        return nbr.cpu().numpy(), eidx.cpu().numpy(), etime.cpu().numpy()

    def find_before(self, src_idx, cut_time):
        """All interactions of `src_idx` strictly before `cut_time` (utils/utils.py:150-161)."""
        c = self.csr
        lo, hi = int(c.rowptr[src_idx]), int(c.rowptr[src_idx + 1])
        ts = c.ts[lo:hi].cpu().numpy()
        i = int(np.searchsorted(ts, cut_time))
        return (c.nbr[lo:lo + i].cpu().numpy().astype(np.int64), c.eidx[lo:lo + i].cpu().numpy().astype(np.int64),
                ts[:i])


def get_neighbor_finder(data, uniform, max_node_idx=None, device="cuda", seed=0):
    """Reference utils/utils.py:117-127: `data` has .sources/.destinations/.edge_idxs/.timestamps."""
    if max_node_idx is None:
        max_node_idx = max(int(np.max(data.sources)), int(np.max(data.destinations)))
    csr = TemporalCSR(data.sources, data.destinations, data.edge_idxs, data.timestamps,
                      n_nodes=int(max_node_idx) + 1, device=device)
    return NeighborFinder(csr, uniform=uniform, seed=seed)
