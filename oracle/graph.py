"""Oracle: temporal adjacency and neighbour lookup (test infrastructure only).

Restates reference utils/utils.py:117-127 (get_neighbor_finder),
:131-148 (NeighborFinder.__init__), :150-161 (find_before) and :163-220
(get_temporal_neighbor) on flat numpy arrays.
"""
import numpy as np

from .philox import philox4x32_10, mulhi32, PURPOSE_NBR


class AdjacencyOracle:
    """Undirected, per-node time-sorted adjacency.

    Every event (s, d, eidx, t) is appended to both endpoints in stream order
    (reference utils/utils.py:120-125) and each node's list is then sorted by
    timestamp with a STABLE sort (Python `sorted`, :139), so equal timestamps
    keep stream order.
    """

    def __init__(self, sources, destinations, edge_idxs, timestamps, n_nodes=None, uniform=False):
        sources = np.asarray(sources, dtype=np.int64)
        destinations = np.asarray(destinations, dtype=np.int64)
        edge_idxs = np.asarray(edge_idxs, dtype=np.int64)
        ts = np.asarray(timestamps, dtype=np.float64)
        E = sources.shape[0]
        if n_nodes is None:
            n_nodes = int(max(sources.max(), destinations.max())) + 1 if E else 1
        # interleave so that the per-node append order equals the reference's loop order
        node = np.stack([sources, destinations], axis=1).ravel()
        other = np.stack([destinations, sources], axis=1).ravel()
        eid = np.repeat(edge_idxs, 2)
        t = np.repeat(ts, 2)
        order = np.lexsort((np.arange(2 * E), t, node))   # by node, then ts, then append order
        self.n_nodes = n_nodes
        self.rowptr = np.zeros(n_nodes + 1, dtype=np.int64)
        np.cumsum(np.bincount(node, minlength=n_nodes), out=self.rowptr[1:])
        self.nbr = other[order]
        self.eidx = eid[order]
        self.ts = t[order]
        self.uniform = uniform

    def count_before(self, node, cut_time):
        """np.searchsorted(ts[node], cut_time) with side='left' -> strictly earlier (utils.py:158)."""
        lo, hi = self.rowptr[node], self.rowptr[node + 1]
        return int(np.searchsorted(self.ts[lo:hi], cut_time, side="left"))

    def get_temporal_neighbor(self, source_nodes, timestamps, n_neighbors=20, call_id=0, seed=0):
        """Most-recent (default) or uniform-with-replacement temporal neighbours.

        Output: neighbors int32[Q,n], edge_idxs int32[Q,n], edge_times float32[Q,n];
        rows are right-aligned and left-padded with zeros (utils.py:216-218);
        n_neighbors == 0 is widened to one all-zero column (:175).

        Uniform mode (utils.py:193-204) draws with replacement; the oracle draws slot j of
        query q as pos = mulhi32(philox(q, call_id, j, PURPOSE_NBR; seed)[0], i) instead of
        numpy's MT19937 (deviation i) and orders the picks by (fp32 time, CSR position),
        a total order standing in for the reference's unstable argsort (deviation ii).
        """
        source_nodes = np.asarray(source_nodes, dtype=np.int64)
        timestamps = np.asarray(timestamps, dtype=np.float64)
        assert source_nodes.shape[0] == timestamps.shape[0]
        Q = source_nodes.shape[0]
        n = n_neighbors if n_neighbors > 0 else 1
        out_n = np.zeros((Q, n), dtype=np.int32)
        out_e = np.zeros((Q, n), dtype=np.int32)
        out_t = np.zeros((Q, n), dtype=np.float32)
        if n_neighbors <= 0:
            return out_n, out_e, out_t
        for q in range(Q):
            lo = self.rowptr[source_nodes[q]]
            i = self.count_before(source_nodes[q], timestamps[q])
            if i == 0:
                continue
            if self.uniform:
                j = np.arange(n, dtype=np.int64)
                x0 = philox4x32_10(q, call_id, j, PURPOSE_NBR, seed & 0xFFFFFFFF, seed >> 32)[0]
                pos = mulhi32(x0, i)
                t32 = self.ts[lo + pos].astype(np.float32)
                o = np.lexsort((pos, t32))
                pos = pos[o]
                out_n[q] = self.nbr[lo + pos]
                out_e[q] = self.eidx[lo + pos]
                out_t[q] = t32[o]
            else:
                k = min(i, n)
                sl = slice(lo + i - k, lo + i)
                out_n[q, n - k:] = self.nbr[sl]
                out_e[q, n - k:] = self.eidx[sl]
                out_t[q, n - k:] = self.ts[sl].astype(np.float32)
        return out_n, out_e, out_t
