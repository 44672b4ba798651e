"""Drop-in for reference model/tgn.py: class TGN with the same constructor, methods, sub-module
names and state_dict keys; the arithmetic runs in libpfo_b200.so through pfotgnrec_b200.engine.

Put `pfotgnrec_b200/overlay` ahead of the reference on sys.path and the reference's unchanged
`main.py` / `evaluation.py` import this class."""
import logging

import numpy as np
import torch

from pfotgnrec_b200 import _lib
from pfotgnrec_b200.containers import (Memory, TimeEncode, get_embedding_module, get_memory_updater,
                                       get_message_aggregator, get_message_function)
from pfotgnrec_b200.engine import TGNEngine, ModelConfig
from pfotgnrec_b200.graph import NeighborFinder, TemporalCSR


def _as_device_finder(nf, device):
    """Accept this package's NeighborFinder or a reference-style one (lists of per-node arrays)."""
    if nf is None or isinstance(nf, NeighborFinder):
        return nf
    if getattr(nf, "_pfo_csr", None) is None:
        node, other, eidx, ts = [], [], [], []
        for v, (nb, ei, tt) in enumerate(zip(nf.node_to_neighbors, nf.node_to_edge_idxs, nf.node_to_edge_timestamps)):
            node.append(np.full(len(nb), v, dtype=np.int64)); other.append(np.asarray(nb, dtype=np.int64))
            eidx.append(np.asarray(ei, dtype=np.int64)); ts.append(np.asarray(tt, dtype=np.float64))
        csr = TemporalCSR.__new__(TemporalCSR)
        node = np.concatenate(node) if node else np.zeros(0, np.int64)
        csr.n_nodes, csr.n_events, csr.device = len(nf.node_to_neighbors), len(node) // 2, torch.device(device)
        rowptr = np.zeros(csr.n_nodes + 1, dtype=np.int64)
        np.cumsum(np.bincount(node, minlength=csr.n_nodes), out=rowptr[1:])
        csr.rowptr = torch.as_tensor(rowptr, device=device)
        csr.nbr = torch.as_tensor(np.concatenate(other).astype(np.int32), device=device)
        csr.eidx = torch.as_tensor(np.concatenate(eidx).astype(np.int32), device=device)
        csr.ts = torch.as_tensor(np.concatenate(ts), device=device)
        nf._pfo_csr = NeighborFinder(csr, uniform=nf.uniform)
    return nf._pfo_csr


class TGN(torch.nn.Module):
    def __init__(self, neighbor_finder, node_features, edge_features, device, n_layers=2,
                 n_heads=2, dropout=0.1, use_memory=False,
                 memory_update_at_start=True,
                 message_dimension=100, memory_dimension=500,
                 embedding_module_type="graph_attention",
                 message_function="mlp",
                 mean_time_shift_src=0, std_time_shift_src=1, mean_time_shift_dst=0,
                 std_time_shift_dst=1, n_neighbors=None, aggregator_type="last",
                 memory_updater_type="gru",
                 use_destination_embedding_in_message=False,
                 use_source_embedding_in_message=False,
                 dyrep=False, gemm_mode="fp32", memory_nodes=None):
        super().__init__()
        _lib.use_device(device)      # main.py:103 hands cuda:{gpu}: the kernels launch on the current device's stream
        if use_memory and not memory_update_at_start:
            raise NotImplementedError("memory_update_at_start=False is never executed by main.py")
        if aggregator_type not in ("last", "mean"):
            raise ValueError("Message aggregator {} not implemented".format(aggregator_type))
        if message_function not in ("identity", "mlp"):
            raise ValueError("Message function {} not implemented".format(message_function))
        # feature tables on the device; edge features z-normalised over ALL rows, the padding row included (tgn.py:38-41)
        if isinstance(edge_features, torch.Tensor):      # extension: tables already on the device (scale configuration)
            ef = edge_features.to(device=device, dtype=torch.float32)
            var, mean = torch.var_mean(ef, dim=0, unbiased=False)
            self.edge_raw_features = ((ef - mean) / var.sqrt()).contiguous()
            self.node_raw_features = node_features.to(device=device, dtype=torch.float32).contiguous()
        else:
            ef = edge_features.astype(np.float32)
            ef = (ef - ef.mean(axis=0)) / ef.std(axis=0)
            self.node_raw_features = torch.from_numpy(node_features.astype(np.float32)).to(device)
            self.edge_raw_features = torch.from_numpy(ef.astype(np.float32)).to(device)
        self.n_nodes, self.n_node_features = self.node_raw_features.shape
        self.n_edge_features = self.edge_raw_features.shape[1]
        # plain attributes the callers (and checkpoints of the reference) know by name
        for name, value in dict(n_layers=n_layers, neighbor_finder=neighbor_finder, device=device,
                                logger=logging.getLogger(__name__), embedding_dimension=self.n_node_features,
                                n_neighbors=n_neighbors, embedding_module_type=embedding_module_type,
                                use_destination_embedding_in_message=use_destination_embedding_in_message,
                                use_source_embedding_in_message=use_source_embedding_in_message, dyrep=dyrep,
                                use_memory=use_memory, mean_time_shift_src=mean_time_shift_src,
                                std_time_shift_src=std_time_shift_src, mean_time_shift_dst=mean_time_shift_dst,
                                std_time_shift_dst=std_time_shift_dst).items():
            setattr(self, name, value)
        # sub-modules in the reference's construction order: the generator draws decide the initial weights
        d = self.n_node_features
        self.time_encoder = TimeEncode(dimension=d)
        self.memory = None
        if use_memory:
            if memory_dimension != d:
                raise ValueError("memory_dimension must equal the node feature dimension (memory + features)")
            self.memory_dimension, self.memory_update_at_start = memory_dimension, memory_update_at_start
            raw_dim = 2 * memory_dimension + self.n_edge_features + self.time_encoder.dimension
            message_dimension = raw_dim if message_function == "identity" else message_dimension
            # memory_nodes (extension, node-sharded mode): this rank holds the memory rows of the nodes it owns only
            self.memory = Memory(n_nodes=int(memory_nodes) if memory_nodes is not None else self.n_nodes,
                                 memory_dimension=memory_dimension, input_dimension=message_dimension,
                                 message_dimension=message_dimension, device=device,
                                 n_edge_features=self.n_edge_features)
            self.message_aggregator = get_message_aggregator(aggregator_type=aggregator_type, device=device)
            self.message_function = get_message_function(module_type=message_function, raw_message_dimension=raw_dim,
                                                          message_dimension=message_dimension)
            self.memory_updater = get_memory_updater(module_type=memory_updater_type, memory=self.memory,
                                                     message_dimension=message_dimension,
                                                     memory_dimension=memory_dimension, device=device)
        self.embedding_module = get_embedding_module(
            module_type=embedding_module_type, node_features=self.node_raw_features,
            edge_features=self.edge_raw_features, memory=self.memory, neighbor_finder=neighbor_finder,
            time_encoder=self.time_encoder, n_layers=n_layers, n_node_features=d,
            n_edge_features=self.n_edge_features, n_time_features=d, embedding_dimension=d, device=device,
            n_heads=n_heads, dropout=dropout, use_memory=use_memory, n_neighbors=n_neighbors)
        self._cfg = ModelConfig(d=self.n_node_features, n_edge_feat=self.n_edge_features, n_layers=n_layers,
                                n_heads=n_heads, use_memory=use_memory, updater=memory_updater_type,
                                embedding=embedding_module_type, dyrep=dyrep,
                                dst_emb_in_msg=use_destination_embedding_in_message,
                                src_emb_in_msg=use_source_embedding_in_message,
                                message_fn=message_function if use_memory else "identity",
                                msg_dim=int(message_dimension), aggregator=aggregator_type,
                                shift=(float(mean_time_shift_src), float(std_time_shift_src),
                                       float(mean_time_shift_dst), float(std_time_shift_dst)),
                                dropout=float(dropout), gemm_mode=gemm_mode)
        self._engine = None

    # ------------------------------------------------------------------ engine plumbing
    def _get_engine(self):
        nf = _as_device_finder(self.embedding_module.neighbor_finder, self.device)
        if self._engine is None:
            state = self.memory.state if self.use_memory else None
            self._engine = TGNEngine(self._cfg, state, self.node_raw_features, self.edge_raw_features, nf)
        eng = self._engine
        eng.nf = nf
        if self.use_memory:
            eng.state = self.memory.state
        return eng

    def _params(self):
        return {k: v for k, v in self.named_parameters(remove_duplicate=False)}

    def _run(self, source_nodes, destination_nodes, extra, edge_times, edge_idxs, n_neighbors):
        dev = self.device
        i32 = lambda a: torch.as_tensor(np.asarray(a).astype(np.int32), device=dev)
        ts = torch.as_tensor(np.asarray(edge_times, dtype=np.float64), device=dev)
        eng = self._get_engine()
        return eng.compute_temporal_embeddings(self._params(), i32(source_nodes), i32(destination_nodes),
                                               [i32(e) for e in extra], ts, i32(edge_idxs), n_neighbors,
                                               train=self.training)

    # ------------------------------------------------------------------ reference API
    def compute_temporal_embeddings_p(self, source_nodes, destination_nodes, p_pos_nodes, p_neg_nodes,
                                      edge_times, edge_idxs, n_neighbors=20):
        """Reference model/tgn.py:102-217: embeddings of [sources | destinations | p_pos | p_neg]."""
        s, d, p, n = self._run(source_nodes, destination_nodes, [p_pos_nodes, p_neg_nodes], edge_times,
                               edge_idxs, n_neighbors)
        return s, d, p, n

    def compute_temporal_embeddings(self, source_nodes, destination_nodes, p_neg_nodes,
                                    edge_times, edge_idxs, n_neighbors=20):
        """Reference model/tgn.py:219-327: embeddings of [sources | destinations | negatives]."""
        s, d, n = self._run(source_nodes, destination_nodes, [p_neg_nodes], edge_times, edge_idxs, n_neighbors)
        return s, d, n

    def set_neighbor_finder(self, neighbor_finder):
        self.neighbor_finder = neighbor_finder
        self.embedding_module.neighbor_finder = neighbor_finder
