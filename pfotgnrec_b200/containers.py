"""Parameter containers of the drop-in overlay.

The reference's modules (model/time_encoding.py, model/temporal_attention.py, modules/*.py, utils/utils.py:4-17) own
both the weights and the arithmetic.  Here the arithmetic lives in libpfo_b200.so behind the step engine, so these
classes only hold the weights -- under the reference's class names, attribute names and CONSTRUCTION ORDER, because
three things the callers rely on derive from exactly that: the `state_dict` keys of a checkpoint, the order of
`named_parameters()` Adam sees, and the initial values drawn from the global torch generator after
`torch.manual_seed` (tests/test_overlay_host.py compares all three with the unmodified reference).
`pfotgnrec_b200/overlay/{model,modules,utils}/*.py` re-export them under the reference's module paths.
"""
from collections import defaultdict
import math

import numpy as np
import torch
from torch import nn

from .engine import TGNState, ModelConfig

_FUSED = "evaluated inside TGN.compute_temporal_embeddings* (libpfo_b200.so); this class only holds the weights"


class _Container(nn.Module):
    def forward(self, *args, **kwargs):
        raise NotImplementedError(f"{type(self).__name__} is {_FUSED}")


class _Exact:
    gemm_mode = "simt"          # standalone sub-module calls use the fp32 FFMA kernel


@torch.no_grad()
def _dense(x, linear, relu=False):
    """y = act(x W^T + b) through pfo_linear_f32 (inference only: the trained path is the fused step)."""
    from . import _lib
    from .engine import _linear
    if x.device.type != "cuda":
        raise _lib.PfoError("sub-module forwards run in libpfo_b200.so on a CUDA device: there is no CPU fallback")
    x = x.contiguous().float()
    M, K = x.shape
    lda = (K + 3) // 4 * 4                   # operand rows are read as 16-byte vectors: pad the row stride like the engine does
    if lda != K:
        x = torch.nn.functional.pad(x, (0, lda - K))
    W, b = linear.weight.detach().contiguous(), linear.bias.detach().contiguous()
    N = W.shape[0]
    ldc = (N + 3) // 4 * 4
    y = torch.empty(M, ldc, device=x.device)
    _linear(_Exact, _lib.ptr(x), lda, None, _lib.ptr(W), K, 0, _lib.ptr(b), _lib.ptr(y), ldc, M, N, K, act=1 if relu else 0)
    return y[:, :N] if ldc != N else y


# ----------------------------------------------------------------------------- model/time_encoding.py:5-25
class TimeEncode(_Container):
    """cos(t * w + b), w_k = 10^(-9k/(d-1)), b = 0, both learnable."""

    def __init__(self, dimension):
        super().__init__()
        self.dimension = dimension
        self.w = nn.Linear(1, dimension)                 # draws from the generator exactly like the reference does
        freq = torch.from_numpy(1 / 10 ** np.linspace(0, 9, dimension)).float()
        self.w.weight = nn.Parameter(freq.reshape(dimension, -1))
        self.w.bias = nn.Parameter(torch.zeros(dimension).float())

    @torch.no_grad()
    def forward(self, t):
        """model/time_encoding.py:17-25 on its own (inference only; inside the TGN step the encoding is fused into the
        message and attention kernels): t [batch, seq] -> cos(t * w + b) [batch, seq, dimension]."""
        from . import _lib
        t = t.contiguous().float()
        out = torch.empty(tuple(t.shape) + (self.dimension,), device=t.device)
        if t.device.type != "cuda":
            raise _lib.PfoError("TimeEncode.forward runs in libpfo_b200.so on a CUDA device: there is no CPU fallback")
        _lib.call("pfo_time_encode", _lib.ptr(t), _lib.ptr(self.w.weight.detach().reshape(-1).contiguous()),
                  _lib.ptr(self.w.bias.detach().contiguous()), t.numel(), self.dimension, 0, _lib.ptr(out), None)
        return out


# ----------------------------------------------------------------------------- utils/utils.py:4-17
class MergeLayer(_Container):
    """fc2(relu(fc1([x1 | x2]))), Xavier-normal weights."""

    def __init__(self, dim1, dim2, dim3, dim4):
        super().__init__()
        self.fc1, self.fc2, self.act = nn.Linear(dim1 + dim2, dim3), nn.Linear(dim3, dim4), nn.ReLU()
        for lin in (self.fc1, self.fc2):
            nn.init.xavier_normal_(lin.weight)

    def forward(self, x1, x2):
        """utils/utils.py:14-17 on its own (inference only): fc2(relu(fc1([x1 | x2])))."""
        return _dense(_dense(torch.cat([x1, x2], dim=1), self.fc1, relu=True), self.fc2)


# ----------------------------------------------------------------------------- model/temporal_attention.py:7-32
class TemporalAttentionLayer(_Container):
    """Merge MLP first, then nn.MultiheadAttention(embed = d + d_time, kdim = vdim = d + d_time + F)."""

    def __init__(self, n_node_features, n_neighbors_features, n_edge_features, time_dim,
                 output_dimension, n_head=2, dropout=0.1):
        super().__init__()
        self.n_head, self.feat_dim, self.time_dim = n_head, n_node_features, time_dim
        self.query_dim = n_node_features + time_dim
        self.key_dim = n_neighbors_features + time_dim + n_edge_features
        self.merger = MergeLayer(self.query_dim, n_node_features, n_node_features, output_dimension)
        self.multi_head_target = nn.MultiheadAttention(embed_dim=self.query_dim, kdim=self.key_dim, vdim=self.key_dim,
                                                       num_heads=n_head, dropout=dropout)


# ----------------------------------------------------------------------------- modules/memory.py:8-75
class Memory(nn.Module):
    """memory / last_update / pending messages over the dense device state (engine.TGNState)."""

    def __init__(self, n_nodes, memory_dimension, input_dimension, message_dimension=None,
                 device="cpu", combination_method='sum', n_edge_features=None):
        super().__init__()
        self.n_nodes, self.memory_dimension = n_nodes, memory_dimension
        self.input_dimension, self.message_dimension = input_dimension, message_dimension
        self.device, self.combination_method = device, combination_method
        F = n_edge_features if n_edge_features is not None else input_dimension - 3 * memory_dimension
        self._state = TGNState(n_nodes, ModelConfig(d=memory_dimension, n_edge_feat=F), device)
        self.__init_memory__()

    def __init_memory__(self):
        """Zero the state; main.py calls this at the start of every epoch."""
        self._state.reset()
        # no-grad parameters sharing the state's storage, so that they are saved with the model
        self.memory = nn.Parameter(self._state.memory, requires_grad=False)
        self.last_update = nn.Parameter(self._state.last_update, requires_grad=False)

    @property
    def state(self):
        # .to(device) may have re-homed the parameters: keep the engine's view on the same storage
        if self.memory.data_ptr() != self._state.memory.data_ptr():
            self._state.memory = self.memory.data
        if self.last_update.data_ptr() != self._state.last_update.data_ptr():
            self._state.last_update = self.last_update.data
        return self._state

    @property
    def messages(self):
        """{node: [(message, timestamp)]} view of the pending table (debugging / compatibility)."""
        st, out = self._state, defaultdict(list)
        for node in torch.nonzero(st.pend_valid).flatten().tolist():
            out[node] = [(st.pend_msg[node, :st.cfg.raw].clone(), st.pend_ts[node].clone())]
        return out

    def store_raw_messages(self, nodes, node_id_to_messages):
        """modules/memory.py:35-37 on the dense table: the LAST message appended for a node is the one the `last`
        aggregator would pick, so it becomes the node's pending row (one stacked copy, no per-row kernel launches)."""
        st = self._state
        keep = [(int(node), node_id_to_messages[node][-1]) for node in nodes if len(node_id_to_messages[node]) > 0]
        if not keep:
            return
        idx = torch.as_tensor([k for k, _ in keep], dtype=torch.long, device=st.pend_msg.device)
        st.pend_msg[idx, :st.cfg.raw] = torch.stack([m[0] for _, m in keep]).to(st.pend_msg)
        st.pend_ts[idx] = torch.stack([torch.as_tensor(m[1]).reshape(()) for _, m in keep]).to(st.pend_ts)
        st.pend_valid[idx] = 1

    def get_memory(self, node_idxs):
        return self.memory[node_idxs, :]

    def set_memory(self, node_idxs, values):
        self.memory[node_idxs, :] = values

    def get_last_update(self, node_idxs):
        return self.last_update[node_idxs]

    def backup_memory(self):
        b = self.state.backup()
        return b[0], b[1], b[2:]

    def restore_memory(self, memory_backup):
        self.state.restore((memory_backup[0], memory_backup[1]) + tuple(memory_backup[2]))

    def detach_memory(self):
        """The dense state never carries autograd history: nothing to detach."""
        return None

    def clear_messages(self, nodes):
        idx = torch.as_tensor(list(nodes), dtype=torch.long, device=self._state.device)
        self._state.pend_valid[idx] = 0


# ----------------------------------------------------------------------------- modules/message_function.py:4-40
class MessageFunction(nn.Module):
    def compute_message(self, raw_messages):
        return None


class IdentityMessageFunction(MessageFunction):
    def compute_message(self, raw_messages):
        return raw_messages


class MLPMessageFunction(MessageFunction):
    """Linear(raw, raw // 2) -> ReLU -> Linear(raw // 2, message_dimension); runs inside the lazy memory update."""

    def __init__(self, raw_message_dimension, message_dimension):
        super().__init__()
        half = raw_message_dimension // 2
        self.mlp = self.layers = nn.Sequential(nn.Linear(raw_message_dimension, half), nn.ReLU(),
                                               nn.Linear(half, message_dimension))

    def compute_message(self, raw_messages):
        """modules/message_function.py:23-26 on its own (inference only; the trained path runs it inside the lazy
        memory update)."""
        return _dense(_dense(raw_messages, self.mlp[0], relu=True), self.mlp[2])


def get_message_function(module_type, raw_message_dimension, message_dimension):
    if module_type == "identity":
        return IdentityMessageFunction()
    if module_type == "mlp":
        return MLPMessageFunction(raw_message_dimension, message_dimension)
    raise ValueError(module_type)


# ----------------------------------------------------------------------------- modules/message_aggregator.py:6-90
class MessageAggregator(nn.Module):
    """`last` = the last-wins scatter of pfo_store_messages, `mean` = the in-order segment mean of
    pfo_store_messages_mean; `aggregate` serves the dict-of-lists compatibility view."""

    reduce = None

    def __init__(self, device):
        super().__init__()
        self.device = device

    def aggregate(self, node_ids, messages):
        if self.reduce is None:
            raise NotImplementedError
        ids = [n for n in sorted({int(x) for x in node_ids}) if len(messages[n]) > 0]
        if not ids:
            return ids, [], []
        return (ids, torch.stack([type(self).reduce(messages[n]) for n in ids]),
                torch.stack([messages[n][-1][1] for n in ids]))


class LastMessageAggregator(MessageAggregator):
    reduce = staticmethod(lambda lst: lst[-1][0])


class MeanMessageAggregator(MessageAggregator):
    reduce = staticmethod(lambda lst: torch.mean(torch.stack([m[0] for m in lst]), dim=0))


def get_message_aggregator(aggregator_type, device):
    kinds = {"last": LastMessageAggregator, "mean": MeanMessageAggregator}
    if aggregator_type not in kinds:
        raise ValueError("Message aggregator {} not implemented".format(aggregator_type))
    return kinds[aggregator_type](device=device)


# ----------------------------------------------------------------------------- modules/memory_updater.py:5-76
class MemoryUpdater(nn.Module):
    def update_memory(self, unique_node_ids, unique_messages, timestamps):
        pass


class SequenceMemoryUpdater(MemoryUpdater):
    cell = None

    def __init__(self, memory, message_dimension, memory_dimension, device):
        super().__init__()
        self.memory = memory
        self.layer_norm = nn.LayerNorm(memory_dimension)         # constructed, never applied (as in the reference)
        self.message_dimension, self.device = message_dimension, device
        if self.cell is not None:
            self.memory_updater = self.cell(input_size=message_dimension, hidden_size=memory_dimension)

    def update_memory(self, unique_node_ids, unique_messages, timestamps):
        raise NotImplementedError(f"the memory is persisted inside the step: {_FUSED}")

    def get_updated_memory(self, unique_node_ids, unique_messages, timestamps):
        raise NotImplementedError(f"the memory is updated lazily inside the step: {_FUSED}")


class GRUMemoryUpdater(SequenceMemoryUpdater):
    cell = nn.GRUCell


class RNNMemoryUpdater(SequenceMemoryUpdater):
    cell = nn.RNNCell


def get_memory_updater(module_type, memory, message_dimension, memory_dimension, device):
    kinds = {"gru": GRUMemoryUpdater, "rnn": RNNMemoryUpdater}
    if module_type not in kinds:
        raise ValueError("Memory updater {} not implemented".format(module_type))
    return kinds[module_type](memory, message_dimension, memory_dimension, device)


# ----------------------------------------------------------------------------- modules/embedding_module.py:9-321
class EmbeddingModule(nn.Module):
    """Holds what the reference's embedding modules hold (the callers set `.neighbor_finder` directly, main.py:427)."""

    def __init__(self, node_features, edge_features, memory, neighbor_finder, time_encoder, n_layers,
                 n_node_features, n_edge_features, n_time_features, embedding_dimension, device, dropout):
        super().__init__()
        self.node_features, self.edge_features = node_features, edge_features
        self.memory, self.neighbor_finder, self.time_encoder = memory, neighbor_finder, time_encoder
        self.n_layers, self.dropout, self.device = n_layers, dropout, device
        self.n_node_features, self.n_edge_features = n_node_features, n_edge_features
        self.n_time_features, self.embedding_dimension = n_time_features, embedding_dimension

    def compute_embedding(self, memory, source_nodes, timestamps, n_layers, n_neighbors=20, time_diffs=None,
                          use_time_proj=True):
        raise NotImplementedError(f"embeddings are {_FUSED}")


class IdentityEmbedding(EmbeddingModule):
    pass


class TimeEmbedding(EmbeddingModule):
    """Jodie: memory * (1 + Linear(1 -> d)(time_diff)), normal(0, 1/sqrt(fan_in)) initialisation."""

    def __init__(self, node_features, edge_features, memory, neighbor_finder, time_encoder, n_layers,
                 n_node_features, n_edge_features, n_time_features, embedding_dimension, device,
                 n_heads=2, dropout=0.1, use_memory=True, n_neighbors=1):
        super().__init__(node_features, edge_features, memory, neighbor_finder, time_encoder, n_layers,
                         n_node_features, n_edge_features, n_time_features, embedding_dimension, device, dropout)

        class NormalLinear(nn.Linear):
            def reset_parameters(self):
                stdv = 1. / math.sqrt(self.weight.size(1))
                self.weight.data.normal_(0, stdv)
                if self.bias is not None:
                    self.bias.data.normal_(0, stdv)

        self.embedding_layer = NormalLinear(1, self.n_node_features)


class GraphEmbedding(EmbeddingModule):
    def __init__(self, node_features, edge_features, memory, neighbor_finder, time_encoder, n_layers,
                 n_node_features, n_edge_features, n_time_features, embedding_dimension, device,
                 n_heads=2, dropout=0.1, use_memory=True):
        super().__init__(node_features, edge_features, memory, neighbor_finder, time_encoder, n_layers,
                         n_node_features, n_edge_features, n_time_features, embedding_dimension, device, dropout)
        self.use_memory = use_memory


class GraphSumEmbedding(GraphEmbedding):
    """linear_1 over [h_nbr | te | e] summed over the slots, linear_2 over [sum | h_q | te(0)] (TGNEngine._sum_forward)."""

    def __init__(self, node_features, edge_features, memory, neighbor_finder, time_encoder, n_layers,
                 n_node_features, n_edge_features, n_time_features, embedding_dimension, device,
                 n_heads=2, dropout=0.1, use_memory=True):
        super().__init__(node_features, edge_features, memory, neighbor_finder, time_encoder, n_layers,
                         n_node_features, n_edge_features, n_time_features, embedding_dimension, device,
                         n_heads, dropout, use_memory)
        d = embedding_dimension
        self.linear_1 = nn.ModuleList([nn.Linear(d + n_time_features + n_edge_features, d) for _ in range(n_layers)])
        self.linear_2 = nn.ModuleList([nn.Linear(d + n_node_features + n_time_features, d) for _ in range(n_layers)])


class GraphAttentionEmbedding(GraphEmbedding):
    def __init__(self, node_features, edge_features, memory, neighbor_finder, time_encoder, n_layers,
                 n_node_features, n_edge_features, n_time_features, embedding_dimension, device,
                 n_heads=2, dropout=0.1, use_memory=True):
        super().__init__(node_features, edge_features, memory, neighbor_finder, time_encoder, n_layers,
                         n_node_features, n_edge_features, n_time_features, embedding_dimension, device,
                         n_heads, dropout, use_memory)
        self.attention_models = nn.ModuleList([TemporalAttentionLayer(
            n_node_features=n_node_features, n_neighbors_features=n_node_features, n_edge_features=n_edge_features,
            time_dim=n_time_features, n_head=n_heads, dropout=dropout, output_dimension=n_node_features)
            for _ in range(n_layers)])


def get_embedding_module(module_type, node_features, edge_features, memory, neighbor_finder,
                         time_encoder, n_layers, n_node_features, n_edge_features, n_time_features,
                         embedding_dimension, device, n_heads=2, dropout=0.1, n_neighbors=None, use_memory=True):
    common = dict(node_features=node_features, edge_features=edge_features, memory=memory,
                  neighbor_finder=neighbor_finder, time_encoder=time_encoder, n_layers=n_layers,
                  n_node_features=n_node_features, n_edge_features=n_edge_features,
                  n_time_features=n_time_features, embedding_dimension=embedding_dimension, device=device)
    if module_type in ("graph_attention", "graph_sum"):
        cls = GraphAttentionEmbedding if module_type == "graph_attention" else GraphSumEmbedding
        return cls(n_heads=n_heads, dropout=dropout, use_memory=use_memory, **common)
    if module_type == "identity":
        return IdentityEmbedding(dropout=dropout, **common)
    if module_type == "time":
        return TimeEmbedding(dropout=dropout, n_neighbors=n_neighbors, **common)
    raise ValueError("Embedding Module {} not supported".format(module_type))
