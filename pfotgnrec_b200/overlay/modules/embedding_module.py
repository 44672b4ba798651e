"""Drop-in for reference modules/embedding_module.py: parameter containers with the reference's
attribute names; the embedding itself is computed by the engine behind TGN (K1 + K3 + K4)."""
import math

import torch
from torch import nn

from model.temporal_attention import TemporalAttentionLayer


class EmbeddingModule(nn.Module):
    def __init__(self, node_features, edge_features, memory, neighbor_finder, time_encoder, n_layers,
                 n_node_features, n_edge_features, n_time_features, embedding_dimension, device, dropout):
        super(EmbeddingModule, self).__init__()
        self.node_features = node_features
        self.edge_features = edge_features
        self.memory = memory
        self.neighbor_finder = neighbor_finder
        self.time_encoder = time_encoder
        self.n_layers = n_layers
        self.n_node_features = n_node_features
        self.n_edge_features = n_edge_features
        self.n_time_features = n_time_features
        self.dropout = dropout
        self.embedding_dimension = embedding_dimension
        self.device = device

    def compute_embedding(self, memory, source_nodes, timestamps, n_layers, n_neighbors=20, time_diffs=None,
                          use_time_proj=True):
        raise NotImplementedError("embeddings are computed inside TGN.compute_temporal_embeddings*")


class IdentityEmbedding(EmbeddingModule):
    pass


class TimeEmbedding(EmbeddingModule):
    def __init__(self, node_features, edge_features, memory, neighbor_finder, time_encoder, n_layers,
                 n_node_features, n_edge_features, n_time_features, embedding_dimension, device,
                 n_heads=2, dropout=0.1, use_memory=True, n_neighbors=1):
        super(TimeEmbedding, self).__init__(node_features, edge_features, memory, neighbor_finder, time_encoder,
                                            n_layers, n_node_features, n_edge_features, n_time_features,
                                            embedding_dimension, device, dropout)

        class NormalLinear(nn.Linear):
            def reset_parameters(self):      # Jodie's init (reference embedding_module.py:47-53)
                stdv = 1. / math.sqrt(self.weight.size(1))
                self.weight.data.normal_(0, stdv)
                if self.bias is not None:
                    self.bias.data.normal_(0, stdv)

        self.embedding_layer = NormalLinear(1, self.n_node_features)


class GraphEmbedding(EmbeddingModule):
    def __init__(self, node_features, edge_features, memory, neighbor_finder, time_encoder, n_layers,
                 n_node_features, n_edge_features, n_time_features, embedding_dimension, device,
                 n_heads=2, dropout=0.1, use_memory=True):
        super(GraphEmbedding, self).__init__(node_features, edge_features, memory, neighbor_finder, time_encoder,
                                             n_layers, n_node_features, n_edge_features, n_time_features,
                                             embedding_dimension, device, dropout)
        self.use_memory = use_memory


class GraphSumEmbedding(GraphEmbedding):
    """Parameter container of reference embedding_module.py:183-219 (same attribute names and init order); the sum over
    the sampled slots and both linear layers run in the step engine (TGNEngine._sum_forward)."""

    def __init__(self, node_features, edge_features, memory, neighbor_finder, time_encoder, n_layers,
                 n_node_features, n_edge_features, n_time_features, embedding_dimension, device,
                 n_heads=2, dropout=0.1, use_memory=True):
        super(GraphSumEmbedding, self).__init__(node_features, edge_features, memory, neighbor_finder, time_encoder,
                                                n_layers, n_node_features, n_edge_features, n_time_features,
                                                embedding_dimension, device, n_heads, dropout, use_memory)
        self.linear_1 = torch.nn.ModuleList([torch.nn.Linear(embedding_dimension + n_time_features + n_edge_features,
                                                             embedding_dimension) for _ in range(n_layers)])
        self.linear_2 = torch.nn.ModuleList([torch.nn.Linear(embedding_dimension + n_node_features + n_time_features,
                                                             embedding_dimension) for _ in range(n_layers)])


class GraphAttentionEmbedding(GraphEmbedding):
    def __init__(self, node_features, edge_features, memory, neighbor_finder, time_encoder, n_layers,
                 n_node_features, n_edge_features, n_time_features, embedding_dimension, device,
                 n_heads=2, dropout=0.1, use_memory=True):
        super(GraphAttentionEmbedding, self).__init__(node_features, edge_features, memory, neighbor_finder,
                                                      time_encoder, n_layers, n_node_features, n_edge_features,
                                                      n_time_features, embedding_dimension, device, n_heads,
                                                      dropout, use_memory)
        self.attention_models = torch.nn.ModuleList([TemporalAttentionLayer(
            n_node_features=n_node_features, n_neighbors_features=n_node_features,
            n_edge_features=n_edge_features, time_dim=n_time_features, n_head=n_heads, dropout=dropout,
            output_dimension=n_node_features) for _ in range(n_layers)])


def get_embedding_module(module_type, node_features, edge_features, memory, neighbor_finder,
                         time_encoder, n_layers, n_node_features, n_edge_features, n_time_features,
                         embedding_dimension, device, n_heads=2, dropout=0.1, n_neighbors=None, use_memory=True):
    common = dict(node_features=node_features, edge_features=edge_features, memory=memory,
                  neighbor_finder=neighbor_finder, time_encoder=time_encoder, n_layers=n_layers,
                  n_node_features=n_node_features, n_edge_features=n_edge_features,
                  n_time_features=n_time_features, embedding_dimension=embedding_dimension, device=device)
    if module_type == "graph_attention":
        return GraphAttentionEmbedding(n_heads=n_heads, dropout=dropout, use_memory=use_memory, **common)
    elif module_type == "graph_sum":
        return GraphSumEmbedding(n_heads=n_heads, dropout=dropout, use_memory=use_memory, **common)
    elif module_type == "identity":
        return IdentityEmbedding(dropout=dropout, **common)
    elif module_type == "time":
        return TimeEmbedding(dropout=dropout, n_neighbors=n_neighbors, **common)
    raise ValueError("Embedding Module {} not supported".format(module_type))
