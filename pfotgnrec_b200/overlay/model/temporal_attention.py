"""Drop-in for reference model/temporal_attention.py: parameter container with the reference's
sub-module names and construction order (MergeLayer first, then nn.MultiheadAttention), so that
initial weights and state_dict keys match.  Forward/backward run in the fused K4 path."""
import torch
from torch import nn

from utils.utils import MergeLayer


class TemporalAttentionLayer(torch.nn.Module):
    def __init__(self, n_node_features, n_neighbors_features, n_edge_features, time_dim,
                 output_dimension, n_head=2, dropout=0.1):
        super(TemporalAttentionLayer, self).__init__()
        self.n_head = n_head
        self.feat_dim = n_node_features
        self.time_dim = time_dim
        self.query_dim = n_node_features + time_dim
        self.key_dim = n_neighbors_features + time_dim + n_edge_features
        self.merger = MergeLayer(self.query_dim, n_node_features, n_node_features, output_dimension)
        self.multi_head_target = nn.MultiheadAttention(embed_dim=self.query_dim, kdim=self.key_dim,
                                                       vdim=self.key_dim, num_heads=n_head, dropout=dropout)

    def forward(self, *args, **kwargs):
        raise NotImplementedError("TemporalAttentionLayer is evaluated inside TGN.compute_temporal_embeddings*")
