"""Shared helpers of the parity tests: golden loading and oracle construction."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def golden_params(z):
    return {k[2:]: torch.tensor(v) for k, v in z.items() if k.startswith("w_")}


def oracle_from_golden(z, requires_grad=True):
    from oracle.graph import AdjacencyOracle
    from oracle.tgn import TGNOracle
    p = golden_params(z)
    if requires_grad:
        for v in p.values():
            v.requires_grad_(True)
    adj = AdjacencyOracle(z["st_sources"], z["st_destinations"], z["st_edge_idxs"],
                          z["st_timestamps"], n_nodes=int(z["st_n_nodes"]))
    sh = z["cfg_shift"]
    o = TGNOracle(p, adj, z["node_feat"], z["st_edge_features"], n_layers=int(z["cfg_n_layers"]),
                  n_heads=2, use_memory=bool(z["cfg_use_memory"]),
                  memory_updater=str(z["cfg_updater"]), embedding=str(z["cfg_embedding"]),
                  dyrep=bool(z["cfg_dyrep"]), use_destination_embedding_in_message=bool(z["cfg_dst_emb"]),
                  use_source_embedding_in_message=bool(z["cfg_src_emb"]) if "cfg_src_emb" in z else False,
                  message_function=str(z["cfg_msg_fn"]) if "cfg_msg_fn" in z else "identity",
                  aggregator=str(z["cfg_aggregator"]) if "cfg_aggregator" in z else "last",
                  mean_time_shift_src=sh[0], std_time_shift_src=sh[1],
                  mean_time_shift_dst=sh[2], std_time_shift_dst=sh[3])
    return o, p


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny): the 'relative error' of the parity contract."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)) if a.size else 0.0


def batch_inputs(z, bi):
    B = int(z["cfg_B"])
    sl = slice(bi * B, (bi + 1) * B)
    extra = ([z[f"b{bi}_ppos"]] if int(z["cfg_with_ppos"]) else []) + [z[f"b{bi}_neg"]]
    return (z["st_sources"][sl], z["st_destinations"][sl], extra, z["st_timestamps"][sl],
            z["st_edge_idxs"][sl])
