"""Oracle: TGN memory / message / embedding path on dense state (test infrastructure only).

A CPU-torch (fp32) restatement of the reference's model path.  Parameters are addressed
by the reference's own `state_dict` key names so that weights dumped from the reference
load directly.  The reference's dict-of-lists message store (modules/memory.py:33-37) is
restated as a dense per-node "pending message" slot; for the `last` aggregator that is
exact because a node's list is cleared whenever it is a positive (model/tgn.py:191) and
only positives receive new messages (:205-206), see SURVEY.md section 7 restatement 1.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- pieces
def time_encode(t, w, b):
    """cos(t * w + b): reference model/time_encoding.py:17-25 (nn.Linear(1, d) then cos)."""
    return torch.cos(F.linear(t.unsqueeze(-1), w, b))


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.GRUCell arithmetic (reference modules/memory_updater.py:60):
    r,z = sigmoid(W_i{r,z} x + b + W_h{r,z} h + b); n = tanh(W_in x + b_in + r*(W_hn h + b_hn));
    h' = (1 - z) * n + z * h.  Gate rows are stacked [r; z; n]."""
    gi = F.linear(x, w_ih, b_ih)
    gh = F.linear(h, w_hh, b_hh)
    i_r, i_z, i_n = gi.chunk(3, dim=1)
    h_r, h_z, h_n = gh.chunk(3, dim=1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return n + z * (h - n)


def rnn_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.RNNCell (tanh) arithmetic (reference modules/memory_updater.py:68)."""
    return torch.tanh(F.linear(x, w_ih, b_ih) + F.linear(h, w_hh, b_hh))


def merge_layer(x1, x2, p, prefix):
    """fc2(relu(fc1([x1 | x2]))): reference utils/utils.py:4-17."""
    x = torch.cat([x1, x2], dim=1)
    h = torch.relu(F.linear(x, p[prefix + "fc1.weight"], p[prefix + "fc1.bias"]))
    return F.linear(h, p[prefix + "fc2.weight"], p[prefix + "fc2.bias"])


def temporal_attention(h_q, te_q, h_nbr, te_nbr, e_nbr, pad_mask, p, prefix, n_heads,
                       drop_mask=None):
    """One temporal attention layer: reference model/temporal_attention.py:34-90.

    h_q [Q,d], te_q [Q,dt], h_nbr [Q,n,d], te_nbr [Q,n,dt], e_nbr [Q,n,F], pad_mask bool [Q,n].
    Query = [h_q | te_q]; key = value = [h_nbr | e_nbr | te_nbr] (:51-52).  Rows whose
    neighbours are all padding get slot 0 un-masked (:60-65) and their attention output is
    zeroed before the merge MLP (:84).  Multi-head attention follows torch.nn.MultiheadAttention
    with separate q/k/v projection weights; `drop_mask` (optional [Q,H,n] of 0 or 1/(1-p))
    multiplies the softmax weights (deviation iii: parity cases use none)."""
    Q, n = pad_mask.shape
    q_in = torch.cat([h_q, te_q], dim=1)
    k_in = torch.cat([h_nbr, e_nbr, te_nbr], dim=2)
    E = q_in.shape[1]
    hd = E // n_heads
    b_in = p[prefix + "multi_head_target.in_proj_bias"]
    invalid = pad_mask.all(dim=1, keepdim=True)
    mask = pad_mask.clone()
    mask[invalid.squeeze(1), 0] = False
    q = F.linear(q_in, p[prefix + "multi_head_target.q_proj_weight"], b_in[:E])
    k = F.linear(k_in, p[prefix + "multi_head_target.k_proj_weight"], b_in[E:2 * E])
    v = F.linear(k_in, p[prefix + "multi_head_target.v_proj_weight"], b_in[2 * E:])
    q = q.view(Q, n_heads, hd) * (1.0 / math.sqrt(hd))
    k = k.view(Q, n, n_heads, hd)
    v = v.view(Q, n, n_heads, hd)
    s = torch.einsum("qhc,qnhc->qhn", q, k)
    s = s.masked_fill(mask.unsqueeze(1), float("-inf"))
    a = torch.softmax(s, dim=2)
    if drop_mask is not None:
        a = a * drop_mask
    o = torch.einsum("qhn,qnhc->qhc", a, v).reshape(Q, E)
    o = F.linear(o, p[prefix + "multi_head_target.out_proj.weight"],
                 p[prefix + "multi_head_target.out_proj.bias"])
    o = o.masked_fill(invalid, 0.0)
    return merge_layer(o, h_q, p, prefix + "merger.")


def bpr_loss(e_u, e_pos, e_neg):
    """-mean_b log sigmoid(mean_k(<e_u,e_pos> - <e_u,e_neg_k>)): reference main.py:321-337.
    e_u [B,d], e_pos [B,d], e_neg [B*k,d] interaction-major."""
    B, d = e_u.shape
    e_neg = e_neg.view(B, -1, d)
    pos = (e_u * e_pos).sum(dim=1, keepdim=True)
    neg = torch.einsum("bd,bkd->bk", e_u, e_neg)
    return -torch.log(torch.sigmoid((pos - neg).mean(dim=1))).mean()


def normalize_edge_features(edge_features):
    """z-normalisation over ALL rows including padding row 0: reference model/tgn.py:38-41."""
    ef = np.asarray(edge_features).astype(np.float32)
    ef = ef - ef.mean(axis=0)
    ef = ef / ef.std(axis=0)
    return ef.astype(np.float32)


# ----------------------------------------------------------------------------- model
class TGNOracle:
    """Dense-state restatement of reference model/tgn.py:15-382 and the modules it drives."""

    def __init__(self, params, adj, node_features, edge_features, n_layers=1, n_heads=2,
                 use_memory=True, memory_updater="gru", embedding="graph_attention",
                 dyrep=False, use_destination_embedding_in_message=False,
                 use_source_embedding_in_message=False, message_function="identity", aggregator="last",
                 mean_time_shift_src=0.0, std_time_shift_src=1.0,
                 mean_time_shift_dst=0.0, std_time_shift_dst=1.0, edge_features_normalized=False):
        self.p = params
        self.adj = adj
        self.node_feat = torch.as_tensor(np.asarray(node_features).astype(np.float32))
        ef = edge_features if edge_features_normalized else normalize_edge_features(edge_features)
        self.edge_feat = torch.as_tensor(np.asarray(ef, dtype=np.float32))
        self.n_nodes, self.d = self.node_feat.shape
        self.F = self.edge_feat.shape[1]
        self.n_layers, self.n_heads = n_layers, n_heads
        self.use_memory, self.updater, self.embedding = use_memory, memory_updater, embedding
        self.dyrep = dyrep
        self.dst_emb_in_msg = use_destination_embedding_in_message
        self.src_emb_in_msg = use_source_embedding_in_message
        self.message_function, self.aggregator = message_function, aggregator
        self.shift = (mean_time_shift_src, std_time_shift_src, mean_time_shift_dst, std_time_shift_dst)
        self.raw_dim = 2 * self.d + self.F + self.d
        self.call_id = 0
        self.nbr_seed = 0
        self.reset_memory()

    # reference modules/memory.py:23-33
    def reset_memory(self):
        N, d = self.n_nodes, self.d
        self.memory = torch.zeros(N, d)
        self.last_update = torch.zeros(N)
        self.pend_valid = torch.zeros(N, dtype=torch.bool)
        self.pend_msg = torch.zeros(N, self.raw_dim)
        self.pend_ts = torch.zeros(N)

    def _cell(self, x, h):
        pre = "memory_updater.memory_updater."
        f = gru_cell if self.updater == "gru" else rnn_cell
        return f(x, h, self.p[pre + "weight_ih"], self.p[pre + "weight_hh"],
                 self.p[pre + "bias_ih"], self.p[pre + "bias_hh"])

    def _te(self, t):
        return time_encode(t, self.p["time_encoder.w.weight"], self.p["time_encoder.w.bias"])

    # reference model/tgn.py:342-354 + modules/memory_updater.py:35-53 + message_aggregator.py:38-55
    def get_updated_memory(self):
        ids = torch.nonzero(self.pend_valid).squeeze(1)
        mem = self.memory.clone()
        lu = self.last_update.clone()
        if ids.numel() > 0:
            x = self.pend_msg[ids]
            if self.message_function == "mlp":                    # modules/message_function.py:13-26 (tgn.py:348)
                pre = "message_function.mlp."
                x = F.linear(torch.relu(F.linear(x, self.p[pre + "0.weight"], self.p[pre + "0.bias"])),
                             self.p[pre + "2.weight"], self.p[pre + "2.bias"])
            mem[ids] = self._cell(x, self.memory[ids])
            lu[ids] = self.pend_ts[ids]
        return mem, lu

    # reference modules/embedding_module.py:76-175 (recursion), :244-258, :57-61
    def _embed(self, memory, nodes, timestamps, n_layers, n_neighbors):
        nodes_t = torch.as_tensor(np.asarray(nodes, dtype=np.int64))
        ts32 = torch.as_tensor(np.asarray(timestamps, dtype=np.float64)).float().unsqueeze(1)
        te_q = self._te(torch.zeros_like(ts32))                       # [Q,1,d]
        h = self.node_feat[nodes_t]
        if self.use_memory:
            h = memory[nodes_t] + h
        if n_layers == 0:
            return h
        h_q = self._embed(memory, nodes, timestamps, n_layers - 1, n_neighbors)
        nbr, eidx, etime = self.adj.get_temporal_neighbor(nodes, timestamps, n_neighbors,
                                                          call_id=self.call_id, seed=self.nbr_seed)
        self.call_id += 1
        self.last_neighbors = (nbr, eidx, etime)
        deltas = np.asarray(timestamps, dtype=np.float64)[:, None] - etime      # fp64, then fp32
        deltas_t = torch.as_tensor(deltas).float()
        n_eff = n_neighbors if n_neighbors > 0 else 1
        h_n = self._embed(memory, nbr.flatten(), np.repeat(timestamps, n_eff), n_layers - 1,
                          n_neighbors).view(len(nodes), n_eff, -1)
        te_n = self._te(deltas_t)
        e_n = self.edge_feat[torch.as_tensor(eidx.astype(np.int64))]
        mask = torch.as_tensor(nbr == 0)
        if self.embedding == "graph_attention":
            prefix = "embedding_module.attention_models.%d." % (n_layers - 1)
            return temporal_attention(h_q, te_q.squeeze(1), h_n, te_n, e_n, mask, self.p, prefix,
                                      self.n_heads)
        if self.embedding == "graph_sum":                              # :205-219
            l1 = "embedding_module.linear_1.%d." % (n_layers - 1)
            l2 = "embedding_module.linear_2.%d." % (n_layers - 1)
            x = torch.cat([h_n, te_n, e_n], dim=2)
            s = torch.relu(F.linear(x, self.p[l1 + "weight"], self.p[l1 + "bias"]).sum(dim=1))
            x2 = torch.cat([s, h_q, te_q.squeeze(1)], dim=1)
            return F.linear(x2, self.p[l2 + "weight"], self.p[l2 + "bias"])
        raise ValueError(self.embedding)

    def _time_diffs(self, last_update, groups, edge_times):
        """reference model/tgn.py:145-153 / :260-266: int64 subtraction, then fp32 (x-mean)/std."""
        ms, ss, md, sd = self.shift
        out = []
        for gi, g in enumerate(groups):
            rep = len(g) // len(edge_times)
            t = torch.as_tensor(np.repeat(np.asarray(edge_times), rep)).long()
            diff = t - last_update[torch.as_tensor(np.asarray(g, dtype=np.int64))].long()
            mean, std = (ms, ss) if gi == 0 else (md, sd)
            out.append((diff - float(mean)) / float(std))
        return torch.cat(out, dim=0)

    # reference model/tgn.py:102-217 (with p_pos group) and :219-327
    def compute_temporal_embeddings(self, source_nodes, destination_nodes, extra_groups,
                                    edge_times, edge_idxs, n_neighbors):
        """extra_groups: list of 1-D id arrays, each k*B long and interaction-major
        ([negatives] for tgn.py:219, [p_pos, p_neg] for tgn.py:102).  Returns the list of
        embeddings [src, dst, *extra] and advances memory / pending messages."""
        src = np.asarray(source_nodes, dtype=np.int64)
        dst = np.asarray(destination_nodes, dtype=np.int64)
        groups = [src, dst] + [np.asarray(g, dtype=np.int64) for g in extra_groups]
        B = len(src)
        nodes = np.concatenate(groups)
        ts = np.asarray(edge_times)
        timestamps = np.concatenate([np.repeat(ts, len(g) // B) for g in groups])
        memory = last_update = None
        if self.use_memory:
            memory, last_update = self.get_updated_memory()
        if self.embedding == "time":
            td = self._time_diffs(last_update, groups, ts)
            w = self.p["embedding_module.embedding_layer.weight"]
            b = self.p["embedding_module.embedding_layer.bias"]
            emb = memory[torch.as_tensor(nodes)] * (1 + F.linear(td.unsqueeze(1), w, b))
        elif self.embedding == "identity":
            emb = memory[torch.as_tensor(nodes)]
        else:
            emb = self._embed(memory, nodes, timestamps, self.n_layers, n_neighbors)
        sizes = [len(g) for g in groups]
        outs = list(torch.split(emb, sizes))
        if self.use_memory:
            positives = torch.as_tensor(np.unique(np.concatenate([src, dst])))
            ids = positives[self.pend_valid[positives]]
            if ids.numel() > 0:                                   # tgn.py:185 persist, :191 clear
                self.memory[ids] = memory[ids].detach()
                self.last_update[ids] = self.pend_ts[ids]
                self.pend_valid[ids] = False
            eidx_t = torch.as_tensor(np.asarray(edge_idxs, dtype=np.int64))
            self._store_messages(src, outs[0], dst, outs[1], ts, eidx_t)
            self._store_messages(dst, outs[1], src, outs[0], ts, eidx_t)
            if self.dyrep:                                        # tgn.py:211-215 / :322-325
                outs = list(torch.split(memory[torch.as_tensor(nodes)], sizes))
        return outs

    # reference model/tgn.py:357-378 (+ last-wins store, memory.py:35-37 / aggregator :49-50)
    def _store_messages(self, a, a_emb, b, b_emb, edge_times, eidx_t):
        a_t, b_t = torch.as_tensor(a), torch.as_tensor(b)
        t32 = torch.as_tensor(np.asarray(edge_times)).float()
        m_a = a_emb.detach() if self.src_emb_in_msg else self.memory[a_t]
        m_b = b_emb.detach() if self.dst_emb_in_msg else self.memory[b_t]
        delta = t32 - self.last_update[a_t]
        msg = torch.cat([m_a, m_b, self.edge_feat[eidx_t], self._te(delta.unsqueeze(1)).view(len(a), -1)],
                        dim=1).detach()
        if self.aggregator == "mean":                             # message_aggregator.py:62-81: a node's list holds
            for node in np.unique(a):                             # exactly its messages of this batch (cleared at
                rows = np.nonzero(a == node)[0]                   # tgn.py:191 before they are appended)
                self.pend_msg[node] = torch.mean(torch.stack([msg[i] for i in rows]), dim=0)
                self.pend_ts[node] = t32[rows[-1]]
                self.pend_valid[node] = True
            return
        for i in range(len(a)):                                   # later occurrences win
            self.pend_msg[a[i]] = msg[i]
            self.pend_ts[a[i]] = t32[i]
            self.pend_valid[a[i]] = True


def eval_scores(e_src, e_dst, e_neg):
    """Scores of the positive and the candidates: reference evaluation.py:107-115.
    Returns float32 [B, 1+N]."""
    B, d = e_src.shape
    pos = (e_src * e_dst).sum(dim=1, keepdim=True)
    neg = (e_src.view(B, 1, d) * e_neg.view(B, -1, d)).sum(dim=2)
    return torch.cat([pos, neg], dim=1)


def eval_ranking(scores):
    """argsort(scores)[::-1] per row with a stable sort: reference evaluation.py:134-138."""
    s = np.asarray(scores)
    return np.argsort(s, axis=1, kind="stable")[:, ::-1]
