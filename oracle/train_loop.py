"""Oracle: the per-batch training / evaluation loop of the reference on CPU (test infrastructure
and the CPU baseline of bench.py -- never part of the product path).

Restates reference main.py:160-394 (one optimiser step per batch: candidate sampling, MV
selection, embeddings, BPR, backward, Adam, detach) and evaluation.py:63-138 (candidates,
embeddings, scores, ranking) on top of oracle.tgn / oracle.sampling / oracle.graph.
"""
import numpy as np
import torch

from .graph import AdjacencyOracle
from .sampling import sample_candidates, mv_select
from .tgn import TGNOracle, bpr_loss, eval_scores, eval_ranking


def init_params(d, F, model="ours", n_layers=1, seed=0):
    """Random parameters with the reference's state_dict names and shapes (values are not the
    reference's initial draws; parity tests inject weights instead)."""
    g = torch.Generator().manual_seed(seed)
    E, Ek, raw = 2 * d, 2 * d + F, 3 * d + F
    G = 3 if model in ("ours", "tgn", "tgat") else 1

    def rnd(*shape, s=0.1):
        return (torch.randn(*shape, generator=g) * s).requires_grad_(True)

    p = {"time_encoder.w.weight": torch.tensor(1 / 10 ** np.linspace(0, 9, d), dtype=torch.float32)
         .reshape(d, 1).requires_grad_(True),
         "time_encoder.w.bias": torch.zeros(d, requires_grad=True)}
    if model != "tgat":
        pre = "memory_updater.memory_updater."
        p.update({pre + "weight_ih": rnd(G * d, raw), pre + "weight_hh": rnd(G * d, d),
                  pre + "bias_ih": rnd(G * d), pre + "bias_hh": rnd(G * d)})
    if model == "jodie":
        p.update({"embedding_module.embedding_layer.weight": rnd(d, 1, s=1.0),
                  "embedding_module.embedding_layer.bias": rnd(d, s=1.0)})
    else:
        for l in range(n_layers):
            a = f"embedding_module.attention_models.{l}."
            p.update({a + "multi_head_target.q_proj_weight": rnd(E, E), a + "multi_head_target.k_proj_weight": rnd(E, Ek),
                      a + "multi_head_target.v_proj_weight": rnd(E, Ek), a + "multi_head_target.in_proj_bias": rnd(3 * E),
                      a + "multi_head_target.out_proj.weight": rnd(E, E), a + "multi_head_target.out_proj.bias": rnd(E),
                      a + "merger.fc1.weight": rnd(d, E + d), a + "merger.fc1.bias": rnd(d),
                      a + "merger.fc2.weight": rnd(d, d), a + "merger.fc2.bias": rnd(d)})
    return p


def time_statistics(sources, destinations, timestamps):
    """mean / std of the per-node inter-event times, sources and destinations separately, the first event of a
    node measured from 0 (reference utils/data.py:75-99; main.py passes them to the TGN constructor)."""
    out = []
    for ids in (sources, destinations):
        last, diffs = {}, np.empty(len(ids), dtype=np.float64)
        for k, (v, t) in enumerate(zip(ids.tolist(), timestamps.tolist())):
            diffs[k] = t - last.get(v, 0)
            last[v] = t
        out += [float(np.mean(diffs)), float(np.std(diffs))]
    return tuple(out)


class OracleTrainer:
    def __init__(self, st, model="ours", bs=512, d=64, n_neighbors=10, n_layers=1, lr=1e-4, num_negatives=20,
                 p_neg_num=3, gamma=2.0, lam=0.5, seed=0, params=None, train_mask=None):
        from pfotgnrec_b200.synth import log_returns      # the synthetic-data helper, not a product kernel
        self.st, self.model, self.bs, self.n, self.seed = st, model, bs, n_neighbors, seed
        self.K, self.p_neg_num, self.gamma, self.lam = num_negatives, p_neg_num, gamma, lam
        tm = st.split()[0] if train_mask is None else train_mask
        tr = np.nonzero(tm)[0]
        uniform = model == "tgat"
        self.adj_train = AdjacencyOracle(st.sources[tr], st.destinations[tr], st.edge_idxs[tr], st.timestamps[tr],
                                         n_nodes=st.n_nodes, uniform=uniform)
        self.adj_full = AdjacencyOracle(st.sources, st.destinations, st.edge_idxs, st.timestamps,
                                        n_nodes=st.n_nodes, uniform=uniform)
        F = st.edge_features.shape[1]
        self.p = params if params is not None else init_params(d, F, model, n_layers, seed)
        node_feat = np.random.RandomState(0).rand(st.n_nodes, d)
        kw = {}
        if model == "jodie":
            kw = dict(memory_updater="rnn", embedding="time")
        elif model == "dyrep":
            kw = dict(memory_updater="rnn", dyrep=True, use_destination_embedding_in_message=True)
        elif model == "tgat":
            kw = dict(use_memory=False)
        ms, ss, md, sd = time_statistics(st.sources, st.destinations, st.timestamps)
        self.tgn = TGNOracle(self.p, self.adj_train, node_feat, st.edge_features, n_layers=n_layers, n_heads=2,
                             mean_time_shift_src=ms, std_time_shift_src=ss, mean_time_shift_dst=md,
                             std_time_shift_dst=sd, **kw)
        self.tgn.nbr_seed = seed
        self.opt = torch.optim.Adam([v for v in self.p.values()], lr=lr)
        self.universe_items = np.unique(st.destinations[tr])
        self.logret = log_returns(st.prices_future)

    def _ports(self, s, e):
        ptr = self.st.port_ptr[s:e + 1]
        return ptr - ptr[0], self.st.port_items[ptr[0]:ptr[-1]]

    def train_step(self, s, e):
        st, U = self.st, self.st.n_users
        src, dst, ts, ei = st.sources[s:e], st.destinations[s:e], st.timestamps[s:e], st.edge_idxs[s:e]
        pptr, pstock = self._ports(s, e)
        self.tgn.adj = self.adj_train
        self.opt.zero_grad()
        if self.model == "ours":
            neg = sample_candidates(ei, self.universe_items, pptr, pstock.astype(np.int64) + U + 1, self.K, self.seed)
            cand = np.concatenate([(dst - U - 1)[:, None], neg - U - 1], axis=1)
            pp, pn = mv_select(self.logret, st.day_idx[s:e], cand, pptr, pstock, self.gamma, self.lam, 1, self.p_neg_num)
            outs = self.tgn.compute_temporal_embeddings(src, dst, [pp + U + 1, pn + U + 1], ts, ei, self.n)
            loss = bpr_loss(outs[0], outs[2], outs[3])
        else:
            neg = sample_candidates(ei, self.universe_items, pptr, pstock.astype(np.int64) + U + 1, self.p_neg_num,
                                    self.seed)
            outs = self.tgn.compute_temporal_embeddings(src, dst, [neg.ravel()], ts, ei, self.n)
            loss = bpr_loss(outs[0], outs[1], outs[2])
        if loss.requires_grad:
            loss.backward()
            self.opt.step()
        return float(loss.item())

    @torch.no_grad()
    def eval_step(self, s, e, n_items=None):
        st, U = self.st, self.st.n_users
        src, dst, ts, ei = st.sources[s:e], st.destinations[s:e], st.timestamps[s:e], st.edge_idxs[s:e]
        pptr, pstock = self._ports(s, e)
        items = np.unique(st.destinations)
        N = len(items) if n_items is None else n_items
        self.tgn.adj = self.adj_full
        cand = sample_candidates(np.arange(e - s), items, pptr, pstock.astype(np.int64) + U + 1, N, 2024)
        outs = self.tgn.compute_temporal_embeddings(src, dst, [cand.ravel()], ts, ei, self.n)
        scores = eval_scores(outs[0], outs[1], outs[2]).numpy()
        rk = eval_ranking(scores)
        return np.argmax(rk == 0, axis=1), rk[:, :5], cand, scores
