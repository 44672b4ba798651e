"""Node-sharded multi-GPU path: one process per GPU, NCCL all-to-all.

The reference is single-process; this module ADDS the sharding the north_star asks for:

* owner(x) = x mod G.  Each rank owns, for its nodes only: memory, last_update, the pending
  message table and the CSR adjacency rows (stored under the local index x // G).  Node and
  edge features, the price table and the weights are replicated (static / tiny).
* The global batch is split by event index (rank r takes the r-th slice); per step the ranks
  exchange, with `all_to_all_single`:
    R1  (node, t) queries -> owners, K1 on the local CSR rows, neighbour lists back;
    R2  unique touched node ids -> owners, lazy GRU/RNN there (de-duplicated over all
        requesters), updated-memory rows + last_update' back;
    R3  (training) d(loss)/d(row) back to the owners, cell backward there;
    R4  message rows built where the events live -> owners, persist + last-wins there, the
        winner chosen by GLOBAL batch position so the result does not depend on G;
  followed by one all-reduce of the ~133 k parameter gradients.
* Every quantity is identical to the 1-GPU path on the same global batch up to fp32 summation
  order (tools/check_sharded.py asserts it on 2 GPUs).

`Router` is pure torch + torch.distributed, so its bucket logic is covered by world_size-2
gloo tests on CPU (tests/test_dist_router.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import ptr
from .engine import TGNEngine, TGNState, ModelConfig, _linear, _wgrad
from .graph import TemporalCSR, NeighborFinder


class Plan:
    __slots__ = ("order", "send", "recv", "n_in", "n_out")


class Router:
    """Variable-size bucket exchange: rows of x go to rank dest[i]; `backward` returns replies
    to the original row order."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def plan(self, dest: torch.Tensor) -> Plan:
        p = Plan()
        dest = dest.long()
        p.order = torch.sort(dest, stable=True).indices
        send = torch.bincount(dest, minlength=self.world)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        p.send, p.recv = send.tolist(), recv.tolist()          # host sync: split sizes live on the host
        p.n_in, p.n_out = int(dest.shape[0]), int(sum(p.recv))
        return p

    def forward(self, p: Plan, x: torch.Tensor) -> torch.Tensor:
        xs = x.index_select(0, p.order).contiguous()
        out = torch.empty((p.n_out,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_to_all_single(out, xs, output_split_sizes=p.recv, input_split_sizes=p.send, group=self.group)
        return out

    def backward(self, p: Plan, y: torch.Tensor) -> torch.Tensor:
        ys = torch.empty((p.n_in,) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device)
        dist.all_to_all_single(ys, y.contiguous(), output_split_sizes=p.send, input_split_sizes=p.recv,
                               group=self.group)
        out = torch.empty_like(ys)
        out.index_copy_(0, p.order, ys)
        return out


def local_csr(st_sources, st_destinations, st_edge_idxs, st_timestamps, n_nodes, rank, world, device):
    """CSR rows of the nodes owned by `rank`, indexed by x // world; neighbour ids stay global."""
    src = np.asarray(st_sources, dtype=np.int64)
    dst = np.asarray(st_destinations, dtype=np.int64)
    E = src.shape[0]
    node = np.stack([src, dst], axis=1).ravel()
    other = np.stack([dst, src], axis=1).ravel()
    eid = np.repeat(np.asarray(st_edge_idxs, dtype=np.int64), 2)
    ts = np.repeat(np.asarray(st_timestamps, dtype=np.float64), 2)
    mine = np.nonzero(node % world == rank)[0]
    n_local = (n_nodes + world - 1) // world
    ln = node[mine] // world
    order = np.lexsort((np.arange(mine.shape[0]), ts[mine], ln))
    csr = TemporalCSR.__new__(TemporalCSR)
    csr.n_nodes, csr.n_events, csr.device = n_local, E, torch.device(device)
    rowptr = np.zeros(n_local + 1, dtype=np.int64)
    np.cumsum(np.bincount(ln, minlength=n_local), out=rowptr[1:])
    csr.rowptr = torch.as_tensor(rowptr, device=device)
    csr.nbr = torch.as_tensor(other[mine][order].astype(np.int32), device=device)
    csr.eidx = torch.as_tensor(eid[mine][order].astype(np.int32), device=device)
    csr.ts = torch.as_tensor(ts[mine][order], device=device)
    return csr


class ShardedNeighborFinder:
    """K1 over the owners' CSR rows: exchange R1."""

    def __init__(self, local_nf: NeighborFinder, router: Router):
        if local_nf.uniform:
            raise NotImplementedError("uniform neighbour sampling is single-GPU only in this round")
        self.local, self.router, self.uniform = local_nf, router, False

    def sample(self, q_nodes, q_ts, n_neighbors, out=None):
        G, R = self.router.world, q_nodes.shape[0]
        n = max(int(n_neighbors), 1)
        plan = self.router.plan(q_nodes % G)
        req = torch.empty(R, 3, dtype=torch.int32, device=q_nodes.device)
        req[:, 0] = torch.div(q_nodes, G, rounding_mode="floor")
        req[:, 1:3] = q_ts.contiguous().view(torch.int32).view(R, 2)
        got = self.router.forward(plan, req)
        ql = got[:, 0].contiguous()
        qt = got[:, 1:3].contiguous().view(torch.float64).view(-1)
        nbr, eidx, _et, dt = self.local.sample(ql, qt, n)
        reply = torch.cat([nbr, eidx, dt.view(torch.int32)], dim=1)
        back = self.router.backward(plan, reply)
        return (back[:, :n].contiguous(), back[:, n:2 * n].contiguous(), None,
                back[:, 2 * n:].contiguous().view(torch.float32))


class _Scratch:
    """Compaction scratch over the GLOBAL id space (requester side)."""

    def __init__(self, n_nodes, device):
        self.bitmap = torch.zeros((n_nodes + 31) // 32, dtype=torch.int32, device=device)
        self.slot_of_node = torch.zeros(n_nodes, dtype=torch.int32, device=device)
        self.compact_ws = torch.zeros(int(_lib.query("pfo_compact_workspace_ints", n_nodes)), dtype=torch.int32,
                                      device=device)
        self.n_unique = torch.zeros(1, dtype=torch.int32, device=device)


class ShardedEngine(TGNEngine):
    """TGNEngine whose node table, its backward and the message store go through the owners."""

    def __init__(self, cfg: ModelConfig, node_feat, edge_feat, local_nf, router: Router, n_nodes_global):
        if cfg.message_fn != "identity" or cfg.aggregator != "last" or cfg.src_emb_in_msg:
            raise NotImplementedError("the node-sharded engine covers the models main.py builds: identity message "
                                      "function, `last` aggregator, memory rows in the messages")
        self.router, self.G, self.rank = router, router.world, router.rank
        self.n_global = n_nodes_global
        n_local = (n_nodes_global + self.G - 1) // self.G
        state = TGNState(n_local, cfg, node_feat.device)
        super().__init__(cfg, state, node_feat, edge_feat, ShardedNeighborFinder(local_nf, router))
        self.req = _Scratch(n_nodes_global, node_feat.device)

    # requester-side unique ids over the global id space (exact count: one host sync)
    def _unique_global(self, id_lists):
        r = self.req
        total = 0
        for ids in id_lists:
            _lib.call("pfo_mark_nodes", ptr(ids), ids.numel(), 1, ptr(r.bitmap))
            total += ids.numel()
        cap = min(total, self.n_global)
        uniq = torch.zeros(cap, dtype=torch.int32, device=self.device)
        _lib.call("pfo_compact_nodes", ptr(r.bitmap), self.n_global, ptr(r.compact_ws), ptr(uniq),
                  ptr(r.slot_of_node), ptr(r.n_unique))
        U = int(r.n_unique.item())
        return uniq[:U].contiguous(), U

    def node_table(self, id_lists, cellW):
        c, st, dev, G = self.cfg, self.state, self.device, self.G
        d = c.d
        uniq, U = self._unique_global(id_lists)
        self.slot_map = self.req.slot_of_node
        n_uniq = torch.tensor([U], dtype=torch.int32, device=dev)
        nf_rows = self.node_feat.index_select(0, uniq.long())
        if not c.use_memory:
            return dict(uniq=uniq, u_max=U, n_uniq=n_uniq, H0=nf_rows, Hnew=None, lu_u=None)
        # R2: ids -> owners
        plan = self.router.plan(uniq % G)
        got = self.router.forward(plan, torch.div(uniq, G, rounding_mode="floor").view(-1, 1))[:, 0].contiguous()
        R = got.shape[0]
        _lib.call("pfo_mark_nodes", ptr(got), R, 0, ptr(st.bitmap))
        u_own = min(max(R, 1), st.n_nodes)
        uo = torch.zeros(u_own, dtype=torch.int32, device=dev)
        _lib.call("pfo_compact_nodes", ptr(st.bitmap), st.n_nodes, ptr(st.compact_ws), ptr(uo),
                  ptr(st.slot_of_node), ptr(st.n_unique))
        n_own = st.n_unique.clone()
        Gd = c.gates * d
        HG = torch.empty(u_own, d, device=dev)
        XG = torch.empty(u_own, c.rawp, device=dev)
        valid_u = torch.empty(u_own, dtype=torch.uint8, device=dev)
        lu_own = torch.zeros(u_own, device=dev)
        _lib.call("pfo_gather_state", ptr(uo), ptr(n_own), u_own, d, c.raw, ptr(st.memory), ptr(st.pend_msg),
                  c.rawp, ptr(st.pend_valid), ptr(st.pend_ts), ptr(st.last_update),
                  ptr(HG), ptr(XG), ptr(valid_u), ptr(lu_own))
        W_ih, W_hh, b_ih, b_hh = cellW
        GI = torch.empty(u_own, Gd, device=dev)
        GH = torch.empty(u_own, Gd, device=dev)
        _linear(c, ptr(XG), c.rawp, None, ptr(W_ih), c.raw, 0, ptr(b_ih), ptr(GI), Gd, u_own, Gd, c.raw, m_dev=ptr(n_own))
        _linear(c, ptr(HG), d, None, ptr(W_hh), d, 0, ptr(b_hh), ptr(GH), Gd, u_own, Gd, d, m_dev=ptr(n_own))
        Hnew_own = torch.zeros(u_own, d, device=dev)
        scratch = torch.empty(u_own, d, device=dev)         # H0 is formed on the requester
        _lib.call("pfo_cell_forward", ptr(uo), ptr(n_own), u_own, d, c.cell, ptr(GI), ptr(GH), Gd, ptr(HG),
                  ptr(valid_u), ptr(self.node_feat), ptr(Hnew_own), ptr(scratch))
        slots_own = torch.empty(R, dtype=torch.int32, device=dev)
        _lib.call("pfo_map_slots", ptr(got), R, 0, ptr(st.slot_of_node), ptr(slots_own))
        reply = torch.empty(R, d + 1, device=dev)
        _lib.call("pfo_gather_rows", ptr(Hnew_own), d, ptr(slots_own), R, d, ptr(reply), d + 1)
        reply[:, d] = lu_own.index_select(0, slots_own.long())
        back = self.router.backward(plan, reply)
        Hnew = back[:, :d].contiguous()
        lu_u = back[:, d].contiguous()
        own = dict(uniq=uo, u_max=u_own, n_uniq=n_own, HG=HG, XG=XG, valid_u=valid_u, GI=GI, GH=GH,
                   Hnew=Hnew_own, slots=slots_own, R=R)
        return dict(uniq=uniq, u_max=U, n_uniq=n_uniq, H0=Hnew + nf_rows, Hnew=Hnew, lu_u=lu_u, plan=plan, own=own)

    def node_table_backward(self, tab, dH0, g_cell):
        # R3: gradient rows -> owners, summed over requesters, then the cell backward of the base class
        own = tab["own"]
        got = self.router.forward(tab["plan"], dH0)
        d = self.cfg.d
        dH_own = torch.zeros(own["u_max"], d, device=self.device)
        _lib.call("pfo_scatter_add_rows", ptr(got), d, ptr(own["slots"]), own["R"], d, ptr(dH_own), d)
        TGNEngine.node_table_backward(self, own, dH_own, g_cell)

    def persist_and_store(self, tab, batch, emb, tw, tb):
        # R4: rows built here, applied at the owners with last-wins by global batch position
        c, st, dev, G = self.cfg, self.state, self.device, self.G
        d, F, B = c.d, c.n_edge_feat, batch["B"]
        src, dst = batch["src"], batch["dst"]
        s_slot, d_slot = self._slots(src), self._slots(dst)
        rows = torch.empty(2 * B, c.raw, device=dev)
        t32 = torch.empty(2 * B, device=dev)
        o_src = o_dst = None
        if c.dst_emb_in_msg:
            o_src, o_dst = emb[B:2 * B].contiguous(), emb[:B].contiguous()
        _lib.call("pfo_build_messages", ptr(s_slot), ptr(d_slot), ptr(batch["eidx"]), ptr(batch["ts"]), B, d, F,
                  ptr(tab["Hnew"]), ptr(tab["lu_u"]), ptr(self.edge_feat), ptr(tw), ptr(tb), ptr(o_src), ptr(o_dst),
                  ptr(rows), c.raw, ptr(t32))
        nodes = torch.cat([src, dst])
        Bg = B * G
        ar = torch.arange(B, dtype=torch.int32, device=dev) + self.rank * B
        key = torch.cat([ar, ar + Bg])
        meta = torch.stack([torch.div(nodes, G, rounding_mode="floor"), key, t32.view(torch.int32)], dim=1)
        plan = self.router.plan(nodes % G)
        got_meta = self.router.forward(plan, meta)
        got_rows = self.router.forward(plan, rows)
        R = got_meta.shape[0]
        own = tab["own"]
        node_l = got_meta[:, 0].contiguous()
        key_r = got_meta[:, 1].contiguous()
        t_r = got_meta[:, 2].contiguous().view(torch.float32)
        _lib.call("pfo_apply_messages", ptr(node_l), ptr(key_r), R, d, c.raw, ptr(st.slot_of_node), ptr(own["Hnew"]),
                  ptr(got_rows), c.raw, ptr(t_r), ptr(st.memory), ptr(st.last_update), ptr(st.pend_msg), c.rawp,
                  ptr(st.pend_ts), ptr(st.pend_valid), ptr(st.last_pos))


class ShardedTrainer:
    """PfoTrainer over G ranks: global batch [s, s + bs*G), rank r trains on its r-th slice."""

    def __init__(self, st, tc, device, rank, world):
        from .trainer import PfoTrainer, StreamOnDevice, load_overlay
        from .sampler import MVSelector, CandidateSampler
        from .synth import log_returns
        self.st, self.tc, self.device, self.rank, self.world = st, tc, torch.device(device), rank, world
        self.router = Router()
        tgn_mod, _ = load_overlay()
        train_mask = st.split()[0]
        tr = np.nonzero(train_mask)[0]
        csr = local_csr(st.sources[tr], st.destinations[tr], st.edge_idxs[tr], st.timestamps[tr], st.n_nodes,
                        rank, world, device)
        node_feat = np.random.RandomState(0).rand(st.n_nodes, tc.d)
        kw = dict(memory_updater_type="gru", embedding_module_type="graph_attention", use_memory=True)
        if tc.model == "jodie":
            kw.update(memory_updater_type="rnn", embedding_module_type="time")
        elif tc.model == "tgat":
            kw.update(use_memory=False)
        ms, ss, md, sd = PfoTrainer._time_statistics(self)
        torch.manual_seed(tc.seed)                        # identical initial weights on every rank
        self.tgn = tgn_mod.TGN(neighbor_finder=None, node_features=node_feat, edge_features=st.edge_features.copy(),
                               device=self.device, n_layers=tc.n_layers, n_heads=tc.n_heads, dropout=tc.dropout,
                               message_dimension=100, memory_dimension=tc.d, message_function="identity",
                               aggregator_type="last", n_neighbors=tc.n_neighbors, mean_time_shift_src=ms,
                               std_time_shift_src=ss, mean_time_shift_dst=md, std_time_shift_dst=sd,
                               gemm_mode=tc.gemm_mode, **kw).to(self.device)
        self.engine = ShardedEngine(self.tgn._cfg, self.tgn.node_raw_features, self.tgn.edge_raw_features,
                                    NeighborFinder(csr), self.router, st.n_nodes)
        self.opt = torch.optim.Adam(self.tgn.parameters(), lr=tc.lr, fused=True)
        self.dev_stream = StreamOnDevice(st, device)
        universe_items = np.unique(st.destinations[tr])
        self.mv = None
        if tc.model == "ours":
            self.mv = MVSelector(log_returns(st.prices_future), universe_items - st.n_users - 1, st.n_users,
                                 gamma=tc.gamma, lam=tc.lambda_mv, n_candidates=tc.num_negatives,
                                 n_pos=tc.p_pos_num, n_neg=tc.p_neg_num, seed=tc.seed, device=device)
        self.neg_sampler = CandidateSampler(universe_items, device=device)
        self.bpr_ws = torch.empty(1024, device=self.device)
        self.params = [p for p in self.tgn.parameters() if p.requires_grad]

    def train_step(self, s, e):
        from .trainer import bpr_loss
        tc, D, G = self.tc, self.dev_stream, self.world
        bs = (e - s) // G
        ls, le = s + self.rank * bs, s + (self.rank + 1) * bs
        b = dict(src=D.src[ls:le], dst=D.dst[ls:le], ts=D.ts[ls:le], eidx=D.eidx[ls:le], ev=D.ev[ls:le],
                 day=D.day[ls:le], port_ptr=D.port_ptr[ls:le + 1])
        self.tgn.train()
        self.opt.zero_grad(set_to_none=True)
        params = self.tgn._params()
        if tc.model == "ours":
            p_pos, p_neg = self.mv.select(b["ev"], b["day"], b["dst"], b["port_ptr"], D.port_items)
            e_s, _, e_p, e_n = self.engine.compute_temporal_embeddings(params, b["src"], b["dst"], [p_pos, p_neg],
                                                                       b["ts"], b["eidx"], tc.n_neighbors, train=True)
        else:
            neg = self.neg_sampler.sample(b["ev"], b["port_ptr"], D.port_items_as_item_ids, tc.p_neg_num,
                                          seed=tc.seed).reshape(-1)
            e_s, e_p, e_n = self.engine.compute_temporal_embeddings(params, b["src"], b["dst"], [neg], b["ts"],
                                                                    b["eidx"], tc.n_neighbors, train=True)
        loss = bpr_loss(e_s, e_p, e_n, self.bpr_ws)
        loss.backward()
        # one all-reduce of the parameter gradients (loss is the mean over the GLOBAL batch)
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat)
        flat.div_(G)
        off = 0
        for p, g in zip(self.params, grads):
            n = g.numel()
            p.grad = flat[off:off + n].view_as(p)
            off += n
        self.opt.step()
        return loss.detach()

    def gather_memory(self):
        """Global [N, d] memory / last_update / pend_valid assembled on every rank (tests)."""
        st, G = self.engine.state, self.world
        outs = []
        for t in (st.memory, st.last_update, st.pend_valid.to(torch.float32)):
            parts = [torch.empty_like(t) for _ in range(G)]
            dist.all_gather(parts, t.contiguous())
            full = torch.stack(parts, dim=1).reshape((-1,) + tuple(t.shape[1:]))[:self.st.n_nodes]
            outs.append(full)
        return outs
