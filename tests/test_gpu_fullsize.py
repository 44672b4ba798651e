"""Parity at BASELINE.json's full single-GPU size (100 000 users x 1 000 stocks x 5 000 000 events, bs 8 192), where
the CPU oracle is too slow to replay everything: size-independent properties of each kernel's output (sortedness,
set equality, idempotence, batch-split independence, conservation) on the whole output, plus the oracle on a random
sample of it."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
BS = 8192


@pytest.fixture(scope="module")
def big():
    from pfotgnrec_b200.synth import make_stream
    from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig
    st = make_stream(n_users=100000, n_items=1000, n_events=5000000, n_days=200, seed=0, ts_mode="nbg")
    tr = PfoTrainer(st, TrainConfig(model="ours", bs=BS, cuda_graph=False), device="cuda:0")
    return st, tr


def test_full_size_neighbor_sampling_properties(big):
    """K1 over the full adjacency (10 M entries), 200 000 queries: right-aligned rows, times ascending and strictly
    before the query time, ids / edge ids / times consistent with the CSR, row length = min(n, #earlier entries)
    (checked against numpy searchsorted for every query), and a 2 000-query sample bit-exact against the oracle."""
    from oracle.graph import AdjacencyOracle
    st, tr = big
    csr = tr.csr_full
    rng = np.random.default_rng(1)
    Q, n = 200000, 10
    nodes = rng.integers(0, st.n_nodes, size=Q)
    nodes[:50000] = st.destinations[rng.integers(0, st.n_events, size=50000)]          # hot items: degree ~ thousands
    ts = st.timestamps[rng.integers(0, st.n_events, size=Q)].copy()
    ts[::3] += 1.0
    nb, ei, et = tr.nf_full.get_temporal_neighbor(nodes, ts, n)
    live = nb != 0
    assert (np.diff(live.astype(np.int8), axis=1) >= 0).all()                           # padding only on the left
    assert (et[~live] == 0).all() and (ei[~live] == 0).all()
    et64 = et.astype(np.float64)
    both = live[:, 1:] & live[:, :-1]
    assert (np.diff(et64, axis=1)[both] >= 0).all()                                     # most recent last
    rowptr, adj_ts = csr.rowptr.cpu().numpy(), csr.ts.cpu().numpy()
    adj_nbr, adj_eidx = csr.nbr.cpu().numpy(), csr.eidx.cpu().numpy()
    cnt = np.empty(Q, dtype=np.int64)
    for q in range(Q):                                                                  # utils/utils.py:158 per query
        lo, hi = rowptr[nodes[q]], rowptr[nodes[q] + 1]
        cnt[q] = np.searchsorted(adj_ts[lo:hi], ts[q], side="left")
    assert np.array_equal(live.sum(axis=1), np.minimum(cnt, n))
    # every returned slot is the CSR entry at its position: entry k of the row = adjacency[lo + cnt - len + k]
    q_idx, k_idx = np.nonzero(live)
    pos = rowptr[nodes[q_idx]] + cnt[q_idx] - live.sum(axis=1)[q_idx] + (k_idx - (n - live.sum(axis=1)[q_idx]))
    assert np.array_equal(nb[q_idx, k_idx], adj_nbr[pos]) and np.array_equal(ei[q_idx, k_idx], adj_eidx[pos])
    assert np.array_equal(et[q_idx, k_idx], adj_ts[pos].astype(np.float32))
    assert (adj_ts[pos] < ts[q_idx]).all()
    sample = rng.choice(Q, size=2000, replace=False)
    adj = AdjacencyOracle(st.sources, st.destinations, st.edge_idxs, st.timestamps, n_nodes=st.n_nodes)
    o = adj.get_temporal_neighbor(nodes[sample], ts[sample], n)
    assert np.array_equal(nb[sample], o[0]) and np.array_equal(ei[sample], o[1]) and np.array_equal(et[sample], o[2])


def test_full_size_compaction_is_sorted_unique_and_invertible(big):
    from pfotgnrec_b200 import _lib
    from pfotgnrec_b200._lib import ptr
    st, tr = big
    state = tr.tgn.memory.state
    rng = np.random.default_rng(2)
    ids = torch.as_tensor(rng.integers(0, st.n_nodes, size=6 * BS * 11).astype(np.int32), device=DEV)
    _lib.call("pfo_mark_nodes", ptr(ids), ids.numel(), 1, ptr(state.bitmap))
    uniq = torch.zeros(min(ids.numel(), st.n_nodes), dtype=torch.int32, device=DEV)
    _lib.call("pfo_compact_nodes", ptr(state.bitmap), st.n_nodes, ptr(state.compact_ws), ptr(uniq),
              ptr(state.slot_of_node), ptr(state.n_unique))
    nu = int(state.n_unique.item())
    u = uniq[:nu].cpu().numpy()
    want = np.unique(ids.cpu().numpy())
    want = want[want != 0]
    assert np.array_equal(u, want)                                                       # ascending, no duplicates
    assert np.array_equal(state.slot_of_node[uniq[:nu].long()].cpu().numpy(), np.arange(nu))
    assert int(state.bitmap.abs().sum().item()) == 0                                     # left clean for the next batch


def test_full_size_mv_selection_properties_and_batch_split_independence(big):
    """K5 on a full batch: candidates distinct, outside the held portfolio, inside the training universe; the selected
    positive / negatives are candidates; the draw depends on the interaction only -- one launch over 8 192
    interactions equals two launches over its halves (the Philox stream is keyed by the interaction id) -- and a
    512-interaction sample is bit-exact against the oracle (candidates, fp64 scores, selected ids)."""
    from oracle import sampling
    st, tr = big
    D = tr.dev_stream
    s, e = 2000000, 2000000 + BS
    sel = lambda a, b: tr.mv.select(D.ev[a:b], D.day[a:b], D.dst[a:b], D.port_ptr[a:b + 1], D.port_items, return_scores=True)
    pp, pn, cand, y = sel(s, e)
    h = (s + e) // 2
    pp1, pn1, cand1, y1 = sel(s, h)
    pp2, pn2, cand2, y2 = sel(h, e)
    assert torch.equal(torch.cat([pp1, pp2]), pp) and torch.equal(torch.cat([pn1, pn2]), pn)
    assert torch.equal(torch.cat([cand1, cand2]), cand) and torch.equal(torch.cat([y1, y2]), y)
    c = cand.cpu().numpy()
    U = st.n_users
    assert np.array_equal(c[:, 0], st.destinations[s:e] - U - 1)
    srt = np.sort(c[:, 1:], axis=1)
    assert (np.diff(srt, axis=1) > 0).all()                                              # without replacement
    universe = tr.universe_items - U - 1
    assert np.isin(c[:, 1:], universe).all()
    for b in range(0, BS, 7):                                                            # not held (utils/utils.py:96)
        assert not np.isin(c[b, 1:], st.portfolio(s + b)).any()
    ppn, pnn = pp.cpu().numpy() - U - 1, pn.cpu().numpy().reshape(BS, 3) - U - 1
    assert (ppn[:, None] == c).any(axis=1).all()
    assert all((pnn[:, k][:, None] == c).any(axis=1).all() for k in range(3))
    B = 512
    pptr = st.port_ptr[s:s + B + 1]
    neg = sampling.sample_candidates(st.edge_idxs[s:s + B], tr.universe_items, pptr - pptr[0],
                                     st.port_items[pptr[0]:pptr[-1]].astype(np.int64) + U + 1, 20, tr.tc.seed)
    oc = np.concatenate([(st.destinations[s:s + B] - U - 1)[:, None], neg - U - 1], axis=1)
    assert np.array_equal(c[:B], oc)
    lr = tr.mv.logret.cpu().numpy()
    oy = sampling.mv_scores(lr, st.day_idx[s:s + B], oc, pptr - pptr[0], st.port_items[pptr[0]:pptr[-1]], tr.tc.gamma)
    assert np.array_equal(y.cpu().numpy()[:B], oy)
    opp, opn = sampling.mv_select(lr, st.day_idx[s:s + B], oc, pptr - pptr[0], st.port_items[pptr[0]:pptr[-1]],
                                  tr.tc.gamma, tr.tc.lambda_mv)
    assert np.array_equal(ppn[:B], opp) and np.array_equal(pnn[:B].ravel(), opn)


def test_full_size_step_state_update_is_last_wins_and_conservative(big):
    """One full-size training step: exactly the batch's positives hold a pending message afterwards (plus whatever
    was pending before), each with the time of its LAST interaction in the batch; rows of nodes outside the batch
    are untouched (memory, last_update, pending table); the loss is finite."""
    st, tr = big
    state = tr.tgn.memory.state
    s, e = 2100000, 2100000 + BS
    tr.train_step(s - BS, s)                                                             # some state to begin with
    before = [t.clone() for t in (state.memory, state.last_update, state.pend_msg, state.pend_ts, state.pend_valid)]
    loss = float(tr.train_step(s, e).item())
    assert np.isfinite(loss)
    nodes = np.concatenate([st.sources[s:e], st.destinations[s:e]])
    t32 = np.concatenate([st.timestamps[s:e], st.timestamps[s:e]]).astype(np.float32)
    last = {}
    for nd, t in zip(nodes, t32):                                                        # later occurrences win
        last[int(nd)] = t
    pos = np.fromiter(last.keys(), dtype=np.int64)
    pv, pt = state.pend_valid.cpu().numpy().astype(bool), state.pend_ts.cpu().numpy()
    assert pv[pos].all()
    assert np.array_equal(pt[pos], np.fromiter(last.values(), dtype=np.float32))
    assert np.array_equal(pv, before[4].cpu().numpy().astype(bool) | np.isin(np.arange(st.n_nodes), pos))
    other = np.ones(st.n_nodes, dtype=bool)
    other[pos] = False
    om = torch.as_tensor(other, device=DEV)
    for now, was in zip((state.memory, state.last_update, state.pend_msg, state.pend_ts), before[:4]):
        assert torch.equal(now[om], was[om])
    # positives that had a pending message were persisted: last_update = that message's time
    had = before[4].cpu().numpy().astype(bool)[pos]
    assert np.array_equal(state.last_update.cpu().numpy()[pos][had], before[3].cpu().numpy()[pos][had])


def test_full_size_eval_step_metric_invariants(big):
    """Full ranking of 512 users against all 1 000 stocks: rank / top-5 consistent with the scores, metric block
    invariants (recall@1 <= @3 <= @5, ndcg <= recall, ndcg@1 == recall@1), running sums == column sums, and a
    64-user sample of the per-user table bit-exact against the numpy oracle."""
    from oracle import eval_metrics as em
    st, tr = big
    s, e = 4600000, 4600000 + 512
    tr.reset_eval_metrics()
    pos_rank, top, cand, scores, pe = (t.cpu().numpy() for t in tr.eval_step(s, e))
    N = cand.shape[1]
    assert N == len(np.unique(st.destinations))
    rk = np.argsort(scores, axis=1, kind="stable")[:, ::-1]
    assert np.array_equal(top, rk[:, :5]) and np.array_equal(pos_rank, np.argmax(rk == 0, axis=1))
    assert (pe[:, 0] <= pe[:, 1]).all() and (pe[:, 1] <= pe[:, 2]).all()
    assert (pe[:, 3:6] <= pe[:, 0:3]).all() and np.array_equal(pe[:, 3], pe[:, 0])
    acc = tr.eval_acc.cpu().numpy()
    assert acc[30] == 512 and np.allclose(acc[:18], pe.sum(axis=0), rtol=1e-12, atol=1e-12)
    assert np.array_equal(acc[18:30], (pe[:, 6:] > 0).sum(axis=0))
    B, U = 64, st.n_users
    pp = st.port_ptr[s:s + B + 1]
    ref, _, _ = em.per_event_metrics(scores[:B], st.destinations[s:s + B] - U - 1, cand[:B] - U - 1, st.day_idx[s:s + B],
                                     pp - pp[0], st.port_items[pp[0]:pp[-1]], st.prices_past, st.prices_future)
    assert np.array_equal(pe[:B], ref)
