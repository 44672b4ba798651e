"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):      python -m oracle.make_golden
The reference modules are imported from /root/reference (never copied); the MV-selection
block, which is inline script code (reference main.py:209-292), is executed by slicing the
reference file's own text at run time.  Documented deviations applied here: stable argsort
for the rank fusion (ii) and dropout = 0 (iii).  Inputs are synthetic (pfotgnrec_b200.synth).
Re-running the script reproduces every integer / index array and every forward quantity bit for bit; a handful
of GRU parameter gradients move by ~1e-7 relative between runs (threaded reductions inside CPU torch's backward),
two orders of magnitude below the tolerance the tests apply to gradients.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
from stage_reference import ref_root      # noqa: E402

# /root/reference in the build container (the committed goldens come from there); the staged byte-for-byte copy
# baseline/_ref on the GPU box (only `--device cuda`, the side-by-side run of tests/test_gpu_reference.py, uses it)
REF = ref_root() or "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def _import_reference():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from model.tgn import TGN                       # noqa: E402
    from utils.utils import get_neighbor_finder, RandEdgeSampler   # noqa: E402
    return TGN, get_neighbor_finder, RandEdgeSampler


def _data(stream, sl=slice(None)):
    return types.SimpleNamespace(sources=stream.sources[sl], destinations=stream.destinations[sl],
                                 edge_idxs=stream.edge_idxs[sl], timestamps=stream.timestamps[sl])


def golden_neighbors(get_neighbor_finder):
    from pfotgnrec_b200.synth import make_stream
    out = {}
    for name, mode in (("small", "small"), ("nbg", "nbg")):
        st = make_stream(n_users=50, n_items=10, n_events=600, n_days=20, seed=3, ts_mode=mode,
                         with_prices=False)
        nf = get_neighbor_finder(_data(st), uniform=False, max_node_idx=st.n_nodes - 1)
        rng = np.random.default_rng(5)
        Q = 300
        nodes = rng.integers(0, st.n_nodes, size=Q)
        # half of the cut times coincide with event times (ties), the rest fall in between
        ts = st.timestamps[rng.integers(0, st.n_events, size=Q)].copy()
        ts[::2] += rng.integers(-2, 3, size=ts[::2].shape[0])
        ts[:5] = [0, st.timestamps[0], st.timestamps[-1] + 1, -5, st.timestamps[-1]]
        for n in (10, 3, 1):
            nb, ei, et = nf.get_temporal_neighbor(nodes, ts, n_neighbors=n)
            out[f"{name}_n{n}_nbr"], out[f"{name}_n{n}_eidx"], out[f"{name}_n{n}_etime"] = nb, ei, et
        out[f"{name}_nodes"], out[f"{name}_ts"] = nodes, ts
        for k in ("sources", "destinations", "edge_idxs", "timestamps"):
            out[f"{name}_{k}"] = getattr(st, k)
        out[f"{name}_n_nodes"] = st.n_nodes
    np.savez_compressed(os.path.join(OUT, "neighbors.npz"), **out)
    print("neighbors.npz", len(out))


def _run_model(TGN, get_neighbor_finder, tag, *, d, n_layers, n_neighbors, use_memory, updater,
               embedding, dyrep, dst_emb, ts_mode, with_ppos, B=24, n_batches=4, n_neg=3, seed=11,
               msg_fn="identity", aggregator="last", src_emb=False, device="cpu", out_dir=None):
    from pfotgnrec_b200.synth import make_stream
    st = make_stream(n_users=40, n_items=12, n_events=B * n_batches + 40, n_days=10, seed=seed,
                     ts_mode=ts_mode, with_prices=False)
    rng = np.random.default_rng(seed + 1)
    torch.manual_seed(seed)
    node_feat = rng.random((st.n_nodes, d))
    nf = get_neighbor_finder(_data(st), uniform=False, max_node_idx=st.n_nodes - 1)
    shift = (3.0, 7.0, 2.0, 5.0)
    tgn = TGN(neighbor_finder=nf, node_features=node_feat, edge_features=st.edge_features.copy(),
              device=torch.device(device), n_layers=n_layers, n_heads=2, dropout=0.0,
              use_memory=use_memory, message_dimension=100, memory_dimension=d,
              memory_update_at_start=True, embedding_module_type=embedding,
              message_function=msg_fn, aggregator_type=aggregator, memory_updater_type=updater,
              n_neighbors=n_neighbors, mean_time_shift_src=shift[0], std_time_shift_src=shift[1],
              mean_time_shift_dst=shift[2], std_time_shift_dst=shift[3],
              use_destination_embedding_in_message=dst_emb, use_source_embedding_in_message=src_emb,
              dyrep=dyrep)
    tgn = tgn.to(torch.device(device))              # main.py:124
    tgn.train()
    out = {"cfg_" + k: v for k, v in dict(d=d, n_layers=n_layers, n_neighbors=n_neighbors,
                                          use_memory=int(use_memory), B=B, n_batches=n_batches,
                                          n_neg=n_neg, dyrep=int(dyrep), dst_emb=int(dst_emb),
                                          with_ppos=int(with_ppos)).items()}
    out["cfg_updater"], out["cfg_embedding"] = updater, embedding
    out["cfg_msg_fn"], out["cfg_aggregator"], out["cfg_src_emb"] = msg_fn, aggregator, int(src_emb)
    out["cfg_shift"] = np.array(shift)
    skip = ("memory.memory", "memory.last_update", "memory_updater.memory.", "embedding_module.memory.")
    for k, v in tgn.state_dict().items():
        if not any(k.startswith(s) or k == s for s in skip):
            out["w_" + k] = v.detach().cpu().numpy().copy()
    out["node_feat"] = node_feat
    for k in ("sources", "destinations", "edge_idxs", "timestamps", "edge_features"):
        out["st_" + k] = getattr(st, k)
    out["st_n_nodes"] = st.n_nodes
    params = [p for p in tgn.parameters() if p.requires_grad]
    names = [k for k, p in tgn.named_parameters() if p.requires_grad]
    for bi in range(n_batches):
        sl = slice(bi * B, (bi + 1) * B)
        src, dst = st.sources[sl], st.destinations[sl]
        ts, ei = st.timestamps[sl], st.edge_idxs[sl]
        neg = rng.integers(st.n_users + 1, st.n_nodes, size=B * n_neg)
        for p in params:
            p.grad = None
        if with_ppos:
            ppos = rng.integers(st.n_users + 1, st.n_nodes, size=B)
            e_s, e_d, e_p, e_n = tgn.compute_temporal_embeddings_p(src, dst, ppos, neg, ts, ei, n_neighbors)
            out[f"b{bi}_ppos"] = ppos
            out[f"b{bi}_emb_ppos"] = e_p.detach().cpu().numpy().copy()
        else:
            e_s, e_d, e_n = tgn.compute_temporal_embeddings(src, dst, neg, ts, ei, n_neighbors)
            e_p = e_d
        # BPR exactly as reference main.py:321-337
        bs = e_s.shape[0]
        s_ = e_s.view(bs, 1, -1)
        pos_scores = torch.sum(s_ * e_p.view(bs, 1, -1), dim=2)
        neg_scores = torch.matmul(s_, e_n.view(bs, n_neg, -1).transpose(1, 2)).squeeze()
        loss = -torch.mean(torch.log(torch.sigmoid(torch.mean(pos_scores - neg_scores, dim=1))))
        if dyrep or not loss.requires_grad:     # main.py:386-387; the identity embedding of batch 0 is plain zero memory
            loss.requires_grad_()
        loss.backward()
        if use_memory:
            tgn.memory.detach_memory()
        out[f"b{bi}_neg"] = neg
        out[f"b{bi}_emb_src"] = e_s.detach().cpu().numpy().copy()
        out[f"b{bi}_emb_dst"] = e_d.detach().cpu().numpy().copy()
        out[f"b{bi}_emb_neg"] = e_n.detach().cpu().numpy().copy()
        out[f"b{bi}_loss"] = float(loss.item())
        for k, p in zip(names, params):
            out[f"b{bi}_g_{k}"] = (p.grad.detach().cpu().numpy().copy() if p.grad is not None
                                   else np.zeros(tuple(p.shape), dtype=np.float32))
        if use_memory:
            out[f"b{bi}_memory"] = tgn.memory.memory.detach().cpu().numpy().copy()
            out[f"b{bi}_last_update"] = tgn.memory.last_update.detach().cpu().numpy().copy()
            N = st.n_nodes
            raw = 3 * d + st.edge_features.shape[1]
            pv = np.zeros(N, dtype=bool)
            pm = np.zeros((N, raw), dtype=np.float32)
            pt = np.zeros(N, dtype=np.float32)
            for node, lst in tgn.memory.messages.items():
                if len(lst) > 0:
                    pv[node] = True
                    pm[node] = (lst[-1][0] if aggregator == "last" else
                                torch.mean(torch.stack([m[0] for m in lst]), dim=0)).detach().cpu().numpy()
                    pt[node] = float(lst[-1][1])
            out[f"b{bi}_pend_valid"], out[f"b{bi}_pend_msg"], out[f"b{bi}_pend_ts"] = pv, pm, pt
    np.savez_compressed(os.path.join(out_dir or OUT, f"tgn_{tag}.npz"), **out)
    print(f"tgn_{tag}.npz", len(out))


def golden_mv_select(RandEdgeSampler):
    """Run the reference's inline MV block (main.py:209-292) on a synthetic batch."""
    import scipy.stats as stats
    from pfotgnrec_b200.synth import make_stream
    st = make_stream(n_users=80, n_items=60, n_events=400, n_days=12, seed=21, ts_mode="nbg")
    lines = open(os.path.join(REF, "main.py")).read().split("\n")
    block = "\n".join(lines[204:292])        # main.py:205-292: p_pos_batch=[] ... p_neg_batch.append
    block = "\n".join(l[8:] if l.startswith("        ") else l for l in block.split("\n"))
    code_of = dict(enumerate(st.codes))
    idx_of = {c: k for k, c in code_of.items()}
    time_feature = {dk: {c: st.prices_future[di, k] for k, c in enumerate(st.codes)}
                    for di, dk in enumerate(st.day_keys)}
    rng = np.random.default_rng(22)
    out = {}
    for ci, (lam, gamma, K) in enumerate(((0.5, 2.0, 20), (0.1, 2.0, 20), (0.9, 3.0, 7))):
        B = 64
        sl = slice(ci * B, (ci + 1) * B)
        cand = np.concatenate([(st.destinations[sl] - st.n_users - 1)[:, None],
                               rng.integers(0, st.n_items, size=(B, K))], axis=1)
        if ci == 0:
            cand[:8, 3] = cand[:8, 0]        # the positive can be drawn as a negative (exact ties)
        portfolios = [[st.codes[k] for k in st.portfolio(e)] or [""] for e in range(sl.start, sl.stop)]
        for stable in (True, False):
            ns = {"np": np if not stable else _StableNumpy(), "stats": stats,
                  "args": types.SimpleNamespace(gamma=gamma, lambda_mv=lam, p_pos_num=1, p_neg_num=3),
                  "time_feature": time_feature,
                  "portfolios_batch": np.array(portfolios, dtype=object),
                  "sources_batch": st.sources[sl], "timestamps_batch": st.timestamps[sl],
                  "destinations_batch": np.vectorize(code_of.get)(cand[:, 0]),
                  "negatives_batch": np.vectorize(code_of.get)(cand[:, 1:])}
            exec(block, ns)
            pp = np.array([idx_of[c] for y in ns["p_pos_batch"] for c in y])
            pn = np.array([idx_of[c] for y in ns["p_neg_batch"] for c in y])
            tag = "stable" if stable else "unstable"
            out[f"c{ci}_ppos_{tag}"], out[f"c{ci}_pneg_{tag}"] = pp, pn
        out[f"c{ci}_cand"], out[f"c{ci}_lam"], out[f"c{ci}_gamma"] = cand, lam, gamma
        out[f"c{ci}_event0"] = sl.start
    for k in ("day_idx", "port_ptr", "port_items", "prices_future"):
        out["st_" + k] = getattr(st, k)
    np.savez_compressed(os.path.join(OUT, "mv_select.npz"), **out)
    print("mv_select.npz", len(out))


def golden_eval_metrics():
    """Run the reference's UNMODIFIED eval_recommendation (evaluation.py:39-265) on a synthetic stream written
    in the reference's on-disk format, with a stub model whose embeddings are rows of a fixed random table: pins
    the metric block (ranking, Recall/NDCG@k, delta-return / delta-Sharpe@k in and out of sample, aggregation)."""
    import tempfile
    from pfotgnrec_b200.synth import make_stream, write_reference_format
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import evaluation as ref_eval                   # noqa: E402  (/root/reference/evaluation.py)
    st = make_stream(n_users=60, n_items=30, n_events=330, n_days=6, seed=31, ts_mode="nbg")
    work = tempfile.mkdtemp(prefix="pfo_golden_eval_")
    write_reference_format(st, work, period="30")
    rng = np.random.default_rng(32)
    table = torch.tensor(rng.standard_normal((st.n_nodes, 16)).astype(np.float32))
    negs = []

    class StubTGN:
        def eval(self):
            return self

        def compute_temporal_embeddings(self, src, dst, neg, ts, eidx, n_neighbors):
            negs.append(np.asarray(neg).copy())
            return table[torch.as_tensor(src)], table[torch.as_tensor(dst)], table[torch.as_tensor(neg)]

    portfolios = np.array([[st.codes[k] for k in st.portfolio(e)] or [""] for e in range(st.n_events)], dtype=object)
    e0, e1, B = 100, 330, 64                        # 3 full batches + a short last one (skipped, evaluation.py:68-69)
    data = types.SimpleNamespace(sources=st.sources[e0:e1], destinations=st.destinations[e0:e1],
                                 timestamps=st.timestamps[e0:e1], edge_idxs=st.edge_idxs[e0:e1],
                                 portfolios=portfolios[e0:e1])
    cwd = os.getcwd()
    os.chdir(work)
    out = {}
    try:
        # the true item is often drawn as a candidate too (equal scores): the ranking depends on how argsort
        # breaks ties, so both variants are recorded -- numpy's default and the stable sort (deviation ii)
        for tag, mod in (("unstable", np), ("stable", _StableNumpy())):
            ref_eval.np = mod
            n0 = len(negs)
            res = ref_eval.eval_recommendation(StubTGN(), data, _data(st), B, 10, st.n_users, "30", False, "val")
            out.update({f"res_{tag}_" + k: float(v) for k, v in res.items()})
        assert all(np.array_equal(a, b) for a, b in zip(negs[:n0], negs[n0:]))    # RandomState(2024) per batch
        del negs[n0:]
    finally:
        ref_eval.np = np
        os.chdir(cwd)
    out["table"] = table.numpy()
    out["negatives"] = np.stack([n.reshape(B, -1) for n in negs])         # [n_batches, B, N_ITEMS] item ids
    out["e0"], out["B"], out["n_batches"] = e0, B, len(negs)
    for k in ("sources", "destinations", "timestamps", "day_idx", "port_ptr", "port_items", "prices_future",
              "prices_past"):
        out["st_" + k] = getattr(st, k)
    out["st_n_users"] = st.n_users
    np.savez_compressed(os.path.join(OUT, "eval_metrics.npz"), **out)
    print("eval_metrics.npz", len(out), "batches", len(negs))


def golden_uniform_neighbors(get_neighbor_finder):
    """The reference's NeighborFinder in UNIFORM mode (utils/utils.py:193-204): per query the support its own
    `find_before` returns (everything strictly before the cut time) and one draw of `get_temporal_neighbor` (numpy's
    global MT19937: the values are not the contract, their structure is -- n picks WITH replacement out of the support,
    re-sorted by fp32 time, all-zero rows when the support is empty)."""
    from pfotgnrec_b200.synth import make_stream
    st = make_stream(n_users=60, n_items=12, n_events=700, n_days=15, seed=51, ts_mode="small", with_prices=False)
    nf = get_neighbor_finder(_data(st), uniform=True, max_node_idx=st.n_nodes - 1)
    rng = np.random.default_rng(52)
    Q, n = 200, 8
    nodes = rng.integers(0, st.n_nodes, size=Q)
    ts = st.timestamps[rng.integers(0, st.n_events, size=Q)] + rng.integers(-1, 2, size=Q)
    ts[:3] = [-1.0, st.timestamps[-1] + 5, st.timestamps[0]]
    smax = max(len(nf.find_before(int(v), float(t))[0]) for v, t in zip(nodes, ts))
    sup_e = np.zeros((Q, smax), dtype=np.int64)
    sup_n = np.zeros((Q, smax), dtype=np.int64)
    sup_len = np.zeros(Q, dtype=np.int64)
    for q, (v, t) in enumerate(zip(nodes, ts)):
        nb, ei, _ = nf.find_before(int(v), float(t))
        sup_len[q] = len(nb)
        sup_n[q, :len(nb)], sup_e[q, :len(ei)] = nb, ei
    np.random.seed(7)
    nb, ei, et = nf.get_temporal_neighbor(nodes, ts, n_neighbors=n)
    out = dict(nodes=nodes, ts=ts, n=n, sup_len=sup_len, sup_nbr=sup_n, sup_eidx=sup_e, ref_nbr=nb, ref_eidx=ei, ref_etime=et,
               n_nodes=st.n_nodes)
    for k in ("sources", "destinations", "edge_idxs", "timestamps"):
        out["st_" + k] = getattr(st, k)
    np.savez_compressed(os.path.join(OUT, "neighbors_uniform.npz"), **out)
    print("neighbors_uniform.npz", len(out))


def golden_sampler_support(RandEdgeSampler):
    """The reference's RandEdgeSampler (utils/utils.py:65-114) on a synthetic batch: its OWN fields after construction
    (`dst_unique`, `portfolio_list`) give the support of every interaction, available_i = setdiff1d(dst_unique,
    portfolio_i) (:96), and its own `sample(size)` output shows the replace rule (:99-111): distinct ids iff
    len(available_i) >= size.  The draws themselves come from numpy's MT19937 and are NOT the contract (the Philox
    stream is); support, replace rule and uniformity are."""
    from pfotgnrec_b200.synth import make_stream
    st = make_stream(n_users=60, n_items=24, n_events=500, n_days=8, seed=41, ts_mode="nbg", with_prices=False)
    map_item_id = {c: k for k, c in enumerate(st.codes)}
    train_dst = st.destinations[:400]
    # drop a few stocks from the training destinations so the universe is a strict subset of the items
    train_dst = train_dst[~np.isin(train_dst - st.n_users - 1, [3, 7, 11])]
    sl = slice(400, 464)
    portfolios = np.array([[st.codes[k] for k in st.portfolio(e)] or [""] for e in range(sl.start, sl.stop)], dtype=object)
    out = {"train_dst": train_dst, "e0": sl.start, "B": sl.stop - sl.start, "n_users": st.n_users}
    for k in ("sources", "destinations", "port_ptr", "port_items", "edge_idxs"):
        out["st_" + k] = getattr(st, k)
    rs = RandEdgeSampler(st.sources[sl], train_dst, portfolios, st.n_users, map_item_id)
    I = len(rs.dst_unique)
    avail = np.zeros((out["B"], I), dtype=bool)
    for i in range(out["B"]):
        a = np.setdiff1d(rs.dst_unique, rs.portfolio_list[i])
        avail[i] = np.isin(rs.dst_unique, a)
    out["dst_unique"], out["available"] = rs.dst_unique, avail
    np.random.seed(5)
    for size in (3, 20, I - 2, I + 9):                   # without replacement ... with replacement (evaluation's case)
        out[f"ref_sample_{size}"] = rs.sample(size)
    out["sizes"] = np.array([3, 20, I - 2, I + 9])
    rs2 = RandEdgeSampler(st.sources[sl], train_dst, portfolios, st.n_users, map_item_id, seed=2024)
    out["ref_sample_seeded"] = rs2.sample(I)
    np.savez_compressed(os.path.join(OUT, "sampler_support.npz"), **out)
    print("sampler_support.npz", len(out))


def golden_state_dict(TGN, get_neighbor_finder):
    """A reference checkpoint: model A (seed 11) trains 2 batches with Adam, `A.state_dict()` is saved with ALL its
    keys (parameters, the `time_encoder` registered under two names, the three aliases of the memory buffers:
    memory.*, memory_updater.memory.*, embedding_module.memory.*); model B (seed 99) loads it (strict) and runs batch 3
    -- the pending raw messages are not part of a state_dict, so B starts from the memory alone, as a reloaded reference
    does.  The drop-in has to load the same dictionary and reproduce B's batch."""
    from pfotgnrec_b200.synth import make_stream
    d, B, n = 32, 24, 10
    st = make_stream(n_users=40, n_items=12, n_events=B * 3 + 40, n_days=10, seed=13, ts_mode="small", with_prices=False)
    rng = np.random.default_rng(14)
    node_feat = rng.random((st.n_nodes, d))
    nf = get_neighbor_finder(_data(st), uniform=False, max_node_idx=st.n_nodes - 1)

    def make(seed):
        torch.manual_seed(seed)
        return TGN(neighbor_finder=nf, node_features=node_feat, edge_features=st.edge_features.copy(),
                   device=torch.device("cpu"), n_layers=1, n_heads=2, dropout=0.0, use_memory=True,
                   message_dimension=100, memory_dimension=d, memory_update_at_start=True,
                   embedding_module_type="graph_attention", message_function="identity", aggregator_type="last",
                   memory_updater_type="gru", n_neighbors=n, mean_time_shift_src=0.0, std_time_shift_src=1.0,
                   mean_time_shift_dst=0.0, std_time_shift_dst=1.0, use_destination_embedding_in_message=False,
                   use_source_embedding_in_message=False, dyrep=False)

    def batch(tgn, bi, neg):
        sl = slice(bi * B, (bi + 1) * B)
        return tgn.compute_temporal_embeddings(st.sources[sl], st.destinations[sl], neg, st.timestamps[sl],
                                               st.edge_idxs[sl], n)

    A = make(11).train()
    opt = torch.optim.Adam(A.parameters(), lr=1e-2)
    negs = [rng.integers(st.n_users + 1, st.n_nodes, size=B * 3) for _ in range(3)]
    for bi in range(2):
        opt.zero_grad()
        e_s, e_d, e_n = batch(A, bi, negs[bi])
        s_ = e_s.view(B, 1, -1)
        loss = -torch.mean(torch.log(torch.sigmoid(torch.mean(
            torch.sum(s_ * e_d.view(B, 1, -1), dim=2) - torch.matmul(s_, e_n.view(B, 3, -1).transpose(1, 2)).squeeze(), dim=1))))
        loss.backward()
        opt.step()
        A.memory.detach_memory()
    sd = A.state_dict()
    out = {"sd_" + k: v.detach().cpu().numpy().copy() for k, v in sd.items()}
    out["sd_keys"] = np.array(list(sd.keys()))
    Bm = make(99)
    Bm.load_state_dict(sd, strict=True)
    Bm.eval()
    with torch.no_grad():
        e_s, e_d, e_n = batch(Bm, 2, negs[2])
    out.update(emb_src=e_s.numpy().copy(), emb_dst=e_d.numpy().copy(), emb_neg=e_n.numpy().copy(), neg=negs[2],
               memory_after=Bm.memory.memory.detach().numpy().copy(),
               last_update_after=Bm.memory.last_update.detach().numpy().copy(), node_feat=node_feat,
               cfg_d=d, cfg_B=B, cfg_n=n)
    for k in ("sources", "destinations", "edge_idxs", "timestamps", "edge_features"):
        out["st_" + k] = getattr(st, k)
    out["st_n_nodes"] = st.n_nodes
    np.savez_compressed(os.path.join(OUT, "state_dict.npz"), **out)
    print("state_dict.npz", len(out))


class _StableNumpy:
    """numpy with argsort(kind='stable') -- documented deviation (ii)."""

    def __getattr__(self, name):
        if name == "argsort":
            return lambda a, *args, **kw: np.argsort(a, *args, kind="stable", **kw)
        return getattr(np, name)


def model_cases():
    """tag -> keyword arguments of `_run_model`: the six models main.py / BASELINE configs build and the API variants."""
    common = dict(n_layers=1, n_neighbors=10, use_memory=True, updater="gru",
                  embedding="graph_attention", dyrep=False, dst_emb=False)
    return {
        "ours": dict(d=64, ts_mode="small", with_ppos=True, **common),
        "ours_nbg": dict(d=32, ts_mode="nbg", with_ppos=True, n_batches=2, **common),
        "tgn": dict(d=32, ts_mode="small", with_ppos=False, **common),
        "jodie": dict(d=32, ts_mode="small", with_ppos=False, **{**common, "updater": "rnn", "embedding": "time"}),
        "dyrep": dict(d=32, ts_mode="small", with_ppos=False,
                      **{**common, "updater": "rnn", "dyrep": True, "dst_emb": True}),
        "tgat2": dict(d=32, ts_mode="small", with_ppos=False,
                      **{**common, "use_memory": False, "n_layers": 2, "n_neighbors": 5}),
        # BASELINE config 3 at its own shape: no memory, 2 layers x 20 neighbours (most-recent mode; the uniform mode
        # draws from numpy's global MT19937 in the reference and is pinned through the Philox oracle instead)
        "tgat2x20": dict(d=32, ts_mode="small", with_ppos=False, n_batches=2, B=16,
                         **{**common, "use_memory": False, "n_layers": 2, "n_neighbors": 20}),
        # API-surface variants no model of main.py builds: MLP message function + mean aggregator,
        # the node's own embedding in its message, graph_sum, and the identity embedding (memory rows as they are)
        "mlp_mean": dict(d=32, ts_mode="small", with_ppos=False, n_batches=3, msg_fn="mlp", aggregator="mean", **common),
        "srcemb": dict(d=32, ts_mode="small", with_ppos=False, n_batches=3, src_emb=True, **{**common, "dst_emb": True}),
        "gsum": dict(d=32, ts_mode="small", with_ppos=False, n_batches=3, **{**common, "embedding": "graph_sum"}),
        "gsum2": dict(d=32, ts_mode="small", with_ppos=False, n_batches=2,
                      **{**common, "embedding": "graph_sum", "use_memory": False, "n_layers": 2, "n_neighbors": 4}),
        "identity": dict(d=32, ts_mode="small", with_ppos=False, n_batches=3, **{**common, "embedding": "identity"}),
    }


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", help="cuda: run the reference's own torch-CUDA path (side-by-side test)")
    ap.add_argument("--out", default=None, help="output directory (default tests/golden; required with --device cuda)")
    ap.add_argument("--only", default=None, help="one of the non-model goldens: sampler_support | state_dict | uniform_neighbors")
    ap.add_argument("--tags", default=None, help="comma-separated model cases (default: all, plus the non-model goldens)")
    a = ap.parse_args(argv)
    if a.device != "cpu" and a.out is None:
        raise SystemExit("--device cuda writes side-by-side vectors: pass --out (the committed goldens are CPU runs)")
    out_dir = a.out or OUT
    os.makedirs(out_dir, exist_ok=True)
    TGN, get_neighbor_finder, RandEdgeSampler = _import_reference()
    cases = model_cases()
    if a.only:
        {"uniform_neighbors": lambda: golden_uniform_neighbors(get_neighbor_finder),
         "sampler_support": lambda: golden_sampler_support(RandEdgeSampler),
         "state_dict": lambda: golden_state_dict(TGN, get_neighbor_finder)}[a.only]()
        return
    for tag in (a.tags.split(",") if a.tags else cases):
        _run_model(TGN, get_neighbor_finder, tag, device=a.device, out_dir=out_dir, **cases[tag])
    if a.tags is None and a.device == "cpu":
        golden_neighbors(get_neighbor_finder)
        golden_mv_select(RandEdgeSampler)
        golden_eval_metrics()
        golden_sampler_support(RandEdgeSampler)
        golden_state_dict(TGN, get_neighbor_finder)
        golden_uniform_neighbors(get_neighbor_finder)


if __name__ == "__main__":
    main()
