"""The oracle against vectors produced by the unmodified reference (CPU, no GPU needed)."""
import numpy as np
import pytest
import torch

from helpers import load_golden, oracle_from_golden, rel_err, batch_inputs
from oracle.graph import AdjacencyOracle
from oracle.tgn import bpr_loss
from oracle import sampling

TOL = 1e-5   # fp32 contract of BASELINE.json north_star


@pytest.mark.parametrize("name", ["small", "nbg"])
@pytest.mark.parametrize("n", [10, 3, 1])
def test_neighbors_bit_exact(name, n):
    z = load_golden("neighbors.npz")
    adj = AdjacencyOracle(z[f"{name}_sources"], z[f"{name}_destinations"], z[f"{name}_edge_idxs"],
                          z[f"{name}_timestamps"], n_nodes=int(z[f"{name}_n_nodes"]))
    nb, ei, et = adj.get_temporal_neighbor(z[f"{name}_nodes"], z[f"{name}_ts"], n)
    assert np.array_equal(nb, z[f"{name}_n{n}_nbr"])
    assert np.array_equal(ei, z[f"{name}_n{n}_eidx"])
    assert np.array_equal(et, z[f"{name}_n{n}_etime"])
    assert nb.dtype == np.int32 and ei.dtype == np.int32 and et.dtype == np.float32


@pytest.mark.parametrize("tag", ["ours", "ours_nbg", "tgn", "jodie", "dyrep", "tgat2", "tgat2x20", "mlp_mean", "srcemb", "gsum", "gsum2",
                                 "identity"])
def test_model_path_matches_reference(tag):
    z = load_golden(f"tgn_{tag}.npz")
    o, p = oracle_from_golden(z)
    n = int(z["cfg_n_neighbors"])
    for bi in range(int(z["cfg_n_batches"])):
        src, dst, extra, ts, ei = batch_inputs(z, bi)
        for v in p.values():
            v.grad = None
        outs = o.compute_temporal_embeddings(src, dst, extra, ts, ei, n)
        names = ["src", "dst"] + (["ppos"] if len(extra) == 2 else []) + ["neg"]
        for nm, e in zip(names, outs):
            assert rel_err(e.detach().numpy(), z[f"b{bi}_emb_{nm}"]) < TOL, (tag, bi, nm)
        e_pos = outs[2] if len(extra) == 2 else outs[1]
        loss = bpr_loss(outs[0], e_pos, outs[-1])
        assert abs(loss.item() - float(z[f"b{bi}_loss"])) < TOL * max(1.0, abs(float(z[f"b{bi}_loss"])))
        if loss.requires_grad:
            loss.backward()
        for k, v in p.items():
            g = v.grad.numpy() if v.grad is not None else np.zeros(tuple(v.shape), np.float32)
            key = f"b{bi}_g_{k}"
            if key in z:
                ref = z[key]
                assert np.abs(g - ref).max() <= 2e-5 * max(np.abs(ref).max(), 1e-3) + 1e-7, (tag, bi, k)
        if bool(z["cfg_use_memory"]):
            assert rel_err(o.memory.numpy(), z[f"b{bi}_memory"]) < TOL
            assert np.array_equal(o.last_update.numpy(), z[f"b{bi}_last_update"])
            assert np.array_equal(o.pend_valid.numpy(), z[f"b{bi}_pend_valid"])
            v = z[f"b{bi}_pend_valid"]
            assert np.array_equal(o.pend_ts.numpy()[v], z[f"b{bi}_pend_ts"][v])
            assert rel_err(o.pend_msg.numpy()[v], z[f"b{bi}_pend_msg"][v]) < TOL


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_mv_select_ids_bit_exact(ci):
    from pfotgnrec_b200.synth import log_returns
    z = load_golden("mv_select.npz")
    cand = z[f"c{ci}_cand"]
    B = cand.shape[0]
    e0 = int(z[f"c{ci}_event0"])
    ptr = z["st_port_ptr"][e0:e0 + B + 1]
    pp, pn = sampling.mv_select(log_returns(z["st_prices_future"]), z["st_day_idx"][e0:e0 + B], cand,
                                ptr - ptr[0], z["st_port_items"][ptr[0]:ptr[-1]],
                                float(z[f"c{ci}_gamma"]), float(z[f"c{ci}_lam"]))
    assert np.array_equal(pp, z[f"c{ci}_ppos_stable"])
    assert np.array_equal(pn, z[f"c{ci}_pneg_stable"])


def test_time_statistics_host_vs_oracle():
    """pfotgnrec_b200.trainer.time_statistics (one sort) == the oracle's restatement of the reference loop
    (utils/data.py:75-99), bit for bit, on a small and an NBG-format stream."""
    from pfotgnrec_b200.synth import make_stream
    from pfotgnrec_b200.trainer import time_statistics as host_stats
    from oracle.train_loop import time_statistics as oracle_stats
    for mode in ("small", "nbg"):
        st = make_stream(n_users=200, n_items=40, n_events=2500, n_days=15, seed=5, ts_mode=mode, with_prices=False)
        assert tuple(host_stats(st.sources, st.destinations, st.timestamps)) == \
            oracle_stats(st.sources, st.destinations, st.timestamps)


def _eval_golden_batches(z):
    """(scores, pos_stock, cand_stock, day_idx, port_ptr, port_items) per batch of tests/golden/eval_metrics.npz;
    scores are formed like reference evaluation.py:107-115 from the stub model's embedding table."""
    table = torch.tensor(z["table"])
    e0, B, U = int(z["e0"]), int(z["B"]), int(z["st_n_users"])
    for bi in range(int(z["n_batches"])):
        s, e = e0 + bi * B, e0 + (bi + 1) * B
        src, dst, neg = z["st_sources"][s:e], z["st_destinations"][s:e], z["negatives"][bi]
        es = table[torch.as_tensor(src)].view(B, 1, -1)
        pos = torch.sum(es * table[torch.as_tensor(dst)].view(B, 1, -1), dim=2)
        negs = torch.sum(es * table[torch.as_tensor(neg.reshape(-1))].view(B, neg.shape[1], -1), dim=2)
        scores = torch.cat([pos, negs], dim=1).numpy()
        ptr = z["st_port_ptr"][s:e + 1]
        yield (scores, dst - U - 1, neg - U - 1, z["st_day_idx"][s:e], ptr - ptr[0],
               z["st_port_items"][ptr[0]:ptr[-1]])


def test_eval_metric_block_matches_reference():
    """oracle/eval_metrics.py vs the dictionary the unmodified reference eval_recommendation returned
    (evaluation.py:127-258) on the same scores: 30 keys, fp64."""
    from oracle import eval_metrics as em
    z = load_golden("eval_metrics.npz")
    rows = [em.per_event_metrics(*b, z["st_prices_past"], z["st_prices_future"])[0] for b in _eval_golden_batches(z)]
    got = em.aggregate(np.concatenate(rows), "val")
    ref = {k[len("res_stable_"):]: float(v) for k, v in z.items() if k.startswith("res_stable_")}
    assert set(got) == set(ref) and len(ref) == 30
    for k, v in ref.items():
        assert abs(got[k] - v) <= 1e-12 * max(1.0, abs(v)), (k, got[k], v)


def test_philox4x32_10_known_answer_vectors():
    """The shared random stream (oracle/philox.py == csrc/philox.cuh) is Philox4x32-10 of Random123: its three published
    known-answer vectors (kat_vectors: zero, all-ones and the pi-digits counter / key)."""
    from oracle.philox import philox4x32_10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox4x32_10(np.array([ctr[0]], dtype=np.int64), ctr[1], ctr[2], ctr[3], key[0], key[1])
        assert tuple(int(x[0]) for x in got) == want


def _support_case(z):
    e0, B, U = int(z["e0"]), int(z["B"]), int(z["n_users"])
    ptr = z["st_port_ptr"][e0:e0 + B + 1]
    held = z["st_port_items"][ptr[0]:ptr[-1]].astype(np.int64) + U + 1
    return e0, B, ptr - ptr[0], held


def check_sampler_support(z, sample_fn):
    """Support, replace rule and uniformity of a candidate sampler against what the reference's RandEdgeSampler holds
    after construction (utils/utils.py:73-81 -> `dst_unique`, `portfolio_list`) and does in `sample` (:93-113):
    candidates come from setdiff1d(dst_unique, portfolio_i); they are distinct iff that set has >= size members (the
    reference's own samples show the same pattern); every available item is equally likely."""
    e0, B, pptr, held = _support_case(z)
    dst_unique, avail = z["dst_unique"], z["available"]
    assert np.array_equal(dst_unique, np.unique(z["train_dst"]))
    ev = z["st_edge_idxs"][e0:e0 + B]
    for size in z["sizes"].tolist():
        ref = z[f"ref_sample_{size}"]
        got = np.asarray(sample_fn(ev, dst_unique, pptr, held, size, 0))
        assert got.shape == ref.shape
        for i in range(B):
            allowed = dst_unique[avail[i]]
            assert np.isin(got[i], allowed).all(), (size, i)                   # support
            assert np.isin(ref[i], allowed).all()
            distinct = len(allowed) >= size                                    # replace rule (:99 vs :107)
            assert (len(set(ref[i].tolist())) == size) == distinct
            assert (len(set(got[i].tolist())) == size) == distinct, (size, i)
    # uniformity over the available items: pooled chi-square of 64 interactions x 400 epochs x 3 draws
    counts = np.zeros(len(dst_unique))
    expect = np.zeros(len(dst_unique))
    for rep in range(400):
        got = np.asarray(sample_fn(ev + 1000 * rep, dst_unique, pptr, held, 3, 0))
        counts += np.bincount(np.searchsorted(dst_unique, got.ravel()), minlength=len(dst_unique))
        expect += (avail / avail.sum(axis=1, keepdims=True)).sum(axis=0) * 3
    chi2 = float(((counts - expect) ** 2 / expect).sum())
    dof = len(dst_unique) - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof), (chi2, dof)                      # ~6 sigma of chi-square(dof)


def test_candidate_sampler_support_and_replace_rule_vs_reference():
    """oracle/sampling.py::sample_candidates against the reference RandEdgeSampler's own support / replace rule."""
    z = load_golden("sampler_support.npz")
    check_sampler_support(z, lambda ev, items, pptr, held, size, seed:
                          sampling.sample_candidates(ev, items, pptr, held, size, seed))


def check_uniform_neighbors(z, sample_fn):
    """Structure of a uniform-mode neighbour sampler against the reference's (utils/utils.py:193-204), on the supports
    the reference's own `find_before` returned: every pick is an (neighbour, edge) pair of the support, n picks with
    replacement (so a support smaller than n repeats), ascending fp32 times, all-zero rows for an empty support -- the
    reference's own draw has the same structure; uniformity over the support is checked in the aggregate."""
    n, Q = int(z["n"]), z["nodes"].shape[0]
    ts_of_edge = np.zeros(int(z["st_edge_idxs"].max()) + 1, dtype=np.float32)
    ts_of_edge[z["st_edge_idxs"]] = z["st_timestamps"].astype(np.float32)
    hist, expect = np.zeros(8), np.zeros(8)
    for call in range(30):
        nb, ei, et = sample_fn(z["nodes"], z["ts"], n, call)
        for q in range(Q):
            L = int(z["sup_len"][q])
            if L == 0:
                assert not nb[q].any() and not ei[q].any() and not et[q].any()
                continue
            pairs = set(zip(z["sup_nbr"][q, :L].tolist(), z["sup_eidx"][q, :L].tolist()))
            assert all((a, b) in pairs for a, b in zip(nb[q].tolist(), ei[q].tolist())), (call, q)
            assert np.all(np.diff(et[q]) >= 0) and np.array_equal(et[q], ts_of_edge[ei[q]])
            if call == 0:
                rp = set(zip(z["ref_nbr"][q].tolist(), z["ref_eidx"][q].tolist()))
                assert rp <= pairs and np.all(np.diff(z["ref_etime"][q]) >= 0)
            if L >= 8:           # position of each pick inside the support, folded to 8 bins
                order = {e: k for k, e in enumerate(z["sup_eidx"][q, :L].tolist())}
                for e in ei[q].tolist():
                    hist[order[e] * 8 // L] += 1
                expect += np.bincount(np.arange(L) * 8 // L, minlength=8) * (n / L)
    assert hist.sum() > 5000 and np.abs(hist - expect).max() < 5 * np.sqrt(expect.max()), (hist, expect)


def test_uniform_neighbors_structure_vs_reference():
    z = load_golden("neighbors_uniform.npz")
    adj = AdjacencyOracle(z["st_sources"], z["st_destinations"], z["st_edge_idxs"], z["st_timestamps"],
                          n_nodes=int(z["n_nodes"]), uniform=True)
    check_uniform_neighbors(z, lambda nodes, ts, n, call: adj.get_temporal_neighbor(nodes, ts, n, call_id=call, seed=3))
