"""Kernel-level parity on the GPU, through the C ABI (ctypes), against the CPU oracle."""
import numpy as np
import pytest
import torch

from helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _stream(U=300, I=40, E=6000, mode="small", seed=1):
    from pfotgnrec_b200.synth import make_stream
    return make_stream(n_users=U, n_items=I, n_events=E, n_days=30, seed=seed, ts_mode=mode)


# ------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("name", ["small", "nbg"])
@pytest.mark.parametrize("n", [10, 3, 1])
def test_neighbors_golden_bit_exact(name, n):
    from pfotgnrec_b200.graph import TemporalCSR, NeighborFinder
    z = load_golden("neighbors.npz")
    csr = TemporalCSR(z[f"{name}_sources"], z[f"{name}_destinations"], z[f"{name}_edge_idxs"],
                      z[f"{name}_timestamps"], n_nodes=int(z[f"{name}_n_nodes"]), device=DEV)
    nf = NeighborFinder(csr)
    nb, ei, et = nf.get_temporal_neighbor(z[f"{name}_nodes"], z[f"{name}_ts"], n)
    assert np.array_equal(nb, z[f"{name}_n{n}_nbr"])
    assert np.array_equal(ei, z[f"{name}_n{n}_eidx"])
    assert np.array_equal(et, z[f"{name}_n{n}_etime"])
    assert nb.dtype == np.int32 and et.dtype == np.float32


@pytest.mark.parametrize("uniform", [False, True])
@pytest.mark.parametrize("mode", ["small", "nbg"])
def test_neighbors_vs_oracle(uniform, mode):
    from oracle.graph import AdjacencyOracle
    from pfotgnrec_b200.graph import TemporalCSR, NeighborFinder
    st = _stream(mode=mode)
    adj = AdjacencyOracle(st.sources, st.destinations, st.edge_idxs, st.timestamps, n_nodes=st.n_nodes, uniform=uniform)
    csr = TemporalCSR(st.sources, st.destinations, st.edge_idxs, st.timestamps, n_nodes=st.n_nodes, device=DEV)
    # the device-built CSR equals the oracle's stable (node, ts, stream-order) sort
    assert np.array_equal(csr.rowptr.cpu().numpy(), adj.rowptr)
    assert np.array_equal(csr.nbr.cpu().numpy(), adj.nbr)
    assert np.array_equal(csr.eidx.cpu().numpy(), adj.eidx)
    assert np.array_equal(csr.ts.cpu().numpy(), adj.ts)
    rng = np.random.default_rng(0)
    Q = 1000 + 7        # ragged: not a multiple of 32
    nodes = rng.integers(0, st.n_nodes, size=Q)
    ts = st.timestamps[rng.integers(0, st.n_events, size=Q)] + rng.integers(-1, 2, size=Q)
    nf = NeighborFinder(csr, uniform=uniform, seed=77)
    for call in range(2):
        for n in (10, 20):
            nf.call_id = call
            a = nf.get_temporal_neighbor(nodes, ts, n)
            b = adj.get_temporal_neighbor(nodes, ts, n, call_id=call, seed=77)
            for x, y in zip(a, b):
                assert np.array_equal(x, y)


def test_neighbors_empty_and_delta():
    from pfotgnrec_b200.graph import TemporalCSR, NeighborFinder
    st = _stream(mode="nbg")
    csr = TemporalCSR(st.sources, st.destinations, st.edge_idxs, st.timestamps, n_nodes=st.n_nodes, device=DEV)
    nf = NeighborFinder(csr)
    qn = torch.as_tensor(st.sources[:500].astype(np.int32), device=DEV)
    qt = torch.as_tensor(st.timestamps[:500], device=DEV)
    nbr, eidx, et, dt = nf.sample(qn, qt, 10)
    ref = (st.timestamps[:500, None] - et.cpu().numpy().astype(np.float64)).astype(np.float32)
    assert np.array_equal(dt.cpu().numpy(), ref)          # fp64 subtract, then fp32 (embedding_module.py:133-135)
    e = nf.get_temporal_neighbor(np.zeros(0, np.int64), np.zeros(0), 10)
    assert e[0].shape == (0, 10)
    z = nf.get_temporal_neighbor(np.array([1, 2]), np.array([5.0, 6.0]), 0)
    assert z[0].shape == (2, 1) and not z[0].any()


@pytest.mark.parametrize("lpq", [1, 4, 8, 32, 0])
def test_neighbors_cooperative_search_every_width(lpq):
    """The (lanes+1)-ary cooperative lower-bound search of K1 at every lanes-per-query width (0 = the launcher's
    choice): golden rows bit-exact, and a stream with heavy rows (a few stocks holding thousands of entries, ties in
    the timestamps, queries before the first / after the last entry, ragged query counts) bit-exact against the
    oracle's np.searchsorted (utils/utils.py:150-220)."""
    from oracle.graph import AdjacencyOracle
    from pfotgnrec_b200.graph import TemporalCSR, NeighborFinder
    z = load_golden("neighbors.npz")
    for name in ("small", "nbg"):
        csr = TemporalCSR(z[f"{name}_sources"], z[f"{name}_destinations"], z[f"{name}_edge_idxs"],
                          z[f"{name}_timestamps"], n_nodes=int(z[f"{name}_n_nodes"]), device=DEV)
        nf = NeighborFinder(csr)
        nf.lanes_per_query = lpq
        for n in (10, 3, 1):
            nb, ei, et = nf.get_temporal_neighbor(z[f"{name}_nodes"], z[f"{name}_ts"], n)
            assert np.array_equal(nb, z[f"{name}_n{n}_nbr"]) and np.array_equal(ei, z[f"{name}_n{n}_eidx"])
            assert np.array_equal(et, z[f"{name}_n{n}_etime"])
    st = _stream(U=300, I=6, E=40000, mode="small", seed=17)        # 6 stocks x ~6 700 entries each, tied timestamps
    adj = AdjacencyOracle(st.sources, st.destinations, st.edge_idxs, st.timestamps, n_nodes=st.n_nodes)
    csr = TemporalCSR(st.sources, st.destinations, st.edge_idxs, st.timestamps, n_nodes=st.n_nodes, device=DEV)
    nf = NeighborFinder(csr)
    nf.lanes_per_query = lpq
    rng = np.random.default_rng(lpq)
    for Q in (1, 31, 33, 4099):
        nodes = np.where(rng.random(Q) < 0.7, rng.integers(st.n_users + 1, st.n_nodes, size=Q),
                         rng.integers(0, st.n_nodes, size=Q))
        ts = st.timestamps[rng.integers(0, st.n_events, size=Q)] + rng.integers(-1, 2, size=Q)
        ts[:1] = -3.0
        ts[-1:] = st.timestamps[-1] + 10
        for n in (10, 20, 1):
            a = nf.get_temporal_neighbor(nodes, ts, n)
            b = adj.get_temporal_neighbor(nodes, ts, n)
            for x, y in zip(a, b):
                assert np.array_equal(x, y), (lpq, Q, n)


def test_uniform_neighbors_structure_vs_reference():
    """K1's uniform mode against the supports / structure of the reference's own uniform NeighborFinder
    (tests/golden/neighbors_uniform.npz, utils/utils.py:193-204) -- the checker the oracle passes on CPU."""
    from test_oracle_golden import check_uniform_neighbors
    from pfotgnrec_b200.graph import TemporalCSR, NeighborFinder
    z = load_golden("neighbors_uniform.npz")
    csr = TemporalCSR(z["st_sources"], z["st_destinations"], z["st_edge_idxs"], z["st_timestamps"],
                      n_nodes=int(z["n_nodes"]), device=DEV)
    nf = NeighborFinder(csr, uniform=True, seed=3)

    def sample(nodes, ts, n, call):
        nf.call_id = call
        return nf.get_temporal_neighbor(nodes, ts, n)

    check_uniform_neighbors(z, sample)


def test_time_encode_cos_paths():
    """TimeEncode (model/time_encoding.py:17-25) through pfo_time_encode: the fp64 quadrant reduction the kernels use
    stays within 2e-7 of the fp64 libm value from day-scale arguments up to ~1e10 rad (NBG-format time deltas); the fp32
    Cody-Waite reduction (mode 2, not used by the product: measured slower once a per-warp choice is added, see
    pfo_math.cuh) agrees with it to <= 1 ulp of 1.0 below its 2^17 limit."""
    from pfotgnrec_b200 import _lib
    from pfotgnrec_b200._lib import ptr
    d = 64
    g = torch.Generator(device="cuda").manual_seed(0)
    w = torch.tensor(1.0 / 10 ** np.linspace(0, 9, d), dtype=torch.float32, device=DEV)
    b = torch.rand(d, device=DEV, generator=g)

    def run(t, mode):
        M = t.shape[0]
        c, s_ = torch.empty(M, d, device=DEV), torch.empty(M, d, device=DEV)
        _lib.call("pfo_time_encode", ptr(t), ptr(w), ptr(b), M, d, mode, ptr(c), ptr(s_))
        return c, s_

    ulp = float(np.spacing(np.float32(1.0)))
    t_small = (torch.rand(20000, device=DEV, generator=g) * 2 - 1) * 131000.0
    c64, s64 = run(t_small, 1)
    c32, s32 = run(t_small, 2)
    ca, sa = run(t_small, 0)
    assert float((c32 - c64).abs().max()) <= 1.0 * ulp + 1e-12
    assert float((s32 - s64).abs().max()) <= 1.0 * ulp + 1e-12
    assert torch.equal(ca, c64) and torch.equal(sa, s64)
    x = torch.addcmul(b.double(), t_small.double()[:, None], w.double()).float().double()     # fmaf(t, w, b)
    assert float((c64.double() - torch.cos(x)).abs().max()) < 2e-7
    assert float((s64.double() - torch.sin(x)).abs().max()) < 2e-7
    assert float((c32.double() - torch.cos(x)).abs().max()) < 2e-7
    ch, sh = run(t_small, 3)                              # cosine-only half-turn form (neighbour forward kernel)
    assert float((ch.double() - torch.cos(x)).abs().max()) < 2e-7 and torch.equal(sh, s64)
    assert float((ch - c64).abs().max()) <= 3.0 * ulp
    sign = torch.where(torch.rand(20000, device=DEV, generator=g) < 0.5, -1.0, 1.0)
    t_big = sign * (1.0e9 + torch.rand(20000, device=DEV, generator=g) * 1.0e10)
    cb, sb = run(t_big, 0)
    xb = torch.addcmul(b.double(), t_big.double()[:, None], w.double()).float().double()
    assert float((cb.double() - torch.cos(xb)).abs().max()) < 2e-7
    assert float((sb.double() - torch.sin(xb)).abs().max()) < 2e-7
    assert float((run(t_big, 3)[0].double() - torch.cos(xb)).abs().max()) < 2e-7
    # the standalone module forward (containers.TimeEncode) goes through the same entry point
    from pfotgnrec_b200.containers import TimeEncode
    te = TimeEncode(d).to(DEV)
    out = te(t_small[:64].view(8, 8))
    ref = torch.cos((t_small[:64].double().view(8, 8, 1) * te.w.weight.double().view(1, 1, d)).float().double())
    assert out.shape == (8, 8, d) and float((out.double() - ref).abs().max()) < 2e-7


# ------------------------------------------------------------------------------ compaction
def test_unique_node_compaction():
    from pfotgnrec_b200 import _lib
    from pfotgnrec_b200._lib import ptr
    for N in (1000, 32768, 32769, 100003, 400003):    # one-CTA path up to 32 768 nodes, three-kernel path beyond
        rng = np.random.default_rng(N)
        ids = rng.integers(0, N, size=5000).astype(np.int32)
        ids[:10] = 0
        t = torch.as_tensor(ids, device=DEV)
        bitmap = torch.zeros((N + 31) // 32, dtype=torch.int32, device=DEV)
        ws = torch.zeros(_lib.query("pfo_compact_workspace_ints", N), dtype=torch.int32, device=DEV)
        uniq = torch.zeros(5000, dtype=torch.int32, device=DEV)
        slot = torch.zeros(N, dtype=torch.int32, device=DEV)
        cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
        _lib.call("pfo_mark_nodes", ptr(t), t.numel(), 1, ptr(bitmap))
        _lib.call("pfo_compact_nodes", ptr(bitmap), N, ptr(ws), ptr(uniq), ptr(slot), ptr(cnt))
        ref = np.unique(ids[ids > 0])
        assert int(cnt.item()) == ref.shape[0]
        assert np.array_equal(uniq.cpu().numpy()[:ref.shape[0]], ref)
        assert not bitmap.any()
        out = torch.empty_like(t)
        _lib.call("pfo_map_slots", ptr(t), t.numel(), 1, ptr(slot), ptr(out))
        o = out.cpu().numpy()
        assert (o[ids == 0] == -1).all()
        assert np.array_equal(ref[o[ids > 0]], ids[ids > 0])


# ------------------------------------------------------------------------------ dense contractions
@pytest.mark.parametrize("M,N,K", [(1000, 192, 193), (257, 64, 64), (4096, 129, 64), (33, 130, 32), (5000, 64, 192)])
@pytest.mark.parametrize("wt", [0, 1])
def test_linear_f32(M, N, K, wt):
    from pfotgnrec_b200 import _lib
    from pfotgnrec_b200._lib import ptr
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M + 50, K + 3, generator=g)
    W = torch.randn(N, K, generator=g)
    b = torch.randn(N, generator=g)
    idx = torch.randint(-1, M + 50, (M,), generator=g, dtype=torch.int32)
    rz = (torch.rand(M, generator=g) < 0.1).to(torch.int32)
    ref_rows = torch.where(idx.unsqueeze(1) >= 0, A[idx.clamp(min=0).long(), :K], torch.zeros(1))
    ref = (ref_rows.double() @ W.double().t() + b.double()) * 0.5
    ref = torch.relu(ref)
    ref[rz != 0] = 0
    Ad, Wd, bd, idxd, rzd = (t.to(DEV) for t in (A, W.t().contiguous() if wt else W, b, idx, rz))
    C = torch.full((M, N + 2), 7.0, device=DEV)
    _lib.call("pfo_linear_f32", ptr(Ad), K + 3, ptr(idxd), ptr(Wd), N if wt else K, wt, ptr(bd), None, 0,
              ptr(C), N + 2, M, None, N, K, 0.5, 1, ptr(rzd), None, 0, 0)
    assert rel_err(C[:, :N].cpu().numpy(), ref.numpy()) < 1e-5
    assert (C[:, N:] == 7.0).all()
    # accumulate + device-side row count
    md = torch.tensor([M - 5], dtype=torch.int32, device=DEV)
    C2 = torch.ones(M, N, device=DEV)
    _lib.call("pfo_linear_f32", ptr(Ad), K + 3, None, ptr(Wd), N if wt else K, wt, None, None, 0,
              ptr(C2), N, M, ptr(md), N, K, 1.0, 0, None, None, 0, 1)
    ref2 = A[:M, :K].double() @ W.double().t() + 1.0
    assert rel_err(C2[:M - 5].cpu().numpy(), ref2[:M - 5].numpy()) < 1e-5
    assert (C2[M - 5:] == 1.0).all()


@pytest.mark.parametrize("M,N,K", [(3000, 192, 193), (100, 64, 130), (70000, 64, 64), (1, 128, 129)])
def test_wgrad_f32(M, N, K):
    from pfotgnrec_b200 import _lib
    from pfotgnrec_b200._lib import ptr
    g = torch.Generator(device="cpu").manual_seed(M)
    G = torch.randn(M, N, generator=g)
    A = torch.randn(M, K, generator=g)
    ref = G.double().t() @ A.double()
    refb = G.double().sum(0)
    Gd, Ad = G.to(DEV), A.to(DEV)
    dW = torch.zeros(N, K, device=DEV)
    db = torch.zeros(N, device=DEV)
    ws = torch.empty(_lib.query("pfo_wgrad_workspace_floats", M, N, K, 1), device=DEV)
    _lib.call("pfo_wgrad_f32", ptr(Gd), N, ptr(Ad), K, None, M, None, N, K, ptr(dW), K, ptr(db), 0, ptr(ws))
    assert rel_err(dW.cpu().numpy(), ref.numpy()) < 2e-5
    assert rel_err(db.cpu().numpy(), refb.numpy()) < 2e-5
    dW2 = dW.clone()
    _lib.call("pfo_wgrad_f32", ptr(Gd), N, ptr(Ad), K, None, M, None, N, K, ptr(dW2), K, None, 1, ptr(ws))
    assert rel_err(dW2.cpu().numpy(), 2 * ref.numpy()) < 2e-5
    # determinism: bit-identical on a re-run
    dW3 = torch.zeros(N, K, device=DEV)
    _lib.call("pfo_wgrad_f32", ptr(Gd), N, ptr(Ad), K, None, M, None, N, K, ptr(dW3), K, None, 0, ptr(ws))
    assert torch.equal(dW3, dW)


# ------------------------------------------------------------------------------ K5
def _mv_inputs(st, sl):
    ptr_ = st.port_ptr[sl.start:sl.stop + 1]
    return ptr_ - ptr_[0], st.port_items[ptr_[0]:ptr_[-1]]


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_mv_select_golden_ids_bit_exact(ci):
    from pfotgnrec_b200.sampler import MVSelector
    from pfotgnrec_b200.synth import log_returns
    z = load_golden("mv_select.npz")
    cand = z[f"c{ci}_cand"]
    B, C = cand.shape
    e0 = int(z[f"c{ci}_event0"])
    ptr_ = z["st_port_ptr"][e0:e0 + B + 1]
    sel = MVSelector(log_returns(z["st_prices_future"]), np.arange(60), n_users=0, gamma=float(z[f"c{ci}_gamma"]),
                     lam=float(z[f"c{ci}_lam"]), n_candidates=C - 1)
    pp, pn = sel.select(np.arange(B), z["st_day_idx"][e0:e0 + B], cand[:, 0] + 1, ptr_ - ptr_[0],
                        z["st_port_items"][ptr_[0]:ptr_[-1]], cand=cand)
    assert np.array_equal(pp.cpu().numpy() - 1, z[f"c{ci}_ppos_stable"])
    assert np.array_equal(pn.cpu().numpy() - 1, z[f"c{ci}_pneg_stable"])


def test_mv_select_sampled_vs_oracle():
    from oracle import sampling
    from pfotgnrec_b200.sampler import MVSelector
    from pfotgnrec_b200.synth import log_returns
    st = _stream(U=500, I=300, E=4000, mode="nbg", seed=5)
    lr = log_returns(st.prices_future)
    universe = np.unique(st.destinations[:3000] - st.n_users - 1)
    sl = slice(1000, 1000 + 777)
    pptr, pitems = _mv_inputs(st, sl)
    ev = st.edge_idxs[sl]
    for K, lam in ((20, 0.5), (31, 0.3), (5, 0.9)):
        sel = MVSelector(lr, universe, n_users=st.n_users, gamma=2.0, lam=lam, n_candidates=K, seed=123)
        pp, pn, cand, y = sel.select(ev, st.day_idx[sl], st.destinations[sl], pptr, pitems, return_scores=True)
        ref_neg = sampling.sample_candidates(ev, universe, pptr, pitems, K, seed=123)
        ref_cand = np.concatenate([(st.destinations[sl] - st.n_users - 1)[:, None], ref_neg], axis=1)
        assert np.array_equal(cand.cpu().numpy(), ref_cand)          # Philox candidates: bit-exact
        ry = sampling.mv_scores(lr, st.day_idx[sl], ref_cand, pptr, pitems, 2.0)
        assert np.array_equal(y.cpu().numpy(), ry)                   # fp64 y_mv: bit-exact (same op order, no fma)
        rp, rn = sampling.mv_select(lr, st.day_idx[sl], ref_cand, pptr, pitems, 2.0, lam)
        assert np.array_equal(pp.cpu().numpy() - st.n_users - 1, rp)
        assert np.array_equal(pn.cpu().numpy() - st.n_users - 1, rn)


@pytest.mark.parametrize("size", [3, 20, 40, 45])
def test_candidate_sampler_vs_oracle(size):
    from oracle import sampling
    from pfotgnrec_b200.sampler import CandidateSampler
    st = _stream(U=200, I=40, E=2000, mode="small", seed=9)     # 40 items: size 40/45 hits the replacement path
    universe = np.unique(st.destinations)
    sl = slice(100, 400)
    pptr, pitems = _mv_inputs(st, sl)
    held_items = pitems.astype(np.int64) + st.n_users + 1
    cs = CandidateSampler(universe)
    out = cs.sample(st.edge_idxs[sl], pptr, held_items, size, seed=2024).cpu().numpy()
    ref = sampling.sample_candidates(st.edge_idxs[sl], universe, pptr, held_items, size, seed=2024)
    assert np.array_equal(out, ref)
    for b in range(out.shape[0]):
        held = set(held_items[pptr[b]:pptr[b + 1]].tolist())
        assert not (set(out[b].tolist()) & held)


def test_candidate_sampler_support_and_replace_rule_vs_reference():
    """K5's sampling kernel against the reference RandEdgeSampler's own support / replace rule / uniformity
    (tests/golden/sampler_support.npz, utils/utils.py:73-113) -- the same checker the oracle passes on CPU."""
    from test_oracle_golden import check_sampler_support
    from pfotgnrec_b200.sampler import CandidateSampler
    z = load_golden("sampler_support.npz")
    cs = {}

    def sample(ev, items, pptr, held, size, seed):
        c = cs.setdefault(id(items), CandidateSampler(items))
        return c.sample(ev, pptr, held, int(size), seed=int(seed)).cpu().numpy().astype(np.int64)

    check_sampler_support(z, sample)


# ------------------------------------------------------------------------------ K6 / eval
def test_bpr_forward_backward():
    from oracle.tgn import bpr_loss
    from pfotgnrec_b200 import _lib
    from pfotgnrec_b200._lib import ptr
    torch.manual_seed(0)
    B, k, d = 517, 3, 64
    eu, ep, en = torch.randn(B, d), torch.randn(B, d), torch.randn(B * k, d)
    for t in (eu, ep, en):
        t.mul_(0.3).requires_grad_(True)
    loss = bpr_loss(eu, ep, en)
    loss.backward()
    du, dp, dn = (torch.empty_like(t, device=DEV) for t in (eu, ep, en))
    out = torch.zeros(1, device=DEV)
    ws = torch.empty(1024, device=DEV)
    eu_d, ep_d, en_d = eu.detach().to(DEV), ep.detach().to(DEV), en.detach().to(DEV)   # keep alive across the launch
    _lib.call("pfo_bpr", ptr(eu_d), ptr(ep_d), ptr(en_d), B, k, d,
              ptr(du), ptr(dp), ptr(dn), ptr(out), 1.0, ptr(ws))
    assert abs(out.item() - loss.item()) < 1e-5 * abs(loss.item())
    assert rel_err(du.cpu().numpy(), eu.grad.numpy()) < 1e-5
    assert rel_err(dp.cpu().numpy(), ep.grad.numpy()) < 1e-5
    assert rel_err(dn.cpu().numpy(), en.grad.numpy()) < 1e-5


def test_eval_score_and_ranking():
    from oracle.tgn import eval_scores, eval_ranking
    from pfotgnrec_b200 import _lib
    from pfotgnrec_b200._lib import ptr
    torch.manual_seed(1)
    B, n_cand, d = 37, 203, 64
    es, ed, ec = torch.randn(B, d), torch.randn(B, d), torch.randn(B * n_cand, d)
    ec.view(B, n_cand, d)[:, 5] = ec.view(B, n_cand, d)[:, 9]        # exact ties among candidates
    ec.view(B, n_cand, d)[::2, 3] = ed[::2]                           # ties with the positive
    ref = eval_scores(es, ed, ec)
    rank = eval_ranking(ref.numpy())
    scores = torch.empty(B, 1 + n_cand, device=DEV)
    pos_rank = torch.empty(B, dtype=torch.int32, device=DEV)
    top = torch.empty(B, 5, dtype=torch.int32, device=DEV)
    es_d, ed_d, ec_d = es.to(DEV), ed.to(DEV), ec.to(DEV)
    _lib.call("pfo_eval_score", ptr(es_d), ptr(ed_d), ptr(ec_d), B, n_cand, d, 5,
              ptr(scores), ptr(pos_rank), ptr(top))
    s = scores.cpu().numpy()
    assert rel_err(s, ref.numpy()) < 1e-5
    # ranking semantics checked on the kernel's own scores (bit-exact integer work)
    rk = eval_ranking(s)
    assert np.array_equal(top.cpu().numpy(), rk[:, :5])
    assert np.array_equal(pos_rank.cpu().numpy(), np.argmax(rk == 0, axis=1))


def _eval_metrics_on_device(scores_inputs, pos_item, cand_item, item_offset, day_idx, port_ptr, port_items, lr_past,
                            lr_future, acc=None):
    """pfo_eval_score -> pfo_eval_metrics on the device; returns (scores, pos_rank, top5, per_event)."""
    from pfotgnrec_b200 import _lib
    from pfotgnrec_b200._lib import ptr
    es, ed, ec = (t.to(DEV).contiguous() for t in scores_inputs)
    B, d = es.shape
    N = ec.shape[0] // B
    scores = torch.empty(B, 1 + N, device=DEV)
    pos_rank = torch.empty(B, dtype=torch.int32, device=DEV)
    top = torch.empty(B, 5, dtype=torch.int32, device=DEV)
    _lib.call("pfo_eval_score", ptr(es), ptr(ed), ptr(ec), B, N, d, 5, ptr(scores), ptr(pos_rank), ptr(top))
    i32 = lambda a: torch.as_tensor(np.ascontiguousarray(a).astype(np.int32), device=DEV)
    pi = i32(port_items if len(port_items) else np.zeros(1))
    pp = torch.as_tensor(np.ascontiguousarray(port_ptr).astype(np.int64), device=DEV)
    pos_d, cand_d, day_d = i32(pos_item), i32(cand_item), i32(day_idx)
    per_event = torch.full((B, 18), float("nan"), dtype=torch.float64, device=DEV)
    _lib.call("pfo_eval_metrics", ptr(pos_rank), ptr(top), 5, ptr(pos_d), ptr(cand_d), N, int(item_offset), ptr(day_d),
              ptr(pp), ptr(pi), ptr(lr_past), ptr(lr_future), lr_past.shape[1], lr_past.shape[2], B, ptr(per_event),
              ptr(acc))
    return scores.cpu().numpy(), pos_rank.cpu().numpy(), top.cpu().numpy(), per_event.cpu().numpy()


def test_eval_metric_block_golden_and_oracle():
    """pfo_eval_metrics (reference evaluation.py:127-258) on the inputs of the reference run behind
    tests/golden/eval_metrics.npz: per-interaction metrics BIT-EXACT against the numpy oracle on the kernel's own
    ranking, the 30 aggregated keys against the dictionary the unmodified reference returned."""
    from oracle import eval_metrics as em
    from pfotgnrec_b200.synth import log_returns
    z = load_golden("eval_metrics.npz")
    table = torch.tensor(z["table"])
    e0, B, U = int(z["e0"]), int(z["B"]), int(z["st_n_users"])
    lrp = torch.as_tensor(log_returns(z["st_prices_past"]), device=DEV).contiguous()
    lrf = torch.as_tensor(log_returns(z["st_prices_future"]), device=DEV).contiguous()
    acc = torch.zeros(31, dtype=torch.float64, device=DEV)
    rows = []
    for bi in range(int(z["n_batches"])):
        s, e = e0 + bi * B, e0 + (bi + 1) * B
        src, dst, neg = z["st_sources"][s:e], z["st_destinations"][s:e], z["negatives"][bi]
        ptr_ = z["st_port_ptr"][s:e + 1]
        held = z["st_port_items"][ptr_[0]:ptr_[-1]]
        emb = (table[torch.as_tensor(src)], table[torch.as_tensor(dst)], table[torch.as_tensor(neg.reshape(-1))])
        sc, rk, top, pe = _eval_metrics_on_device(emb, dst, neg, U + 1, z["st_day_idx"][s:e], ptr_ - ptr_[0], held,
                                                  lrp, lrf, acc)
        ref, ork, otop = em.per_event_metrics(sc, dst - U - 1, neg - U - 1, z["st_day_idx"][s:e], ptr_ - ptr_[0], held,
                                              z["st_prices_past"], z["st_prices_future"])
        assert np.array_equal(rk, ork) and np.array_equal(top, otop)
        assert np.array_equal(pe, ref), np.abs(pe - ref).max()          # fp64, rounded like numpy: bit-exact
        rows.append(pe)
    allrows = np.concatenate(rows)
    a = acc.cpu().numpy()
    assert a[30] == allrows.shape[0]
    assert np.allclose(a[:18], allrows.sum(axis=0), rtol=1e-13, atol=1e-15)
    assert np.array_equal(a[18:30], (allrows[:, 6:] > 0).sum(axis=0))
    got = em.aggregate(allrows, "val")
    for k, v in z.items():
        if k.startswith("res_stable_"):
            assert abs(got[k[len("res_stable_"):]] - float(v)) <= 1e-12 * max(1.0, abs(float(v))), k


def test_eval_metric_block_edge_cases():
    """Empty portfolios for every interaction, portfolios of 1..5 stocks, the true item ranked first / last,
    ragged batch (not a multiple of the warps per CTA), T < 8 and T = 32 return columns (numpy's two summation
    regimes)."""
    from oracle import eval_metrics as em
    rng = np.random.default_rng(7)
    for T1, B, I, maxp in ((30, 77, 25, 5), (6, 33, 12, 0), (33, 130, 40, 3)):
        D = 4
        prices = [np.exp(np.cumsum(rng.standard_normal((D, I, T1)) * 0.02, axis=2)) * 50.0 for _ in range(2)]
        plen = rng.integers(0, maxp + 1, size=B)
        pp = np.r_[0, np.cumsum(plen)].astype(np.int64)
        held = np.concatenate([rng.choice(I, size=n, replace=False) for n in plen] + [np.zeros(0, np.int64)]).astype(np.int64)
        day = rng.integers(0, D, size=B)
        N, d = 19, 16
        es, ed, ec = torch.randn(B, d), torch.randn(B, d), torch.randn(B * N, d)
        ed[:5] = es[:5] * 10.0                     # true item first
        ed[5:10] = -es[5:10] * 10.0                # true item last
        pos = rng.integers(0, I, size=B) + 100
        cand = rng.integers(0, I, size=(B, N)) + 100
        lrp, lrf = (torch.as_tensor(np.log(p[..., 1:] / p[..., :-1]), device=DEV).contiguous() for p in prices)
        sc, rk, top, pe = _eval_metrics_on_device((es, ed, ec), pos, cand, 100, day, pp, held, lrp, lrf)
        ref, ork, otop = em.per_event_metrics(sc, pos - 100, cand - 100, day, pp, held, prices[0], prices[1])
        assert np.array_equal(rk, ork) and np.array_equal(top, otop)
        assert (rk[:5] == 0).all() and (rk[5:10] == N).all()
        assert np.array_equal(pe, ref), (T1, np.abs(pe - ref).max())


# ------------------------------------------------------------------------------ bf16 tcgen05 path
@pytest.mark.parametrize("M,N,K", [(1000, 192, 193), (257, 64, 64), (4096, 129, 64), (33, 130, 32), (5000, 64, 192),
                                   (300, 128, 128), (70000, 192, 64)])
@pytest.mark.parametrize("wt", [0, 1])
def test_linear_bf16_tcgen05(M, N, K, wt):
    from pfotgnrec_b200 import _lib
    from pfotgnrec_b200._lib import ptr
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    lda = K + 3 if K % 2 else K          # exercise both the scalar and the float4 staging path
    A = torch.randn(M + 50, lda, generator=g)
    W = torch.randn(N, K, generator=g)
    b = torch.randn(N, generator=g)
    idx = torch.randint(-1, M + 50, (M,), generator=g, dtype=torch.int32)
    rz = (torch.rand(M, generator=g) < 0.1).to(torch.int32)
    rows = torch.where(idx.unsqueeze(1) >= 0, A[idx.clamp(min=0).long(), :K], torch.zeros(1))
    # operands rounded to bf16, exact products, fp32-ish accumulation: a sharp check of the layouts
    ref = (rows.bfloat16().double() @ W.bfloat16().double().t() + b.double()) * 0.5
    ref = torch.relu(ref)
    ref[rz != 0] = 0
    full = torch.relu((rows.double() @ W.double().t() + b.double()) * 0.5)
    full[rz != 0] = 0
    Ad, Wd, bd, idxd, rzd = (t.to(DEV) for t in (A, W.t().contiguous() if wt else W, b, idx, rz))
    C = torch.full((M, N + 2), 7.0, device=DEV)
    _lib.call("pfo_linear_bf16", ptr(Ad), lda, ptr(idxd), ptr(Wd), N if wt else K, wt, ptr(bd), None, 0,
              ptr(C), N + 2, M, None, N, K, 0.5, 1, ptr(rzd), None, 0, 0)
    got = C[:, :N].cpu().numpy()
    assert rel_err(got, ref.numpy()) < 1e-5
    assert rel_err(got, full.numpy()) < 2e-2           # the bf16 contract of BASELINE.json
    assert (C[:, N:] == 7.0).all()


# ------------------------------------------------------------------------------ TMA + tcgen05 (tf32 / 3xtf32) path
def _tf32_trunc(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("M,N,K", [(1000, 192, 193), (257, 64, 64), (4096, 129, 64), (33, 130, 32), (5000, 64, 192),
                                   (300, 128, 128), (70000, 192, 64), (1, 16, 1), (129, 256, 320), (40000, 64, 130),
                                   (3000, 264, 64), (3000, 64, 328), (3000, 328, 64), (3000, 64, 264)])
@pytest.mark.parametrize("wt", [0, 1])
@pytest.mark.parametrize("passes", [3, 1])
def test_linear_tf32_tma(M, N, K, wt, passes):
    from pfotgnrec_b200 import _lib
    from pfotgnrec_b200._lib import ptr
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    lda = (K + 3) // 4 * 4 + 4                      # strided rows, 16-byte aligned: the TMA path
    A = torch.randn(M + 50, lda, generator=g)
    W = torch.randn(N, K, generator=g)
    b = torch.randn(N, generator=g)
    rz = (torch.rand(M, generator=g) < 0.1).to(torch.int32)
    brs = torch.rand(M, generator=g)
    full = (A[:M, :K].double() @ W.double().t() + b.double() * brs.double().unsqueeze(1)) * 0.5
    full = torch.relu(full)
    full[rz != 0] = 0
    Ad, Wd, bd, rzd, brsd = (t.to(DEV) for t in (A, W.t().contiguous() if wt else W, b, rz, brs))
    C = torch.full((M, N + 4), 7.0, device=DEV)
    _lib.call("pfo_linear_tf32", ptr(Ad), lda, None, ptr(Wd), N if wt else K, wt, ptr(bd), ptr(brsd), 1,
              ptr(C), N + 4, M, None, N, K, 0.5, 1, ptr(rzd), None, 0, 0, passes)
    got = C[:, :N].cpu().numpy()
    assert rel_err(got, full.numpy()) < (1e-5 if passes == 3 else 2e-3)
    assert (C[:, N:] == 7.0).all()
    # accumulate + relu gate + device-side row count, unaligned output stride (scalar store path)
    live = max(M - 5, 1)
    md = torch.tensor([live], dtype=torch.int32, device=DEV)
    gate = torch.randn(M, N + 1, generator=g)
    C2 = torch.ones(M, N + 1, device=DEV)
    _lib.call("pfo_linear_tf32", ptr(Ad), lda, None, ptr(Wd), N if wt else K, wt, None, None, 0,
              ptr(C2), N + 1, M, ptr(md), N, K, 1.0, 0, None, ptr(gate.to(DEV)), N + 1, 1, passes)
    ref2 = A[:M, :K].double() @ W.double().t()
    ref2[gate[:, :N] <= 0] = 0
    ref2 = ref2 + 1.0
    assert rel_err(C2[:live, :N].cpu().numpy(), ref2[:live].numpy()) < (1e-5 if passes == 3 else 2e-3)
    assert (C2[live:] == 1.0).all() and (C2[:, N] == 1.0).all()


@pytest.mark.parametrize("M,N,K", [(3000, 192, 193), (100, 64, 130), (70000, 64, 64), (1, 128, 129), (5000, 128, 128),
                                   (777, 64, 192), (49152, 128, 64), (31, 192, 64)])
@pytest.mark.parametrize("passes", [3, 1])
def test_wgrad_tf32_tma(M, N, K, passes):
    from pfotgnrec_b200 import _lib
    from pfotgnrec_b200._lib import ptr
    g = torch.Generator(device="cpu").manual_seed(M)
    ldg, lda = N + 4, (K + 3) // 4 * 4
    G = torch.randn(M + 40, ldg, generator=g)
    A = torch.randn(M + 40, lda, generator=g)
    G[M:] = float("nan")                             # rows past M must never be read into the sums
    A[M:] = float("nan")
    ref = G[:M, :N].double().t() @ A[:M, :K].double()
    refb = G[:M, :N].double().sum(0)
    tol = 2e-5 if passes == 3 else 5e-3
    Gd, Ad = G.to(DEV), A.to(DEV)
    dW = torch.zeros(N, K, device=DEV)
    db = torch.zeros(N, device=DEV)
    ws = torch.empty(_lib.query("pfo_wgrad_tf32_workspace_floats", M + 40, N, K, 1), device=DEV)
    _lib.call("pfo_wgrad_tf32", ptr(Gd), ldg, ptr(Ad), lda, None, M, None, N, K, ptr(dW), K, ptr(db), 0, ptr(ws), passes)
    assert rel_err(dW.cpu().numpy(), ref.numpy()) < tol
    assert rel_err(db.cpu().numpy(), refb.numpy()) < tol
    # accumulate, no bias, live row count on the device (capacity M + 40, rows past M hold NaN)
    md = torch.tensor([M], dtype=torch.int32, device=DEV)
    dW2 = dW.clone()
    _lib.call("pfo_wgrad_tf32", ptr(Gd), ldg, ptr(Ad), lda, None, M + 40, ptr(md), N, K, ptr(dW2), K, None, 1, ptr(ws), passes)
    assert rel_err(dW2.cpu().numpy(), 2 * ref.numpy()) < tol
    # determinism: bit-identical on a re-run
    dW3 = torch.zeros(N, K, device=DEV)
    _lib.call("pfo_wgrad_tf32", ptr(Gd), ldg, ptr(Ad), lda, None, M, None, N, K, ptr(dW3), K, None, 0, ptr(ws), passes)
    assert torch.equal(dW3, dW)


# ------------------------------------------------------------------------------ attention operand folding
def _fold_reference(cfg, Wq, Wk, Wv, b_in, Wo, bo, W1, b1, tb):
    """fp64 torch restatement of the fold (differentiable): what the per-query kernels assume about Wqk / cqk / Wc1T,
    derived from nn.MultiheadAttention + MergeLayer (model/temporal_attention.py:52-90, utils/utils.py:14-17)."""
    import math
    d, E, Ek, H, ekp = cfg.d, cfg.E, cfg.Ek, cfg.n_heads, cfg.ekp
    hd = E // H
    scale = 1.0 / math.sqrt(hd)
    D = torch.float64
    Wq, Wk, Wv, b_in, Wo, bo, W1, b1, tb = (t.to(D) for t in (Wq, Wk, Wv, b_in, Wo, bo, W1, b1, tb))
    te0 = torch.cos(tb)
    cq = Wq[:, d:] @ te0 + b_in[:E]
    WkT = Wk.view(H, hd, Ek).transpose(1, 2)
    A = scale * (WkT @ Wq[:, :d].reshape(H, hd, d))
    cA = scale * (WkT @ cq.view(H, hd, 1)).squeeze(2)
    Wqk = torch.nn.functional.pad(A, (0, 0, 0, ekp - Ek)).reshape(H * ekp, d)
    cqk = torch.nn.functional.pad(cA, (0, ekp - Ek)).reshape(H * ekp)
    WvA = torch.cat([Wv.view(H, hd, Ek), b_in[2 * E:].view(H, hd, 1)], dim=2)
    WoH = Wo.view(E, H, hd).permute(1, 0, 2)
    W1a, W1b = W1[:, :E], W1[:, E:]
    Bh = W1a @ (WoH @ WvA)
    tail0 = torch.cat([(W1a @ bo).unsqueeze(0), b1.unsqueeze(0), Bh.new_zeros(ekp - Ek - 3, d)], dim=0)
    tailz = Bh.new_zeros(ekp - Ek - 1, d)
    rows = []
    for h in range(H):
        rows += [Bh[h].t(), tail0 if h == 0 else tailz]
    return Wqk, cqk, torch.cat(rows + [W1b.t()], dim=0)


@pytest.mark.parametrize("d,F,H", [(64, 1, 2), (32, 3, 2), (64, 1, 4), (32, 2, 1)])
def test_fold_attention_forward_and_adjoint(d, F, H):
    """pfo_fold_attention_fwd / _bwd against the fp64 torch restatement and its autograd gradients."""
    from pfotgnrec_b200.engine import ModelConfig, _FoldAttention
    cfg = ModelConfig(d=d, n_edge_feat=F, n_heads=H)
    g = torch.Generator(device="cpu").manual_seed(d + F + H)
    E, Ek = 2 * d, 2 * d + F
    shapes = [(E, E), (E, Ek), (E, Ek), (3 * E,), (E, E), (E,), (d, E + d), (d,), (d,)]
    w = [(torch.randn(*s, generator=g) * 0.3).to(DEV).requires_grad_(True) for s in shapes]
    w64 = [t.detach().double().requires_grad_(True) for t in w]
    ref = _fold_reference(cfg, *w64)
    got = _FoldAttention.apply(cfg, *w)
    for a, b in zip(got, ref):
        assert a.shape == b.shape
        assert rel_err(a.detach().cpu().numpy(), b.detach().cpu().numpy()) < 1e-6
    cot = [torch.randn(t.shape, generator=g).to(DEV) for t in got]
    torch.autograd.backward(got, cot)
    torch.autograd.backward(ref, [c.double() for c in cot])
    names = ["Wq", "Wk", "Wv", "b_in", "Wo", "bo", "W1", "b1", "tb"]
    for nm, a, b in zip(names, w, w64):
        assert rel_err(a.grad.cpu().numpy(), b.grad.cpu().numpy()) < 1e-6, nm


# ------------------------------------------------------------------------------ K4 neighbour kernels
def _nbr_restatement(QK, T, idx, eidx, dt, ef, tw, tb, d, F, H, ekp):
    """fp64 torch restatement of the neighbour-level kernel (csrc/attention_kernels.cu header): per query and head,
    x_j = [T[idx_j] | ef[eidx_j] | cos(dt_j w + b)], masked softmax of qk_h . x_j over the live slots,
    XB_h = [sum_j p_j x_j (as h | e | te) | sum_j p_j | valid | one | 0..]."""
    Q, n = idx.shape
    live = idx >= 0
    h = T[idx.clamp(min=0).long()]                                    # [Q, n, d]
    e = ef[eidx.long()]                                               # [Q, n, F]
    te = torch.cos(dt.unsqueeze(-1) * tw + tb)                         # [Q, n, d]
    x = torch.cat([h, e, te], dim=-1) * live.unsqueeze(-1)             # [Q, n, 2d + F]
    s = torch.einsum("qhk,qnk->qhn", QK[:, :, :2 * d + F], x)
    s = s.masked_fill(~live.unsqueeze(1), float("-inf"))
    any_live = live.any(dim=1)
    p = torch.softmax(s, dim=-1)
    p = torch.where(any_live.view(Q, 1, 1), p, torch.zeros_like(p))
    xb = torch.einsum("qhn,qnk->qhk", p, x)
    tail = torch.zeros(Q, H, ekp - (2 * d + F), dtype=xb.dtype, device=xb.device)
    tail[:, :, 0] = p.sum(-1)
    tail[:, :, 1] = any_live.to(xb.dtype).view(Q, 1)
    tail[:, :, 2] = 1.0
    return torch.cat([xb, tail], dim=-1), p


@pytest.mark.parametrize("d,F,H,n", [(64, 1, 2, 10), (32, 3, 2, 7), (64, 1, 4, 20), (128, 2, 1, 5)])
def test_attn_nbr_forward_backward_vs_torch(d, F, H, n):
    """pfo_attn_nbr_fwd / _bwd against an fp64 torch restatement and its autograd (1e-5 of max-norm), the shared-row
    entry point pfo_attn_nbr_fwd_rows bit-identical to per-query rows, strided feature tables and a strided XB."""
    from pfotgnrec_b200 import _lib
    from pfotgnrec_b200._lib import ptr
    g = torch.Generator(device="cuda").manual_seed(d + F + H + n)
    Q, R, NE = 1037, 400, 77
    ekp = (2 * d + F + 3 + 3) // 4 * 4
    ldt, ldxb = d + 4, H * ekp + d                                    # table rows and XB rows with a stride
    Tfull = torch.randn(R, ldt, device=DEV, generator=g)
    T = Tfull[:, :d]
    QK = torch.randn(Q, H, ekp, device=DEV, generator=g) * 0.3
    idx = torch.randint(-1, R, (Q, n), device=DEV, generator=g, dtype=torch.int32)
    idx[5] = -1                                                       # a query without neighbours
    idx[6, : n - 1] = -1
    eidx = torch.randint(0, NE, (Q, n), device=DEV, generator=g, dtype=torch.int32)
    dt = torch.rand(Q, n, device=DEV, generator=g) * 10               # small arguments: fp32 fmaf(dt, w, b) ~ the fp64 one
    ef = torch.randn(NE, F, device=DEV, generator=g)
    tw, tb = torch.rand(d, device=DEV, generator=g), torch.rand(d, device=DEV, generator=g)

    def fwd(qk, rows=None):
        XB = torch.full((Q, ldxb), 7.0, device=DEV)
        P = torch.empty(Q, H, n, device=DEV)
        inv = torch.empty(Q, dtype=torch.int32, device=DEV)
        if rows is None:
            _lib.call("pfo_attn_nbr_fwd", ptr(qk), ptr(Tfull), ldt, ptr(idx), ptr(eidx), ptr(dt), ptr(ef), ptr(tw), ptr(tb),
                      Q, n, d, F, H, ekp, 0.0, 0, 0, None, ptr(XB), ldxb, ptr(P), ptr(inv))
        else:
            _lib.call("pfo_attn_nbr_fwd_rows", ptr(qk), ptr(rows), ptr(Tfull), ldt, ptr(idx), ptr(eidx), ptr(dt), ptr(ef),
                      ptr(tw), ptr(tb), Q, n, d, F, H, ekp, 0.0, 0, 0, None, ptr(XB), ldxb, ptr(P), ptr(inv))
        return XB, P, inv

    XB, P, inv = fwd(QK)
    leaf = [t.double().detach().requires_grad_(True) for t in (QK, T, tw, tb)]
    ref, pref = _nbr_restatement(leaf[0], leaf[1], idx, eidx, dt.double(), ef.double(), leaf[2], leaf[3], d, F, H, ekp)
    got = XB[:, :H * ekp].view(Q, H, ekp)
    assert rel_err(got.cpu().numpy(), ref.detach().cpu().numpy()) < 1e-5
    assert rel_err(P.cpu().numpy(), pref.detach().cpu().numpy()) < 1e-5
    assert torch.equal(inv.bool(), ~(idx >= 0).any(dim=1))
    assert float((XB[:, H * ekp:] - 7.0).abs().max()) == 0.0            # the h_query slot of the row is not touched
    # shared rows: 50 distinct query operands, query q reads row q % 50
    tab = QK[:50].contiguous()
    rows = (torch.arange(Q, device=DEV) % 50).to(torch.int32)
    XBr, Pr, _ = fwd(tab, rows)
    XBe, Pe, _ = fwd(tab[rows.long()].contiguous())
    assert torch.equal(XBr, XBe) and torch.equal(Pr, Pe)
    # backward against autograd of the restatement
    G = torch.randn(Q, ldxb, device=DEV, generator=g)
    (ref * G[:, :H * ekp].view(Q, H, ekp).double()).sum().backward()
    dQK = torch.empty(Q, H, ekp, device=DEV)
    dT = torch.zeros(R, ldt, device=DEV)
    gwb = torch.zeros(2 * d, device=DEV)
    ws = torch.empty(int(_lib.query("pfo_attn_nbr_bwd_workspace_floats", d)), device=DEV)
    _lib.call("pfo_attn_nbr_bwd", ptr(QK), ptr(G), ldxb, ptr(P), ptr(inv), ptr(Tfull), ldt, ptr(idx), ptr(eidx), ptr(dt),
              ptr(ef), ptr(tw), ptr(tb), Q, n, d, F, H, ekp, 0.0, 0, 0, None, ptr(dQK), ptr(dT), ldt, ptr(gwb), 0, ptr(ws))
    k = 2 * d + F
    assert rel_err(dQK[:, :, :k].cpu().numpy(), leaf[0].grad[:, :, :k].cpu().numpy()) < 1e-5
    assert float(dQK[:, :, k:].abs().max()) == 0.0
    assert rel_err(dT[:, :d].cpu().numpy(), leaf[1].grad.cpu().numpy()) < 1e-5
    assert float(dT[:, d:].abs().max()) == 0.0
    assert rel_err(gwb[:d].cpu().numpy(), leaf[2].grad.cpu().numpy()) < 2e-5
    assert rel_err(gwb[d:].cpu().numpy(), leaf[3].grad.cpu().numpy()) < 2e-5
