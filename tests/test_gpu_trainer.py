"""GPU tests of the training / evaluation driver (pfotgnrec_b200/trainer.py): the step loop of reference
main.py:179-394 against the CPU oracle loop, and the CUDA-graph replay against kernel-by-kernel launches."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import rel_err


def _stream(seed=1):
    from pfotgnrec_b200.synth import make_stream
    return make_stream(n_users=300, n_items=60, n_events=3000, n_days=20, seed=seed, ts_mode="small")


@pytest.mark.parametrize("model", ["ours", "tgn", "jodie"])
def test_trainer_steps_match_oracle_loop(model):
    """Per-step parity of loss / memory / pending flags with weights injected from the CUDA model (Adam turns
    1e-10 gradient noise into 1e-4 updates, so the weights are re-synchronised every step)."""
    from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig
    from oracle.train_loop import OracleTrainer
    st = _stream()
    tc = TrainConfig(model=model, bs=128, lr=1e-4, cuda_graph=False)
    tr = PfoTrainer(st, tc, device="cuda:0")
    p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in tr.tgn.named_parameters()}
    orc = OracleTrainer(st, model, bs=128, params=p, seed=tc.seed)
    for i in range(4):
        s, e = 1000 + i * 128, 1000 + (i + 1) * 128
        for k, v in tr.tgn.named_parameters():
            p[k].data.copy_(v.detach().cpu())
        la = float(tr.train_step(s, e).item())
        lb = orc.train_step(s, e)
        assert abs(la - lb) < 1e-4 * max(1.0, abs(lb)), (i, la, lb)
    assert rel_err(tr.tgn.memory.memory.detach().cpu().numpy(), orc.tgn.memory.numpy()) < 1e-4
    assert np.array_equal(tr.tgn.memory.state.pend_valid.cpu().numpy().astype(bool), orc.tgn.pend_valid.numpy())


@pytest.mark.parametrize("model", ["ours", "jodie"])
def test_cuda_graph_replay_matches_eager(model):
    """One CUDA graph per batch size (captured on the third step) == kernel-by-kernel launches: same losses,
    same memory, bit-identical pending-message flags and timestamps, same weights after Adam."""
    from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig
    st = _stream(seed=2)
    out = []
    for graph in (False, True):
        tr = PfoTrainer(st, TrainConfig(model=model, bs=128, lr=1e-3, cuda_graph=graph), device="cuda:0")
        losses = []
        for i in range(8):
            s = 600 + i * 128
            losses.append(float(tr.train_step(s, s + 128).item()))
        if graph:
            assert tr._graphs[128].graph is not None and tr._graphs[128].launches > 10
        sd = tr.tgn.memory.state
        out.append((losses, sd.memory.cpu().numpy(), sd.pend_valid.cpu().numpy(), sd.pend_ts.cpu().numpy(),
                    {k: v.detach().cpu().numpy() for k, v in tr.tgn.named_parameters()}))
    (la, ma, va, ta, pa), (lb, mb, vb, tb, pb) = out
    assert np.allclose(la, lb, rtol=1e-5, atol=1e-6), (la, lb)
    assert rel_err(mb, ma) < 1e-5
    assert np.array_equal(va, vb) and np.array_equal(ta, tb)
    for k in pa:
        assert rel_err(pb[k], pa[k]) < 1e-4, k


def test_cuda_graph_from_host_batches():
    """train_step_host (pinned host buffers -> static device buffers -> graph replay) == train_step on the
    device-resident stream."""
    from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig
    st = _stream(seed=3)
    res = []
    for host in (False, True):
        tr = PfoTrainer(st, TrainConfig(model="ours", bs=64, cuda_graph=True), device="cuda:0")
        hbs = tr.make_host_batches(500, 6, 64)
        ls = []
        for i in range(6):
            l = tr.train_step_host(hbs[i]) if host else tr.train_step(500 + i * 64, 500 + (i + 1) * 64)
            ls.append(float(l.item()))
        res.append((ls, tr.tgn.memory.state.memory.cpu().numpy()))
    assert np.allclose(res[0][0], res[1][0], rtol=1e-5, atol=1e-6)
    assert rel_err(res[1][1], res[0][1]) < 1e-5


def test_dropout_stream_is_keyed_by_the_device_step_counter():
    """Attention dropout draws from Philox keyed by (query, head * n + slot, step + *step_dev): the host part
    and the device part of the step are interchangeable, and a bumped counter gives a fresh mask."""
    from pfotgnrec_b200 import _lib
    from pfotgnrec_b200._lib import ptr
    g = torch.Generator(device="cuda").manual_seed(0)
    Q, n, d, F, H = 257, 10, 64, 1, 2
    ekp = (2 * d + F + 3 + 3) // 4 * 4
    T = torch.randn(500, d, device="cuda", generator=g)
    QK = torch.randn(Q, H, ekp, device="cuda", generator=g) * 0.2
    idx = torch.randint(-1, 500, (Q, n), device="cuda", generator=g, dtype=torch.int32)
    eidx = torch.randint(0, 50, (Q, n), device="cuda", generator=g, dtype=torch.int32)
    dt = torch.rand(Q, n, device="cuda", generator=g) * 100
    ef = torch.randn(50, F, device="cuda", generator=g)
    tw, tb = torch.rand(d, device="cuda", generator=g), torch.rand(d, device="cuda", generator=g)

    def run(p_drop, step, ctr):
        XB = torch.empty(Q, H, ekp, device="cuda")
        P = torch.empty(Q, H, n, device="cuda")
        inv = torch.empty(Q, dtype=torch.int32, device="cuda")
        c = torch.tensor([ctr], dtype=torch.int32, device="cuda")
        _lib.call("pfo_attn_nbr_fwd", ptr(QK), ptr(T), d, ptr(idx), ptr(eidx), ptr(dt), ptr(ef), ptr(tw), ptr(tb),
                  Q, n, d, F, H, ekp, float(p_drop), 7, step, ptr(c), ptr(XB), H * ekp, ptr(P), ptr(inv))
        torch.cuda.synchronize()
        return XB.cpu().numpy()

    base = run(0.0, 0, 0)
    a = run(0.5, 3, 16)
    assert np.array_equal(a, run(0.5, 19, 0))            # host and device parts of the step add up
    assert not np.array_equal(a, run(0.5, 3, 32))        # next batch: fresh mask
    assert not np.array_equal(a, base)
    psum = a[:, :, 2 * d + F]                            # kept softmax mass, scaled by 1/(1-p): mean ~ 1
    live = (idx.cpu().numpy() >= 0).any(axis=1)
    assert abs(psum[live].mean() - 1.0) < 0.1


def test_dropout_mask_placement_and_expectation():
    """Placement and scaling of the attention dropout against nn.MultiheadAttention's (temporal_attention.py:28-32,
    torch's multi_head_attention_forward: softmax -> dropout(p) on the WEIGHTS -> AV, no renormalisation):
      * every dropped-path weight is either 0 or the undropped softmax weight / (1 - p), and a fraction ~p is dropped;
      * the per-head weighted neighbour sum is linear in the weights, so its mean over many independent masks converges
        to the undropped one (E[mask / (1 - p)] = 1) at the Monte-Carlo rate."""
    from pfotgnrec_b200 import _lib
    from pfotgnrec_b200._lib import ptr
    g = torch.Generator(device="cuda").manual_seed(1)
    Q, n, d, F, H, p = 192, 10, 64, 1, 2, 0.3
    ekp = (2 * d + F + 3 + 3) // 4 * 4
    T = torch.randn(300, d, device="cuda", generator=g)
    QK = torch.randn(Q, H, ekp, device="cuda", generator=g) * 0.1
    idx = torch.randint(0, 300, (Q, n), device="cuda", generator=g, dtype=torch.int32)
    eidx = torch.randint(0, 50, (Q, n), device="cuda", generator=g, dtype=torch.int32)
    dt = torch.rand(Q, n, device="cuda", generator=g) * 100
    ef = torch.randn(50, F, device="cuda", generator=g)
    tw, tb = torch.rand(d, device="cuda", generator=g), torch.rand(d, device="cuda", generator=g)

    def run(p_drop, ctr, idx_=None):
        XB = torch.empty(Q, H, ekp, device="cuda")
        P = torch.empty(Q, H, n, device="cuda")
        inv = torch.empty(Q, dtype=torch.int32, device="cuda")
        c = torch.tensor([ctr], dtype=torch.int32, device="cuda")
        ix = idx if idx_ is None else idx_
        _lib.call("pfo_attn_nbr_fwd", ptr(QK), ptr(T), d, ptr(ix), ptr(eidx), ptr(dt), ptr(ef), ptr(tw), ptr(tb),
                  Q, n, d, F, H, ekp, float(p_drop), 3, 1, ptr(c), ptr(XB), H * ekp, ptr(P), ptr(inv))
        return XB, P

    psum_col = 2 * d + F
    # (a) one live neighbour per query: its softmax weight is 1, so after dropout the kept mass of a head is exactly
    # 0 or 1 / (1 - p), and the head's neighbour sum is exactly 0 or x_j / (1 - p)
    one = torch.full((Q, n), -1, dtype=torch.int32, device="cuda")
    one[:, n - 1] = idx[:, n - 1]
    XB1, _ = run(0.0, 0, one)
    XBd, _ = run(p, 16, one)
    mass = XBd[:, :, psum_col]
    kept = mass != 0
    assert torch.allclose(mass[kept], torch.full_like(mass[kept], 1.0 / (1.0 - p)), rtol=1e-6, atol=0)
    assert mass[kept].unique().numel() == 1
    assert 0.15 < float((~kept).float().mean()) < 0.45                                   # ~p of the 2Q (query, head) weights
    assert torch.allclose(XBd[:, :, :d][kept], XB1[:, :, :d][kept] / (1.0 - p), rtol=1e-6, atol=1e-7)
    assert float(XBd[:, :, :psum_col][~kept].abs().max()) == 0.0
    # (b) the saved softmax weights do not depend on the mask; the dropped path is unbiased
    XB0, P0 = run(0.0, 0)
    assert torch.allclose(P0.sum(dim=2), torch.ones(Q, H, device="cuda"), atol=1e-5)
    R = 400
    acc = torch.zeros_like(XB0, dtype=torch.float64)
    for r in range(R):
        XB, P = run(p, 16 * (r + 1))
        assert torch.equal(P, P0)
        acc += XB.double()
    mean = acc / R
    assert abs(float(mean[:, :, psum_col].mean()) - 1.0) < 5e-3                           # E[sum_j p'_j] = 1
    ref = XB0[:, :, :psum_col].double()
    err = (mean[:, :, :psum_col] - ref).abs().mean() / ref.abs().mean()
    assert err < 0.05, float(err)


def test_eval_step_graph_replay_matches_eager_and_oracle():
    """Evaluation step (candidates, embeddings, scores, ranking): CUDA-graph replay == eager launches, and the
    first batch == the oracle's evaluation step (ranks bit-exact, scores to 1e-5)."""
    from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig
    from oracle.train_loop import OracleTrainer
    st = _stream(seed=4)
    res = []
    for graph in (False, True):
        tr = PfoTrainer(st, TrainConfig(model="ours", bs=64, cuda_graph=graph), device="cuda:0")
        outs = []
        for i in range(5):
            r = tr.eval_step(2400 + i * 64, 2400 + (i + 1) * 64, n_items=30)
            outs.append([t.clone().cpu().numpy() for t in r])
        res.append((outs, tr))
    for a, b in zip(res[0][0], res[1][0]):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        assert rel_err(b[3], a[3]) < 1e-6
    tr = res[0][1]
    p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in tr.tgn.named_parameters()}
    orc = OracleTrainer(st, "ours", bs=64, params=p, seed=0)
    rk, top5, cand, scores = orc.eval_step(2400, 2464, n_items=30)
    a = res[0][0][0]
    assert np.array_equal(a[2], cand - 0) and rel_err(a[3], scores) < 1e-5
    assert np.array_equal(a[0], rk) and np.array_equal(a[1], top5)
    # the metric block of the same step (evaluation.py:127-207) against the numpy oracle, on the kernel's ranking
    from oracle import eval_metrics as em
    U = st.n_users
    pp = st.port_ptr[2400:2465]
    ref, _, _ = em.per_event_metrics(a[3], st.destinations[2400:2464] - U - 1, a[2] - U - 1, st.day_idx[2400:2464],
                                     pp - pp[0], st.port_items[pp[0]:pp[-1]], st.prices_past, st.prices_future)
    assert np.array_equal(a[4], ref)
    for outs, t in res:                                   # running sums on the device == sum of the per-batch tables
        allrows = np.concatenate([o[4] for o in outs])
        acc = t.eval_acc.cpu().numpy()
        assert acc[30] == allrows.shape[0] and np.allclose(acc[:18], allrows.sum(axis=0), rtol=1e-12, atol=1e-14)
        summ = t.eval_summary("test")
        assert abs(summ["test_recall_avg_5"] - allrows[:, 2].mean()) < 1e-12 and len(summ) == 30


def test_evaluate_loop_skips_last_batch_like_reference():
    """PfoTrainer.evaluate == the loop of reference evaluation.py:63-69: full batches only, the last one skipped."""
    from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig
    st = _stream(seed=5)
    tr = PfoTrainer(st, TrainConfig(model="ours", bs=64, cuda_graph=False), device="cuda:0")
    out = tr.evaluate(2400, 2400 + 64 * 3 + 10, bs=64, n_items=30, EVAL="val")
    assert float(tr.eval_acc[30].item()) == 64 * 3
    out2 = tr.evaluate(2700, 2700 + 64 * 2, bs=64, n_items=30, EVAL="val")      # exact multiple: last batch skipped too
    assert float(tr.eval_acc[30].item()) == 64
    assert set(out) == set(out2) and all(np.isfinite(v) for v in out.values())


def test_overlay_eval_recommendation_matches_trainer_loop(tmp_path, monkeypatch):
    """The drop-in `evaluation.eval_recommendation` (reference signature, data read from the reference's on-disk
    files: pickled price dictionaries, stock codes, 'YYYYMMDD' day keys) == PfoTrainer.evaluate on the in-memory
    stream: same 30 keys, same values."""
    import types
    from pfotgnrec_b200.synth import make_stream, write_reference_format
    from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig
    st = make_stream(n_users=300, n_items=60, n_events=3000, n_days=20, seed=6, ts_mode="nbg")
    write_reference_format(st, str(tmp_path), period="30")
    monkeypatch.chdir(tmp_path)
    tc = TrainConfig(model="ours", bs=64, cuda_graph=False)
    a = PfoTrainer(st, tc, device="cuda:0")
    for i in range(3):                           # some history in memory / pending messages first
        a.train_step(2000 + i * 64, 2000 + (i + 1) * 64)
    s, e = 2400, 2400 + 64 * 3 + 20
    state0 = a.tgn.memory.state.backup()         # both loops start from the same state and weights
    want = a.evaluate(s, e, bs=64, EVAL="val")
    mem_want = a.tgn.memory.memory.detach().clone()
    a.tgn.memory.state.restore(state0)
    import evaluation as ev_mod                  # pfotgnrec_b200/overlay/evaluation.py (load_overlay put it on sys.path)
    assert "overlay" in ev_mod.__file__
    ev_mod._TABLES.clear()
    portfolios = np.array([[st.codes[k] for k in st.portfolio(i)] or [""] for i in range(st.n_events)], dtype=object)
    ns = lambda sl: types.SimpleNamespace(sources=st.sources[sl], destinations=st.destinations[sl],
                                          timestamps=st.timestamps[sl], edge_idxs=st.edge_idxs[sl],
                                          portfolios=portfolios[sl])
    a.tgn.set_neighbor_finder(a.nf_full)
    got = ev_mod.eval_recommendation(a.tgn, ns(slice(s, e)), ns(slice(None)), 64, 10, st.n_users, "30", False, "val")
    assert set(got) == set(want) and len(got) == 30
    for k in want:
        assert got[k] == want[k], (k, got[k], want[k])
    assert torch.equal(a.tgn.memory.memory.detach(), mem_want)


def test_fit_epoch_loop_matches_manual_loop():
    """PfoTrainer.fit (the epoch loop of reference main.py:144-443) == the same calls made by hand: memory re-initialised
    per epoch, training batches over the training split (short last batch included), validation then test evaluation
    on the full graph with the memory carried over.  lr = 0 keeps the weights fixed, so the two runs are bit-identical."""
    from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig
    st = _stream(seed=8)
    tc = TrainConfig(model="ours", bs=200, lr=0.0, cuda_graph=False)
    a, b = PfoTrainer(st, tc, device="cuda:0"), PfoTrainer(st, tc, device="cuda:0")
    hist = a.fit(epochs=2, bs=200)
    assert len(hist) == 2 and all(len(h) == 62 for h in hist)
    (t0, t1), (v0, v1), (e0, e1) = b.split_ranges()
    assert (t0, e1) == (0, st.n_events) and t1 == v0 and v1 == e0 and (t1 - t0) % 200 != 0      # a short last batch
    for epoch in range(2):
        b.tgn.memory.__init_memory__()
        b.epoch = epoch                      # the candidate streams are keyed by the epoch (fresh draws, main.py:194-195)
        losses = [float(b.train_step(s, min(t1, s + 200)).item()) for s in range(t0, t1, 200)]
        want = {"epoch": epoch, "loss": float(np.mean(np.asarray(losses, dtype=np.float32)))}
        want.update(b.evaluate(v0, v1, bs=200, EVAL="valid"))
        want.update(b.evaluate(e0, e1, bs=200, EVAL="test"))
        got = hist[epoch]
        assert set(got) == set(want)
        for k, v in want.items():
            assert abs(got[k] - v) <= 1e-6 * max(1.0, abs(v)), (epoch, k, got[k], v)
        assert 0.0 <= got["valid_recall_avg_5"] <= 1.0 and np.isfinite(got["loss"])
    # same weights, memory reset -- but every epoch draws fresh MV candidates, like the reference (main.py:194-195)
    assert hist[0]["loss"] != hist[1]["loss"] and abs(hist[0]["loss"] - hist[1]["loss"]) < 0.1


def test_every_epoch_draws_fresh_candidates():
    """The reference draws new candidates every epoch (np.random.choice from the unseeded global stream,
    main.py:194-195): the Philox stream id of an interaction folds the epoch in, so epoch 1 differs from epoch 0 while
    each epoch stays reproducible and independent of the batch split."""
    from pfotgnrec_b200.synth import make_stream
    from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig
    st = make_stream(n_users=300, n_items=60, n_events=3000, n_days=20, seed=1, ts_mode="small")
    tr = PfoTrainer(st, TrainConfig(model="ours", bs=128), device="cuda")

    def cands(epoch, s, e):
        tr.epoch = epoch
        b = tr._batch(s, e)
        return tr.mv.select(b["ev"], b["day"], b["dst"], b["port_ptr"], tr.dev_stream.port_items,
                            return_scores=True)[2].cpu().numpy()

    c0, c1 = cands(0, 1000, 1128), cands(1, 1000, 1128)
    assert np.array_equal(c0[:, 0], c1[:, 0])                       # column 0 is the true destination
    assert (c0[:, 1:] != c1[:, 1:]).mean() > 0.5                    # fresh draws
    assert np.array_equal(c1, cands(1, 1000, 1128))                 # reproducible
    assert np.array_equal(c1[:64], cands(1, 1000, 1064))            # batch-split independent
    # the static buffers of the captured graph carry the same offset
    tr.epoch = 1
    sg = tr._step_graph(128)
    tr._fill_static(sg, 1000, 1128)
    assert torch.equal(sg.static["ev"], tr._batch(1000, 1128)["ev"])
    tr.epoch = 0
    # baselines: the negatives of the BPR loss
    tr2 = PfoTrainer(st, TrainConfig(model="tgn", bs=128), device="cuda")
    def negs(epoch):
        tr2.epoch = epoch
        b = tr2._batch(1000, 1128)
        return tr2.neg_sampler.sample(b["ev"], b["port_ptr"], tr2.dev_stream.port_items_as_item_ids, 3, seed=0).cpu().numpy()
    assert (negs(0) != negs(1)).mean() > 0.5


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs a second GPU")
def test_trainer_on_a_non_default_device():
    """PfoTrainer(device='cuda:1') selects the device before any kernel is launched: same losses as on cuda:0."""
    from pfotgnrec_b200.synth import make_stream
    from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig
    st = make_stream(n_users=300, n_items=60, n_events=3000, n_days=20, seed=1, ts_mode="small")
    losses = []
    for dev in ("cuda:0", "cuda:1"):
        tr = PfoTrainer(st, TrainConfig(model="ours", bs=128, cuda_graph=False), device=dev)
        losses.append([float(tr.train_step(1000 + 128 * i, 1128 + 128 * i).item()) for i in range(3)])
        assert tr.tgn.memory.memory.device == torch.device(dev)
    torch.cuda.set_device(0)
    assert np.allclose(losses[0], losses[1], rtol=1e-5)
