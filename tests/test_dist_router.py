"""world_size-2 gloo tests (CPU) of the all-to-all bucket routing used by the node-sharded path."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pfotgnrec_b200.dist import Exchange
    try:
        ex = Exchange("cpu")
        g = torch.Generator().manual_seed(100 + rank)
        # a sharded table: node x lives on rank x % world at row x // world; value = f(x)
        N, d = 1000, 5
        n_local = (N + world - 1) // world
        local_ids = torch.arange(n_local) * world + rank
        table = (local_ids.float().unsqueeze(1) * 10 + torch.arange(d).float())
        for step, R in enumerate((1, 257, 64, 257)):
            ids = torch.randint(0, N, (R,), generator=g).to(torch.int32)
            ids[::7] = -1                                        # rows that are not sent
            n_valid = torch.tensor([R - 3], dtype=torch.int32)   # device-side count: the tail is ignored too
            live = (ids >= 0) & (torch.arange(R) < R - 3)
            plan = ex.plan(("t",), ids, R, n_valid=n_valid)
            assert plan.cap == ex.cap_for(("t", R), R)
            assert torch.equal(plan.slot >= 0, live)
            assert torch.equal(plan.local[live].long(), ids[live].long() // world)
            # request: owner-local ids into their slots, empty slots stay -1
            req = ex.buffer(plan, 1, fill=-1)
            ex.scatter(plan, plan.local.view(-1, 1), req)
            got = ex.all_to_all(req).view(-1)
            assert int((got >= 0).sum()) <= got.numel()
            # reply in the same slots: rows of the owner's table
            reply = torch.zeros(got.numel(), d)
            ok = got >= 0
            reply[ok] = table[got[ok].long()]
            back = ex.all_to_all(reply)
            out = torch.empty(R, d)
            ex.gather(plan, back, 0, d, out)
            expect = ids.float().unsqueeze(1) * 10 + torch.arange(d).float()
            assert torch.equal(out[live], expect[live]), (rank, R)
            assert float(out[~live].abs().sum()) == 0.0
            # gradient direction: rows travel to the owners along the same slots and are summed there
            send = ex.buffer(plan, d, dtype=torch.float32)
            ex.scatter(plan, torch.ones(R, d), send)
            recv = ex.all_to_all(send)
            acc = torch.zeros(n_local, d).index_add_(0, got[ok].long(), recv[ok])
            total = acc.sum()
            dist.all_reduce(total)
            cnt = torch.tensor([float(live.sum())])
            dist.all_reduce(cnt)
            assert float(total) == float(cnt) * d
            ex.collect()                                         # eager step: bucket counts -> max over ranks
            if step == 1:
                ex.freeze()                                      # R = 257 is calibrated from here on
        key = ("t", 257)
        assert key in ex.frozen and ex.frozen[key] % ex.quantum == 0 and ex.frozen[key] >= ex.observed[key]
        ex.check_overflow()
        # a frozen capacity that is too small drops rows and raises the flag on every rank
        ex.frozen[("tiny", 64)] = 2
        plan = ex.plan(("tiny",), torch.arange(64, dtype=torch.int32) * world, 64)         # all rows to rank 0
        assert int((plan.slot >= 0).sum()) == 2
        try:
            ex.check_overflow()
            raise AssertionError("overflow not reported")
        except RuntimeError:
            pass
        # ragged slices (the short last batch of an epoch): this rank has 5 or 4 interactions x 6 rows, the buffers are
        # sized for the largest slice on every rank, so the equal-split all-to-all still lines up
        b_loc = 5 - rank
        ex.batch = (b_loc, 5)
        ids = (torch.arange(6 * b_loc, dtype=torch.int32) * 3 + rank) % N
        plan = ex.plan(("ragged",), ids, 6 * b_loc)
        assert plan.cap == ex.cap_for(("ragged", 30), 30)
        req = ex.buffer(plan, 1, fill=-1)
        ex.scatter(plan, plan.local.view(-1, 1), req)
        got = ex.all_to_all(req).view(-1)
        n_got = torch.tensor([float((got >= 0).sum())])
        dist.all_reduce(n_got)
        assert float(n_got) == 6 * 5 + 6 * 4
        ex.collect()
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_exchange_round_trip_gloo_world2():
    """Fixed-capacity bucket exchange of the node-sharded path (pfotgnrec_b200/dist.py::Exchange): slots, request /
    reply in the same slots, gradient direction, capacity calibration and the overflow flag, on 2 gloo ranks."""
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}


def test_local_csr_partition_matches_global():
    """Every rank's CSR rows are the global CSR rows of the nodes it owns (CPU, numpy only)."""
    sys.path.insert(0, ROOT)
    import numpy as np
    from oracle.graph import AdjacencyOracle
    from pfotgnrec_b200.dist import local_csr
    from pfotgnrec_b200.synth import make_stream
    st = make_stream(n_users=200, n_items=30, n_events=3000, n_days=10, seed=7, ts_mode="small", with_prices=False)
    adj = AdjacencyOracle(st.sources, st.destinations, st.edge_idxs, st.timestamps, n_nodes=st.n_nodes)
    for world in (2, 3):
        for rank in range(world):
            c = local_csr(st.sources, st.destinations, st.edge_idxs, st.timestamps, st.n_nodes, rank, world, "cpu")
            rp = c.rowptr.numpy()
            for x in range(rank, st.n_nodes, world):
                lo, hi = adj.rowptr[x], adj.rowptr[x + 1]
                l = x // world
                assert np.array_equal(c.nbr.numpy()[rp[l]:rp[l + 1]], adj.nbr[lo:hi])
                assert np.array_equal(c.eidx.numpy()[rp[l]:rp[l + 1]], adj.eidx[lo:hi])
                assert np.array_equal(c.ts.numpy()[rp[l]:rp[l + 1]], adj.ts[lo:hi])


def _dp_worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pfotgnrec_b200.trainer import replica_slice, allreduce_sum_
    try:
        # slices tile the global batch in rank order
        ls, le = replica_slice(1000, 1000 + 64 * world, rank, world)
        assert (ls, le) == (1000 + 64 * rank, 1000 + 64 * (rank + 1))
        # the gradient bucket: every rank back-propagates loss_r / world, the sum over ranks is the gradient of the
        # global-batch mean loss; parameter .grad tensors are views into the bucket
        torch.manual_seed(0)
        w = [torch.randn(5, 3, requires_grad=True), torch.randn(7, requires_grad=True)]
        sizes = [p.numel() for p in w]
        flat = torch.zeros(sum(sizes))
        for p, g in zip(w, flat.split(sizes)):
            p.grad = g.view_as(p)
        x = torch.arange(4 * world * 3, dtype=torch.float32).view(4 * world, 3) / 10.0
        xs = x[4 * rank:4 * (rank + 1)]
        loss = ((xs @ w[0].t()).pow(2).mean() + w[1].sum() * xs.mean())
        (loss / world).backward()
        allreduce_sum_(flat)
        wg = [p.detach().clone().requires_grad_(True) for p in w]
        full = sum(((x[4 * r:4 * (r + 1)] @ wg[0].t()).pow(2).mean() + wg[1].sum() * x[4 * r:4 * (r + 1)].mean())
                   for r in range(world)) / world
        full.backward()
        for p, q in zip(w, wg):
            assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-6)
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_replicated_dp_gradient_bucket_gloo_world2():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    port = 31500 + os.getpid() % 2000
    mp.spawn(_dp_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}


def test_replica_slice_covers_ragged_batches():
    """The short last batch of an epoch (main.py:180) rarely divides by the world size: the slices stay consecutive,
    cover the batch exactly once and differ by at most one interaction; fewer interactions than ranks is refused."""
    sys.path.insert(0, ROOT)
    from pfotgnrec_b200.trainer import replica_slice
    for n, world in ((130, 4), (128, 4), (7, 2), (9, 8)):
        cuts = [replica_slice(1000, 1000 + n, r, world) for r in range(world)]
        assert cuts[0][0] == 1000 and cuts[-1][1] == 1000 + n
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        sizes = [b - a for a, b in cuts]
        assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
    with pytest.raises(ValueError):
        replica_slice(0, 3, 0, 4)


def _metric_worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    from pfotgnrec_b200.evalmetrics import EvalMetricBlock
    from pfotgnrec_b200.trainer import allreduce_sum_, replica_slice
    try:
        # every rank holds the running sums of ITS slice of the users (replicated evaluation splits by user)
        rng = np.random.default_rng(7)
        per_event = rng.standard_normal((64, 18))
        per_event[:, :6] = (per_event[:, :6] > 0)
        ls, le = replica_slice(0, 64, rank, world)
        mine = per_event[ls:le]
        blk = EvalMetricBlock(np.zeros((1, 1, 29)), np.zeros((1, 1, 29)), 1, device="cpu")
        blk.acc[:18] = torch.as_tensor(mine.sum(axis=0))
        blk.acc[18:30] = torch.as_tensor((mine[:, 6:] > 0).sum(axis=0).astype(np.float64))
        blk.acc[30] = mine.shape[0]
        got = blk.summary("val", reduce=allreduce_sum_)
        assert abs(got["val_recall_avg_3"] - per_event[:, 1].mean()) < 1e-12
        assert abs(got["val_sharpe_avg_5_"] - per_event[:, 17].mean()) < 1e-12
        assert abs(got["val_return_percent_1"] - (per_event[:, 6] > 0).mean()) < 1e-12
        assert len(got) == 30
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_eval_metric_sums_all_reduce_gloo_world2():
    """The 31 running sums of the evaluation metric block, held per rank for its slice of the users, reduce to the
    whole-run dictionary with one all-reduce (reference evaluation.py:209-258 over all interactions)."""
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    port = 31500 + os.getpid() % 2000
    mp.spawn(_metric_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}


def test_chunked_device_csr_build_matches_the_host_build():
    """build_local_csr (the chunked torch build used on the GPU, also for the 10^9-interaction procedural stream) ==
    the numpy build of the owned rows == the rows of the global stable (node, time, stream order) sort."""
    sys.path.insert(0, ROOT)
    import numpy as np
    from pfotgnrec_b200.dist import local_csr, local_csr_from_device_stream, _local_csr_device
    from pfotgnrec_b200.graph import TemporalCSR
    from pfotgnrec_b200.synth_device import DeviceStream
    ds = DeviceStream(n_users=700, n_items=40, n_events=6000, n_days=12, seed=3, device="cpu")
    st = ds.materialise()
    n_tr = ds.n_train()
    for world in (2, 3):
        for rank in range(world):
            ref = local_csr(st.sources[:n_tr], st.destinations[:n_tr], st.edge_idxs[:n_tr], st.timestamps[:n_tr],
                            st.n_nodes, rank, world, "cpu")
            got = local_csr_from_device_stream(ds, n_tr, rank, world, chunk=1000)      # 5 chunks
            n_local = (st.n_nodes + world - 1) // world
            blank = TemporalCSR.__new__(TemporalCSR)
            blank.n_nodes, blank.n_events, blank.device = n_local, n_tr, torch.device("cpu")
            t = lambda a, dt: torch.as_tensor(np.asarray(a[:n_tr], dtype=dt))
            one = _local_csr_device(blank, t(st.sources, np.int64), t(st.destinations, np.int64),
                                    t(st.edge_idxs, np.int64), t(st.timestamps, np.float64), n_local, rank, world)
            for c in (got, one):
                for k in ("rowptr", "nbr", "eidx", "ts"):
                    assert torch.equal(getattr(c, k).long() if k != "ts" else getattr(c, k),
                                       getattr(ref, k).long() if k != "ts" else getattr(ref, k)), (world, rank, k)
