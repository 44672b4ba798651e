"""Procedural (GPU-resident) stream generator of the scale configuration -- host logic, runs on CPU tensors."""
import numpy as np
import torch

from pfotgnrec_b200.synth_device import DeviceStream, hash64


def test_hash_is_a_function_of_its_inputs_only():
    i = torch.arange(1000, dtype=torch.int64)
    a, b = hash64(i, 3, 1), hash64(i, 3, 1)
    assert torch.equal(a, b) and not torch.equal(a, hash64(i, 3, 2)) and not torch.equal(a, hash64(i, 4, 1))
    assert torch.equal(hash64(i[500:], 3, 1), a[500:])
    bits = ((a.view(-1, 1) >> torch.arange(64)) & 1).float().mean()
    assert abs(float(bits) - 0.5) < 0.01


def test_device_stream_columns_are_range_independent_and_well_formed():
    ds = DeviceStream(n_users=5000, n_items=300, n_events=40000, n_days=20, seed=7, device="cpu")
    full = ds.columns(0, 40000)
    part = ds.columns(12345, 23456)
    for k in ("src", "dst", "ts", "eidx", "ev", "day"):
        assert torch.equal(part[k], full[k][12345:23456]), k
    src, dst, ts = full["src"].numpy(), full["dst"].numpy(), full["ts"].numpy()
    assert src.min() >= 1 and src.max() <= 5000 and dst.min() >= 5001 and dst.max() <= 5300
    assert np.all(np.diff(ts) >= 0)                                 # chronological (utils/data.py:48-50 splits by time)
    assert [str(int(t))[:8] for t in ts[[0, -1]]] == [ds.day_keys[0], ds.day_keys[-1]]          # main.py:212
    day = full["day"].numpy()
    assert np.array_equal(np.array([ds.day_keys[d] for d in day[::997]]), np.array([str(int(t))[:8] for t in ts[::997]]))
    hh = (ts.astype(np.int64) // 10000) % 100
    assert hh.min() >= 9 and hh.max() <= 15
    # Zipf popularity: the head is heavy, every id is reachable in principle
    cnt = np.sort(np.bincount(dst - 5001, minlength=300))[::-1]
    assert cnt[0] > 8 * cnt[150] > 0
    # portfolio CSR: 0..5 distinct stocks, padded capacity untouched beyond nnz
    ptr_, items = part["port_ptr"].numpy(), part["port_items"].numpy()
    lens = np.diff(ptr_)
    assert lens.min() == 0 and lens.max() == 5 and items.shape[0] == 5 * (23456 - 12345) + 1
    for b in range(0, len(lens), 53):
        row = items[ptr_[b]:ptr_[b + 1]]
        assert len(set(row.tolist())) == len(row) and (row >= 0).all() and (row < 300).all()
    # the 80 % split point found by bisection == the count the reference's quantile mask gives
    n_tr = ds.n_train()
    assert n_tr == int((ts <= np.quantile(ts, 0.8)).sum())
    ef = ds.edge_feature(torch.arange(40000, dtype=torch.int64)).numpy()
    assert abs(ef.mean()) < 0.02 and abs(ef.std() - 1.0) < 0.02


def test_device_stream_materialises_into_the_host_stream_type():
    ds = DeviceStream(n_users=400, n_items=50, n_events=3000, n_days=10, seed=1, device="cpu")
    st = ds.materialise()
    assert st.n_events == 3000 and st.n_nodes == 451 and st.edge_features.shape == (3001, 1)
    assert st.port_ptr[-1] == st.port_items.shape[0] and st.prices_future.shape == (10, 50, 30)
    tr, va, te = st.split()
    assert int(tr.sum()) == ds.n_train() and tr.sum() + va.sum() + te.sum() == 3000
