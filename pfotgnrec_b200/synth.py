"""Synthetic investor-stock interaction streams.

The NBG dataset of the reference is not available offline, so every test and
benchmark runs on a deterministic synthetic stream with the same id
conventions as the reference's ETL (reference utils/preprocess_data.py:47-73):
users are 1..U, items are U+1..U+I, edge idxs are 1..E, id 0 is padding.

`make_stream` builds the in-memory stream; `write_reference_format` writes
the on-disk files the unchanged reference `main.py` loads (reference
utils/data.py:20-25, main.py:88-89, evaluation.py:41-43).
"""
from __future__ import annotations

import json
import os
import pickle
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Stream:
    """One chronological interaction stream plus the price tables of MV sampling."""
    n_users: int
    n_items: int
    sources: np.ndarray          # int64[E]  in 1..U
    destinations: np.ndarray     # int64[E]  in U+1..U+I
    timestamps: np.ndarray       # float64[E] non-decreasing
    edge_idxs: np.ndarray        # int64[E]  1..E
    edge_features: np.ndarray    # float64[E+1, F], row 0 zeros
    day_idx: np.ndarray          # int32[E] row of the price table for the event's day
    port_ptr: np.ndarray         # int64[E+1] CSR over portfolios
    port_items: np.ndarray       # int32[nnz] 0-based stock index (item idx - U - 1)
    prices_future: np.ndarray    # float64[D, I, 30]
    prices_past: np.ndarray      # float64[D, I, 30]
    day_keys: list = field(default_factory=list)   # 'YYYYMMDD' strings, len D
    codes: list = field(default_factory=list)      # 6-digit stock codes, len I

    @property
    def n_events(self) -> int:
        return int(self.sources.shape[0])

    @property
    def n_nodes(self) -> int:
        return self.n_users + self.n_items + 1

    @property
    def upper_u(self) -> int:
        return self.n_users

    def portfolio(self, e: int) -> np.ndarray:
        return self.port_items[self.port_ptr[e]:self.port_ptr[e + 1]]

    def split(self, q=(0.8, 0.9)):
        """Chronological 80/10/10 split by timestamp quantile (reference utils/data.py:27,48-50)."""
        val_time, test_time = np.quantile(self.timestamps, list(q))
        train = self.timestamps <= val_time
        val = (self.timestamps <= test_time) & (self.timestamps > val_time)
        test = self.timestamps > test_time
        return train, val, test


def _zipf_probs(n: int, a: float) -> np.ndarray:
    p = 1.0 / np.arange(1, n + 1, dtype=np.float64) ** a
    return p / p.sum()


def _day_keys(n_days: int) -> list:
    d0 = np.datetime64("2020-01-01")
    return [str(d0 + np.timedelta64(i, "D")).replace("-", "") for i in range(n_days)]


def make_stream(n_users=2000, n_items=200, n_events=20000, n_days=200, seed=0,
                ts_mode="nbg", n_edge_feat=1, max_port=5, zipf=0.8,
                with_prices=True) -> Stream:
    """Deterministic synthetic stream.

    ts_mode: "nbg"  -> float64 YYYYMMDDhhmmss (what the reference's str(ts)[:8] expects,
                       reference main.py:212); ill-conditioned for fp32 time encoding.
             "small" -> small non-negative integers with ties (well-conditioned floats,
                       SURVEY.md section 7 hard part 1).
    """
    rng = np.random.default_rng(seed)
    U, I, E, D = n_users, n_items, n_events, n_days
    # user / item popularity: Zipf over a random permutation of ids
    pu = _zipf_probs(U, zipf)[rng.permutation(U)]
    pi = _zipf_probs(I, zipf)[rng.permutation(I)]
    src = rng.choice(U, size=E, p=pu).astype(np.int64) + 1
    dst0 = rng.choice(I, size=E, p=pi).astype(np.int64)
    dst = dst0 + U + 1
    day = np.sort(rng.integers(0, D, size=E)).astype(np.int32)
    if ts_mode == "nbg":
        sec = rng.integers(9 * 3600, 15 * 3600, size=E)
        order = np.lexsort((sec, day))
        day, sec = day[order], sec[order]
        keys = _day_keys(D)
        ymd = np.array([int(k) for k in keys], dtype=np.int64)[day]
        hh, mm, ss = sec // 3600, (sec // 60) % 60, sec % 60
        ts = (ymd * 1000000 + hh * 10000 + mm * 100 + ss).astype(np.float64)
    elif ts_mode == "small":
        # ~E/2 distinct integer ticks -> frequent ties, strictly chronological days
        per_day = max(1, E // (2 * D))
        tick = rng.integers(0, per_day, size=E)
        order = np.lexsort((tick, day))
        day, tick = day[order], tick[order]
        ts = (day.astype(np.int64) * per_day + tick).astype(np.float64)
        keys = _day_keys(D)
    else:
        raise ValueError(ts_mode)
    eidx = np.arange(1, E + 1, dtype=np.int64)
    efeat = np.zeros((E + 1, n_edge_feat), dtype=np.float64)
    efeat[1:] = rng.standard_normal((E, n_edge_feat))
    # portfolios: 0..max_port distinct stocks per event
    plen = rng.integers(0, max_port + 1, size=E)
    pptr = np.zeros(E + 1, dtype=np.int64)
    np.cumsum(plen, out=pptr[1:])
    pit = np.empty(int(pptr[-1]), dtype=np.int32)
    # distinct stocks within an event: draw, then redraw only the rows that hold a duplicate
    for L in range(1, max_port + 1):
        ev = np.nonzero(plen == L)[0]
        if ev.size == 0:
            continue
        pick = np.sort(rng.integers(0, I, size=(ev.size, L)), axis=1)
        while L > 1:
            bad = np.nonzero((pick[:, 1:] == pick[:, :-1]).any(axis=1))[0]
            if bad.size == 0:
                break
            pick[bad] = np.sort(rng.integers(0, I, size=(bad.size, L)), axis=1)
        pos = pptr[ev][:, None] + np.arange(L)[None, :]
        pit[pos.ravel()] = pick.ravel().astype(np.int32)
    codes = ["%06d" % (100000 + 7 * k) for k in range(I)]
    if with_prices:
        # 30-day geometric random walks, sigma = 2 %/day, per (day, stock)
        def walks():
            r = rng.standard_normal((D, I, 29)) * 0.02
            p0 = rng.uniform(10.0, 200.0, size=(D, I, 1))
            return np.concatenate([p0, p0 * np.exp(np.cumsum(r, axis=2))], axis=2)
        pf, pp = walks(), walks()
    else:
        pf = pp = np.ones((1, 1, 30))
    return Stream(U, I, src, dst, ts, eidx, efeat, day, pptr, pit, pf, pp, keys, codes)


def log_returns(prices: np.ndarray) -> np.ndarray:
    """29 daily log-returns of 30 prices (reference main.py:218,226-227)."""
    return np.log(prices[..., 1:] / prices[..., :-1])


def write_reference_format(stream: Stream, root: str, period: str = "30") -> str:
    """Write `root/data/period_{p}/...` exactly as the reference loads it."""
    d = os.path.join(root, "data", f"period_{period}")
    os.makedirs(d, exist_ok=True)
    recs = []
    for e in range(stream.n_events):
        port = [stream.codes[k] for k in stream.portfolio(e)] or [""]
        recs.append({"u": int(stream.sources[e]), "i": int(stream.destinations[e]),
                     "ts": float(stream.timestamps[e]), "label": "0",
                     "idx": int(stream.edge_idxs[e]), "portfolio": port})
    with open(os.path.join(d, "ml_transaction.json"), "w") as f:
        json.dump(recs, f)
    np.save(os.path.join(d, "ml_transaction.npy"), stream.edge_features)
    np.save(os.path.join(d, "ml_transaction_node.npy"), np.zeros((stream.n_nodes, 172)))
    with open(os.path.join(d, "map_item_id.pkl"), "wb") as f:
        pickle.dump({c: k for k, c in enumerate(stream.codes)}, f)
    for name, arr in (("future", stream.prices_future), ("past", stream.prices_past)):
        tf = {dk: {c: arr[di, k] for k, c in enumerate(stream.codes)}
              for di, dk in enumerate(stream.day_keys)}
        with open(os.path.join(d, f"time_feature_{name}_{period}.pkl"), "wb") as f:
            pickle.dump(tf, f)
    return d


def read_reference_format(root: str, period: str = "30") -> Stream:
    """Load `root/data/period_{p}/...` -- the files the reference's `get_data` / `main.py` / `evaluation.py` read
    (utils/data.py:20-25, main.py:88-89, evaluation.py:41-43; formats in SURVEY 8f-2) -- into a `Stream`:
    ml_transaction.json (records u, i, ts, label, idx, portfolio), ml_transaction.npy (edge features, row 0 = padding),
    map_item_id.pkl (stock code -> 0-based index), time_feature_{past,future}_{p}.pkl (day 'YYYYMMDD' -> code -> prices).
    Interactions are kept in file order (the reference never re-sorts them); the day of an interaction is
    str(ts)[:8] like main.py:212 / evaluation.py:151; an empty portfolio is the reference's ['']."""
    d = os.path.join(root, "data", f"period_{period}")
    with open(os.path.join(d, "ml_transaction.json")) as f:
        recs = json.load(f)
    if isinstance(recs, dict):                                   # pandas' column-oriented layout
        cols = {k: [v[i] for i in sorted(v, key=int)] for k, v in recs.items()}
        recs = [dict(zip(cols, row)) for row in zip(*cols.values())]
    E = len(recs)
    src = np.fromiter((r["u"] for r in recs), dtype=np.int64, count=E)
    dst = np.fromiter((r["i"] for r in recs), dtype=np.int64, count=E)
    ts = np.fromiter((r["ts"] for r in recs), dtype=np.float64, count=E)
    eidx = np.fromiter((r["idx"] for r in recs), dtype=np.int64, count=E)
    efeat = np.load(os.path.join(d, "ml_transaction.npy")).astype(np.float64)
    with open(os.path.join(d, "map_item_id.pkl"), "rb") as f:
        map_item_id = pickle.load(f)
    codes = [c for c, _ in sorted(map_item_id.items(), key=lambda kv: kv[1])]
    n_users, n_items = int(src.max()), len(codes)
    tables = {}
    for name in ("future", "past"):
        with open(os.path.join(d, f"time_feature_{name}_{period}.pkl"), "rb") as f:
            tables[name] = pickle.load(f)
    day_keys = sorted(set(tables["future"]) | set(tables["past"]))
    day_of = {k: i for i, k in enumerate(day_keys)}
    T1 = len(next(iter(next(iter(tables["future"].values())).values())))

    def dense(tf):
        a = np.full((len(day_keys), n_items, T1), np.nan)
        for k, row in tf.items():
            for code, prices in row.items():
                j = map_item_id.get(code)
                if j is not None:
                    a[day_of[k], j] = prices
        return a

    day_idx = np.fromiter((day_of[str(t)[:8]] for t in ts), dtype=np.int32, count=E)
    held = [[] if "" in r["portfolio"] else [map_item_id[c] for c in r["portfolio"]] for r in recs]
    pptr = np.zeros(E + 1, dtype=np.int64)
    np.cumsum([len(h) for h in held], out=pptr[1:])
    pit = np.fromiter((x for h in held for x in h), dtype=np.int32, count=int(pptr[-1]))
    return Stream(n_users, n_items, src, dst, ts, eidx, efeat, day_idx, pptr, pit, dense(tables["future"]),
                  dense(tables["past"]), day_keys, codes)
