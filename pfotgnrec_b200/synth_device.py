"""GPU-resident synthetic interaction stream for the scale configuration (BASELINE config 4: 10^7 users x 5 000
stocks x 10^9 events), SURVEY.md section 8f-2.

`synth.make_stream` builds its stream with numpy on the host, which is fine up to ~10^7 events; a 10^9-event stream
is ~50 GB of columns and cannot be drawn, sorted or copied through the host per process.  Here every column of
interaction i is a pure FUNCTION of (seed, i) -- a splitmix64-style integer hash evaluated with torch integer ops on
the device -- so any range [s, e) can be materialised on any rank in any order without communication, identically:

    user / stock   inverse-CDF of a Zipf(0.8) popularity over a hashed permutation of the ids (torch.searchsorted)
    timestamp      YYYYMMDDhhmmss (the reference's NBG format, main.py:212), strictly chronological by construction:
                   interaction i falls on day floor(i * D / E) at 09:00:00 + its rank inside the day scaled to 6 h
    edge feature   one standard-normal-ish value per interaction (sum of uniforms), row 0 = padding
    portfolio      0..5 DISTINCT stocks: (start + k * stride) mod I, k < len, all from the hash of i
    prices         30-day geometric random walks per (day, stock) -- small (D * I * 30), drawn on the host like
                   synth.make_stream

Same id conventions as the reference's ETL (utils/preprocess_data.py:47-73): users 1..U, items U+1..U+I, edge idxs
1..E, id 0 = padding.  `DeviceStream.columns(s, e)` returns the device columns of a range; `materialise()` returns a
host `synth.Stream` (small sizes only: tests compare the trainer on both representations).
"""
from __future__ import annotations

import numpy as np
import torch

from .synth import Stream, _day_keys, _zipf_probs

_M64 = (1 << 64) - 1


def _i64(x):
    """Python int (mod 2^64) -> the int64 with the same bits."""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(x, k):
    """Logical right shift of an int64 tensor."""
    return (x >> k) & ((1 << (64 - k)) - 1)


def hash64(i: torch.Tensor, seed: int, stream: int) -> torch.Tensor:
    """splitmix64 finaliser of (i, seed, stream): int64 tensor -> int64 tensor of well-mixed bits (wrapping multiply)."""
    z = i + _i64(0x9E3779B97F4A7C15 * (2 * seed + 1) + 0xD1B54A32D192ED03 * (stream + 1))
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def _unit(h: torch.Tensor) -> torch.Tensor:
    """Top 53 bits of a hash -> float64 in [0, 1)."""
    return _lsr(h, 11).to(torch.float64) * (1.0 / (1 << 53))


class DeviceStream:
    """Procedural stream: columns of any interaction range on demand, on `device`."""

    def __init__(self, n_users, n_items, n_events, n_days=200, seed=0, device="cuda", max_port=5, zipf=0.8):
        self.n_users, self.n_items, self.n_events, self.n_days = int(n_users), int(n_items), int(n_events), int(n_days)
        self.seed, self.max_port = int(seed), int(max_port)
        self.device = torch.device(device)
        dev = self.device
        # popularity: Zipf over a hashed permutation of the ids (the CDFs are the only per-id tables: 8 B per id)
        for name, n, strm in (("u", self.n_users, 11), ("i", self.n_items, 12)):
            w = 1.0 / torch.arange(1, n + 1, dtype=torch.float64, device=dev) ** zipf
            order = torch.argsort(hash64(torch.arange(n, dtype=torch.int64, device=dev), self.seed, strm))
            p = torch.empty_like(w)
            p[order] = w                                        # id `order[r]` gets the r-th largest weight
            cdf = torch.cumsum(p / p.sum(), 0)
            cdf[-1] = 1.0
            setattr(self, f"cdf_{name}", cdf)
        self.day_keys = _day_keys(self.n_days)
        self.ymd = torch.tensor([int(k) for k in self.day_keys], dtype=torch.int64, device=dev)
        self.codes = ["%06d" % (100000 + 7 * k) for k in range(self.n_items)]
        rng = np.random.default_rng(self.seed)

        def walks():
            D, I = self.n_days, self.n_items
            r = rng.standard_normal((D, I, 29)) * 0.02
            p0 = rng.uniform(10.0, 200.0, size=(D, I, 1))
            return np.concatenate([p0, p0 * np.exp(np.cumsum(r, axis=2))], axis=2)

        self.prices_future, self.prices_past = walks(), walks()

    @property
    def n_nodes(self):
        return self.n_users + self.n_items + 1

    @property
    def upper_u(self):
        return self.n_users

    # ---- per-interaction columns (i = 0-based interaction index, int64 tensor on the device)
    def src_dst(self, i):
        u = torch.searchsorted(self.cdf_u, _unit(hash64(i, self.seed, 1))).clamp_(max=self.n_users - 1)
        v = torch.searchsorted(self.cdf_i, _unit(hash64(i, self.seed, 2))).clamp_(max=self.n_items - 1)
        return (u + 1).to(torch.int32), (v + self.n_users + 1).to(torch.int32)

    def day(self, i):
        return torch.div(i * self.n_days, self.n_events, rounding_mode="floor").to(torch.int32)

    def timestamps(self, i):
        """float64 YYYYMMDDhhmmss, non-decreasing in i."""
        E, D = self.n_events, self.n_days
        d = torch.div(i * D, E, rounding_mode="floor")
        first = torch.div(d * E + D - 1, D, rounding_mode="floor")          # first interaction of day d
        nxt = torch.div((d + 1) * E + D - 1, D, rounding_mode="floor")
        sec = 9 * 3600 + torch.div((i - first) * (6 * 3600), (nxt - first).clamp_(min=1), rounding_mode="floor")
        hh, mm, ss = torch.div(sec, 3600, rounding_mode="floor"), torch.div(sec, 60, rounding_mode="floor") % 60, sec % 60
        return (self.ymd[d] * 1000000 + hh * 10000 + mm * 100 + ss).to(torch.float64)

    def edge_feature(self, i):
        """~N(0, 1): sum of four uniforms, centred and scaled (one feature per interaction, like the real data)."""
        h = hash64(i, self.seed, 3)
        parts = [(_lsr(h, 16 * k) & 0xFFFF).to(torch.float32) for k in range(4)]
        u = (parts[0] + parts[1] + parts[2] + parts[3]) * (1.0 / 65536.0)
        return (u - 2.0) * float(np.sqrt(3.0))

    def portfolio(self, i):
        """(length int64[n], items int32[n, max_port]): `length` leading entries of each row are the held stocks."""
        I, P = self.n_items, self.max_port
        h = hash64(i, self.seed, 4)
        plen = _lsr(h, 40) % (P + 1)
        start = _lsr(h, 8) % I
        stride = 1 + (h & 0xFF) % max(1, (I - 1) // max(P, 1))
        k = torch.arange(P, dtype=torch.int64, device=i.device).view(1, P)
        items = (start.view(-1, 1) + k * stride.view(-1, 1)) % I
        return plen, items.to(torch.int32)

    def columns(self, s, e):
        """Device columns of interactions [s, e): the batch dictionary `PfoTrainer` consumes (portfolio CSR padded to
        a static capacity of max_port * (e - s) + 1 entries: no host sync)."""
        dev = self.device
        i = torch.arange(s, e, dtype=torch.int64, device=dev)
        src, dst = self.src_dst(i)
        plen, items = self.portfolio(i)
        ptr_ = torch.zeros(e - s + 1, dtype=torch.int64, device=dev)
        torch.cumsum(plen, 0, out=ptr_[1:])
        cap = self.max_port * (e - s) + 1
        k = torch.arange(self.max_port, dtype=torch.int64, device=dev).view(1, -1)
        pos = torch.where(k < plen.view(-1, 1), ptr_[:-1].view(-1, 1) + k, torch.full_like(items, cap - 1, dtype=torch.int64))
        port_items = torch.zeros(cap, dtype=torch.int32, device=dev)
        port_items.scatter_(0, pos.reshape(-1), items.reshape(-1))
        return dict(src=src, dst=dst, ts=self.timestamps(i), eidx=(i + 1).to(torch.int32), ev=i + 1, day=self.day(i),
                    port_ptr=ptr_, port_items=port_items)

    def n_train(self, q=0.8):
        """Interactions with timestamp <= the q-quantile of all timestamps (utils/data.py:27,48): found by bisection on
        the monotone timestamp column, no 10^9-element quantile."""
        E = self.n_events
        tq = float(self.timestamps(torch.tensor([min(E - 1, int(q * (E - 1)))], dtype=torch.int64, device=self.device))[0])
        lo, hi = 0, E
        while lo < hi:                                          # first index whose timestamp exceeds tq
            mid = (lo + hi) // 2
            if float(self.timestamps(torch.tensor([mid], dtype=torch.int64, device=self.device))[0]) <= tq:
                lo = mid + 1
            else:
                hi = mid
        return lo

    # ---- host view (small sizes)
    def materialise(self) -> Stream:
        E = self.n_events
        c = self.columns(0, E)
        plen, _ = self.portfolio(torch.arange(E, dtype=torch.int64, device=self.device))
        nnz = int(plen.sum().item())
        ef = np.zeros((E + 1, 1), dtype=np.float64)
        ef[1:, 0] = self.edge_feature(torch.arange(E, dtype=torch.int64, device=self.device)).double().cpu().numpy()
        return Stream(self.n_users, self.n_items, c["src"].cpu().numpy().astype(np.int64),
                      c["dst"].cpu().numpy().astype(np.int64), c["ts"].cpu().numpy(),
                      np.arange(1, E + 1, dtype=np.int64), ef, c["day"].cpu().numpy().astype(np.int32),
                      c["port_ptr"].cpu().numpy(), c["port_items"][:nnz].cpu().numpy().astype(np.int32),
                      self.prices_future, self.prices_past, self.day_keys, self.codes)
