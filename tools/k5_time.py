"""Time pfo_mv_select alone at the bench shape (B = 8192, K = 20, T = 29, 1000 stocks, portfolios of 0-5 stocks), L2 flushed
between launches, for several CTAs-per-SM caps of its grid (PFO_MV_CTAS; unset = occupancy API).  Prints one line each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pfotgnrec_b200.sampler import MVSelector

B, I, D, T, U = 8192, 1000, 200, 29, 100000
rng = np.random.default_rng(0)
logret = rng.normal(0, 0.02, (D, I, T))
mv = MVSelector(logret, np.arange(I), U, n_candidates=20, device="cuda")
ev = torch.arange(B, dtype=torch.int64, device="cuda") + 12345
day = torch.as_tensor(rng.integers(0, D, B).astype(np.int32), device="cuda")
dst = torch.as_tensor((U + 1 + rng.integers(0, I, B)).astype(np.int32), device="cuda")
cnt = rng.integers(0, 6, B)
pp = torch.as_tensor(np.r_[0, np.cumsum(cnt)].astype(np.int64), device="cuda")
items = np.concatenate([rng.choice(I, c, replace=False) for c in cnt] + [np.zeros(1, np.int64)]).astype(np.int32)
pi = torch.as_tensor(items, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for cap in ("", "6", "7", "8", "10"):
    if cap:
        os.environ["PFO_MV_CTAS"] = cap
    else:
        os.environ.pop("PFO_MV_CTAS", None)
    for _ in range(3):
        out = mv.select(ev, day, dst, pp, pi)
    ts = []
    for _ in range(10):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = mv.select(ev, day, dst, pp, pi); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    if ref is None:
        ref = [t.clone() for t in out]
    same = all(torch.equal(a, b) for a, b in zip(out, ref))
    print(f"PFO_MV_CTAS={cap or 'auto'}: median {np.median(ts):.1f} us  min {min(ts):.1f} us  same ids {same}", flush=True)
