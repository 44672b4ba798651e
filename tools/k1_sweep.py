"""K1 width sweep: time pfo_neighbor_sample per lanes-per-query on the bench stream's training graph (tools/, not product).
Graph-replayed launches between CUDA events, L2 flushed in between; prints us per launch and the roofline fraction."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pfotgnrec_b200 import _lib
from pfotgnrec_b200.graph import TemporalCSR, NeighborFinder
from pfotgnrec_b200.synth import make_stream

st = make_stream(n_users=100000, n_items=1000, n_events=5000000, n_days=200, seed=0, ts_mode="nbg", with_prices=False)
n_tr = int(st.split()[0].sum())
csr = TemporalCSR(st.sources[:n_tr], st.destinations[:n_tr], st.edge_idxs[:n_tr], st.timestamps[:n_tr], n_nodes=st.n_nodes, device="cuda")
nf = NeighborFinder(csr)
deg = np.diff(csr.rowptr.cpu().numpy())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rng = np.random.default_rng(0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6545.9
for bs in (8192, 65536):
    s0 = 2_000_000
    src, dst, ts = st.sources[s0:s0 + bs], st.destinations[s0:s0 + bs], st.timestamps[s0:s0 + bs]
    items = rng.integers(st.n_users + 1, st.n_nodes, size=4 * bs)
    nodes = np.concatenate([src, dst, items])
    qts = np.concatenate([ts, ts, np.repeat(ts, 4)])
    qn = torch.as_tensor(nodes.astype(np.int32), device="cuda")
    qt = torch.as_tensor(qts, device="cuda")
    Q = qn.shape[0]
    alg = float(np.sum(16 + 8 * np.ceil(np.log2(deg[nodes] + 1.0)) + 28 * 10))
    for lpq in (1, 4, 8, 32, 0):
        nf.lanes_per_query = lpq
        out = nf.sample(qn, qt, 10)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            nf.sample(qn, qt, 10, out=out)
        ms = []
        for _ in range(20):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        us = float(np.median(ms)) * 1e3
        print(f"bs {bs} Q {Q} lanes/query {lpq:2d}: {us:7.1f} us  frac {alg / (us * 1e-6) / 1e9 / peak:.3f}", flush=True)
