"""Drop-in for reference evaluation.py: `eval_recommendation` with the reference's signature, loop structure and
returned dictionary (evaluation.py:39-265), scored, ranked and reduced on the device.

Per batch the reference samples N_ITEMS candidates per interaction on the host, embeds them, copies every score
to the host and loops over interactions in Python (ranking, Recall/NDCG@k, six return_sharpe_at_k calls).  Here
the candidates are drawn by the K5 sampling kernel (Philox stream, seed 2024 keyed by position in the batch like
the reference's per-batch RandomState), embeddings come from the step engine, and `pfo_eval_score` +
`pfo_eval_metrics` leave 31 running sums on the device; one 248-byte copy at the end builds the dictionary.
The host only converts the batch's timestamps / portfolio codes to table indices.
"""
import math
import pickle

import numpy as np
import torch

from pfotgnrec_b200.evalmetrics import EvalMetricBlock
from pfotgnrec_b200.sampler import CandidateSampler

_TABLES = {}        # period -> (day index of 'YYYYMMDD', logret_past, logret_future) built once from the pickles


def _price_tables(period, map_item_id):
    if period not in _TABLES:
        past = pickle.load(open(f'data/period_{period}/time_feature_past_{period}.pkl', 'rb'))
        future = pickle.load(open(f'data/period_{period}/time_feature_future_{period}.pkl', 'rb'))
        days = sorted(set(past) | set(future))
        day_of = {k: i for i, k in enumerate(days)}
        n_stocks = max(map_item_id.values()) + 1
        T1 = len(next(iter(next(iter(past.values())).values())))

        def dense(tf):
            a = np.full((len(days), n_stocks, T1), np.nan)
            for k, row in tf.items():
                di = day_of[k]
                for code, prices in row.items():
                    j = map_item_id.get(code)
                    if j is not None:
                        a[di, j] = prices
            return np.log(a[..., 1:] / a[..., :-1])          # evaluation.py:29,165,172: daily log-returns

        _TABLES[period] = (day_of, dense(past), dense(future))
    return _TABLES[period]


def _csr(rows, dev):
    ptr_ = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in rows], out=ptr_[1:])
    flat = np.asarray([x for r in rows for x in r] or [0], dtype=np.int32)
    return torch.as_tensor(ptr_, device=dev), torch.as_tensor(flat, device=dev)


def eval_recommendation(tgn, data, full_data, batch_size, n_neighbors, upper_u, period, is_test_run, EVAL):
    map_item_id = pickle.load(open(f'data/period_{period}/map_item_id.pkl', 'rb'))
    day_of, lr_past, lr_future = _price_tables(period, map_item_id)
    dev = tgn.device
    block = EvalMetricBlock(lr_past, lr_future, int(upper_u) + 1, device=dev)
    items = np.unique(full_data.destinations)                 # evaluation.py:81-82 (hoisted out of the batch loop)
    N_ITEMS = len(items)
    sampler = CandidateSampler(items, device=dev)
    i32 = lambda a: torch.as_tensor(np.asarray(a).astype(np.int32), device=dev)
    with torch.no_grad():
        tgn = tgn.eval()
        eng = tgn._get_engine()
        TEST_BATCH_SIZE = batch_size
        num_test_instance = len(data.sources)
        num_test_batch = math.ceil(num_test_instance / TEST_BATCH_SIZE)
        for batch in range(num_test_batch):
            s_idx = batch * TEST_BATCH_SIZE
            e_idx = min(num_test_instance, s_idx + TEST_BATCH_SIZE)
            if e_idx == num_test_instance:                     # evaluation.py:68-69: the last batch is skipped
                continue
            if is_test_run and batch == 2:
                break
            B = e_idx - s_idx
            portfolios = data.portfolios[s_idx:e_idx]
            # portfolio of the metric block: empty when it holds '' (evaluation.py:155); the sampler instead drops
            # the '' entries one by one (utils/utils.py:76-81) -- the two differ only for mixed lists
            port_ptr, port_items = _csr([[] if '' in p else [map_item_id[c] for c in p] for p in portfolios], dev)
            held_ptr, held_items = _csr([[map_item_id[c] for c in p if c] for p in portfolios], dev)
            day_idx = i32([day_of[str(t)[:8]] for t in data.timestamps[s_idx:e_idx]])     # evaluation.py:150-151
            src, dst = i32(data.sources[s_idx:e_idx]), i32(data.destinations[s_idx:e_idx])
            ts = torch.as_tensor(np.asarray(data.timestamps[s_idx:e_idx], dtype=np.float64), device=dev)
            eidx = i32(data.edge_idxs[s_idx:e_idx])
            ev = torch.arange(B, dtype=torch.int64, device=dev)      # RandomState(2024) per batch (evaluation.py:88)
            cand = sampler.sample(ev, held_ptr, held_items + (int(upper_u) + 1), N_ITEMS, seed=2024)
            e_s, e_d, e_c = eng.compute_temporal_embeddings(tgn._params(), src, dst, [cand.reshape(-1)], ts, eidx,
                                                            n_neighbors, train=False)
            block.step(e_s, e_d, e_c, dst, cand, day_idx, port_ptr, port_items)
    return block.summary(EVAL)
