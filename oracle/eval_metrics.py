"""Oracle: the evaluation metric block (test infrastructure only).

Restates reference evaluation.py:127-258 -- per interaction: ranking of the true item among the
scored candidates, Recall@k / NDCG@k (`recall_at_k` :11-13, `ndcg_at_k` :15-21), and the change of
annualised return / Sharpe ratio of the held portfolio when the top-k recommended stocks are
appended (`return_sharpe_at_k` :23-36), in sample (`time_feature_past`) and out of sample
(`time_feature_future`), for k in (1, 3, 5); then the means and fraction-positive over all
interactions (:209-258).  numpy fp64, one interaction at a time like the reference; `argsort` is
stable (documented deviation ii).
"""
import numpy as np

KS = (1, 3, 5)
N_PER_EVENT = 18        # recall x3 | ndcg x3 | return_in x3 | sharpe_in x3 | return_out x3 | sharpe_out x3


def ranking_of(scores_row):
    """evaluation.py:137-138: positions sorted by descending score (stable ascending sort, reversed)."""
    return np.argsort(scores_row, kind="stable")[::-1]


def _perf(prices):
    """(annualised mean daily log-return, Sharpe) of an equally weighted portfolio of price rows
    (evaluation.py:165-169 and :31-34)."""
    logret = np.log(prices[:, 1:] / prices[:, :-1])
    daily = np.mean(logret, axis=0)
    ret = np.mean(daily) * 251
    sharpe = (np.mean(daily) * 251) / (np.std(daily) * np.sqrt(251))
    return ret, sharpe


def return_sharpe_at_k(port_prices, ret0, sharpe0, top_prices, k):
    """evaluation.py:23-36: append the top-k price rows to the portfolio rows."""
    new = np.concatenate([port_prices, top_prices[:k]], axis=0)
    r, s = _perf(new)
    return r - ret0, s - sharpe0


def per_event_metrics(scores, pos_stock, cand_stock, day_idx, port_ptr, port_items, prices_past, prices_future):
    """scores float[B, 1+N] (column 0 = the true item); pos_stock int[B], cand_stock int[B, N] 0-based stock
    indices; portfolio CSR over 0-based stocks (an empty row = the reference's ['']); prices_* float64[D, I, 30].
    Returns (float64[B, 18] in the column order of N_PER_EVENT, pos_rank int[B], top5 int[B, 5])."""
    B, N1 = scores.shape
    out = np.zeros((B, N_PER_EVENT), dtype=np.float64)
    pos_rank = np.zeros(B, dtype=np.int64)
    top5 = np.zeros((B, 5), dtype=np.int64)
    for b in range(B):
        ranking = ranking_of(scores[b])
        pos_rank[b] = int(np.nonzero(ranking == 0)[0][0])
        top5[b, :min(5, N1)] = ranking[:5]
        for j, k in enumerate(KS):                                          # evaluation.py:141-144, one relevant item
            hit = pos_rank[b] < k
            out[b, j] = 1.0 if hit else 0.0
            out[b, 3 + j] = 1.0 / np.log2(pos_rank[b] + 2) if hit else 0.0
        stocks = np.concatenate([[pos_stock[b]], cand_stock[b]])[ranking]   # evaluation.py:178-182
        held = port_items[port_ptr[b]:port_ptr[b + 1]]
        for s, prices in enumerate((prices_past, prices_future)):
            day = prices[day_idx[b]]
            if held.shape[0] == 0:                                          # evaluation.py:155-159
                ret0 = sharpe0 = 0.0
                port_rows = np.empty((0, day.shape[1]))
            else:                                                           # :161-176
                port_rows = day[held]
                ret0, sharpe0 = _perf(port_rows)
            top_rows = day[stocks[:5]]
            for j, k in enumerate(KS):                                      # :185-191
                dr, ds = return_sharpe_at_k(port_rows, ret0, sharpe0, top_rows, k)
                out[b, 6 + 6 * s + j] = dr
                out[b, 9 + 6 * s + j] = ds
    return out, pos_rank, top5


def aggregate(per_event, EVAL="val"):
    """evaluation.py:209-258: means over all interactions and the fraction of interactions with a positive
    change; same dictionary keys as the reference."""
    m = per_event.mean(axis=0)
    frac = (per_event > 0).mean(axis=0)
    d = {}
    for j, k in enumerate(KS):
        d[f"{EVAL}_recall_avg_{k}"] = m[j]
        d[f"{EVAL}_ndcg_avg_{k}"] = m[3 + j]
        for s, suf in enumerate(("", "_")):
            d[f"{EVAL}_return_avg_{k}{suf}"] = m[6 + 6 * s + j]
            d[f"{EVAL}_return_percent_{k}{suf}"] = frac[6 + 6 * s + j]
            d[f"{EVAL}_sharpe_avg_{k}{suf}"] = m[9 + 6 * s + j]
            d[f"{EVAL}_sharpe_percent_{k}{suf}"] = frac[9 + 6 * s + j]
    return d
