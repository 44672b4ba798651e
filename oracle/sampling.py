"""Oracle: candidate sampling and mean-variance efficient selection (test infrastructure only).

* `sample_candidates`  restates reference utils/utils.py:65-114 (RandEdgeSampler) with the
  shared Philox stream in place of numpy's MT19937 (documented deviation i).
* `mv_select` restates the inline block reference main.py:197-304: 29 log-returns per
  candidate, y_mv from the mean / (co)variance against the held portfolio, rank fusion with
  the preference order, top-1 positive and bottom-3 negatives; `argsort` is stable
  (documented deviation ii).
"""
import numpy as np

from .philox import philox4x32_10, mulhi32, PURPOSE_NEG, PURPOSE_NEG_REPL, PURPOSE_NEG_SEQ


def sample_candidates(event_ids, items_sorted, port_ptr, port_items, size, seed):
    """Uniform candidates from (all train items \\ portfolio of the event).

    event_ids  int64[B]   global, batch-independent id of each interaction (its edge idx)
    items_sorted int64[M] np.unique(all train destinations)       (utils.py:73)
    port_ptr/port_items   CSR of the held item ids per interaction (utils.py:76-81)
    Returns int64[B, size] item ids.

    Without replacement when enough items are available (utils.py:107-111), by one of two exact schemes chosen
    from the sizes alone:
      * size <= 32 and 8 * size <= n_available -- sequential rejection: slot j keeps the first draw
        available[mulhi32(philox(g_lo, g_hi, j | a << 16, PURPOSE_NEG_SEQ)[0], n_available)], a = 0, 1, ..,
        that no slot < j holds;
      * otherwise every available item at position p of `items_sorted` gets the key
        (philox(g_lo, g_hi, p, PURPOSE_NEG)[0], p) and the sample is the `size` smallest keys in ascending order
        (the 32-bit draw first, the position breaks ties).
    With replacement when fewer items are available than requested (:99-105): draw j is
    available[mulhi32(philox(g_lo, g_hi, j, PURPOSE_NEG_REPL)[0], n_available)].
    """
    event_ids = np.asarray(event_ids, dtype=np.int64)
    items_sorted = np.asarray(items_sorted, dtype=np.int64)
    B, M = event_ids.shape[0], items_sorted.shape[0]
    out = np.zeros((B, size), dtype=np.int64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    pos_all = np.arange(M, dtype=np.int64)
    for b in range(B):
        g = int(event_ids[b])
        held = port_items[port_ptr[b]:port_ptr[b + 1]]
        avail_mask = ~np.isin(items_sorted, held)
        pos = pos_all[avail_mask]
        n_av = pos.shape[0]
        if n_av < size:
            j = np.arange(size, dtype=np.int64)
            x0 = philox4x32_10(g & 0xFFFFFFFF, (g >> 32) & 0xFFFFFFFF, j, PURPOSE_NEG_REPL, k0, k1)[0]
            out[b] = items_sorted[pos[mulhi32(x0, n_av)]]
        elif size <= 32 and 8 * size <= n_av:
            glo, ghi = g & 0xFFFFFFFF, (g >> 32) & 0xFFFFFFFF
            q = mulhi32(philox4x32_10(glo, ghi, np.arange(size, dtype=np.int64), PURPOSE_NEG_SEQ, k0, k1)[0], n_av)
            if np.unique(q).shape[0] != size:                  # rare: resolve the collisions slot by slot
                for j in range(1, size):
                    a = 0
                    while q[j] in q[:j]:
                        a += 1
                        q[j] = mulhi32(philox4x32_10(glo, ghi, j | (a << 16), PURPOSE_NEG_SEQ, k0, k1)[0], n_av)
            out[b] = items_sorted[pos[q]]
        else:
            x0 = philox4x32_10(g & 0xFFFFFFFF, (g >> 32) & 0xFFFFFFFF, pos, PURPOSE_NEG, k0, k1)[0]
            o = np.lexsort((pos, x0))[:size]
            out[b] = items_sorted[pos[o]]
    return out


def _seq_sum(x):
    """Left-to-right sum over the last axis (the order the CUDA kernel uses)."""
    acc = np.zeros(x.shape[:-1], dtype=np.float64)
    for t in range(x.shape[-1]):
        acc = acc + x[..., t]
    return acc


def mv_scores(logret, day_idx, cand, port_ptr, port_items, gamma):
    """y_mv of every candidate (reference main.py:243-271), fp64.

    logret float64[D, I, T] log-returns; cand int[B, C] 0-based stock index (column 0 = the
    true destination); portfolio CSR over 0-based stock indices.  Closed form:
        mu = mean(r_c); var = sum((r_c-mu)^2)/(T-1)
        empty portfolio: y = (mu/gamma) / var                                  (:246-254)
        else: y = (mu/gamma - 0.5 * (1/nP) * sum_p cov(r_c, r_p)) / var         (:259-271)
    where sum_p cov(r_c, r_p) = cov(r_c, sum_p r_p) (np.cov uses ddof=1).
    """
    B, C = cand.shape
    T = logret.shape[2]
    y = np.zeros((B, C), dtype=np.float64)
    for b in range(B):
        r = logret[day_idx[b], cand[b], :]                      # [C, T]
        mu = _seq_sum(r) / T
        dc = r - mu[:, None]
        var = _seq_sum(dc * dc) / (T - 1)
        held = port_items[port_ptr[b]:port_ptr[b + 1]]
        nP = held.shape[0]
        if nP == 0:
            y[b] = (mu / gamma) / var
        else:
            S = np.zeros(T, dtype=np.float64)
            for p in held:
                S = S + logret[day_idx[b], p, :]
            mS = _seq_sum(S) / T
            cov = _seq_sum(dc * (S - mS)[None, :]) / (T - 1)
            y[b] = (mu / gamma - 0.5 * (cov / nP)) / var
    return y


def rank_average(v):
    """scipy.stats.rankdata(v) (method='average', ascending), fp64 (main.py:282-283)."""
    v = np.asarray(v)
    less = (v[None, :] < v[:, None]).sum(axis=1)
    eq = (v[None, :] == v[:, None]).sum(axis=1)
    return less + (eq + 1) / 2.0


def mv_select(logret, day_idx, cand, port_ptr, port_items, gamma, lam, n_pos=1, n_neg=3):
    """Rank-fuse y_mv with the preference order and pick positives / negatives (main.py:282-292).

    Returns (p_pos int64[B*n_pos], p_neg int64[B*n_neg]) as 0-based stock indices,
    interaction-major; p_neg is [n_neg-th lowest, ..., lowest] fused rank.
    """
    y = mv_scores(logret, day_idx, cand, port_ptr, port_items, gamma)
    B, C = cand.shape
    pref = np.arange(C, 0, -1, dtype=np.float64)             # rankdata([C-1..0]) = [C..1]
    p_pos = np.zeros((B, n_pos), dtype=np.int64)
    p_neg = np.zeros((B, n_neg), dtype=np.int64)
    for b in range(B):
        invest = rank_average(y[b])
        new_rank = invest * lam + pref * (1 - lam)           # mul, mul, add; no fma
        order = np.argsort(new_rank, kind="stable")[::-1]
        p_pos[b] = cand[b, order[:n_pos]]
        p_neg[b] = cand[b, order[-n_neg:]]
    return p_pos.ravel(), p_neg.ravel()
