// K1: temporal neighbour sampling over a time-sorted CSR adjacency, and the
// touched-node compaction that turns per-batch node ids into a dense "unique node" table.
//
// Replaces reference utils/utils.py:150-220 (NeighborFinder.find_before /
// get_temporal_neighbor) -- a Python loop over queries with np.searchsorted.
// HBM / L2-latency-bound integer work.  A query is a chain of dependent loads: node id -> row bounds -> the
// probes of the lower-bound search -> the gather of its n most recent entries.  The search is COOPERATIVE: LPQ lanes
// of a warp share one query and probe LPQ split points of the current interval at once (a (LPQ+1)-ary search, one
// ballot per round), so a stock row with 10^5 entries takes 4 dependent rounds with a whole warp per query, 6 with
// 8 lanes, 17 with one lane (plain binary search).  More lanes per query shorten the chain but load more sectors,
// so the launcher picks LPQ from the batch size: a whole warp per query while the queries do not fill the machine,
// down to one lane per query (minimum traffic) once they do.  The warp then emits the n output slots of its queries
// cooperatively so that the stores to the [Q, n] outputs are coalesced.
#include "common.cuh"
#include "philox.cuh"

namespace {

__device__ __forceinline__ int64_t lower_bound_ts(const double* __restrict__ ts, int64_t lo, int64_t hi, double t) {
    // first index in [lo, hi) whose timestamp is >= t  (np.searchsorted side='left', utils.py:158)
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(ts + mid) < t) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// (LPQ+1)-ary lower bound shared by the LPQ consecutive lanes of a query group; every lane of the group returns the
// same index.  `gmask` = the lanes of the calling group, `sub` = lane index inside it.  All 32 lanes of the warp call
// this together (groups whose query is out of range pass lo == hi and fall through).
template <int LPQ>
__device__ __forceinline__ int64_t lower_bound_coop(const double* __restrict__ ts, int64_t lo, int64_t hi, double t,
                                                    int sub, int gshift) {
    if (LPQ == 1) return lower_bound_ts(ts, lo, hi, t);
    const unsigned gmask = (LPQ == 32) ? 0xffffffffu : (((1u << LPQ) - 1u) << gshift);
    // rounds are warp-uniform in count only per group, so the loop condition is voted over the whole warp
    while (true) {
        const int64_t s = hi - lo;
        const bool wide = s > LPQ;
        if (!__any_sync(0xffffffffu, wide)) break;
        // split points p_j = lo + s (j + 1) / (LPQ + 1), j = 0 .. LPQ-1: strictly inside [lo, hi) and increasing when s > LPQ
        const int64_t pj = lo + (s * (sub + 1)) / (LPQ + 1);
        const bool pred = wide && (__ldg(ts + pj) < t);
        const unsigned bal = __ballot_sync(0xffffffffu, pred) & gmask;
        if (wide) {
            const int c = __popc(bal);                      // predicates are monotone: the first c split points are < t
            const int64_t nlo = c ? lo + (s * c) / (LPQ + 1) + 1 : lo;
            const int64_t nhi = c < LPQ ? lo + (s * (c + 1)) / (LPQ + 1) : hi;
            lo = nlo; hi = nhi;
        }
    }
    // at most LPQ candidates left: one probe each
    const bool pred = (lo + sub < hi) && (__ldg(ts + lo + sub) < t);
    const unsigned bal = __ballot_sync(0xffffffffu, pred) & gmask;
    return lo + __popc(bal);
}

template <int LPQ>
__global__ void __launch_bounds__(256)
neighbor_recent_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ adj_nbr,
                       const int32_t* __restrict__ adj_eidx, const double* __restrict__ adj_ts,
                       const int32_t* __restrict__ q_nodes, const double* __restrict__ q_ts,
                       int64_t Q, int n, int64_t ld,
                       int32_t* __restrict__ out_nbr, int32_t* __restrict__ out_eidx,
                       float* __restrict__ out_etime, float* __restrict__ out_dt) {
    pfo_pdl_prologue();
    constexpr int QPW = 32 / LPQ;                     // queries per warp-iteration
    const int lane = threadIdx.x & 31;
    const int grp = lane / LPQ, sub = lane % LPQ;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp * QPW; base < Q; base += nwarps * QPW) {
        const int64_t q = base + grp;
        int64_t lo = 0, hi = 0;
        double t = 0.0;
        if (q < Q) {
            const int node = __ldg(q_nodes + q);
            t = __ldg(q_ts + q);
            if (node >= 0) { lo = __ldg(rowptr + node); hi = __ldg(rowptr + node + 1); }   // node < 0: an empty routing slot
        }
        const long long end = lower_bound_coop<LPQ>(adj_ts, lo, hi, t, sub, grp * LPQ);
        const int64_t c = end - lo;
        const int cnt = c > n ? n : (int)c;           // only the most recent n are ever taken (utils.py:207-209)
        const int nq = (Q - base) < QPW ? (int)(Q - base) : QPW;
        const int total = nq * n;
        for (int e0 = 0; e0 < total; e0 += 32) {
            const int e = e0 + lane;
            const bool live = e < total;
            const int ql = live ? e / n : 0;
            const long long end_q = __shfl_sync(0xffffffffu, end, ql * LPQ);
            const int cnt_q = __shfl_sync(0xffffffffu, cnt, ql * LPQ);
            const double t_q = __shfl_sync(0xffffffffu, t, ql * LPQ);
            if (live) {
                const int j = e - ql * n;
                const int jj = j - (n - cnt_q);      // right-aligned, left zero-padded (utils.py:216-218)
                int nb = 0, ei = 0;
                float tt = 0.0f;
                if (jj >= 0) {
                    const int64_t idx = end_q - cnt_q + jj;
                    nb = __ldg(adj_nbr + idx);
                    ei = __ldg(adj_eidx + idx);
                    tt = (float)__ldg(adj_ts + idx);   // fp32 edge time (utils.py:179-180)
                }
                const int64_t o = (base + ql) * ld + j;
                out_nbr[o] = nb;
                out_eidx[o] = ei;
                if (out_etime) out_etime[o] = tt;
                // fp64 subtraction of the fp32-rounded edge time, then fp32 (embedding_module.py:133-135)
                out_dt[o] = (float)(t_q - (double)tt);
            }
        }
    }
}

#define PFO_MAX_UNIFORM_NBR 64

__global__ void __launch_bounds__(128)
neighbor_uniform_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ adj_nbr,
                        const int32_t* __restrict__ adj_eidx, const double* __restrict__ adj_ts,
                        const int32_t* __restrict__ q_nodes, const double* __restrict__ q_ts,
                        int64_t Q, int n, uint32_t k0, uint32_t k1, uint32_t call_id_host,
                        const uint32_t* __restrict__ call_ctr, const int32_t* __restrict__ q_ids, int64_t ld,
                        int32_t* __restrict__ out_nbr, int32_t* __restrict__ out_eidx,
                        float* __restrict__ out_etime, float* __restrict__ out_dt) {
    pfo_pdl_prologue();
    // uniform-with-replacement mode (utils.py:193-204); slot j of query q draws
    // pos = mulhi32(philox(q, call_id, j, PURPOSE_NBR).x, i) and the picks are ordered by
    // (fp32 time, position): see oracle/graph.py for the contract.  call_id = host part + device counter, so a
    // captured CUDA graph of the step draws a fresh stream on every replay (the counter is bumped on the stream).
    const uint32_t call_id = call_id_host + (call_ctr ? *call_ctr : 0u);
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < Q; q += (int64_t)gridDim.x * blockDim.x) {
        const int node = q_nodes[q];
        const double t = q_ts[q];
        int64_t lo = 0, hi = 0;
        if (node >= 0) { lo = rowptr[node]; hi = rowptr[node + 1]; }        // node < 0: an empty routing slot
        const uint32_t cnt = (uint32_t)(lower_bound_ts(adj_ts, lo, hi, t) - lo);
        uint32_t pos[PFO_MAX_UNIFORM_NBR];
        float tt[PFO_MAX_UNIFORM_NBR];
        const int64_t o = q * ld;
        // the stream is keyed by the query's id: its position in this launch, or -- when the queries were routed here
        // from other ranks -- the position it has in the un-sharded query list (q_ids)
        const uint32_t qid = q_ids ? (uint32_t)q_ids[q] : (uint32_t)q;
        if (cnt == 0) {
            for (int j = 0; j < n; ++j) {
                out_nbr[o + j] = 0; out_eidx[o + j] = 0; if (out_etime) out_etime[o + j] = 0.0f; out_dt[o + j] = (float)t;
            }
            continue;
        }
        for (int j = 0; j < n; ++j) {
            const uint32_t p = mulhi32(philox4x32_10(qid, call_id, (uint32_t)j, PFO_PURPOSE_NBR, k0, k1).x, cnt);
            const float tj = (float)__ldg(adj_ts + lo + p);
            int k = j;                                   // insertion sort by (time, position)
            while (k > 0 && (tt[k - 1] > tj || (tt[k - 1] == tj && pos[k - 1] > p))) {
                tt[k] = tt[k - 1]; pos[k] = pos[k - 1]; --k;
            }
            tt[k] = tj; pos[k] = p;
        }
        for (int j = 0; j < n; ++j) {
            out_nbr[o + j] = __ldg(adj_nbr + lo + pos[j]);
            out_eidx[o + j] = __ldg(adj_eidx + lo + pos[j]);
            if (out_etime) out_etime[o + j] = tt[j];
            out_dt[o + j] = (float)(t - (double)tt[j]);
        }
    }
}

// ---------------------------------------------------------------- touched-node compaction
constexpr int kMarkSlots = 4096;           // per-CTA word cache (tag + bits: 32 KiB of shared memory)
constexpr int kMarkThreads = 1024;

// Hot nodes (a few hundred stocks, one 128-byte line of the bitmap) repeat tens of thousands of times per batch
// and would serialise in L2 as same-line atomics.  Each CTA therefore ORs its ids into a direct-mapped shared-memory
// cache of bitmap words first (slot = word mod 4096, claimed with a CAS on its tag) and flushes every non-empty
// slot with ONE global atomic at the end; an id whose slot is held by another word goes to global memory directly
// (cold words: distinct addresses, no contention).  ~20 instructions per id -- the previous version combined equal
// words inside a warp with __match_any_sync, which the compiler expands into a ~200-instruction loop per warp.
__global__ void __launch_bounds__(kMarkThreads)
mark_nodes_kernel(const int32_t* __restrict__ ids, int64_t count, int skip_zero, uint32_t* __restrict__ bitmap) {
    pfo_pdl_prologue();
    __shared__ int tag[kMarkSlots];
    __shared__ uint32_t bits[kMarkSlots];
    for (int i = threadIdx.x; i < kMarkSlots; i += kMarkThreads) { tag[i] = -1; bits[i] = 0u; }
    __syncthreads();
    const int64_t chunk = (count + gridDim.x - 1) / gridDim.x;
    const int64_t begin = (int64_t)blockIdx.x * chunk;
    const int64_t end = begin + chunk < count ? begin + chunk : count;
    for (int64_t i = begin + threadIdx.x; i < end; i += kMarkThreads) {
        const int v = ids[i];
        if (v > 0 || (v == 0 && !skip_zero)) {
            const int word = v >> 5;
            const uint32_t bit = 1u << (v & 31);
            const int slot = word & (kMarkSlots - 1);
            const int old = atomicCAS(&tag[slot], -1, word);
            if (old == -1 || old == word) atomicOr(&bits[slot], bit);
            else atomicOr(bitmap + word, bit);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kMarkSlots; i += kMarkThreads)
        if (bits[i] != 0u) atomicOr(bitmap + tag[i], bits[i]);
}

constexpr int kCompactBlock = 256;
constexpr int kWordsPerThread = 4;
constexpr int kWordsPerBlock = kCompactBlock * kWordsPerThread;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    // 256 threads; returns the exclusive prefix of v, *total = block sum
    __shared__ int warp_tot[kCompactBlock / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < kCompactBlock / 32; ++i) {
        int x = warp_tot[i];
        if (i < w) off += x;
        tot += x;
    }
    __syncthreads();
    *total = tot;
    return off + inc - v;
}

__global__ void __launch_bounds__(kCompactBlock)
compact_count_kernel(const uint32_t* __restrict__ bitmap, int64_t n_words, int32_t* __restrict__ block_sums) {
    pfo_pdl_prologue();
    const int64_t w0 = (int64_t)blockIdx.x * kWordsPerBlock + (int64_t)threadIdx.x * kWordsPerThread;
    int c = 0;
#pragma unroll
    for (int k = 0; k < kWordsPerThread; ++k)
        if (w0 + k < n_words) c += __popc(bitmap[w0 + k]);
    int tot;
    block_exclusive_scan(c, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kCompactBlock)
compact_scan_kernel(int32_t* __restrict__ block_sums, int n_blocks, int32_t* __restrict__ n_unique) {
    pfo_pdl_prologue();
    // single CTA: exclusive scan of the per-block counts, in place
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += kCompactBlock) {
        const int i = base + threadIdx.x;
        const int v = i < n_blocks ? block_sums[i] : 0;
        int tot;
        const int ex = block_exclusive_scan(v, &tot);
        const int c = carry;
        if (i < n_blocks) block_sums[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_unique = carry;
}

__global__ void __launch_bounds__(kCompactBlock)
compact_emit_kernel(uint32_t* __restrict__ bitmap, int64_t n_words, const int32_t* __restrict__ block_offs,
                    int32_t* __restrict__ uniq_ids, int32_t* __restrict__ slot_of_node) {
    pfo_pdl_prologue();
    const int64_t w0 = (int64_t)blockIdx.x * kWordsPerBlock + (int64_t)threadIdx.x * kWordsPerThread;
    uint32_t words[kWordsPerThread];
    int c = 0;
#pragma unroll
    for (int k = 0; k < kWordsPerThread; ++k) {
        words[k] = (w0 + k < n_words) ? bitmap[w0 + k] : 0u;
        c += __popc(words[k]);
    }
    int tot;
    int pos = block_offs[blockIdx.x] + block_exclusive_scan(c, &tot);
#pragma unroll
    for (int k = 0; k < kWordsPerThread; ++k) {
        uint32_t w = words[k];
        if (w) bitmap[w0 + k] = 0u;            // leave the bitmap clean for the next batch
        while (w) {
            const int b = __ffs(w) - 1;
            w &= w - 1;
            const int node = (int)((w0 + k) * 32 + b);
            uniq_ids[pos] = node;              // ascending node id -> deterministic order
            slot_of_node[node] = pos;
            ++pos;
        }
    }
}

// Small id spaces (<= kSmallWords bitmap words = 32 768 nodes: an owner's share of a sharded stream, the test streams):
// count, scan and emit in ONE CTA instead of three launches.  Measured: beyond that the single SM's scattered stores of
// the emit phase cost more than the two launches saved (bench stream, 101 001 nodes: 0.872 -> 0.885 ms per step with a
// 262 144-node limit), so larger spaces keep the three-kernel path.
constexpr int kSmallThreads = 1024;
constexpr int kSmallPerThread = 1;
constexpr int kSmallWords = kSmallThreads * kSmallPerThread;

__global__ void __launch_bounds__(kSmallThreads)
compact_small_kernel(uint32_t* __restrict__ bitmap, int64_t n_words, int32_t* __restrict__ uniq_ids,
                     int32_t* __restrict__ slot_of_node, int32_t* __restrict__ n_unique) {
    pfo_pdl_prologue();
    __shared__ int warp_tot[kSmallThreads / 32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int per = (int)((n_words + kSmallThreads - 1) / kSmallThreads);       // consecutive words per thread (<= 8)
    const int64_t w0 = (int64_t)t * per;
    uint32_t words[kSmallPerThread];
    int c = 0;
#pragma unroll
    for (int k = 0; k < kSmallPerThread; ++k) {
        words[k] = (k < per && w0 + k < n_words) ? bitmap[w0 + k] : 0u;
        c += __popc(words[k]);
    }
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0) {                                        // scan of the 32 warp totals
        int v = warp_tot[lane], iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += y;
        }
        warp_tot[lane] = iv - v;                         // exclusive offset of warp `lane`
        if (lane == 31) *n_unique = iv;
    }
    __syncthreads();
    int pos = warp_tot[w] + inc - c;
#pragma unroll
    for (int k = 0; k < kSmallPerThread; ++k) {
        uint32_t x = words[k];
        if (x) bitmap[w0 + k] = 0u;                      // leave the bitmap clean for the next batch
        while (x) {
            const int b = __ffs(x) - 1;
            x &= x - 1;
            const int node = (int)((w0 + k) * 32 + b);
            uniq_ids[pos] = node;                        // ascending node id -> deterministic order
            slot_of_node[node] = pos;
            ++pos;
        }
    }
}

__global__ void map_slots_kernel(const int32_t* __restrict__ ids, int64_t count, int skip_zero,
                                 const int32_t* __restrict__ slot_of_node, int32_t* __restrict__ out) {
    pfo_pdl_prologue();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = ids[i];
        out[i] = (v > 0 || (v == 0 && !skip_zero)) ? slot_of_node[v] : -1;
    }
}

}  // namespace

template <int LPQ>
static void launch_recent(const int64_t* rowptr, const int32_t* adj_nbr, const int32_t* adj_eidx, const double* adj_ts,
                          const int32_t* q_nodes, const double* q_ts, int64_t Q, int n, int64_t ld, int32_t* out_nbr,
                          int32_t* out_eidx, float* out_etime, float* out_dt, cudaStream_t s) {
    const int64_t threads = Q * LPQ;
    pfo_launch(neighbor_recent_kernel<LPQ>, pfo_grid(threads, 256, 8), 256, 0, s, 
        rowptr, adj_nbr, adj_eidx, adj_ts, q_nodes, q_ts, Q, n, ld, out_nbr, out_eidx, out_etime, out_dt);
}

PFO_API int pfo_neighbor_sample(const int64_t* rowptr, const int32_t* adj_nbr, const int32_t* adj_eidx,
                                const double* adj_ts, const int32_t* q_nodes, const double* q_ts,
                                int64_t n_queries, int n_neighbors, int uniform, uint64_t seed, uint32_t call_id,
                                const uint32_t* call_ctr, int lanes_per_query, const int32_t* q_ids, int64_t ld_out,
                                int32_t* out_nbr, int32_t* out_eidx, float* out_etime, float* out_dt,
                                void* stream) {
    if (n_queries <= 0) return 0;
    if (n_neighbors <= 0) return (int)cudaErrorInvalidValue;
    const int64_t ld = ld_out > 0 ? ld_out : n_neighbors;
    if (ld < n_neighbors) return (int)cudaErrorInvalidValue;
    cudaStream_t s = (cudaStream_t)stream;
    if (uniform) {
        if (n_neighbors > PFO_MAX_UNIFORM_NBR) return (int)cudaErrorInvalidValue;
        pfo_launch(neighbor_uniform_kernel, pfo_grid(n_queries, 128, 8), 128, 0, s, 
            rowptr, adj_nbr, adj_eidx, adj_ts, q_nodes, q_ts, n_queries, n_neighbors,
            (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), call_id, call_ctr, q_ids, ld, out_nbr, out_eidx,
            out_etime, out_dt);
    } else {
        // lanes per query: as many as keep every warp of the launch resident at once (148 SMs x 64 warps), so the
        // dependent-probe chain is as short as the batch allows; one lane per query (least traffic) for large batches
        int lpq = lanes_per_query;
        if (lpq <= 0) {
            const int64_t resident = (int64_t)pfo_num_sms() * 64 * 32;      // threads
            lpq = 32;
            while (lpq > 1 && n_queries * lpq > resident) lpq >>= 1;
            if (lpq == 16) lpq = 8;
            if (lpq == 2) lpq = 1;
        }
#define PFO_K1(L) launch_recent<L>(rowptr, adj_nbr, adj_eidx, adj_ts, q_nodes, q_ts, n_queries, n_neighbors, ld, \
                                   out_nbr, out_eidx, out_etime, out_dt, s)
        switch (lpq) {
            case 32: PFO_K1(32); break;
            case 8: PFO_K1(8); break;
            case 4: PFO_K1(4); break;
            case 1: PFO_K1(1); break;
            default: return (int)cudaErrorInvalidValue;
        }
#undef PFO_K1
    }
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_abi_version(void) { return PFO_ABI_VERSION; }

PFO_API int pfo_mark_nodes(const int32_t* ids, int64_t count, int skip_zero, uint32_t* bitmap, void* stream) {
    if (count <= 0) return 0;
    int64_t grid = (count + 2 * kMarkThreads - 1) / (2 * kMarkThreads);      // >= 2 ids per thread, at most one CTA per SM
    if (grid > pfo_num_sms()) grid = pfo_num_sms();
    pfo_launch(mark_nodes_kernel, (int)grid, kMarkThreads, 0, (cudaStream_t)stream, ids, count, skip_zero, bitmap);
    PFO_LAUNCH_CHECK();
}

PFO_API int64_t pfo_compact_workspace_ints(int64_t n_nodes) {
    const int64_t n_words = (n_nodes + 31) / 32;
    return (n_words + kWordsPerBlock - 1) / kWordsPerBlock + 1;
}

PFO_API int pfo_compact_nodes(uint32_t* bitmap, int64_t n_nodes, int32_t* workspace, int32_t* uniq_ids,
                              int32_t* slot_of_node, int32_t* n_unique, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n_words = (n_nodes + 31) / 32;
    if (n_words <= kSmallWords) {
        pfo_launch(compact_small_kernel, 1, kSmallThreads, 0, s, bitmap, n_words, uniq_ids, slot_of_node, n_unique);
        PFO_LAUNCH_CHECK();
    }
    const int n_blocks = (int)((n_words + kWordsPerBlock - 1) / kWordsPerBlock);
    pfo_launch(compact_count_kernel, n_blocks, kCompactBlock, 0, s, bitmap, n_words, workspace);
    pfo_launch(compact_scan_kernel, 1, kCompactBlock, 0, s, workspace, n_blocks, n_unique);
    pfo_launch(compact_emit_kernel, n_blocks, kCompactBlock, 0, s, bitmap, n_words, workspace, uniq_ids, slot_of_node);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_map_slots(const int32_t* ids, int64_t count, int skip_zero, const int32_t* slot_of_node,
                          int32_t* out, void* stream) {
    if (count <= 0) return 0;
    pfo_launch(map_slots_kernel, pfo_grid(count, 256, 8), 256, 0, (cudaStream_t)stream, ids, count, skip_zero, slot_of_node, out);
    PFO_LAUNCH_CHECK();
}
