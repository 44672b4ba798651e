// K5: candidate sampling + mean-variance efficient selection.  Compiled with -fmad=false so
// that the fp64 arithmetic rounds exactly like the numpy oracle (mul and add never fuse).
//
// Replaces reference utils/utils.py:65-114 (RandEdgeSampler: np.unique + setdiff1d +
// np.random.choice per interaction) and the inline per-interaction / per-candidate Python
// loops of main.py:197-304 (np.cov x21, rankdata x2, argsort x2 per interaction).
// One warp per interaction: the K candidates are drawn without replacement by sequential rejection
// (slot j takes the first draw of its own Philox stream that no earlier slot holds -- ~K Philox calls per
// interaction; when K is not small against the available items the draw falls back to "K smallest of one
// Philox key per item"), lanes 0..K then own one candidate each for y_mv and the rank fusion.
#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr int kCap = 256;          // threshold-select buffer per warp
constexpr int kWarps = 4;
constexpr int kStage = 8;          // candidate rows of the return table in flight per warp while staging

__device__ __forceinline__ int find_pos(const int32_t* __restrict__ items, int M, int v) {
    int lo = 0, hi = M;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (items[mid] < v) lo = mid + 1; else hi = mid; }
    return (lo < M && items[lo] == v) ? lo : -1;
}

__device__ __forceinline__ bool is_held(const int32_t* __restrict__ held, int nP, int v) {
    for (int k = 0; k < nP; ++k) if (held[k] == v) return true;
    return false;
}

// number of distinct held items that belong to the universe (np.setdiff1d semantics)
__device__ int count_held_in_universe(const int32_t* held, int nP, const int32_t* items, int M, int lane) {
    int c = 0;
    for (int k = lane; k < nP; k += 32) {
        const int v = held[k];
        bool first = true;
        for (int k2 = 0; k2 < k; ++k2) if (held[k2] == v) { first = false; break; }
        if (first && find_pos(items, M, v) >= 0) ++c;
    }
    return warp_sum_i(c);
}

// r-th available universe position when drawing with replacement (utils.py:99-105)
__device__ int nth_available(const int32_t* items, int M, const int32_t* held, int nP, int r) {
    // walk the universe positions of the held items in ascending order
    int pos = r, last = -1;
    for (;;) {
        int best = 0x7fffffff;
        for (int k = 0; k < nP; ++k) {
            const int hp = find_pos(items, M, held[k]);
            if (hp > last && hp < best) best = hp;
        }
        if (best == 0x7fffffff || best > pos) break;
        ++pos; last = best;
    }
    return pos;
}

// positions (in the sorted universe) of the distinct held items that belong to it, ascending, in hs[0..nH)
// (np.setdiff1d semantics, utils.py:96); warp-cooperative, nP <= 32.  Returns nH.
__device__ int held_positions(const int32_t* held, int nP, const int32_t* items, int M, int lane, int* hs) {
    int hp = -1, v = 0;
    if (lane < nP) { v = held[lane]; hp = find_pos(items, M, v); }
    for (int k = 0; k < nP; ++k) {                      // duplicates: keep the first occurrence
        const int vk = __shfl_sync(0xffffffffu, v, k);
        if (lane > k && lane < nP && v == vk) hp = -1;
    }
    int rank = 0;
    for (int k = 0; k < nP; ++k) {
        const int hk = __shfl_sync(0xffffffffu, hp, k);
        rank += (hk >= 0 && hk < hp) ? 1 : 0;
    }
    __syncwarp();
    if (hp >= 0) hs[rank] = hp;
    __syncwarp();
    return __popc(__ballot_sync(0xffffffffu, hp >= 0));
}

// universe position of the q-th available item given the ascending held positions
__device__ __forceinline__ int skip_held(const int* hs, int nH, int q) {
    int pos = q;
    for (int i = 0; i < nH; ++i) if (hs[i] <= pos) ++pos;
    return pos;
}

// `size` (<= 32) distinct indices in [0, n_av), slot j in lane j: slot j keeps the first draw
// mulhi32(philox(event, j | attempt << 16, NEG_SEQ), n_av), attempt = 0, 1, .., that no slot < j holds
// (sequential rejection = uniform sampling without replacement; shared with oracle/sampling.py)
__device__ int draw_distinct(uint32_t g_lo, uint32_t g_hi, uint32_t k0, uint32_t k1, int size, int n_av, int lane) {
    int q = -1;
    uint32_t attempt = 0;
    if (lane < size) q = (int)mulhi32(philox4x32_10(g_lo, g_hi, (uint32_t)lane, PFO_PURPOSE_NEG_SEQ, k0, k1).x, (uint32_t)n_av);
    for (int j = 1; j < size; ++j) {
        for (;;) {
            const int qj = __shfl_sync(0xffffffffu, q, j);
            if (!__any_sync(0xffffffffu, lane < j && q == qj)) break;
            if (lane == j) {
                ++attempt;
                q = (int)mulhi32(philox4x32_10(g_lo, g_hi, (uint32_t)j | (attempt << 16), PFO_PURPOSE_NEG_SEQ, k0, k1).x,
                                 (uint32_t)n_av);
            }
        }
    }
    return q;
}

// the sampler's rule, a function of the sizes only: sequential rejection when the sample is small against the pool
__device__ __forceinline__ bool use_rejection(int size, int n_av) { return size <= 32 && 8 * size <= n_av; }

struct MvArgs {
    const int64_t* event_ids; const int32_t* day_idx; const int32_t* pos_stock;
    const int64_t* port_ptr; const int32_t* port_items;
    const int32_t* items; int M;
    const double* logret; int n_stocks; int T;
    int B; int K; double gamma; double lam; int n_pos; int n_neg; uint32_t k0, k1; int sample;
    int32_t* cand; double* y_out; int32_t* p_pos; int32_t* p_neg; int item_offset;
};

__global__ void __launch_bounds__(kWarps * 32)
mv_select_kernel(const MvArgs p) {
    pfo_pdl_prologue();
    __shared__ unsigned long long keys[kWarps][kCap];
    __shared__ int counts[kWarps];
    __shared__ int hs[kWarps][32];
    __shared__ double Ssum[kWarps][32];
    extern __shared__ double rows_all[];                // [kWarps][C][T | 1] candidate rows of the return table
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * kWarps + wib;
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    const int K = p.K, C = K + 1, M = p.M, T = p.T;
    const int TS = T | 1;                               // odd row stride (doubles): lane-per-row walks stay conflict-free
    for (int64_t b = warp; b < p.B; b += nwarps) {
        const int64_t g = p.event_ids[b];
        const uint32_t g_lo = (uint32_t)(g & 0xffffffffll), g_hi = (uint32_t)((g >> 32) & 0xffffffffll);
        const int64_t pb = p.port_ptr[b];
        const int nP = (int)(p.port_ptr[b + 1] - pb);
        const int32_t* held = p.port_items + pb;
        const bool small_p = nP <= 32;                  // held positions fit the per-warp table
        int nH = 0;
        if (p.sample) nH = small_p ? held_positions(held, nP, p.items, M, lane, hs[wib])
                                   : count_held_in_universe(held, nP, p.items, M, lane);
        const int n_av = M - nH;
        int my_cand = 0;
        if (lane == 0) my_cand = p.pos_stock[b] - p.item_offset;
        if (!p.sample) {
            if (lane < C) my_cand = p.cand[b * C + lane];
        } else if (n_av < K) {
            if (lane >= 1 && lane <= K) {
                const int j = lane - 1;
                const uint32_t x0 = philox4x32_10(g_lo, g_hi, (uint32_t)j, PFO_PURPOSE_NEG_REPL, p.k0, p.k1).x;
                const int q = (int)mulhi32(x0, (uint32_t)n_av);
                my_cand = p.items[small_p ? skip_held(hs[wib], nH, q) : nth_available(p.items, M, held, nP, q)];
            }
        } else if (use_rejection(K, n_av)) {
            const int q = __shfl_up_sync(0xffffffffu, draw_distinct(g_lo, g_hi, p.k0, p.k1, K, n_av, lane), 1);
            if (lane >= 1 && lane <= K)
                my_cand = p.items[small_p ? skip_held(hs[wib], nH, q) : nth_available(p.items, M, held, nP, q)];
        } else {
            // threshold select: expected count ~ K + 4 sqrt(K) + 8 keys below the cut
            double frac = ((double)K + 4.0 * sqrt((double)K) + 8.0) / (double)n_av;
            unsigned long long cut = frac >= 1.0 ? 0x100000000ull : (unsigned long long)(frac * 4294967296.0);
            for (;;) {
                if (lane == 0) counts[wib] = 0;
                __syncwarp();
                for (int base = 0; base < M; base += 32) {
                    const int pos = base + lane;
                    bool take = false;
                    uint32_t x0 = 0;
                    if (pos < M && !is_held(held, nP, p.items[pos])) {
                        x0 = philox4x32_10(g_lo, g_hi, (uint32_t)pos, PFO_PURPOSE_NEG, p.k0, p.k1).x;
                        take = (unsigned long long)x0 < cut;
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, take);
                    if (m) {
                        int start = 0;
                        if (lane == 0) { start = counts[wib]; counts[wib] = start + __popc(m); }
                        start = __shfl_sync(0xffffffffu, start, 0);
                        if (take) {
                            const int slot = start + __popc(m & ((1u << lane) - 1u));
                            if (slot < kCap) keys[wib][slot] = ((unsigned long long)x0 << 32) | (unsigned)pos;
                        }
                    }
                }
                __syncwarp();
                const int cnt = counts[wib];
                if (cnt > kCap) { cut >>= 1; continue; }
                if (cnt >= K || cut >= 0x100000000ull) break;
                cut <<= 1;
                if (cut > 0x100000000ull) cut = 0x100000000ull;
            }
            const int cnt = counts[wib];
            // rank by counting: the K smallest keys, ascending (keys are unique: position is the low word)
            for (int i = lane; i < cnt; i += 32) {
                const unsigned long long ki = keys[wib][i];
                int rank = 0;
                for (int t = 0; t < cnt; ++t) rank += keys[wib][t] < ki ? 1 : 0;
                if (rank < K) p.cand[b * C + 1 + rank] = p.items[(int)(ki & 0xffffffffull)];
            }
            __syncwarp();
            __threadfence_block();
            if (lane >= 1 && lane <= K) my_cand = p.cand[b * C + lane];
        }
        if (lane < C) p.cand[b * C + lane] = my_cand;

        // ---- y_mv (main.py:243-271), closed form, left-to-right fp64 sums ----
        const double* lr = p.logret + (int64_t)p.day_idx[b] * p.n_stocks * T;
        // the C candidate rows of the day's return table, staged in shared memory with coalesced loads (lane = return
        // column, kStage rows in flight): lane c then walks ITS row from shared memory in the same left-to-right order
        // (bit-identical sums).  Walking the rows straight from global memory cost one 8-byte sector per lane and
        // step -- 4 passes x T dependent round trips to L1 / L2 per interaction.
        double* rows = rows_all + (size_t)wib * C * TS;
        for (int c0 = 0; c0 < C; c0 += kStage) {
            double v[kStage];
#pragma unroll
            for (int u = 0; u < kStage; ++u) {
                const int cu = __shfl_sync(0xffffffffu, my_cand, (c0 + u) & 31);
                v[u] = (c0 + u < C && lane < T) ? __ldg(lr + (int64_t)cu * T + lane) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kStage; ++u)
                if (c0 + u < C && lane < T) rows[(c0 + u) * TS + lane] = v[u];
        }
        if (lane < T) {
            double s = 0.0;
            for (int k = 0; k < nP; ++k) s = s + lr[(int64_t)held[k] * T + lane];
            Ssum[wib][lane] = s;
        }
        __syncwarp();
        double y = 0.0;
        if (lane < C) {
            const double* r = rows + lane * TS;
            double acc = 0.0;
            for (int t = 0; t < T; ++t) acc = acc + r[t];
            const double mu = acc / (double)T;
            double v = 0.0;
            for (int t = 0; t < T; ++t) { const double dc = r[t] - mu; v = v + dc * dc; }
            const double var = v / (double)(T - 1);
            if (nP == 0) {
                y = (mu / p.gamma) / var;
            } else {
                double ms = 0.0;
                for (int t = 0; t < T; ++t) ms = ms + Ssum[wib][t];
                ms = ms / (double)T;
                double cv = 0.0;
                for (int t = 0; t < T; ++t) { const double dc = r[t] - mu; cv = cv + dc * (Ssum[wib][t] - ms); }
                const double cov = cv / (double)(T - 1);
                y = (mu / p.gamma - 0.5 * (cov / (double)nP)) / var;
            }
            if (p.y_out) p.y_out[b * C + lane] = y;
        }
        // ---- rank fusion (main.py:282-292): rankdata(average) + stable argsort ----
        int less = 0, eq = 0;
        for (int c = 0; c < C; ++c) {
            const double yc = __shfl_sync(0xffffffffu, y, c);
            less += yc < y ? 1 : 0;
            eq += yc == y ? 1 : 0;
        }
        const double invest = (double)less + (double)(eq + 1) / 2.0;
        const double pref = (double)(C - lane);
        const double a1 = invest * p.lam;
        const double a2 = pref * (1.0 - p.lam);
        const double nr = a1 + a2;
        int position = 0;
        for (int c = 0; c < C; ++c) {
            const double nc = __shfl_sync(0xffffffffu, nr, c);
            position += (nc < nr || (nc == nr && c < lane)) ? 1 : 0;
        }
        if (lane < C) {
            const int desc = K - position;            // index in argsort(new_rank)[::-1]
            if (desc < p.n_pos) p.p_pos[b * p.n_pos + desc] = my_cand + p.item_offset;
            if (desc >= C - p.n_neg) p.p_neg[b * p.n_neg + (desc - (C - p.n_neg))] = my_cand + p.item_offset;
        }
        __syncwarp();
    }
}

// ---- general candidate sampler (evaluation: size = N_ITEMS) --------------------------
// one CTA per interaction; without replacement = full bitonic sort of (draw, position) keys.
__global__ void __launch_bounds__(256)
sample_candidates_kernel(const int64_t* __restrict__ event_ids, const int64_t* __restrict__ port_ptr,
                         const int32_t* __restrict__ port_items, const int32_t* __restrict__ items, int M,
                         int size, uint32_t k0, uint32_t k1, int pow2, int32_t* __restrict__ out) {
    pfo_pdl_prologue();
    extern __shared__ unsigned long long sk[];      // [pow2]
    __shared__ int n_held_s;
    __shared__ int hs[32];
    const int b = blockIdx.x;
    const int64_t g = event_ids[b];
    const uint32_t g_lo = (uint32_t)(g & 0xffffffffll), g_hi = (uint32_t)((g >> 32) & 0xffffffffll);
    const int64_t pb = port_ptr[b];
    const int nP = (int)(port_ptr[b + 1] - pb);
    const int32_t* held = port_items + pb;
    const bool small_p = nP <= 32;
    if (threadIdx.x < 32) {
        const int c = small_p ? held_positions(held, nP, items, M, threadIdx.x, hs)
                              : count_held_in_universe(held, nP, items, M, threadIdx.x);
        if (threadIdx.x == 0) n_held_s = c;
    }
    __syncthreads();
    const int nH = n_held_s, n_av = M - nH;
    if (n_av < size) {
        for (int j = threadIdx.x; j < size; j += blockDim.x) {
            const uint32_t x0 = philox4x32_10(g_lo, g_hi, (uint32_t)j, PFO_PURPOSE_NEG_REPL, k0, k1).x;
            const int q = (int)mulhi32(x0, (uint32_t)n_av);
            out[(int64_t)b * size + j] = items[small_p ? skip_held(hs, nH, q) : nth_available(items, M, held, nP, q)];
        }
        return;
    }
    if (use_rejection(size, n_av)) {
        if (threadIdx.x < 32) {
            const int q = draw_distinct(g_lo, g_hi, k0, k1, size, n_av, threadIdx.x);
            if (threadIdx.x < size)
                out[(int64_t)b * size + threadIdx.x] = items[small_p ? skip_held(hs, nH, q) : nth_available(items, M, held, nP, q)];
        }
        return;
    }
    for (int pos = threadIdx.x; pos < pow2; pos += blockDim.x) {
        unsigned long long k = ~0ull;
        if (pos < M && !is_held(held, nP, items[pos]))
            k = ((unsigned long long)philox4x32_10(g_lo, g_hi, (uint32_t)pos, PFO_PURPOSE_NEG, k0, k1).x << 32) | (unsigned)pos;
        sk[pos] = k;
    }
    __syncthreads();
    for (int kk = 2; kk <= pow2; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < pow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = sk[i], c = sk[ixj];
                    const bool up = (i & kk) == 0;
                    if ((a > c) == up) { sk[i] = c; sk[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int j = threadIdx.x; j < size; j += blockDim.x)
        out[(int64_t)b * size + j] = items[(int)(sk[j] & 0xffffffffull)];
}

}  // namespace

PFO_API int pfo_mv_select(const int64_t* event_ids, const int32_t* day_idx, const int32_t* pos_stock,
                          const int64_t* port_ptr, const int32_t* port_items,
                          const int32_t* items_sorted, int n_items_universe,
                          const double* logret, int n_stocks, int n_returns,
                          int B, int K, double gamma, double lam, int n_pos, int n_neg, uint64_t seed, int sample,
                          int32_t* cand, double* y_out, int32_t* p_pos, int32_t* p_neg, int item_offset, void* stream) {
    if (B <= 0) return 0;
    if (K < 1 || K > 31 || n_returns > 32 || n_returns < 2 || n_pos + n_neg > K + 1) return (int)cudaErrorInvalidValue;
    MvArgs a{event_ids, day_idx, pos_stock, port_ptr, port_items, items_sorted, n_items_universe,
             logret, n_stocks, n_returns, B, K, gamma, lam, n_pos, n_neg,
             (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), sample, cand, y_out, p_pos, p_neg, item_offset};
    // candidate rows in shared memory: kWarps x (K + 1) rows of (T | 1) doubles (<= 34 KB, inside the default 48 KB
    // with the 10 KB of static tables); the grid is capped at the CTAs that are resident at once
    const size_t smem = (size_t)kWarps * (K + 1) * (n_returns | 1) * sizeof(double);
    // exactly the CTAs that are resident (occupancy API: 7 per SM at K = 20, T = 29).  A guessed 6 left 3 552 warps for
    // 8 192 interactions -- three rounds for a third of the warps -- where 7 x 148 x 4 = 4 144 warps need two (the
    // final capture shows the staging took the long-scoreboard stalls from 7.9 to 1.9 cycles per issue, but the
    // launch time did not move: profiles/r2f_stall_reasons.txt)
    int per_sm = pfo_resident(mv_select_kernel, kWarps * 32, smem);
    if (const char* e = getenv("PFO_MV_CTAS")) { const int v = atoi(e); if (v > 0) per_sm = v; }    // tools/k5_time.py
    pfo_launch(mv_select_kernel, pfo_grid((int64_t)B * 32, kWarps * 32, per_sm), kWarps * 32, smem, (cudaStream_t)stream, a);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_sample_candidates(const int64_t* event_ids, const int64_t* port_ptr, const int32_t* port_items,
                                  const int32_t* items_sorted, int n_items_universe, int B, int size,
                                  uint64_t seed, int32_t* out, void* stream) {
    if (B <= 0 || size <= 0) return 0;
    int pow2 = 1;
    while (pow2 < n_items_universe) pow2 <<= 1;
    if (pow2 > 16384) return (int)cudaErrorInvalidValue;
    const size_t smem = (size_t)pow2 * sizeof(unsigned long long);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(sample_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr_set = true;
    }
    pfo_launch(sample_candidates_kernel, B, 256, smem, (cudaStream_t)stream, 
        event_ids, port_ptr, port_items, items_sorted, n_items_universe, size,
        (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), pow2, out);
    PFO_LAUNCH_CHECK();
}
