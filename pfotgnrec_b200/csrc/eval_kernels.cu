// Evaluation metric block: Recall/NDCG@{1,3,5} and the change of annualised return / Sharpe ratio of the held
// portfolio when the top-k recommended stocks are appended, in and out of sample -- reference
// evaluation.py:127-207 (the per-interaction Python loop: np.log / np.mean / np.std on a handful of 30-price rows,
// six return_sharpe_at_k calls, :23-36) and :209-258 (means and fraction-positive over the interactions).
// The reference copies every score to the host and loops in Python; here the ranking (pfo_eval_score) and this
// block stay on the device and only 31 doubles per evaluation run ever cross to the host.
//
// Compiled with -fmad=false and written to round exactly like numpy fp64: rows are added one after another
// (np.mean(axis=0) over the [n_stocks, T] block), sums over the T daily returns follow numpy's pairwise_sum
// (8 strided accumulators, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), remainder added in order), np.std is
// sqrt(mean((x - mean)^2)).  The daily log-returns log(p[1:]/p[:-1]) are computed once on the host with numpy.
// One warp per interaction, lane t owns daily return t (T <= 32).
#include "common.cuh"

namespace {

constexpr int kCols = 18;      // recall x3 | ndcg x3 | return_in x3 | sharpe_in x3 | return_out x3 | sharpe_out x3
constexpr int kAcc = 31;       // 18 sums | 12 positive counts (columns 6..17) | number of interactions

__device__ __forceinline__ double bcast(double v, int lane) { return __shfl_sync(0xffffffffu, v, lane); }

// numpy pairwise_sum over lanes 0..T-1 (T <= 32 < PW_BLOCKSIZE); every lane returns the sum
__device__ double np_sum(double v, int T) {
    if (T < 8) {
        double r = 0.0;
        for (int i = 0; i < T; ++i) r += bcast(v, i);
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = bcast(v, j);
    int i = 8;
    for (; i < T - (T % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += bcast(v, i + j);
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < T; ++i) res += bcast(v, i);
    return res;
}

// (mean * 251, (mean * 251) / (std * sqrt(251))) of the daily returns held one per lane (evaluation.py:31-34)
__device__ void perf(double daily, int T, double* ret, double* sharpe) {
    const double mean = np_sum(daily, T) / (double)T;
    const double dev = daily - mean;
    const double var = np_sum(dev * dev, T) / (double)T;
    const double sd = sqrt(var);
    *ret = mean * 251.0;
    *sharpe = (mean * 251.0) / (sd * 15.84297951775486);      // np.sqrt(251)
}

__global__ void __launch_bounds__(128)
eval_metrics_kernel(const int32_t* __restrict__ pos_rank, const int32_t* __restrict__ top_idx, int topk,
                    const int32_t* __restrict__ pos_item, const int32_t* __restrict__ cand, int n_cand,
                    int item_offset, const int32_t* __restrict__ day_idx, const int64_t* __restrict__ port_ptr,
                    const int32_t* __restrict__ port_items, const double* __restrict__ lr_past,
                    const double* __restrict__ lr_future, int n_stocks, int T, int B,
                    double* __restrict__ per_event) {
    pfo_pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    // 1 / np.log2(rank + 2), rank = 0..4 (evaluation.py:15-21 with a single relevant item: idcg = 1)
    const double inv_log2[5] = {1.0, 0.6309297535714575, 0.5, 0.43067655807339306, 0.38685280723454163};
    const int ks[3] = {1, 3, 5};
    for (int b = warp; b < B; b += n_warps) {
        double* out = per_event + (int64_t)b * kCols;
        const int rank = pos_rank[b];
        if (lane < 3) {
            const bool hit = rank < ks[lane];
            out[lane] = hit ? 1.0 : 0.0;
            out[3 + lane] = hit ? inv_log2[rank] : 0.0;
        }
        // stocks of the five best-ranked candidates (position 0 = the true item, evaluation.py:178-182)
        int top_stock = 0;
        if (lane < 5) {
            const int p = top_idx[(int64_t)b * topk + lane];
            top_stock = (p == 0 ? pos_item[b] : cand[(int64_t)b * n_cand + (p - 1)]) - item_offset;
        }
        const int64_t p0 = port_ptr[b];
        const int nP = (int)(port_ptr[b + 1] - p0);
        const int64_t day_off = (int64_t)day_idx[b] * n_stocks;
        for (int s = 0; s < 2; ++s) {
            const double* lr = (s == 0 ? lr_past : lr_future) + day_off * T;
            double sum = 0.0;                                   // rows added one after another (np.mean(axis=0))
            for (int j = 0; j < nP; ++j) {
                const double v = lane < T ? lr[(int64_t)port_items[p0 + j] * T + lane] : 0.0;
                sum = j == 0 ? v : sum + v;
            }
            double ret0 = 0.0, sharpe0 = 0.0;                   // empty portfolio: evaluation.py:155-159
            if (nP > 0) perf(sum / (double)nP, T, &ret0, &sharpe0);
            int rows = nP, kk = 0;
            for (int i = 0; i < 5; ++i) {
                const int stock = __shfl_sync(0xffffffffu, top_stock, i);
                const double v = lane < T ? lr[(int64_t)stock * T + lane] : 0.0;
                sum = rows == 0 ? v : sum + v;
                ++rows;
                if (i + 1 == ks[kk]) {
                    double r, sh;
                    perf(sum / (double)rows, T, &r, &sh);
                    if (lane == 0) {
                        out[6 + 6 * s + kk] = r - ret0;
                        out[9 + 6 * s + kk] = sh - sharpe0;
                    }
                    ++kk;
                }
            }
        }
    }
}

// acc[c] += sum_b per_event[b, c] (c < 18), acc[18 + c - 6] += #{b : per_event[b, c] > 0} (c = 6..17),
// acc[30] += B; one CTA, fixed summation order (deterministic)
__global__ void __launch_bounds__(1024)
eval_metrics_reduce_kernel(const double* __restrict__ per_event, int B, double* __restrict__ acc) {
    pfo_pdl_prologue();
    __shared__ double sh_sum[32][kCols];
    __shared__ double sh_pos[32][kCols];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int c = lane; c < kCols; c += 32) {
        double s = 0.0, p = 0.0;
        for (int b = w; b < B; b += nw) {
            const double v = per_event[(int64_t)b * kCols + c];
            s += v;
            p += v > 0.0 ? 1.0 : 0.0;
        }
        sh_sum[w][c] = s;
        sh_pos[w][c] = p;
    }
    __syncthreads();
    if (threadIdx.x < kCols) {
        const int c = threadIdx.x;
        double s = 0.0, p = 0.0;
        for (int i = 0; i < nw; ++i) { s += sh_sum[i][c]; p += sh_pos[i][c]; }
        acc[c] += s;
        if (c >= 6) acc[kCols + c - 6] += p;
        if (c == 0) acc[kAcc - 1] += (double)B;
    }
}

}  // namespace

PFO_API int pfo_eval_metrics(const int32_t* pos_rank, const int32_t* top_idx, int topk, const int32_t* pos_item,
                             const int32_t* cand, int n_cand, int item_offset, const int32_t* day_idx,
                             const int64_t* port_ptr, const int32_t* port_items, const double* logret_past,
                             const double* logret_future, int n_stocks, int n_returns, int B,
                             double* per_event, double* acc, void* stream) {
    if (B <= 0) return 0;
    if (n_returns < 1 || n_returns > 32 || topk < 5 || n_cand + 1 < 5) return (int)cudaErrorInvalidValue;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = pfo_grid((int64_t)B * 32, 128, 16);
    pfo_launch(eval_metrics_kernel, grid, 128, 0, s, pos_rank, top_idx, topk, pos_item, cand, n_cand, item_offset, day_idx,
                                             port_ptr, port_items, logret_past, logret_future, n_stocks, n_returns, B,
                                             per_event);
    if (acc != nullptr) pfo_launch(eval_metrics_reduce_kernel, 1, 1024, 0, s, per_event, B, acc);
    PFO_LAUNCH_CHECK();
}
