// Parameter-side folding of the temporal-attention layer (reference model/temporal_attention.py:26-90 and
// utils/utils.py:4-17): the ten tensors of nn.MultiheadAttention + MergeLayer become the three GEMM operands
// the per-query kernels consume, and the gradients of those operands are carried back to the ten tensors.
//
//   Wqk  [H*ekp, d], cqk [H*ekp] : qk_h = A_h h_q + cA_h,  A_h = s Wk_h^T Wq_h[:, :d],  cA_h = s Wk_h^T cq_h,
//                                  cq = Wq[:, d:] cos(tb) + bq  (TimeEncode(0), embedding_module.py:92), s = 1/sqrt(hd)
//   Wc1T [H*ekp + d, d]          : rows of head h = (W1a Wo_h [Wv_h | bv_h])^T, row `valid` = W1a bo and row
//                                  `one` = b1 in head 0, zero padding rows, then W1b^T
// Everything here is O(weights) -- a few MFLOP per step -- so the arithmetic is done in fp64 (the folded operands
// are then at least as accurate as the reference's chained fp32 products)
// as a handful of 32x32-tiled fp64 GEMM jobs per launch (one CTA per tile, the leftover matvecs / copies in the last
// CTA).  Two launches forward, two backward, instead of ~95 tiny framework kernels per step.
#include "common.cuh"

namespace {

struct FoldArgs {
    const float *Wq, *Wk, *Wv, *b_in, *Wo, *bo, *W1, *b1, *tb;
    int d, F, H, ekp, E, Ek, hd;
    double scale;
    double* ws;               // [te0 (d) | cq (E) | T (H, E, Ek+1) | gcq (E) | gT (H, Ek+1, E)]
    float *Wqk, *cqk, *Wc1T;  // forward outputs
    const float *gWqk, *gcqk, *gWc1T;   // backward inputs
    float *gWq, *gWk, *gWv, *gb_in, *gWo, *gbo, *gW1, *gb1, *gtb;
};

__device__ __forceinline__ double* ws_te0(const FoldArgs& p) { return p.ws; }
__device__ __forceinline__ double* ws_cq(const FoldArgs& p) { return p.ws + p.d; }
__device__ __forceinline__ double* ws_T(const FoldArgs& p) { return p.ws + p.d + p.E; }
__device__ __forceinline__ double* ws_gcq(const FoldArgs& p) { return ws_T(p) + (size_t)p.H * p.E * (p.Ek + 1); }
__device__ __forceinline__ double* ws_gT(const FoldArgs& p) { return ws_gcq(p) + p.E; }

// [Wv_h | bv_h](i, r)
__device__ __forceinline__ double wva(const FoldArgs& p, int h, int i, int r) {
    const int row = h * p.hd + i;
    return r < p.Ek ? (double)p.Wv[(size_t)row * p.Ek + r] : (double)p.b_in[2 * p.E + row];
}

// ---- one 32x32 output tile of C[m, n] = sum_k A(m, k) B(k, n) in fp64: 256 threads, 2x2 outputs each, K in chunks
// of 32 through shared memory.  A_MC / B_NC say which index of the operand is contiguous in memory so that the
// tile loads coalesce (A_MC: m contiguous, else k; B_NC: n contiguous, else k).
constexpr int TS = 32;

template <bool A_MC, bool B_NC, class FA, class FB, class FC>
__device__ __forceinline__ void gemm_tile(double (*As)[TS + 1], double (*Bs)[TS + 1], int M, int N, int K, int tm, int tn,
                                          FA fa, FB fb, FC fc) {
    const int t = threadIdx.x;
    const int ty = t >> 4, tx = t & 15;
    const int m0 = tm * TS, n0 = tn * TS;
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int k0 = 0; k0 < K; k0 += TS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = t + 256 * j;
            const int fast = idx & 31, slow = idx >> 5;
            {
                const int m = A_MC ? fast : slow, k = A_MC ? slow : fast;
                As[m][k] = (m0 + m < M && k0 + k < K) ? fa(m0 + m, k0 + k) : 0.0;
            }
            {
                const int n = B_NC ? fast : slow, k = B_NC ? slow : fast;
                Bs[k][n] = (n0 + n < N && k0 + k < K) ? fb(k0 + k, n0 + n) : 0.0;
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < TS; ++k) {
            const double a0 = As[ty * 2][k], a1 = As[ty * 2 + 1][k];
            const double b0 = Bs[k][tx * 2], b1 = Bs[k][tx * 2 + 1];
            acc[0][0] = fma(a0, b0, acc[0][0]); acc[0][1] = fma(a0, b1, acc[0][1]);
            acc[1][0] = fma(a1, b0, acc[1][0]); acc[1][1] = fma(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int m = m0 + ty * 2 + i, n = n0 + tx * 2 + j;
            if (m < M && n < N) fc(m, n, acc[i][j]);
        }
}

__host__ __device__ __forceinline__ int tiles(int x) { return (x + TS - 1) / TS; }

// a CTA takes the job `job` of a group of `batch` GEMMs with tiles(M) x tiles(N) tiles each; returns false (and
// rebases job) when the job belongs to a later group
#define FOLD_GROUP(batch, M, N)                                                        \
    const int tmn_ = tiles(M) * tiles(N);                                              \
    if (job >= (batch) * tmn_) { job -= (batch) * tmn_; } else

// ---- forward stage 1: T_h = Wo_h [Wv_h | bv_h] | Wqk_h = s Wk_h^T Wq_h[:, :d] | te0, cq, padding rows of Wqk
__global__ void __launch_bounds__(256) fold_fwd1_kernel(const FoldArgs p) {
    pfo_pdl_prologue();
    const int d = p.d, E = p.E, Ek = p.Ek, H = p.H, hd = p.hd, ekp = p.ekp, R = Ek + 1;
    __shared__ double As[TS][TS + 1], Bs[TS][TS + 1];   // A tile [m][k], B tile [k][n]
    int job = blockIdx.x;
    { FOLD_GROUP(H, E, R) {
        const int h = job / tmn_, t = job % tmn_;
        double* T = ws_T(p) + (size_t)h * E * R;
        gemm_tile<false, true>(As, Bs, E, R, hd, t / tiles(R), t % tiles(R),
            [&](int m, int k) { return (double)p.Wo[(size_t)m * E + h * hd + k]; },
            [&](int k, int n) { return wva(p, h, k, n); },
            [&](int m, int n, double v) { T[(size_t)m * R + n] = v; });
        return; } }
    { FOLD_GROUP(H, Ek, d) {
        const int h = job / tmn_, t = job % tmn_;
        gemm_tile<true, true>(As, Bs, Ek, d, hd, t / tiles(d), t % tiles(d),
            [&](int m, int k) { return (double)p.Wk[(size_t)(h * hd + k) * Ek + m]; },
            [&](int k, int n) { return (double)p.Wq[(size_t)(h * hd + k) * E + n]; },
            [&](int m, int n, double v) { p.Wqk[(size_t)(h * ekp + m) * d + n] = (float)(p.scale * v); });
        return; } }
    // last CTA: te0, cq (warp per output, lanes along k) and the zero padding rows of Wqk
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* te0 = &As[0][0];                                       // d <= 128 doubles; the GEMM tiles are idle here
    for (int k = threadIdx.x; k < d; k += 256) { te0[k] = cos((double)p.tb[k]); ws_te0(p)[k] = te0[k]; }
    __syncthreads();
    for (int e = w; e < E; e += 8) {
        double s = 0.0;
        for (int k = lane; k < d; k += 32) s += (double)p.Wq[(size_t)e * E + d + k] * te0[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) ws_cq(p)[e] = s + (double)p.b_in[e];
    }
    const int pad = ekp - Ek;
    for (int i = threadIdx.x; i < H * pad * d; i += 256) {
        const int c = i % d, r = Ek + (i / d) % pad, h = i / (d * pad);
        p.Wqk[(size_t)(h * ekp + r) * d + c] = 0.0f;
    }
}
static int fold_fwd1_jobs(int d, int F, int H) { const int E = 2 * d, Ek = E + F; return H * tiles(E) * tiles(Ek + 1) + H * tiles(Ek) * tiles(d) + 1; }

// ---- forward stage 2: Wc1T head rows = (W1a T_h)^T | cqk, `valid` / `one` / padding rows, W1b^T rows
__global__ void __launch_bounds__(256) fold_fwd2_kernel(const FoldArgs p) {
    pfo_pdl_prologue();
    const int d = p.d, E = p.E, Ek = p.Ek, H = p.H, hd = p.hd, ekp = p.ekp, R = Ek + 1;
    __shared__ double As[TS][TS + 1], Bs[TS][TS + 1];   // A tile [m][k], B tile [k][n]
    int job = blockIdx.x;
    { FOLD_GROUP(H, d, R) {
        const int h = job / tmn_, t = job % tmn_;
        const double* T = ws_T(p) + (size_t)h * E * R;
        gemm_tile<false, true>(As, Bs, d, R, E, t / tiles(R), t % tiles(R),
            [&](int m, int k) { return (double)p.W1[(size_t)m * (E + d) + k]; },
            [&](int k, int n) { return T[(size_t)k * R + n]; },
            [&](int m, int n, double v) { p.Wc1T[(size_t)(h * ekp + n) * d + m] = (float)v; });
        return; } }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // cqk[h, r] = s sum_t Wk[h*hd+t, r] cq[h*hd+t]: thread per r (coalesced along r), 8 partial sums over t
    double (*part)[TS + 1] = As;                                   // the GEMM tiles are idle in the last CTA
    for (int base = 0; base < H * ekp; base += TS) {
        const int i = base + lane;
        const int h = i / ekp, r = i % ekp;
        double s = 0.0;
        if (i < H * ekp && r < Ek)
            for (int t = w; t < hd; t += 8) s += (double)p.Wk[(size_t)(h * hd + t) * Ek + r] * ws_cq(p)[h * hd + t];
        part[w][lane] = s;
        __syncthreads();
        if (w == 0 && i < H * ekp) {
            double tsum = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) tsum += part[k][lane];
            p.cqk[i] = (float)(p.scale * tsum);
        }
        __syncthreads();
    }
    // `valid` row of head 0 = W1a bo (warp per output column c, lanes along e)
    for (int c = w; c < d; c += 8) {
        double s = 0.0;
        for (int e = lane; e < E; e += 32) s += (double)p.W1[(size_t)c * (E + d) + e] * (double)p.bo[e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) p.Wc1T[(size_t)R * d + c] = (float)s;
    }
    // `one` row of head 0 = b1; remaining tail rows zero; W1b^T
    const int tail = ekp - R;
    for (int i = threadIdx.x; i < H * tail * d; i += 256) {
        const int c = i % d, r = R + (i / d) % tail, h = i / (d * tail);
        if (h == 0 && r == R) continue;
        p.Wc1T[(size_t)(h * ekp + r) * d + c] = (h == 0 && r == R + 1) ? p.b1[c] : 0.0f;
    }
    for (int i = threadIdx.x; i < d * d; i += 256) {
        const int c = i % d, k = i / d;
        p.Wc1T[(size_t)(H * ekp + k) * d + c] = p.W1[(size_t)c * (E + d) + E + k];
    }
}
static int fold_fwd2_jobs(int d, int F, int H) { const int Ek = 2 * d + F; return H * tiles(d) * tiles(Ek + 1) + 1; }

// ---- backward stage 1 (needs T and cq of the forward):
//   gW1a | gT_h = (W1a^T gB_h)^T (layout [h][r][e]) | gWq[:, :d] | gWk | gbo, gcq, gW1b, gb1
__global__ void __launch_bounds__(256) fold_bwd1_kernel(const FoldArgs p) {
    pfo_pdl_prologue();
    const int d = p.d, E = p.E, Ek = p.Ek, H = p.H, hd = p.hd, ekp = p.ekp, R = Ek + 1;
    const float* gc1 = p.gWc1T + (size_t)R * d;                   // head 0, `valid` row
    __shared__ double As[TS][TS + 1], Bs[TS][TS + 1];   // A tile [m][k], B tile [k][n]
    int job = blockIdx.x;
    { FOLD_GROUP(1, d, E) {                                        // gW1a[c, e], K = (h, r) pairs + the rank-1 term gc1 bo^T
        const int t = job;
        gemm_tile<true, false>(As, Bs, d, E, H * R + 1, t / tiles(E), t % tiles(E),
            [&](int m, int k) { return k < H * R ? (double)p.gWc1T[(size_t)((k / R) * ekp + (k % R)) * d + m] : (double)gc1[m]; },
            [&](int k, int n) { return k < H * R ? ws_T(p)[((size_t)(k / R) * E + n) * R + (k % R)] : (double)p.bo[n]; },
            [&](int m, int n, double v) { p.gW1[(size_t)m * (E + d) + n] = (float)v; });
        return; } }
    { FOLD_GROUP(H, R, E) {                                        // gT[h][r][e]
        const int h = job / tmn_, t = job % tmn_;
        double* gT = ws_gT(p) + (size_t)h * R * E;
        gemm_tile<false, true>(As, Bs, R, E, d, t / tiles(E), t % tiles(E),
            [&](int m, int k) { return (double)p.gWc1T[(size_t)(h * ekp + m) * d + k]; },
            [&](int k, int n) { return (double)p.W1[(size_t)k * (E + d) + n]; },
            [&](int m, int n, double v) { gT[(size_t)m * E + n] = v; });
        return; } }
    { FOLD_GROUP(H, hd, d) {                                       // gWq[h*hd + m, c]
        const int h = job / tmn_, t = job % tmn_;
        gemm_tile<false, true>(As, Bs, hd, d, Ek, t / tiles(d), t % tiles(d),
            [&](int m, int k) { return (double)p.Wk[(size_t)(h * hd + m) * Ek + k]; },
            [&](int k, int n) { return (double)p.gWqk[(size_t)(h * ekp + k) * d + n]; },
            [&](int m, int n, double v) { p.gWq[(size_t)(h * hd + m) * E + n] = (float)(p.scale * v); });
        return; } }
    { FOLD_GROUP(H, hd, Ek) {                                      // gWk[h*hd + m, r], K = c plus the rank-1 term cq gcqk^T
        const int h = job / tmn_, t = job % tmn_;
        gemm_tile<false, false>(As, Bs, hd, Ek, d + 1, t / tiles(Ek), t % tiles(Ek),
            [&](int m, int k) { return k < d ? (double)p.Wq[(size_t)(h * hd + m) * E + k] : ws_cq(p)[h * hd + m]; },
            [&](int k, int n) { return k < d ? (double)p.gWqk[(size_t)(h * ekp + n) * d + k] : (double)p.gcqk[h * ekp + n]; },
            [&](int m, int n, double v) { p.gWk[(size_t)(h * hd + m) * Ek + n] = (float)(p.scale * v); });
        return; } }
    // last CTA: matvecs and copies
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double (*part)[TS + 1] = As;                                   // the GEMM tiles are idle in the last CTA
    for (int base = 0; base < E; base += TS) {                     // gbo[e] = sum_c W1[c, e] gc1[c]: thread per e
        const int e = base + lane;
        double s = 0.0;
        if (e < E) for (int c = w; c < d; c += 8) s += (double)p.W1[(size_t)c * (E + d) + e] * (double)gc1[c];
        part[w][lane] = s;
        __syncthreads();
        if (w == 0 && e < E) {
            double tsum = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) tsum += part[k][lane];
            p.gbo[e] = (float)tsum;
        }
        __syncthreads();
    }
    for (int row = w; row < E; row += 8) {                         // gcq[row] = s sum_r Wk[row, r] gcqk[h, r]: warp per row
        const int h = row / hd;
        double s = 0.0;
        for (int r = lane; r < Ek; r += 32) s += (double)p.Wk[(size_t)row * Ek + r] * (double)p.gcqk[h * ekp + r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) ws_gcq(p)[row] = p.scale * s;
    }
    for (int i = threadIdx.x; i < d * d; i += 256) {               // gW1b[c, k] = gWc1T[H*ekp + k, c]
        const int c = i % d, k = i / d;
        p.gW1[(size_t)c * (E + d) + E + k] = p.gWc1T[(size_t)(H * ekp + k) * d + c];
    }
    for (int c = threadIdx.x; c < d; c += 256) p.gb1[c] = p.gWc1T[(size_t)(R + 1) * d + c];   // head 0, `one` row
}
static int fold_bwd1_jobs(int d, int F, int H) {
    const int E = 2 * d, Ek = E + F, R = Ek + 1, hd = E / H;
    return tiles(d) * tiles(E) + H * tiles(R) * tiles(E) + H * tiles(hd) * tiles(d) + H * tiles(hd) * tiles(Ek) + 1;
}

// ---- backward stage 2 (needs gT and gcq): gWo | gWv, gbv | gb_in (q, k parts), gWq[:, d:], gtb
__global__ void __launch_bounds__(256) fold_bwd2_kernel(const FoldArgs p) {
    pfo_pdl_prologue();
    const int d = p.d, E = p.E, Ek = p.Ek, H = p.H, hd = p.hd, R = Ek + 1;
    __shared__ double As[TS][TS + 1], Bs[TS][TS + 1];   // A tile [m][k], B tile [k][n]
    int job = blockIdx.x;
    { FOLD_GROUP(H, E, hd) {                                       // gWo[e, h*hd + t]
        const int h = job / tmn_, t = job % tmn_;
        const double* gT = ws_gT(p) + (size_t)h * R * E;
        gemm_tile<true, false>(As, Bs, E, hd, R, t / tiles(hd), t % tiles(hd),
            [&](int m, int k) { return gT[(size_t)k * E + m]; },
            [&](int k, int n) { return wva(p, h, n, k); },
            [&](int m, int n, double v) { p.gWo[(size_t)m * E + h * hd + n] = (float)v; });
        return; } }
    { FOLD_GROUP(H, hd, R) {                                       // g[Wv | bv][h*hd + t, r]
        const int h = job / tmn_, t = job % tmn_;
        const double* gT = ws_gT(p) + (size_t)h * R * E;
        gemm_tile<true, false>(As, Bs, hd, R, E, t / tiles(R), t % tiles(R),
            [&](int m, int k) { return (double)p.Wo[(size_t)k * E + h * hd + m]; },
            [&](int k, int n) { return gT[(size_t)n * E + k]; },
            [&](int m, int n, double v) {
                const int row = h * hd + m;
                if (n < Ek) p.gWv[(size_t)row * Ek + n] = (float)v; else p.gb_in[2 * E + row] = (float)v;
            });
        return; } }
    // last CTA
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < 2 * E; e += 256)                 // query part = gcq, key part = 0 (cancels in the softmax)
        p.gb_in[e] = e < E ? (float)ws_gcq(p)[e] : 0.0f;
    for (int i = threadIdx.x; i < E * d; i += 256) {               // gWq[e, d + k] = gcq[e] te0[k]
        const int k = i % d, e = i / d;
        p.gWq[(size_t)e * E + d + k] = (float)(ws_gcq(p)[e] * ws_te0(p)[k]);
    }
    double (*part)[TS + 1] = As;                                   // the GEMM tiles are idle in the last CTA
    for (int base = 0; base < d; base += TS) {                     // gtb[k] = -sin(tb[k]) sum_e Wq[e, d+k] gcq[e]: thread per k
        const int k = base + lane;
        double s = 0.0;
        if (k < d) for (int e = w; e < E; e += 8) s += (double)p.Wq[(size_t)e * E + d + k] * ws_gcq(p)[e];
        part[w][lane] = s;
        __syncthreads();
        if (w == 0 && k < d) {
            double tsum = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) tsum += part[j][lane];
            p.gtb[k] = (float)(-sin((double)p.tb[k]) * tsum);
        }
        __syncthreads();
    }
}
static int fold_bwd2_jobs(int d, int F, int H) {
    const int E = 2 * d, Ek = E + F, R = Ek + 1, hd = E / H;
    return H * tiles(E) * tiles(hd) + H * tiles(hd) * tiles(R) + 1;
}

int fill(FoldArgs& a, const float* Wq, const float* Wk, const float* Wv, const float* b_in, const float* Wo,
         const float* bo, const float* W1, const float* b1, const float* tb, int d, int F, int H, int ekp, double* ws) {
    if (d <= 0 || F < 0 || H <= 0 || (2 * d) % H != 0 || ekp < 2 * d + F + 3 || ws == nullptr) return 1;
    a.Wq = Wq; a.Wk = Wk; a.Wv = Wv; a.b_in = b_in; a.Wo = Wo; a.bo = bo; a.W1 = W1; a.b1 = b1; a.tb = tb;
    a.d = d; a.F = F; a.H = H; a.ekp = ekp; a.E = 2 * d; a.Ek = 2 * d + F; a.hd = 2 * d / H;
    a.scale = 1.0 / sqrt((double)a.hd);
    a.ws = ws;
    return 0;
}

}  // namespace

PFO_API int64_t pfo_fold_attention_workspace_doubles(int d, int F, int H) {
    const int64_t E = 2 * d, R = 2 * d + F + 1;
    return d + E + H * E * R + E + H * R * E;
}

PFO_API int pfo_fold_attention_fwd(const float* Wq, const float* Wk, const float* Wv, const float* b_in, const float* Wo,
                                   const float* bo, const float* W1, const float* b1, const float* tb,
                                   int d, int F, int H, int ekp, double* workspace,
                                   float* Wqk, float* cqk, float* Wc1T, void* stream) {
    FoldArgs a{};
    if (fill(a, Wq, Wk, Wv, b_in, Wo, bo, W1, b1, tb, d, F, H, ekp, workspace)) return (int)cudaErrorInvalidValue;
    a.Wqk = Wqk; a.cqk = cqk; a.Wc1T = Wc1T;
    cudaStream_t s = (cudaStream_t)stream;
    pfo_launch(fold_fwd1_kernel, fold_fwd1_jobs(d, F, H), 256, 0, s, a);
    pfo_launch(fold_fwd2_kernel, fold_fwd2_jobs(d, F, H), 256, 0, s, a);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_fold_attention_bwd(const float* Wq, const float* Wk, const float* Wv, const float* b_in, const float* Wo,
                                   const float* bo, const float* W1, const float* b1, const float* tb,
                                   int d, int F, int H, int ekp, double* workspace,
                                   const float* gWqk, const float* gcqk, const float* gWc1T,
                                   float* gWq, float* gWk, float* gWv, float* gb_in, float* gWo, float* gbo,
                                   float* gW1, float* gb1, float* gtb, void* stream) {
    FoldArgs a{};
    if (fill(a, Wq, Wk, Wv, b_in, Wo, bo, W1, b1, tb, d, F, H, ekp, workspace)) return (int)cudaErrorInvalidValue;
    a.gWqk = gWqk; a.gcqk = gcqk; a.gWc1T = gWc1T;
    a.gWq = gWq; a.gWk = gWk; a.gWv = gWv; a.gb_in = gb_in; a.gWo = gWo; a.gbo = gbo; a.gW1 = gW1; a.gb1 = gb1; a.gtb = gtb;
    cudaStream_t s = (cudaStream_t)stream;
    pfo_launch(fold_bwd1_kernel, fold_bwd1_jobs(d, F, H), 256, 0, s, a);
    pfo_launch(fold_bwd2_kernel, fold_bwd2_jobs(d, F, H), 256, 0, s, a);
    PFO_LAUNCH_CHECK();
}
