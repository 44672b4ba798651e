// Tall-skinny dense contractions of the hot path in exact fp32 (FFMA) arithmetic.
//
// Every dense contraction of the reference's path -- GRUCell/RNNCell (reference
// modules/memory_updater.py:31,47), the multi-head-attention projections
// (model/temporal_attention.py:70) and the merge MLP (utils/utils.py:14-17) -- has a huge M
// (rows = unique nodes or queries) and tiny N, K (<= 320).  Two kernels cover all of them:
//   linear : C[m, :] = epi(alpha * (A[row(m), :K] . W^T + bias))        (forward and dgrad)
//   wgrad  : dW[n, k] = sum_m G[m, n] * A[row(m), k],  db[n] = sum_m G[m, n]
// A rows may be gathered through an index (row(m) = a_idx[m]; negative = zero row) so that
// node tables are read in place.  M may live on the device (m_dev) so that data-dependent
// row counts (unique touched nodes) never cost a host sync.
// This is the 1e-5 "exact" mode; the bf16 tcgen05 path lives in linear_tc.cu.
#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 16, LDA_S = BM + 4;

struct LinearArgs {
    const float* A; int64_t lda; const int32_t* a_idx;
    const float* W; int64_t ldw; int w_transposed;      // W(n,k) = w_transposed ? W[k*ldw+n] : W[n*ldw+k]
    const float* bias;                                   // [N] or null
    const float* bias_row_scale; int64_t ld_brs;         // optional per-row multiplier of the bias
    float* C; int64_t ldc;
    int64_t M; const int32_t* m_dev; int N; int K;
    float alpha; int act;                                // act: 0 none, 1 relu
    const int32_t* row_zero;                             // rows with row_zero[m] != 0 are written as 0
    const float* relu_gate; int64_t ld_gate;             // output zeroed where gate <= 0 (relu backward)
    int accumulate;                                      // C += result
};

template <int TN>
__global__ void __launch_bounds__(256)
linear_kernel(const LinearArgs p) {
    pfo_pdl_prologue();
    constexpr int BN = 16 * TN;
    constexpr int LDW_S = BN + 4;
    __shared__ __align__(16) float As[2][BK][LDA_S];
    __shared__ __align__(16) float Ws[2][BK][LDW_S];
    int64_t M = p.M;
    if (p.m_dev) { int64_t md = *p.m_dev; M = md < M ? md : M; }
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    if (m0 >= M) return;
    const int n0 = blockIdx.y * BN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int K = p.K, N = p.N;

    // global -> register staging maps
    const int a_k = tid & 15, a_r = tid >> 4;            // rows a_r + 16*i, i < 8
    const float* a_ptr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + a_r + 16 * i;
        a_ptr[i] = nullptr;
        if (m < M) {
            int64_t row = m;
            if (p.a_idx) row = p.a_idx[m];
            if (row >= 0) a_ptr[i] = p.A + row * p.lda;
        }
    }
    float a_reg[8], w_reg[TN];
    auto load_chunk = [&](int k0) {
        const int k = k0 + a_k;
#pragma unroll
        for (int i = 0; i < 8; ++i) a_reg[i] = (a_ptr[i] && k < K) ? __ldg(a_ptr[i] + k) : 0.0f;
        if (p.w_transposed) {
            const int kk = k0 + (tid >> 4);              // 16 k rows, n = tx + 16*i
#pragma unroll
            for (int i = 0; i < TN; ++i) {
                const int n = n0 + tx + 16 * i;
                w_reg[i] = (kk < K && n < N) ? __ldg(p.W + (int64_t)kk * p.ldw + n) : 0.0f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < TN; ++i) {
                const int n = n0 + (tid >> 4) + 16 * i;  // k = tx
                w_reg[i] = (k < K && n < N) ? __ldg(p.W + (int64_t)n * p.ldw + k) : 0.0f;
            }
        }
    };
    auto store_chunk = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[buf][a_k][a_r + 16 * i] = a_reg[i];
        if (p.w_transposed) {
#pragma unroll
            for (int i = 0; i < TN; ++i) Ws[buf][tid >> 4][tx + 16 * i] = w_reg[i];
        } else {
#pragma unroll
            for (int i = 0; i < TN; ++i) Ws[buf][tx][(tid >> 4) + 16 * i] = w_reg[i];
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    load_chunk(0);
    store_chunk(0);
    __syncthreads();
    const int n_chunks = (K + BK - 1) / BK;
    for (int c = 0; c < n_chunks; ++c) {
        const int buf = c & 1;
        if (c + 1 < n_chunks) load_chunk((c + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[8], b[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
            for (int j4 = 0; j4 < TN / 4; ++j4) {
                const float4 bv = *reinterpret_cast<const float4*>(&Ws[buf][kk][j4 * 64 + tx * 4]);
                b[j4 * 4 + 0] = bv.x; b[j4 * 4 + 1] = bv.y; b[j4 * 4 + 2] = bv.z; b[j4 * 4 + 3] = bv.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (c + 1 < n_chunks) {
            store_chunk(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + ty * 8 + i;
        if (m >= M) continue;
        const bool zero_row = p.row_zero && p.row_zero[m] != 0;
        const float brs = p.bias_row_scale ? p.bias_row_scale[m * p.ld_brs] : 1.0f;
#pragma unroll
        for (int j4 = 0; j4 < TN / 4; ++j4) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int n = n0 + j4 * 64 + tx * 4 + c;
                if (n >= N) continue;
                float v = acc[i][j4 * 4 + c];
                if (p.bias) v += p.bias[n] * brs;
                v *= p.alpha;
                if (p.act == 1) v = fmaxf(v, 0.0f);
                if (p.relu_gate && p.relu_gate[m * p.ld_gate + n] <= 0.0f) v = 0.0f;
                if (zero_row) v = 0.0f;
                float* dst = p.C + m * p.ldc + n;
                if (p.accumulate) v += *dst;
                *dst = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------ wgrad
constexpr int WT = 64, WR = 16, LDG_S = WT + 4;

struct WgradArgs {
    const float* G; int64_t ldg;                 // [M, N]
    const float* A; int64_t lda; const int32_t* a_idx;   // [*, K] (gathered by a_idx if given)
    float* partial;                              // [S, N, Kaug]
    int64_t M; const int32_t* m_dev; int N; int K; int Kaug;   // Kaug = K + 1 when the bias column is on
    int S;
};

__global__ void __launch_bounds__(256)
wgrad_kernel(const WgradArgs p) {
    pfo_pdl_prologue();
    __shared__ __align__(16) float Gs[WR][LDG_S];
    __shared__ __align__(16) float As[WR][LDG_S];
    int64_t M = p.M;
    if (p.m_dev) { int64_t md = *p.m_dev; M = md < M ? md : M; }
    const int tiles_k = (p.Kaug + WT - 1) / WT;
    const int tn = blockIdx.x / tiles_k, tk = blockIdx.x % tiles_k;
    const int n0 = tn * WT, k0 = tk * WT;
    const int64_t slab = (M + p.S - 1) / p.S;
    const int64_t r_begin = (int64_t)blockIdx.y * slab;
    int64_t r_end = r_begin + slab;
    if (r_end > M) r_end = M;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lc = tid & 63, lr = tid >> 6;      // loader: column lc, rows lr + 4*i
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int64_t r0 = r_begin; r0 < r_end; r0 += WR) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = lr + 4 * i;
            const int64_t m = r0 + rr;
            float g = 0.0f, a = 0.0f;
            if (m < r_end) {
                const int n = n0 + lc, k = k0 + lc;
                if (n < p.N) g = __ldg(p.G + m * p.ldg + n);
                if (k < p.K) {
                    int64_t row = m;
                    if (p.a_idx) row = p.a_idx[m];
                    if (row >= 0) a = __ldg(p.A + row * p.lda + k);
                } else if (k == p.K && p.Kaug > p.K) {
                    a = 1.0f;                    // virtual ones column -> bias gradient
                }
            }
            Gs[rr][lc] = g;
            As[rr][lc] = a;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < WR; ++rr) {
            const float4 g4 = *reinterpret_cast<const float4*>(&Gs[rr][ty * 4]);
            const float4 a4 = *reinterpret_cast<const float4*>(&As[rr][tx * 4]);
            const float g[4] = {g4.x, g4.y, g4.z, g4.w};
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(g[i], a[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* out = p.partial + (int64_t)blockIdx.y * p.N * p.Kaug;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= p.N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < p.Kaug) out[(int64_t)n * p.Kaug + k] = acc[i][j];
        }
    }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int S, int N, int K, int Kaug,
                                    float* __restrict__ dW, int64_t lddw, float* __restrict__ db, int accumulate) {
    pfo_pdl_prologue();
    const int total = N * Kaug;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float s = 0.0f;
        for (int y = 0; y < S; ++y) s += partial[(int64_t)y * total + i];   // fixed order: deterministic
        const int n = i / Kaug, k = i - n * Kaug;
        if (k < K) {
            float* d = dW + (int64_t)n * lddw + k;
            *d = accumulate ? *d + s : s;
        } else if (db) {
            db[n] = accumulate ? db[n] + s : s;
        }
    }
}

template <int TN>
int launch_linear(const LinearArgs& a, cudaStream_t s) {
    dim3 grid((unsigned)((a.M + BM - 1) / BM), (unsigned)((a.N + 16 * TN - 1) / (16 * TN)));
    pfo_launch(linear_kernel<TN>, grid, 256, 0, s, a);
    PFO_LAUNCH_CHECK();
}

}  // namespace

// also the landing path of pfo_linear_tf32 for operand layouts TMA cannot describe (linear_tma.cu)
int pfo_linear_f32_impl(const float* A, int64_t lda, const int32_t* a_idx,
                        const float* W, int64_t ldw, int w_transposed,
                        const float* bias, const float* bias_row_scale, int64_t ld_brs,
                        float* C, int64_t ldc, int64_t M, const int32_t* m_dev, int N, int K,
                        float alpha, int act, const int32_t* row_zero,
                        const float* relu_gate, int64_t ld_gate, int accumulate, void* stream) {
    if (M <= 0 || N <= 0) return 0;
    LinearArgs a{A, lda, a_idx, W, ldw, w_transposed, bias, bias_row_scale, ld_brs, C, ldc, M, m_dev, N, K,
                 alpha, act, row_zero, relu_gate, ld_gate, accumulate};
    cudaStream_t s = (cudaStream_t)stream;
    if (N <= 64) return launch_linear<4>(a, s);
    if (N <= 128) return launch_linear<8>(a, s);
    return launch_linear<12>(a, s);      // BN = 192; wider N loops over grid.y
}

PFO_API int pfo_linear_f32(const float* A, int64_t lda, const int32_t* a_idx,
                           const float* W, int64_t ldw, int w_transposed,
                           const float* bias, const float* bias_row_scale, int64_t ld_brs,
                           float* C, int64_t ldc, int64_t M, const int32_t* m_dev, int N, int K,
                           float alpha, int act, const int32_t* row_zero,
                           const float* relu_gate, int64_t ld_gate, int accumulate, void* stream) {
    return pfo_linear_f32_impl(A, lda, a_idx, W, ldw, w_transposed, bias, bias_row_scale, ld_brs, C, ldc, M, m_dev,
                               N, K, alpha, act, row_zero, relu_gate, ld_gate, accumulate, stream);
}

PFO_API int64_t pfo_wgrad_workspace_floats(int64_t M, int N, int K, int with_bias) {
    const int Kaug = K + (with_bias ? 1 : 0);
    const int tiles = ((N + WT - 1) / WT) * ((Kaug + WT - 1) / WT);
    int64_t S = (M + 255) / 256;
    int64_t cap = (4 * 148 + tiles - 1) / tiles;
    if (S > cap) S = cap;
    if (S < 1) S = 1;
    return S * N * Kaug;
}

// also the landing path of pfo_wgrad_tf32 for operand layouts TMA cannot describe (wgrad_tma.cu)
int pfo_wgrad_f32_impl(const float* G, int64_t ldg, const float* A, int64_t lda, const int32_t* a_idx,
                       int64_t M, const int32_t* m_dev, int N, int K,
                       float* dW, int64_t lddw, float* db, int accumulate, float* workspace, void* stream) {
    if (N <= 0 || K <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const int Kaug = K + (db ? 1 : 0);
    const int tiles = ((N + WT - 1) / WT) * ((Kaug + WT - 1) / WT);
    int64_t S = (M + 255) / 256;
    int64_t cap = (4 * 148 + tiles - 1) / tiles;
    if (S > cap) S = cap;
    if (S < 1) S = 1;
    WgradArgs a{G, ldg, A, lda, a_idx, workspace, M, m_dev, N, K, Kaug, (int)S};
    pfo_launch(wgrad_kernel, dim3(tiles, (unsigned)S), 256, 0, s, a);
    const int total = N * Kaug;
    pfo_launch(wgrad_reduce_kernel, (total + 255) / 256, 256, 0, s, workspace, (int)S, N, K, Kaug, dW, lddw, db, accumulate);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_wgrad_f32(const float* G, int64_t ldg, const float* A, int64_t lda, const int32_t* a_idx,
                          int64_t M, const int32_t* m_dev, int N, int K,
                          float* dW, int64_t lddw, float* db, int accumulate, float* workspace, void* stream) {
    return pfo_wgrad_f32_impl(G, ldg, A, lda, a_idx, M, m_dev, N, K, dW, lddw, db, accumulate, workspace, stream);
}
