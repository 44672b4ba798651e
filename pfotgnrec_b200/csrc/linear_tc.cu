// bf16 tcgen05 / TMEM path of the tall-skinny contractions (the 2e-2 "bf16 GEMM mode").
//
// Same contract as pfo_linear_f32 (see linear_simt.cu): C[m,:N] = epi(alpha*(A[row(m),:K].W^T + bias)).
// Operands live in HBM as fp32 (possibly gathered rows), so they are converted to bf16 while
// being staged into shared memory in the UMMA canonical K-major, no-swizzle layout (8x8 core
// matrices of 128 B); one elected thread issues tcgen05.mma (M=128, N=ceil16(N), K=16 per
// instruction, fp32 accumulate in TMEM), completion is signalled through tcgen05.commit on an
// mbarrier, and the four warps read their 32 TMEM lanes back with tcgen05.ld for the fused
// epilogue.  The weight matrix is staged once per CTA and stays resident while the CTA walks
// its row tiles (persistent grid).
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int TM = 128;              // rows per tile = UMMA_M (cta_group::1)
constexpr int THREADS = 128;

struct TcArgs {
    const float* A; int64_t lda; const int32_t* a_idx;
    const float* W; int64_t ldw; int w_transposed;
    const float* bias; const float* bias_row_scale; int64_t ld_brs;
    float* C; int64_t ldc;
    int64_t M; const int32_t* m_dev; int N; int K;
    float alpha; int act; const int32_t* row_zero; const float* relu_gate; int64_t ld_gate; int accumulate;
    int NP; int KP; int tmem_cols;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ uint4 pack8(const float* v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]);
    __nv_bfloat162 d = __floats2bfloat162_rn(v[6], v[7]);
    uint4 r;
    r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
    r.z = *reinterpret_cast<uint32_t*>(&c); r.w = *reinterpret_cast<uint32_t*>(&d);
    return r;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
linear_tc_kernel(const TcArgs p) {
    pfo_pdl_prologue();
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KP = p.KP, NP = p.NP, K = p.K, N = p.N;
    uint8_t* sA = smem;                                   // [KP/8][16 row groups][8 rows][16 B]
    uint8_t* sW = smem + (size_t)TM * KP * 2;             // [KP/8][NP/8][8][16 B]
    int64_t M = p.M;
    if (p.m_dev) { int64_t md = *p.m_dev; M = md < M ? md : M; }
    const int64_t n_tiles = (M + TM - 1) / TM;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&tmem_base_s)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- stage W once: fp32 -> bf16, canonical K-major core-matrix layout
    const int kchunks = KP >> 3;
    for (int i = tid; i < NP * kchunks; i += THREADS) {
        int n, k8;
        if (p.w_transposed) { n = i % NP; k8 = i / NP; } else { k8 = i % kchunks; n = i / kchunks; }
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k8 * 8 + j;
            float x = 0.0f;
            if (n < N && k < K) x = p.w_transposed ? __ldg(p.W + (int64_t)k * p.ldw + n) : __ldg(p.W + (int64_t)n * p.ldw + k);
            v[j] = x;
        }
        *reinterpret_cast<uint4*>(sW + (size_t)k8 * (NP * 16) + (size_t)(n >> 3) * 128 + (n & 7) * 16) = pack8(v);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    const bool vec_ok = (p.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
    uint32_t phase = 0;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ---- stage the A tile: thread r owns row r of the tile
        const int64_t m = tile * TM + tid;
        const float* rowp = nullptr;
        if (m < M) {
            int64_t row = m;
            if (p.a_idx) row = p.a_idx[m];
            if (row >= 0) rowp = p.A + row * p.lda;
        }
        for (int k8 = 0; k8 < kchunks; ++k8) {
            float v[8];
            const int k0 = k8 * 8;
            if (rowp && vec_ok && k0 + 8 <= K) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(rowp + k0));
                const float4 b = __ldg(reinterpret_cast<const float4*>(rowp + k0 + 4));
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = (rowp && k0 + j < K) ? __ldg(rowp + k0 + j) : 0.0f;
            }
            *reinterpret_cast<uint4*>(sA + (size_t)k8 * 2048 + (size_t)tid * 16) = pack8(v);
        }
        // generic-proxy writes -> visible to the async (tensor-core) proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_base = smem_u32(sA), w_base = smem_u32(sW);
            for (int k16 = 0; k16 < (KP >> 4); ++k16) {
                const uint64_t da = make_desc(a_base + (uint32_t)k16 * 2 * 2048, 2048, 128);
                const uint64_t db = make_desc(w_base + (uint32_t)k16 * 2 * (NP * 16), (uint32_t)NP * 16, 128);
                const uint32_t acc = k16 > 0 ? 1u : 0u;
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                    :: "r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                         :: "r"(smem_u32(&mbar)) : "memory");
        }
        mbar_wait(smem_u32(&mbar), phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue: warp w owns TMEM lanes [32w, 32w+32) = rows of the tile
        const int64_t mo = tile * TM + warp * 32 + lane;
        const bool row_ok = mo < M;
        const bool zero_row = row_ok && p.row_zero && p.row_zero[mo] != 0;
        const float brs = (row_ok && p.bias_row_scale) ? p.bias_row_scale[mo * p.ld_brs] : 1.0f;
        for (int c0 = 0; c0 < NP; c0 += 16) {
            uint32_t r[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = c0 + j;
                    if (n >= N) break;
                    float v = __uint_as_float(r[j]);
                    if (p.bias) v += p.bias[n] * brs;
                    v *= p.alpha;
                    if (p.act == 1) v = fmaxf(v, 0.0f);
                    if (p.relu_gate && p.relu_gate[mo * p.ld_gate + n] <= 0.0f) v = 0.0f;
                    if (zero_row) v = 0.0f;
                    float* dst = p.C + mo * p.ldc + n;
                    if (p.accumulate) v += *dst;
                    *dst = v;
                }
            }
        }
        // TMEM and the A buffer are reused by the next tile
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     :: "r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

}  // namespace

PFO_API int pfo_linear_bf16(const float* A, int64_t lda, const int32_t* a_idx, const float* W, int64_t ldw,
                            int w_transposed, const float* bias, const float* bias_row_scale, int64_t ld_brs,
                            float* C, int64_t ldc, int64_t M, const int32_t* m_dev, int N, int K,
                            float alpha, int act, const int32_t* row_zero, const float* relu_gate, int64_t ld_gate,
                            int accumulate, void* stream) {
    if (M <= 0 || N <= 0) return 0;
    if (K > 512) return (int)cudaErrorInvalidValue;
    if (N > 256) {          // one accumulator holds 256 columns: wider outputs go panel by panel
        const int panels = (N + 255) / 256;
        const int np = ((N + panels - 1) / panels + 15) / 16 * 16;
        for (int n0 = 0; n0 < N; n0 += np) {
            const int nn = N - n0 < np ? N - n0 : np;
            int rc = pfo_linear_bf16(A, lda, a_idx, w_transposed ? W + n0 : W + (int64_t)n0 * ldw, ldw, w_transposed,
                                     bias ? bias + n0 : nullptr, bias_row_scale, ld_brs, C + n0, ldc, M, m_dev, nn, K,
                                     alpha, act, row_zero, relu_gate ? relu_gate + n0 : nullptr, ld_gate, accumulate,
                                     stream);
            if (rc) return rc;
        }
        return 0;
    }
    TcArgs a{A, lda, a_idx, W, ldw, w_transposed, bias, bias_row_scale, ld_brs, C, ldc, M, m_dev, N, K,
             alpha, act, row_zero, relu_gate, ld_gate, accumulate, 0, 0, 0};
    a.NP = (N + 15) / 16 * 16;
    a.KP = (K + 15) / 16 * 16;
    int cols = 32;
    while (cols < a.NP) cols <<= 1;
    a.tmem_cols = cols;
    const size_t smem = (size_t)TM * a.KP * 2 + (size_t)a.NP * a.KP * 2;
    if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    const int64_t tiles = (M + TM - 1) / TM;
    // resident CTAs per SM are bounded by shared memory and by TMEM columns (512 per SM)
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > 512 / cols) per_sm = 512 / cols;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int64_t grid = (int64_t)pfo_num_sms() * per_sm;
    if (grid > tiles) grid = tiles;
    pfo_launch(linear_tc_kernel, (unsigned)grid, THREADS, smem, (cudaStream_t)stream, a);
    PFO_LAUNCH_CHECK();
}
