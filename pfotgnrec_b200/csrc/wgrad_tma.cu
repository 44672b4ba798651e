// TMA-fed tcgen05 / TMEM weight-gradient contraction.
//
//     dW[n, k] = sum_m G[m, n] * A[m, k],     db[n] = sum_m G[m, n]
// (autograd of every nn.Linear / GRUCell weight on the reference's path: modules/memory_updater.py:60,
// model/temporal_attention.py:26-32, utils/utils.py:7-8).  The contraction runs over the huge
// row dimension m, so both operands are "MN-major" for the tensor core: a TMA box of R rows x 32
// fp32 columns (128-byte swizzle, 32-byte atoms) is exactly one canonical MN-major swizzle atom column, the MMA K
// step (8 for tf32) is two 4-row groups of the box, and successive 32-column boxes sit LBO bytes
// apart.  D[n (128 lanes), k (<= 256 columns)] accumulates in TMEM over the CTA's slab of rows;
// the bias gradient rides along as one extra B atom holding a column of ones.  Each CTA writes
// its partial [N, K+1] to the workspace and a second kernel reduces the slabs in a fixed order
// (deterministic, like pfo_wgrad_f32).
//
//   warp 0     TMA producer (G boxes + A boxes per stage, mbarrier full/empty ring)
//   warp 1     MMA issuer (one lane): passes = 1 plain TF32, passes = 3 error-compensated 3xTF32
//   warps 2-9  stage fix-up: zero the rows past the live row count (m_dev), split hi/lo in
//              3-pass mode; afterwards the same warps drain TMEM (epilogue)
#include "common.cuh"
#include <cuda.h>

int pfo_make_tensor_map_f32(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                            int atom32);
int pfo_wgrad_f32_impl(const float* G, int64_t ldg, const float* A, int64_t lda, const int32_t* a_idx,
                       int64_t M, const int32_t* m_dev, int N, int K, float* dW, int64_t lddw, float* db,
                       int accumulate, float* workspace, void* stream);

namespace {

constexpr int TN = 128;                  // output rows (n) per CTA = UMMA_M
constexpr int R = 32;                    // stream rows (m) per stage = 4 MMA K steps
constexpr int BOX_BYTES = R * 128;       // one R x 32 fp32 box
constexpr int G_BOXES = TN / 32;
constexpr int MAX_STAGES = 8;
constexpr int EPI_LD = 36;
constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;
constexpr int SMEM_LIMIT = 232448;
constexpr int FIX_WARPS = 8;                // stage fix-up warps (two per scheduler: one could not hide its own latency)
constexpr int WG_THREADS = 64 + 32 * FIX_WARPS;

struct WgTmaArgs {
    float* partial;            // [S, N, Kaug]
    int64_t M; const int32_t* m_dev; int N; int K; int Kaug;
    int S;                     // slabs = gridDim.x
    int nkb;                   // 32-column boxes of A
    int b_boxes;               // nkb + (bias ? 1 : 0)
    int NB;                    // MMA N = accumulator columns
    int stages;
    int tmem_cols;
    int fused;                 // reduce the slabs inside this launch (grid barrier) instead of a second kernel
    float* dW; int64_t lddw; float* db; int accumulate;
};

__device__ unsigned g_wgrad_sync[2];   // arrivals / departures of the fused reduction's grid barrier (self-resetting)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
        :: "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// MN-major tf32 operand: the tensor core wants the 128-byte swizzle with 32-byte atoms
// (descriptor layout SWIZZLE_128B_BASE32B = 1; TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 32 contiguous
// fp32 along M/N per row, 4-row groups along K (SBO = 512 B), successive 32-column atoms LBO bytes apart.
// Established on hardware with tools/umma_probe.cu (plain SWIZZLE_128B yields zeros for MN-major tf32).
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr, uint32_t lbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// one lane of a converged warp (see linear_tma.cu: the issue loops are walked by the whole warp so that their state
// stays in uniform registers; only the TMA / MMA instruction is predicated)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}

template <int PASSES>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmA, const WgTmaArgs p) {
    pfo_pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler (role dispatch)
    const int stages = p.stages, nkb = p.nkb, b_boxes = p.b_boxes;
    const int n0 = blockIdx.y * TN;
    const int boxes = G_BOXES + b_boxes;                       // per stage: [G x4][A x nkb][ones]
    const uint32_t stage_bytes = (uint32_t)boxes * BOX_BYTES;

    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
    uint8_t* sHi = smem_raw + pad;                             // [stages][boxes][R x 128 B]
    uint8_t* sLo = sHi + (size_t)stages * stage_bytes;         // twin ring with the residuals (3-pass)
    float* sEpi = reinterpret_cast<float*>(sLo + (PASSES == 3 ? (size_t)stages * stage_bytes : 0));
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sEpi) + EPI_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto ready_bar = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
    auto empty_bar = [&](int s) { return bar0 + 8u * (2 * MAX_STAGES + s); };
    const uint32_t done_bar = bar0 + 8u * (3 * MAX_STAGES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 1);

    int64_t M = p.M;
    if (p.m_dev) { int64_t md = *p.m_dev; M = md < M ? md : M; }
    // slab of this CTA: a whole number of R-row stages
    int64_t slab = (M + p.S - 1) / p.S;
    slab = (slab + R - 1) / R * R;
    const int64_t r_begin = (int64_t)blockIdx.x * slab;
    int64_t r_end = r_begin + slab;
    if (r_end > M) r_end = M;
    const int n_iters = r_end > r_begin ? (int)((r_end - r_begin + R - 1) / R) : 0;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(ready_bar(s), FIX_WARPS);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the bias atom: column 0 of the extra B box is 1 (hi ring) / 0 (lo ring); TMA never touches it
    if (b_boxes > nkb) {
        for (int i = tid; i < stages * (BOX_BYTES / 16); i += blockDim.x) {
            const int s = i / (BOX_BYTES / 16), u = i % (BOX_BYTES / 16);
            const int row = u >> 3, unit = u & 7;
            const bool first = unit == 2 * (row & 3);              // logical column 0: 32-byte chunk (row & 3)
            const size_t off = (size_t)s * stage_bytes + (size_t)(G_BOXES + nkb) * BOX_BYTES + (size_t)u * 16;
            *reinterpret_cast<float4*>(sHi + off) = make_float4(first ? 1.0f : 0.0f, 0.f, 0.f, 0.f);
            if (PASSES == 3) *reinterpret_cast<float4*>(sLo + off) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (n_iters > 0) {
            if (lane == 0) {
                asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmG)) : "memory");
                asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
            }
            __syncwarp();
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; it < n_iters; ++it) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx(full_bar(stage), (uint32_t)(G_BOXES + nkb) * BOX_BYTES);
                    const int row = (int)(r_begin + (int64_t)it * R);
                    const uint32_t dst = smem_u32(sHi + (size_t)stage * stage_bytes);
#pragma unroll
                    for (int b = 0; b < G_BOXES; ++b)
                        tma_load_2d(dst + (uint32_t)b * BOX_BYTES, &tmG, n0 + b * 32, row, full_bar(stage));
                    for (int b = 0; b < nkb; ++b)
                        tma_load_2d(dst + (uint32_t)(G_BOXES + b) * BOX_BYTES, &tmA, b * 32, row, full_bar(stage));
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (n_iters > 0) {
            // D = f32, A = B = tf32, both MN-major (bits 15, 16), N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(p.NB >> 3) << 17) | ((uint32_t)(TN >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; it < n_iters; ++it) {
                mbar_wait(ready_bar(stage), phase);
                tc_fence_after();
                const uint32_t hi = smem_u32(sHi + (size_t)stage * stage_bytes);
                const uint32_t lo = smem_u32(sLo + (size_t)stage * stage_bytes);
                if (elect_one()) {
#pragma unroll
                    for (int s = 0; s < R / 8; ++s) {
                        const uint32_t koff = (uint32_t)s * 1024u;
                        const uint64_t dg = desc_mn_sw128(hi + koff, BOX_BYTES);
                        const uint64_t da = desc_mn_sw128(hi + (uint32_t)G_BOXES * BOX_BYTES + koff, BOX_BYTES);
                        if (PASSES == 3) {
                            const uint64_t dgl = desc_mn_sw128(lo + koff, BOX_BYTES);
                            const uint64_t dal = desc_mn_sw128(lo + (uint32_t)G_BOXES * BOX_BYTES + koff, BOX_BYTES);
                            tc_mma_tf32(tmem_base, dgl, da, idesc, (it | s) ? 1u : 0u);
                            tc_mma_tf32(tmem_base, dg, dal, idesc, 1u);
                            tc_mma_tf32(tmem_base, dg, da, idesc, 1u);
                        } else {
                            tc_mma_tf32(tmem_base, dg, da, idesc, (it | s) ? 1u : 0u);
                        }
                    }
                    tc_commit(empty_bar(stage));
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
            if (elect_one()) tc_commit(done_bar);
        }
        __syncwarp();
    } else {
        // ===== stage fix-up (mask the tail rows, split hi / lo), then the epilogue
        const int t = tid - 64;                                    // 0 .. 32 * FIX_WARPS - 1
        const int units = (G_BOXES + nkb) * (BOX_BYTES / 16);     // 16-byte units TMA wrote per stage
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < n_iters; ++it) {
            mbar_wait(full_bar(stage), phase);
            const int64_t row0 = r_begin + (int64_t)it * R;
            const int live = (r_end - row0) < R ? (int)(r_end - row0) : R;   // rows of this stage inside the slab
            if (PASSES == 3 || live < R) {
                float4* hi = reinterpret_cast<float4*>(sHi + (size_t)stage * stage_bytes);
                float4* lo = reinterpret_cast<float4*>(sLo + (size_t)stage * stage_bytes);
#pragma unroll 4
                for (int u = t; u < units; u += 32 * FIX_WARPS) {
                    const int row = (u % (BOX_BYTES / 16)) >> 3;
                    float4 v = hi[u];
                    if (row >= live) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (PASSES == 3) {
                        float4 h, l;
                        h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
                        l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                        hi[u] = h;
                        lo[u] = l;
                    } else {
                        hi[u] = v;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(ready_bar(stage));
            if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        // ---- epilogue: partial[slab][n][k]; warp q owns TMEM lanes [32q, 32q+32) = n rows (warps 2-5 only)
        if (warp < 6) {
        const int q = warp & 3;
        float* buf = sEpi + (size_t)(warp - 2) * 32 * EPI_LD;
        const int r_sub = lane >> 3, c4 = (lane & 7) * 4;
        float* out = p.partial + (int64_t)blockIdx.x * p.N * p.Kaug;
        if (n_iters > 0) {
            mbar_wait(done_bar, 0);
            tc_fence_after();
        }
        const int bias_col = nkb * 32;
        for (int c0 = 0; c0 < p.NB; c0 += 32) {
            const int ncols = (p.NB - c0) < 32 ? (p.NB - c0) : 32;
            uint32_t r[32];
            if (n_iters > 0) {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
                tmem_ld16(taddr, r);
                if (ncols > 16) tmem_ld16(taddr + 16u, r + 16);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0u;            // empty slab: its partial is zero
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j * 4 < ncols)
                    *reinterpret_cast<float4*>(buf + lane * EPI_LD + j * 4) =
                        make_float4(__uint_as_float(r[j * 4]), __uint_as_float(r[j * 4 + 1]),
                                    __uint_as_float(r[j * 4 + 2]), __uint_as_float(r[j * 4 + 3]));
            }
            __syncwarp();
            if (c4 < ncols) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = i * 4 + r_sub;
                    const int n = n0 + q * 32 + row;
                    if (n >= p.N) continue;
                    const float4 a4 = *reinterpret_cast<const float4*>(buf + row * EPI_LD + c4);
                    const float v[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int col = c0 + c4 + e;
                        if (col < p.K) out[(int64_t)n * p.Kaug + col] = v[e];
                        else if (col == bias_col && p.Kaug > p.K) out[(int64_t)n * p.Kaug + p.K] = v[e];
                    }
                }
            }
            __syncwarp();
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     :: "r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
    if (!p.fused) return;
    // ---- fused slab reduction (replaces the second launch).  Grid barrier: the grid is at most one CTA per SM and a
    // CTA holds an SM's shared memory alone, so once the grids ahead in the stream have drained every CTA of this one
    // is resident (a dependent grid launched programmatically cannot start before all of these CTAs have -- they
    // trigger it in their first instruction -- and it cannot take their place, it waits for this grid to complete).
    // Then every CTA sums its share of the N x Kaug outputs over the S slabs in slab order (deterministic).  The two
    // counters reset themselves: the last CTA to leave zeroes them for the next launch.  One instance at a time per
    // device: launches are serialised on the caller's stream.
    const unsigned n_ctas = gridDim.x * gridDim.y;
    if (tid == 0) {
        __threadfence();                                           // this CTA's partial is visible device-wide
        atomicAdd(&g_wgrad_sync[0], 1u);
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&g_wgrad_sync[0]) : "memory");
            if (seen < n_ctas) __nanosleep(32);
        } while (seen < n_ctas);
    }
    __syncthreads();
    {
        const int total = p.N * p.Kaug;
        const int cta = (int)(blockIdx.y * gridDim.x + blockIdx.x);
        const int stride = (int)n_ctas * WG_THREADS;
        for (int i = cta * WG_THREADS + tid; i < total; i += stride) {
            const float* src = p.partial + i;
            float s0 = 0.0f;
            int y = 0;
            for (; y + 4 <= p.S; y += 4) {                         // four loads in flight, one running sum (slab order)
                const float a0 = __ldcg(src + (int64_t)y * total), a1 = __ldcg(src + (int64_t)(y + 1) * total);
                const float a2 = __ldcg(src + (int64_t)(y + 2) * total), a3 = __ldcg(src + (int64_t)(y + 3) * total);
                s0 = (((s0 + a0) + a1) + a2) + a3;
            }
            for (; y < p.S; ++y) s0 += __ldcg(src + (int64_t)y * total);
            const int n = i / p.Kaug, k = i - n * p.Kaug;
            if (k < p.K) {
                float* d = p.dW + (int64_t)n * p.lddw + k;
                *d = p.accumulate ? *d + s0 : s0;
            } else if (p.db) {
                p.db[n] = p.accumulate ? p.db[n] + s0 : s0;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned left = atomicAdd(&g_wgrad_sync[1], 1u);
        if (left == n_ctas - 1) {                                  // everyone is past the spin: safe to reset
            g_wgrad_sync[1] = 0u;
            __threadfence();
            g_wgrad_sync[0] = 0u;
        }
    }
}

// dW (+)= sum over the slabs, in a fixed order (deterministic).  A CTA owns 32 consecutive outputs; its 8 warps
// stride the slabs (8 independent chains of coalesced 128-byte reads), then the 8 sums are added in warp order.
__global__ void __launch_bounds__(256)
wgrad_tma_reduce_kernel(const float* __restrict__ partial, int S, int N, int K, int Kaug,
                        float* __restrict__ dW, int64_t lddw, float* __restrict__ db, int accumulate) {
    pfo_pdl_prologue();
    __shared__ float sm[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int total = N * Kaug;
    const int i = blockIdx.x * 32 + lane;
    float s = 0.0f;
    if (i < total) {
        // four independent chains per warp (fixed association, still deterministic): the loads of a single running sum
        // would queue behind each other's latency
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
        int y = w;
        for (; y + 24 < S; y += 32) {
            s0 += partial[(int64_t)y * total + i];
            s1 += partial[(int64_t)(y + 8) * total + i];
            s2 += partial[(int64_t)(y + 16) * total + i];
            s3 += partial[(int64_t)(y + 24) * total + i];
        }
        for (; y < S; y += 8) s0 += partial[(int64_t)y * total + i];
        s = (s0 + s1) + (s2 + s3);
    }
    sm[w][lane] = s;
    __syncthreads();
    if (w == 0 && i < total) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sm[k][lane];
        const int n = i / Kaug, k = i - n * Kaug;
        if (k < K) {
            float* d = dW + (int64_t)n * lddw + k;
            *d = accumulate ? *d + t : t;
        } else if (db) {
            db[n] = accumulate ? db[n] + t : t;
        }
    }
}

int slabs_for(int64_t M, int n_ntiles) {
    int64_t S = (M + 8 * R - 1) / (8 * R);                        // at least 8 stages of work per slab
    int64_t cap = pfo_num_sms() / n_ntiles;
    if (cap < 1) cap = 1;
    if (S > cap) S = cap;
    if (S < 1) S = 1;
    return (int)S;
}

// PFO_WGRAD_FUSED=1 reduces the slabs inside the contraction launch.  Measured SLOWER than the second launch at the
// bench shapes (step 0.872 ms against 0.845 ms, 220 us against 190 us in the five weight gradients of a step,
// profiles/r2_bench_wgrad_fused.json): every CTA idles at the barrier until the slowest slab is done and the reduction
// then runs on 148 CTAs x 320 threads instead of up to 1 164 CTAs.  Parity-green (the GPU suite ran with it), kept as
// an opt-in.
bool wgrad_fused_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("PFO_WGRAD_FUSED");
        on = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return on != 0;
}

template <int PASSES>
int launch_wg(const CUtensorMap& mg, const CUtensorMap& ma, const WgTmaArgs& a, dim3 grid, size_t smem, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tma_kernel<PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    pfo_launch(wgrad_tma_kernel<PASSES>, grid, WG_THREADS, smem, s, mg, ma, a);
    return (int)cudaGetLastError();
}

}  // namespace

PFO_API int64_t pfo_wgrad_tf32_workspace_floats(int64_t M, int N, int K, int with_bias) {
    const int Kaug = K + (with_bias ? 1 : 0);
    const int n_ntiles = (N + TN - 1) / TN;
    int64_t a = (int64_t)slabs_for(M, n_ntiles) * N * Kaug;
    int64_t b = pfo_wgrad_workspace_floats(M, N, K, with_bias);    // the FFMA landing path shares the buffer
    return a > b ? a : b;
}

PFO_API int pfo_wgrad_tf32(const float* G, int64_t ldg, const float* A, int64_t lda, const int32_t* a_idx,
                           int64_t M, const int32_t* m_dev, int N, int K, float* dW, int64_t lddw, float* db,
                           int accumulate, float* workspace, int passes, void* stream) {
    if (N <= 0 || K <= 0) return 0;
    if (passes != 1 && passes != 3) return (int)cudaErrorInvalidValue;
    const int nkb = (K + 31) / 32;
    const int b_boxes = nkb + (db ? 1 : 0);
    const int NB = nkb * 32 + (db ? 16 : 0);
    CUtensorMap mg, ma;
    const bool tma_ok = M > 0 && a_idx == nullptr && NB <= 256 && (ldg % 4 == 0) && (lda % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(G) & 15) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                        pfo_make_tensor_map_f32(&mg, G, M, N, ldg, R, 1) == 0 &&
                        pfo_make_tensor_map_f32(&ma, A, M, K, lda, R, 1) == 0;
    if (!tma_ok)
        return pfo_wgrad_f32_impl(G, ldg, A, lda, a_idx, M, m_dev, N, K, dW, lddw, db, accumulate, workspace, stream);
    const int mult = passes == 3 ? 2 : 1;
    const int stage_bytes = (G_BOXES + b_boxes) * BOX_BYTES * mult;
    const int fixed = EPI_BYTES + 8 * (3 * MAX_STAGES + 2) + 16 + 1024;
    int stages = (SMEM_LIMIT - fixed) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return pfo_wgrad_f32_impl(G, ldg, A, lda, a_idx, M, m_dev, N, K, dW, lddw, db, accumulate, workspace, stream);
    const int n_ntiles = (N + TN - 1) / TN;
    WgTmaArgs a{};
    a.partial = workspace; a.M = M; a.m_dev = m_dev; a.N = N; a.K = K; a.Kaug = K + (db ? 1 : 0);
    a.S = slabs_for(M, n_ntiles); a.nkb = nkb; a.b_boxes = b_boxes; a.NB = NB; a.stages = stages;
    int cols = 32;
    while (cols < NB) cols <<= 1;
    a.tmem_cols = cols;
    const size_t smem = (size_t)stages * stage_bytes + fixed;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((unsigned)a.S, (unsigned)n_ntiles);
    a.fused = wgrad_fused_enabled() && (int)(grid.x * grid.y) <= pfo_num_sms() ? 1 : 0;
    a.dW = dW; a.lddw = lddw; a.db = db; a.accumulate = accumulate;
    int rc = passes == 3 ? launch_wg<3>(mg, ma, a, grid, smem, s) : launch_wg<1>(mg, ma, a, grid, smem, s);
    if (rc || a.fused) return rc;
    const int total = N * a.Kaug;
    pfo_launch(wgrad_tma_reduce_kernel, (total + 31) / 32, 256, 0, s, workspace, a.S, N, K, a.Kaug, dW, lddw, db, accumulate);
    PFO_LAUNCH_CHECK();
}
