// TMA-fed tcgen05 / TMEM path of the tall-skinny contractions (forward and dgrad).
//
// Same contract as pfo_linear_f32 (linear_simt.cu):
//     C[m, :N] = epi(alpha * (A[m, :K] . W^T + bias * brs[m]))
// for the GRU/RNN cell (reference modules/memory_updater.py:31,47), the attention projections
// (model/temporal_attention.py:70) and the merge MLP (utils/utils.py:14-17): M = unique nodes or
// queries (10^4..10^7 rows), N, K <= 320.  The activations live in HBM as fp32, so the tensor
// cores run kind::tf32 straight on what TMA delivers -- no conversion pass:
//
//   warp 0     TMA producer: one 128-row x 32-column fp32 box (16 KiB, SWIZZLE_128B) per K chunk
//              into a ring of shared-memory stages (mbarrier full/empty handshake)
//   warp 1     MMA issuer: the whole warp walks the (tile, chunk) loops, one elected lane issues tcgen05.mma
//              (M=128, N=NT, K=8 per instruction), accumulating in TMEM; tcgen05.commit releases the stage /
//              publishes the accumulator
//   warps 2-9  epilogue: tcgen05.ld of their 32 TMEM lanes (lane = output row), bias / scale / relu / row-mask in
//              registers, a swizzled staging tile per 16-column slab, then coalesced 16-byte stores (relu gate
//              and accumulate applied there); two warps per lane quarter so the schedulers can hide latency
//   warps 10-17 (3-pass mode only) operand splitters, see below
//
// passes = 1: plain TF32 (10-bit mantissa), the "fast" mode (2e-2 contract, in practice ~1e-3).
// passes = 3: error-compensated 3xTF32 for the 1e-5 contract: a = a_hi + a_lo with a_hi =
//   rna_tf32(a); D = A_lo.W_hi + A_hi.W_lo + A_hi.W_hi; the dropped terms are ~2^-22 relative.
//   The splitter warps (one thread per tile row) rewrite each landed stage in place (hi) and put the residual
//   into the stage's TMEM slot: the A_lo.W_hi MMA takes its A operand from tensor memory, so the ring needs no
//   shared-memory twin.  TMEM columns: [2 accumulators x NT][stages x 32 A-lo columns] <= 512.
// The weight slice W[n0:n0+NT, :K] is staged once per CTA (canonical no-swizzle K-major core
// matrices, split hi/lo in 3-pass mode) and stays resident while the CTA walks its row tiles;
// two TMEM accumulators let the epilogue of tile i overlap the MMAs of tile i+1.
#include "common.cuh"
#include <cuda.h>

int pfo_linear_f32_impl(const float* A, int64_t lda, const int32_t* a_idx, const float* W, int64_t ldw,
                        int w_transposed, const float* bias, const float* bias_row_scale, int64_t ld_brs,
                        float* C, int64_t ldc, int64_t M, const int32_t* m_dev, int N, int K,
                        float alpha, int act, const int32_t* row_zero, const float* relu_gate, int64_t ld_gate,
                        int accumulate, void* stream);

namespace {

constexpr int TM = 128;                      // rows per tile = UMMA_M (cta_group::1)
constexpr int CK = 32;                       // fp32 columns per K chunk = one 128-byte swizzle row
constexpr int CHUNK_BYTES = TM * CK * 4;     // 16 KiB
constexpr int MAX_STAGES = 8;
constexpr int EPI_BYTES = 8 * 32 * 64;        // per epilogue warp: one 32-row x 64-byte swizzled staging tile
constexpr int SMEM_LIMIT = 232448;           // 227 KiB opt-in maximum per CTA

struct TmaLinArgs {
    const float* W; int64_t ldw; int w_transposed;
    const float* bias; const float* bias_row_scale; int64_t ld_brs;
    float* C; int64_t ldc;
    int64_t M; const int32_t* m_dev; int N; int K;
    float alpha; int act; const int32_t* row_zero; const float* relu_gate; int64_t ld_gate; int accumulate;
    int NT;            // output columns per CTA (multiple of 16, <= 256)
    int KP8;           // K rounded up to the MMA K step (8)
    int n_chunks;      // ceil(K / 32)
    int stages;
    int tmem_cols;     // allocated TMEM columns (power of two >= 2 * NT [+ 32 * stages in 3-pass mode])
    int acc_stride;    // TMEM columns between the two accumulators
    int lo_col;        // 3-pass mode: first TMEM column of the A-lo slots (32 columns per stage)
    int vec_ok;        // C / relu_gate rows are 16-byte aligned
    int dbg;           // timing experiments only (wrong numerics): 1 = splitters skip their work, 2 = one MMA per K step
    long long* trace;  // diagnostic timeline of CTA (0,0) (tools/linear_trace.py); null in normal use
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// the clock read carries a memory clobber: without it the compiler hoists the read above barriers (observed: the
// "total" stamp after the teardown __syncthreads was taken before the epilogue had finished)
__device__ __forceinline__ long long pfo_clock() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory");
    return t;
}
#define PFO_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[(slot)] = pfo_clock(); } while (0)

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
// A operand from tensor memory (lane = row, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
        :: "r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
// K-major operand in the 128-byte-swizzled layout TMA writes (8-row x 128 B atoms, SBO = 1024 B)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// K-major operand in the canonical no-swizzle layout: 8-row x 16-byte core matrices,
// LBO = distance between core matrices along K, SBO = along M/N
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
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

// One lane of a converged warp.  The TMA / MMA issue loops are executed by the WHOLE warp and only the instruction
// itself is predicated on the elected lane: loop counters and descriptors then stay warp-uniform for the compiler
// (uniform registers feed UTMALDG / UTCHMMA directly).  Under `if (lane == 0)` every operand lived in a vector
// register and each tcgen05.mma cost an ELECT + 4x R2UR.BROADCAST waterfall loop (~110 cycles per MMA, measured).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}

constexpr int EPI_WARPS = 8;                         // two per TMEM lane quarter
constexpr int SPLIT_WARPS = 8;                       // operand splitters of the 3-pass mode
constexpr int THREADS_1 = 32 * (2 + EPI_WARPS), THREADS_3 = THREADS_1 + 32 * SPLIT_WARPS;

// A-operand producer state: walks (tile, chunk) in order; shared by the pre-issue before the weight staging
// and the steady-state loop after it
struct Producer {
    int stage; uint32_t phase; int64_t tile; int c;
};

template <int PASSES>
__global__ void __launch_bounds__(PASSES == 3 ? THREADS_3 : THREADS_1, 1)
linear_tma_kernel(const __grid_constant__ CUtensorMap tmA, const TmaLinArgs p) {
    pfo_pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler (role dispatch)
    const int NT = p.NT, KP8 = p.KP8, stages = p.stages, n_chunks = p.n_chunks;
    const int n0 = blockIdx.y * NT;
    if (tid == 0) PFO_TRACE(0);
    if (p.trace && tid == 0) {                                     // diagnostic: wall-clock window of every CTA
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[128 + 2 * (blockIdx.y * gridDim.x + blockIdx.x)] = (long long)t;
    }

    // ---- carve shared memory (stages and the store staging need 1024-byte alignment for the 128-byte swizzle)
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
    uint8_t* base = smem_raw + pad;
    uint8_t* sA = base;                                            // [stages][16 KiB]  A (hi in 3-pass mode; lo lives in TMEM)
    uint8_t* sOut = sA + (size_t)stages * CHUNK_BYTES;             // [EPI_WARPS][32 rows x 64 B]
    uint8_t* sW = sOut + EPI_BYTES;
    const uint32_t w_bytes = (uint32_t)(KP8 >> 2) * ((uint32_t)NT * 16u + 16u);   // [KP8/4][NT/8 core matrices (8 rows x 16 B) + 16 B pad]
    uint8_t* sWlo = sW + w_bytes;
    float* sBias = reinterpret_cast<float*>(sWlo + (PASSES == 3 ? w_bytes : 0));   // [NT]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + NT);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto ready_bar = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
    auto empty_bar = [&](int s) { return bar0 + 8u * (2 * MAX_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (3 * MAX_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (3 * MAX_STAGES + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 4);

    int64_t M = p.M;
    if (p.m_dev) { int64_t md = *p.m_dev; M = md < M ? md : M; }
    const int64_t n_tiles = (M + TM - 1) / TM;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(ready_bar(s), SPLIT_WARPS);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();                                               // barriers initialised
    if (tid == 64) PFO_TRACE(3);

    // ---- the activation stream does not depend on the weights: fill the ring before staging them
    Producer pr{0, 0u, (int64_t)blockIdx.x, 0};
    auto produce = [&](int max_chunks) {                           // executed by the whole warp 0
        int issued = 0;
        while (pr.tile < n_tiles && issued < max_chunks) {
            mbar_wait(empty_bar(pr.stage), pr.phase ^ 1u);
            if (elect_one()) {
                mbar_arrive_expect_tx(full_bar(pr.stage), CHUNK_BYTES);
                tma_load_2d(smem_u32(sA + (size_t)pr.stage * CHUNK_BYTES), &tmA, pr.c * CK, (int)(pr.tile * TM), full_bar(pr.stage));
            }
            ++issued;
            if (++pr.stage == stages) { pr.stage = 0; pr.phase ^= 1u; }
            if (++pr.c == n_chunks) { pr.c = 0; pr.tile += gridDim.x; }
        }
        __syncwarp();
    };
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
            PFO_TRACE(8);
        }
        __syncwarp();
        produce(stages);
    }
    // ---- stage the weight slice once: rna(w) (and the residual) into core-matrix order (8 rows x 16 bytes).
    // The warp's lanes run along the CONTIGUOUS axis of W in global memory -- K for row-major weights, N for
    // transposed ones -- so a load instruction touches 4 lines, not 32 (the uncoalesced version spent 2-6 us of a
    // 10-40 us launch in LSU replays).  Core matrices of consecutive K units are w_lbo = NT * 16 + 16 bytes apart:
    // the 16-byte pad walks the banks, so the K-major store pattern is conflict-free too.
    {
        const int kcs = KP8 >> 2;                                  // 16-byte units along K
        const bool vec = !p.w_transposed && (p.ldw % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.W) & 15) == 0);
        const int units = NT * kcs;
        const uint32_t lbo = (uint32_t)NT * 16u + 16u;
        constexpr int U = 4;                                       // units in flight per thread
        for (int i0 = tid; i0 < units; i0 += U * blockDim.x) {
            float w[U][4];
            int nn[U], kk[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * blockDim.x;
                if (p.w_transposed) { nn[u] = i % NT; kk[u] = i / NT; } else { kk[u] = i % kcs; nn[u] = i / kcs; }
                const int ng = n0 + nn[u], kc = kk[u];
                w[u][0] = w[u][1] = w[u][2] = w[u][3] = 0.f;
                if (i < units && ng < p.N) {
                    if (vec && kc * 4 + 3 < p.K) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(p.W + (int64_t)ng * p.ldw + kc * 4));
                        w[u][0] = t.x; w[u][1] = t.y; w[u][2] = t.z; w[u][3] = t.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int k = kc * 4 + j;
                            if (k < p.K)
                                w[u][j] = p.w_transposed ? __ldg(p.W + (int64_t)k * p.ldw + ng) : __ldg(p.W + (int64_t)ng * p.ldw + k);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (i0 + u * (int)blockDim.x >= units) break;
                float hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { hi[j] = rna_tf32(w[u][j]); lo[j] = w[u][j] - hi[j]; }
                const size_t off = (size_t)kk[u] * lbo + (size_t)(nn[u] >> 3) * 128 + (size_t)(nn[u] & 7) * 16;
                *reinterpret_cast<float4*>(sW + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                if (PASSES == 3) *reinterpret_cast<float4*>(sWlo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        if (tid == 64) PFO_TRACE(4);
        for (int n = tid; n < NT; n += blockDim.x) sBias[n] = (p.bias && n0 + n < p.N) ? __ldg(p.bias + n0 + n) : 0.0f;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc_stride = (uint32_t)p.acc_stride;
    if (tid == 0) PFO_TRACE(1);
    int tr_tile = 0;

    if (warp == 0) {
        // ===== TMA producer (steady state) =====
        produce(0x7fffffff);
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp walks the loops, the elected lane issues) =====
        {
            // D = f32, A = B = tf32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            const uint32_t a_base = smem_u32(sA), alo_tmem = tmem_base + (uint32_t)p.lo_col;
            const uint32_t w_lbo = (uint32_t)NT * 16u + 16u;
            const uint64_t da0 = desc_sw128(a_base);               // + byte offset >> 4 in the low word
            const uint64_t dw0 = desc_nosw(smem_u32(sW), w_lbo, 128u);
            const uint64_t dwlo0 = desc_nosw(smem_u32(sWlo), w_lbo, 128u);
            const bool one_mma = PASSES == 3 && (p.dbg & 2);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_stride;
                for (int c = 0; c < n_chunks; ++c) {
                    mbar_wait(PASSES == 3 ? ready_bar(stage) : full_bar(stage), phase);
                    tc_fence_after();
                    if (c == 0 && lane == 0 && tr_tile < 8) PFO_TRACE(8 + 8 * tr_tile + 2);
                    int ksteps = (KP8 - c * CK) >> 3;
                    if (ksteps > 4) ksteps = 4;
                    const uint32_t a_off = ((uint32_t)stage * CHUNK_BYTES) >> 4;
                    const uint32_t w_off = ((uint32_t)(c * 4) * 2u * w_lbo) >> 4;
                    const uint32_t w_step = (2u * w_lbo) >> 4;
                    if (elect_one()) {
#pragma unroll 4
                        for (int s = 0; s < ksteps; ++s) {
                            const uint64_t da = da0 + (uint64_t)(a_off + (uint32_t)s * 2u);
                            const uint64_t dw = dw0 + (uint64_t)(w_off + (uint32_t)s * w_step);
                            const uint32_t first = (c | s) ? 1u : 0u;
                            if (PASSES == 3 && !one_mma) {
                                const uint64_t dwlo = dwlo0 + (uint64_t)(w_off + (uint32_t)s * w_step);
                                // A_lo . W_hi: A from its TMEM slot (lane = row, columns = the chunk's 32 K elements)
                                tc_mma_tf32_ts(d_tmem, alo_tmem + (uint32_t)stage * 32u + (uint32_t)s * 8u, dw, idesc, first);
                                tc_mma_tf32(d_tmem, da, dwlo, idesc, 1u);      // small terms first
                                tc_mma_tf32(d_tmem, da, dw, idesc, 1u);
                            } else {
                                tc_mma_tf32(d_tmem, da, dw, idesc, first);
                            }
                        }
                        tc_commit(empty_bar(stage));               // stage free once these MMAs retire
                    }
                    __syncwarp();
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
                if (elect_one()) tc_commit(tfull_bar(acc));        // accumulator complete
                __syncwarp();
                if (lane == 0 && tr_tile < 8) PFO_TRACE(8 + 8 * tr_tile + 3);
                ++tr_tile;
                acc ^= 1; if (acc == 0) acc_phase ^= 1u;
            }
        }
        __syncwarp();
    } else if (warp < 2 + EPI_WARPS) {
        // ===== epilogue: EPI_WARPS warps; warp owns TMEM lanes [32q, 32q+32) = 32 rows of the tile and every
        // other 16-column slab.  A slab goes registers (lane = row: bias, scale, relu, row mask) -> 64-byte-swizzled
        // staging tile -> coalesced 16-byte global stores (4 lanes per row; relu gate and accumulate applied
        // there, where their loads coalesce too).  Two things this shape fixes, both measured (ncu + clock64
        // trace): the first version was instruction-fetch bound (unrolled 45 KB of SASS), the second ran one
        // epilogue warp per scheduler and could not hide its own dependent-issue latency.
        const int q = warp & 3, half = (warp - 2) >> 2;
        uint8_t* stg = sOut + (size_t)(warp - 2) * 2048;
        const uint32_t sw = (uint32_t)((lane >> 1) & 3);
        const int r_sub = lane >> 2, ch = lane & 3;
        int acc = 0; uint32_t acc_phase = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            if (warp == 2 && lane == 0 && tr_tile < 8) PFO_TRACE(8 + 8 * tr_tile + 4);
            const int64_t m0 = tile * TM + q * 32;
            const int64_t mc = m0 + lane < M ? m0 + lane : M - 1;  // clamp the side loads of rows past M
            const bool zero_row = p.row_zero && p.row_zero[mc] != 0;
            const float brs = p.bias_row_scale ? p.bias_row_scale[mc * p.ld_brs] : 1.0f;
            for (int c0 = half * 16; c0 < NT; c0 += 32) {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_stride + (uint32_t)c0;
                const int nb = n0 + c0 + ch * 4;
                const bool full4 = p.vec_ok && nb + 3 < p.N;
                // the relu gate and the accumulate operand of the slab's four row groups are loaded up front: inside
                // the store loop each load sat right in front of its store, four dependent DRAM round trips per
                // slab (the gated dgrad launch took 26 us against 14 us for the ungated one of the same shape)
                float4 g4[4], o4[4];
                if (full4 && (p.relu_gate || p.accumulate)) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int64_t m = m0 + i * 8 + r_sub;
                        g4[i] = make_float4(1.f, 1.f, 1.f, 1.f);
                        o4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (m < M) {
                            if (p.relu_gate) g4[i] = *reinterpret_cast<const float4*>(p.relu_gate + m * p.ld_gate + nb);
                            if (p.accumulate) o4[i] = *reinterpret_cast<const float4*>(p.C + m * p.ldc + nb);
                        }
                    }
                }
                uint32_t r[16];
                tmem_ld16(taddr, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                __syncwarp();                                      // the previous slab has been read out of stg
                uint8_t* dst = stg + (size_t)lane * 64;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 b4 = *reinterpret_cast<const float4*>(sBias + c0 + j * 4);
                    float4 v;
                    v.x = fmaf(b4.x, brs, __uint_as_float(r[j * 4])) * p.alpha;
                    v.y = fmaf(b4.y, brs, __uint_as_float(r[j * 4 + 1])) * p.alpha;
                    v.z = fmaf(b4.z, brs, __uint_as_float(r[j * 4 + 2])) * p.alpha;
                    v.w = fmaf(b4.w, brs, __uint_as_float(r[j * 4 + 3])) * p.alpha;
                    if (p.act == 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    if (zero_row) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(dst + (((uint32_t)j ^ sw) << 4)) = v;
                }
                __syncwarp();
                if (nb < p.N) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = i * 8 + r_sub;
                        const int64_t m = m0 + row;
                        if (m >= M) break;
                        const float4 t = *reinterpret_cast<const float4*>(stg + row * 64 + (((uint32_t)ch ^ (uint32_t)((row >> 1) & 3)) << 4));
                        float v[4] = {t.x, t.y, t.z, t.w};
                        float* cp = p.C + m * p.ldc + nb;
                        if (full4) {
                            if (p.relu_gate) {
                                if (g4[i].x <= 0.f) v[0] = 0.f;
                                if (g4[i].y <= 0.f) v[1] = 0.f;
                                if (g4[i].z <= 0.f) v[2] = 0.f;
                                if (g4[i].w <= 0.f) v[3] = 0.f;
                            }
                            if (p.accumulate) { v[0] += o4[i].x; v[1] += o4[i].y; v[2] += o4[i].z; v[3] += o4[i].w; }
                            *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (nb + e >= p.N) break;
                                float x = v[e];
                                if (p.relu_gate && p.relu_gate[m * p.ld_gate + nb + e] <= 0.f) x = 0.f;
                                cp[e] = p.accumulate ? cp[e] + x : x;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));           // TMEM accumulator drained
            if (warp == 2 && lane == 0 && tr_tile < 8) PFO_TRACE(8 + 8 * tr_tile + 5);
            ++tr_tile;
            acc ^= 1; if (acc == 0) acc_phase ^= 1u;
        }
    } else if (PASSES == 3) {
        // ===== operand splitters: hi = rna_tf32(a) in place, lo = a - hi into the stage's TMEM slot (the A operand of
        // the A_lo . W_hi MMA comes from tensor memory, so the ring holds twice the stages a shared-memory twin allowed).
        // Thread = one row of the tile (TMEM lane), warp pair per lane quarter, 16 of the chunk's 32 columns each.
        const int q = warp & 3, half = (warp - (2 + EPI_WARPS)) >> 2;
        const int row = q * 32 + lane;
        const uint32_t row_off = (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
        const uint32_t lo_taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)p.lo_col + (uint32_t)half * 16u;
        int stage = 0; uint32_t phase = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(full_bar(stage), phase);
                if (c == 0 && tid == THREADS_1 && tr_tile < 8) PFO_TRACE(8 + 8 * tr_tile + 1);
                uint8_t* st = sA + (size_t)stage * CHUNK_BYTES + row_off;
                uint32_t lo[16];
                if (p.dbg & 1) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(ready_bar(stage));
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                    continue;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4* ptr4 = reinterpret_cast<float4*>(st + ((((uint32_t)(half * 4 + j)) ^ sw) << 4));
                    const float4 v = *ptr4;
                    float4 h;
                    h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
                    lo[j * 4 + 0] = __float_as_uint(v.x - h.x); lo[j * 4 + 1] = __float_as_uint(v.y - h.y);
                    lo[j * 4 + 2] = __float_as_uint(v.z - h.z); lo[j * 4 + 3] = __float_as_uint(v.w - h.w);
                    *ptr4 = h;
                }
                tmem_st16(lo_taddr + (uint32_t)stage * 32u, lo);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(ready_bar(stage));
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
            ++tr_tile;
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (tid == 0) PFO_TRACE(2);
    if (p.trace && tid == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[128 + 2 * (blockIdx.y * gridDim.x + blockIdx.x) + 1] = (long long)t;
    }
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     :: "r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

template <int PASSES>
int launch_tma(const CUtensorMap& map, const TmaLinArgs& a, dim3 grid, size_t smem, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(linear_tma_kernel<PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    pfo_launch(linear_tma_kernel<PASSES>, grid, PASSES == 3 ? THREADS_3 : THREADS_1, smem, s, map, a);
    PFO_LAUNCH_CHECK();
}

long long* g_linear_trace = nullptr;
int g_linear_min_stages = 5;      // swept 2..5 at the bench shapes after the MMA-issue fix: 0.841 -> 0.831 ms/step
int g_linear_dbg = 0;

}  // namespace

// Diagnostic hook (not part of the product ABI, not declared in include/pfo_b200.h): when set to a device buffer
// of >= 80 int64, CTA (0,0) of every pfo_linear_tf32 launch records clock64() at its pipeline milestones.
PFO_API void pfo_debug_set_linear_trace(void* device_buffer) { g_linear_trace = static_cast<long long*>(device_buffer); }

PFO_API void pfo_debug_set_linear_min_stages(int n) { g_linear_min_stages = n; }
PFO_API void pfo_debug_set_linear_dbg(int flags) { g_linear_dbg = flags; }

// Builds the 2-D tensor map of a row-major fp32 matrix [rows, cols] (row stride ld floats) with
// 32-column x box_rows boxes and the 128-byte swizzle (atom32: 32-byte swizzle atoms, the form the tensor
// core needs for MN-major tf32 operands).  Shared with wgrad_tma.cu.
int pfo_make_tensor_map_f32(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                            int atom32) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return (int)cudaErrorNotSupported;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4u};
    cuuint32_t box[2] = {(cuuint32_t)CK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

PFO_API int pfo_linear_tf32(const float* A, int64_t lda, const int32_t* a_idx, const float* W, int64_t ldw,
                            int w_transposed, const float* bias, const float* bias_row_scale, int64_t ld_brs,
                            float* C, int64_t ldc, int64_t M, const int32_t* m_dev, int N, int K,
                            float alpha, int act, const int32_t* row_zero, const float* relu_gate, int64_t ld_gate,
                            int accumulate, int passes, void* stream) {
    if (M <= 0 || N <= 0) return 0;
    if (passes != 1 && passes != 3) return (int)cudaErrorInvalidValue;
    // layouts TMA cannot describe (gathered rows, rows that are not 16-byte aligned) take the FFMA kernel
    const bool tma_ok = a_idx == nullptr && K >= 1 && (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                        get_encode_tiled() != nullptr;
    if (!tma_ok)
        return pfo_linear_f32_impl(A, lda, a_idx, W, ldw, w_transposed, bias, bias_row_scale, ld_brs, C, ldc, M, m_dev,
                                   N, K, alpha, act, row_zero, relu_gate, ld_gate, accumulate, stream);
    TmaLinArgs a{};
    a.W = W; a.ldw = ldw; a.w_transposed = w_transposed; a.bias = bias; a.bias_row_scale = bias_row_scale;
    a.ld_brs = ld_brs; a.C = C; a.ldc = ldc; a.M = M; a.m_dev = m_dev; a.N = N; a.K = K; a.alpha = alpha; a.act = act;
    a.row_zero = row_zero; a.relu_gate = relu_gate; a.ld_gate = ld_gate; a.accumulate = accumulate;
    a.trace = g_linear_trace;
    a.dbg = g_linear_dbg;
    a.KP8 = (K + 7) / 8 * 8;
    a.n_chunks = (K + CK - 1) / CK;
    a.vec_ok = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
               (!relu_gate || ((ld_gate % 4 == 0) && ((reinterpret_cast<uintptr_t>(relu_gate) & 15) == 0)));
    const int mult = passes == 3 ? 2 : 1;        // weight copies in shared memory (hi, lo); the A-lo operand lives in TMEM
    const int stage_bytes = CHUNK_BYTES;
    // output staging, barriers + TMEM slot, alignment slack; the bias slice (4 * NT bytes) is counted with the weights
    const int fixed = EPI_BYTES + 8 * (3 * MAX_STAGES + 4) + 16 + 1024;
    const int budget = SMEM_LIMIT - fixed;
    // Output columns per CTA (whole 32-column store slabs).  The ring must be deep enough to keep the HBM latency
    // covered (Little: ~100 KB in flight per SM), so N is cut into more column tiles -- whose CTAs re-read the row
    // tile from L2 -- until `want` stages fit beside the resident weight slice; 3-pass mode also needs
    // 2 * NT + 32 * stages <= 512 TMEM columns.
    static const int env_want = [] { const char* e = getenv("PFO_LINEAR_MIN_STAGES"); return e ? atoi(e) : 0; }();
    const int want = env_want > 0 ? env_want : g_linear_min_stages;
    int n_ntiles = (N + 255) / 256;
    int NT, stages;
    for (;;) {
        NT = ((N + n_ntiles - 1) / n_ntiles + 31) / 32 * 32;
        const int64_t wb = (int64_t)(a.KP8 / 4) * (NT * 16 + 16) * mult + 4 * NT;
        stages = wb < budget ? (int)((budget - wb) / stage_bytes) : 0;
        if (stages > MAX_STAGES) stages = MAX_STAGES;
        if (passes == 3 && stages > (512 - 2 * NT) / 32) stages = (512 - 2 * NT) / 32;
        if (stages >= want || NT <= 32) break;
        ++n_ntiles;
    }
    if (stages < 1) return (int)cudaErrorInvalidValue;
    a.NT = NT;
    n_ntiles = (N + NT - 1) / NT;
    const int w_bytes = (a.KP8 / 4) * (NT * 16 + 16) * mult + 4 * NT;
    a.stages = stages;
    int cols = 32;
    const int need_cols = passes == 3 ? 2 * NT + 32 * stages : 2 * NT;
    while (cols < need_cols) cols <<= 1;
    a.tmem_cols = cols;
    a.acc_stride = passes == 3 ? NT : cols >> 1;
    a.lo_col = 2 * NT;
    CUtensorMap map;
    int rc = pfo_make_tensor_map_f32(&map, A, M, K, lda, TM, 0);
    if (rc) return rc;
    const size_t smem = (size_t)stages * stage_bytes + w_bytes + fixed;
    const int64_t m_tiles = (M + TM - 1) / TM;
    int gx = pfo_num_sms() / n_ntiles;
    if (gx < 1) gx = 1;
    if (gx > m_tiles) gx = (int)m_tiles;
    dim3 grid((unsigned)gx, (unsigned)n_ntiles);
    cudaStream_t s = (cudaStream_t)stream;
    return passes == 3 ? launch_tma<3>(map, a, grid, smem, s) : launch_tma<1>(map, a, grid, smem, s);
}
