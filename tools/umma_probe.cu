// Stand-alone probe of tcgen05 kind::tf32 with MN-major operands loaded by TMA (SWIZZLE_128B boxes of
// 32 rows x 32 fp32 columns).  Sweeps shared-memory descriptor variants and reports which one computes
// D[n, k] = sum_m G[m, n] * A[m, k] exactly on small-integer inputs.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o pfotgnrec_b200/build/umma_probe tools/umma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

constexpr int R = 32, GN = 128, AK = 64, BOX = R * 128;

struct Variant { uint32_t layout, lbo, sbo, kstride, a_major, b_major, lbo_mode, n; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, const Variant& v) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((v.lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((v.sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(v.lbo_mode & 1) << 52) |
           ((uint64_t)(v.layout & 7) << 61);
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmA, Variant v, float* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t full = smem_u32(&bars[0]), done = smem_u32(&bars[1]);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(full) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(done) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(full), "r"((uint32_t)(6 * BOX)) : "memory");
        for (int b = 0; b < 4; ++b)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         :: "r"(smem_u32(base + b * BOX)), "l"(reinterpret_cast<uint64_t>(&tmG)), "r"(b * 32), "r"(0), "r"(full) : "memory");
        for (int b = 0; b < 2; ++b)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         :: "r"(smem_u32(base + (4 + b) * BOX)), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(b * 32), "r"(0), "r"(full) : "memory");
        mbar_wait(full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (v.a_major << 15) | (v.b_major << 16) |
                               ((v.n >> 3) << 17) | ((128u >> 4) << 24);
        for (int s = 0; s < R / 8; ++s) {
            const uint64_t dg = mk_desc(smem_u32(base) + s * v.kstride, v);
            const uint64_t da = mk_desc(smem_u32(base + 4 * BOX) + s * v.kstride, v);
            const uint32_t acc = s ? 1u : 0u;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                         :: "r"(tmem), "l"(dg), "l"(da), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(done) : "memory");
    }
    __syncwarp();
    mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < (int)v.n; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no encode\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    std::vector<float> G(R * GN), A(R * AK), D(128 * 64), ref(128 * 64);
    float *dG, *dA, *dO;
    cudaMalloc(&dG, G.size() * 4); cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dO, D.size() * 4);
    auto mk = [&](CUtensorMap* m, float* p, int cols, CUtensorMapSwizzle swz) {
        cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)R}; cuuint64_t str[1] = {(cuuint64_t)cols * 4};
        cuuint32_t box[2] = {32, 32}; cuuint32_t es[2] = {1, 1};
        return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUtensorMap mg, ma, mg32, ma32;
    if (mk(&mg, dG, GN, CU_TENSOR_MAP_SWIZZLE_128B) != CUDA_SUCCESS || mk(&ma, dA, AK, CU_TENSOR_MAP_SWIZZLE_128B) != CUDA_SUCCESS ||
        mk(&mg32, dG, GN, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) != CUDA_SUCCESS ||
        mk(&ma32, dA, AK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const Variant vars[] = {
        {2, 4096, 1024, 1024, 1, 1, 0, 64},   // 0: as implemented in wgrad_tma.cu
        {2, 1024, 4096, 1024, 1, 1, 0, 64},   // 1: LBO / SBO swapped
        {1, 4096, 1024, 1024, 1, 1, 0, 64},   // 2: SWIZZLE_128B_BASE32B
        {2, 4096, 1024, 1024, 1, 1, 1, 64},   // 3: lbo_mode = 1
        {2, 4096, 128, 1024, 1, 1, 0, 64},    // 4: SBO = row pitch
        {2, 4096, 1024, 1024, 0, 0, 0, 64},   // 5: K-major bits with the MN data (control: must be wrong)
        {2, 4096, 1024, 1024, 1, 0, 0, 64},   // 6: only A MN-major
        {2, 4096, 1024, 1024, 0, 1, 0, 64},   // 7: only B MN-major
        {1, 4096, 512, 1024, 1, 1, 0, 64},    // 8..: ATOM_32B tensor maps + SWIZZLE_128B_BASE32B descriptors
        {1, 4096, 1024, 1024, 1, 1, 0, 64},   // 9
        {1, 4096, 512, 512, 1, 1, 0, 64},     // 10
        {1, 512, 4096, 1024, 1, 1, 0, 64},    // 11
        {1, 4096, 512, 1024, 1, 1, 0, 48},    // 12: N = 48
    };
    for (int pat = 0; pat < 2; ++pat) {
        for (auto& x : G) x = 0; for (auto& x : A) x = 0;
        if (pat == 0) {
            srand(1);
            for (auto& x : G) x = (float)(rand() % 7 - 3);
            for (auto& x : A) x = (float)(rand() % 5 - 2);
        } else {            // one-hot: m0 = 9, n0 = 37, k0 = 41
            G[9 * GN + 37] = 1.0f; A[9 * AK + 41] = 1.0f;
        }
        for (int n = 0; n < 128; ++n) for (int k = 0; k < 64; ++k) {
            float s = 0; for (int m = 0; m < R; ++m) s += G[m * GN + n] * A[m * AK + k];
            ref[n * 64 + k] = s;
        }
        cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        for (int vi = 0; vi < (int)(sizeof(vars) / sizeof(vars[0])); ++vi) {
            cudaMemset(dO, 0xff, D.size() * 4);
            probe_kernel<<<1, 128, 32 * 1024, 0>>>(vi >= 8 ? mg32 : mg, vi >= 8 ? ma32 : ma, vars[vi], dO);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("pat %d variant %d: CUDA error %s\n", pat, vi, cudaGetErrorString(e)); return 2; }
            cudaMemcpy(D.data(), dO, D.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0, nz = 0;
            for (int i = 0; i < 128 * 64; ++i) {
                if ((i % 64) >= (int)vars[vi].n) continue;
                if (D[i] != ref[i]) ++bad;
                if (D[i] != 0.0f) ++nz;
            }
            printf("pat %d variant %d: mismatches %d / %d, nonzeros %d\n", pat, vi, bad, 128 * 64, nz);
            if (pat == 1) for (int i = 0; i < 128 * 64; ++i) if ((i % 64) < (int)vars[vi].n && D[i] != 0.0f) printf("    D[n=%d][k=%d] = %g\n", i / 64, i % 64, D[i]);
        }
    }
    return 0;
}
