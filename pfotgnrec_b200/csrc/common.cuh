// Shared device helpers for the pfotgnrec_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/pfo_b200.h"

#define PFO_API extern "C" __attribute__((visibility("default")))

// Every entry point returns 0 on success or the cudaError_t of the failed launch.
#define PFO_LAUNCH_CHECK() do { cudaError_t e__ = cudaGetLastError(); return (int)e__; } while (0)

static inline int pfo_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// grid for a grid-stride kernel: enough CTAs to cover `work` items, capped at a whole number of waves
static inline int pfo_grid(int64_t work, int block, int ctas_per_sm) {
    int64_t need = (work + block - 1) / block;
    int64_t cap = (int64_t)pfo_num_sms() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// CTAs of `kernel` resident per SM at this block size and dynamic shared-memory size (occupancy API; a few cached
// entries).  A grid-stride kernel whose grid is a guessed multiple of the SM count runs a short second wave when fewer
// CTAs fit than guessed (10 CTAs per SM asked, 6 resident: the last 4 run at two thirds of the occupancy); sizing the
// grid to exactly one resident wave keeps every SM at full occupancy until the end.
template <typename... KArgs>
static inline int pfo_resident(void (*kernel)(KArgs...), int block, size_t smem) {
    struct Entry { const void* k; int block; size_t smem; int occ; };
    static Entry cache[16];
    static int used = 0;
    const void* kp = reinterpret_cast<const void*>(kernel);
    for (int i = 0; i < used; ++i)
        if (cache[i].k == kp && cache[i].block == block && cache[i].smem == smem) return cache[i].occ;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, smem) != cudaSuccess || occ < 1) occ = 1;
    if (used < 16) cache[used++] = Entry{kp, block, smem, occ};
    return occ;
}

// ---- programmatic dependent launch.  A step is ~60 kernels of 5-200 us chained through one stream (one CUDA graph):
// with plain launches every kernel boundary costs the full drain -> schedule -> ramp-up gap.  Every kernel of this
// library is therefore launched with the programmatic-stream-serialization attribute and begins with
//     griddepcontrol.wait;               (returns once the preceding grid has completed and its writes are visible)
//     griddepcontrol.launch_dependents;  (lets the NEXT grid's CTAs be scheduled while this one still runs)
// so the next kernel's launch latency and CTA ramp-up overlap this kernel's execution, while no kernel reads or
// writes global memory before its predecessor is complete (the wait comes first, so the relaxation is never
// transitive: when grid C starts, grid B has passed its own wait, i.e. grid A is complete).  Kernels of other
// libraries in between (torch, NCCL) neither wait nor trigger and keep plain stream order.  PFO_PDL=0 in the
// environment turns the attribute off (the two instructions are then no-ops).
__device__ __forceinline__ void pfo_pdl_prologue() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

static inline bool pfo_pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("PFO_PDL");
        on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}

template <typename... KArgs, typename... Args>
static inline void pfo_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pfo_pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // errors surface through PFO_LAUNCH_CHECK
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 16-byte vector reduction into global memory (sm_90+): one L2 op instead of four
__device__ __forceinline__ void red_add_f32x4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void red_add_f32x2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(addr), "f"(a), "f"(b) : "memory");
}

// base + idx * stride_bytes as ONE IMAD.WIDE (signed 32 x 32 + 64).  Written as PTX because the compiler, given
// `base + (int64_t)idx * stride`, re-derives the base from its parts and widens the stride (6 integer instructions per
// gathered row in the neighbour kernels' SASS instead of 1).
__device__ __forceinline__ const char* pfo_row_ptr(const char* base, int idx, int stride_bytes) {
    unsigned long long a;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(a) : "r"(idx), "r"(stride_bytes), "l"((unsigned long long)base));
    return reinterpret_cast<const char*>(a);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
