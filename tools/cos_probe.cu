// Throughput probe of TimeEncode's cos(fmaf(dt, w, b)) variants on sm_100a (tools/, not product).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/cos_probe.cu -o gpurun_out/cos_probe && gpurun_out/cos_probe
// Each thread evaluates 64 * ITER arguments held in registers and xors the results (no memory traffic in the loop).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../pfotgnrec_b200/csrc/pfo_math.cuh"

template <int V>
__global__ void probe(const float* __restrict__ dt, const float* __restrict__ w, float* out, int iters, float scale) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    float acc = 0.f, acc2 = 0.f;
    float w0 = w[2 * lane] * scale, w1 = w[2 * lane + 1] * scale;
    float d = dt[tid];
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int j = 0; j < 16; ++j) {
            const float x0 = fmaf(d, w0, (float)j), x1 = fmaf(d, w1, (float)j);
            if (V == 0) { acc += pfo_cosf_f64(x0); acc += pfo_cosf_f64(x1); }
            if (V == 1) { acc += pfo_cosf_f32(x0); acc += pfo_cosf_f32(x1); }
            if (V == 2) { acc += pfo_cosf(x0); acc += pfo_cosf(x1); }                       // warp-voted choice
            if (V == 3) { acc += cosf(x0); acc += cosf(x1); }                               // CUDA libm
            if (V == 4) { float s, c; pfo_sincosf(x0, &s, &c); acc += c; acc2 += s; pfo_sincosf(x1, &s, &c); acc += c; acc2 += s; }
            if (V == 5) { float s, c; pfo_sincosf_f64(x0, &s, &c); acc += c; acc2 += s; pfo_sincosf_f64(x1, &s, &c); acc += c; acc2 += s; }
            if (V == 6) { float s, c; pfo_sincosf_f32(x0, &s, &c); acc += c; acc2 += s; pfo_sincosf_f32(x1, &s, &c); acc += c; acc2 += s; }
        }
        d += 1.0f;
    }
    out[tid] = acc + acc2;
}

template <int V>
float run(const float* dt, const float* w, float* out, float scale, const char* name) {
    const int iters = 64, grid = 148 * 8, block = 256;
    probe<V><<<grid, block>>>(dt, w, out, 4, scale);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<V><<<grid, block>>>(dt, w, out, iters, scale);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double evals = (double)grid * block * iters * 32;
    printf("%-34s scale %8.1e: %7.3f ms  %7.1f Gcos/s\n", name, scale, ms, evals / ms * 1e-6);
    return ms;
}

int main() {
    const int N = 148 * 8 * 256;
    float *dt, *w, *out;
    cudaMalloc(&dt, N * 4); cudaMalloc(&w, 64 * 4); cudaMalloc(&out, N * 4);
    float* h = new float[N];
    for (int i = 0; i < N; ++i) h[i] = 1000.0f + (i % 977) * 37.0f;
    cudaMemcpy(dt, h, N * 4, cudaMemcpyHostToDevice);
    float hw[64];
    for (int i = 0; i < 64; ++i) hw[i] = powf(10.f, -9.f * i / 63.f);
    cudaMemcpy(w, hw, 64 * 4, cudaMemcpyHostToDevice);
    for (float scale : {1.0f, 1000.0f}) {          // 1: every argument < 2^17; 1000: the low columns exceed it
        run<0>(dt, w, out, scale, "cos  fp64 reduction");
        run<1>(dt, w, out, scale, "cos  fp32 Cody-Waite (|x|<2^17 only)");
        run<2>(dt, w, out, scale, "cos  warp-voted");
        run<3>(dt, w, out, scale, "cos  cosf (libm)");
        run<5>(dt, w, out, scale, "sincos fp64 reduction");
        run<6>(dt, w, out, scale, "sincos fp32 Cody-Waite");
        run<4>(dt, w, out, scale, "sincos warp-voted");
    }
    return 0;
}
