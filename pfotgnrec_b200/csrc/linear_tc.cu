// bf16 tcgen05 / TMEM path of the tall-skinny contractions (2e-2 contract).
// NOTE: under construction in this round -- the entry point reports "not supported" so that the
// host can never silently fall back; gemm_mode="bf16" raises until the kernel lands.
#include "common.cuh"

PFO_API int pfo_linear_bf16(const float* A, int64_t lda, const int32_t* a_idx, const float* W, int64_t ldw,
                            int w_transposed, const float* bias, const float* bias_row_scale, int64_t ld_brs,
                            float* C, int64_t ldc, int64_t M, const int32_t* m_dev, int N, int K,
                            float alpha, int act, const int32_t* row_zero, const float* relu_gate, int64_t ld_gate,
                            int accumulate, void* stream) {
    return (int)cudaErrorNotSupported;
}
