// Full-range sin/cos of an fp32 argument without the Payne-Hanek slow path.
//
// TimeEncode arguments dt*w+b reach 1e13 rad on YYYYMMDDhhmmss timestamps (SURVEY.md hard part 1),
// far beyond the 105615 rad where CUDA's cosf switches to a ~100-instruction, divergent,
// local-memory slow path (measured: 133 Gcos/s against 1 100 on such arguments, tools/cos_probe.cu).  Two quadrant
// reductions k = rint(x * 2/pi), r = x - k*pi/2 feed the same classic single-precision minimax polynomials on
// [-pi/4, pi/4]:
//   * fp64 (any |x| < 2^44) -- THE PRODUCT PATH: the fp32 argument is exact in fp64; two fp64 fmas with a hi/lo split
//     of pi/2 give r to ~1e-16 before it is rounded to fp32.  Two conversions and four fp64 operations per argument.
//   * fp32 Cody-Waite (|x| < 2^17): three fp32 fmas with a three-way split of pi/2; within 1 ulp of 1.0 of the fp64
//     path (tests/test_gpu_kernels.py::test_time_encode_cos_paths).
// Measured on B200 (profiles/r2_cos_probe.txt): 1 088 Gcos/s for the fp64 reduction, 1 341 for the fp32 one -- fp64
// FMAs issue at half the fp32 rate here, so the fp64 part is ~20 % of a cosine, not its bulk -- and choosing between
// the two per warp (vote + branch) ran SLOWER than the fp64 path alone (1 049 Gcos/s with every argument small, 835 on
// mixed warps; attention fwd / bwd 133 -> 149 us / 180 -> 204 us).  The kernels therefore always take the fp64
// reduction; the fp32 one stays as `mode 2` of pfo_time_encode for the comparison test.
// Absolute error <= ~1.5e-7 (about 2 ulp of 1.0), the same class as cosf itself.
#pragma once

#define PFO_COS_FP32_LIMIT 131072.0f

__device__ __forceinline__ void pfo_sincos_poly(float rf, int q, float* s, float* c) {
    const float r2 = rf * rf;
    const float sn = fmaf(rf * r2, fmaf(r2, fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f), rf);
    const float cs = fmaf(r2 * r2, fmaf(r2, fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f),
                                        4.166664568298827e-2f), fmaf(-0.5f, r2, 1.0f));
    const float s0 = (q & 1) ? cs : sn;
    const float c0 = (q & 1) ? sn : cs;
    *s = (q & 2) ? -s0 : s0;
    *c = ((q + 1) & 2) ? -c0 : c0;
}

__device__ __forceinline__ void pfo_sincosf_f64(float x, float* s, float* c) {
    const double xd = (double)x;
    // k = rint(x * 2/pi) by the magic-number trick (|k| < 2^51): the biased sum holds k in its low mantissa bits, so
    // the quadrant comes from a register move instead of an F2I.S64 and the rounding from a DADD instead of an FRND
    // (conversions issue at a quarter of the fp64 FMA rate)
    const double t = fma(xd, 0.63661977236758134308, 6755399441055744.0);   // 2/pi, 1.5 * 2^52
    const int q = __double2loint(t) & 3;
    const double kd = t - 6755399441055744.0;
    double r = fma(-kd, 1.57079632679489655800e+00, xd);            // pi/2 (hi)
    r = fma(-kd, 6.12323399573676603587e-17, r);                    // pi/2 (lo)
    pfo_sincos_poly((float)r, q, s, c);
}

__device__ __forceinline__ void pfo_sincosf_f32(float x, float* s, float* c) {
    // valid for |x| < PFO_COS_FP32_LIMIT (k < 2^17, well inside the 2^22 range of the fp32 magic number)
    const float t = fmaf(x, 0.63661977236758134308f, 12582912.0f);          // 2/pi, 1.5 * 2^23
    const int q = __float_as_int(t) & 3;
    const float kf = t - 12582912.0f;
    float r = fmaf(-kf, 1.57079601e+00f, x);                                // pi/2 split three ways (Cody-Waite)
    r = fmaf(-kf, 3.13916473e-07f, r);
    r = fmaf(-kf, 5.39030253e-15f, r);
    pfo_sincos_poly(r, q, s, c);
}

// Cosine alone (the neighbour forward kernel needs no sine): half-turn reduction k = rint(x / pi), r = x - k pi in
// [-pi/2, pi/2] with the same fp64 magic-number rounding and hi / lo split, then ONE even polynomial of degree 10 in r
// (interpolated at the Chebyshev nodes of r^2; 2.2e-10 from cos on the interval, 1.1e-7 with its fp32 Horner rounding --
// the class of the quadrant form, tests/test_gpu_kernels.py::test_time_encode_cos_paths) and the sign of (-1)^k as an
// XOR.  14 instructions against ~24 for the sine / cosine pair + quadrant selects.
__device__ __forceinline__ float pfo_cosf_half(float x) {
    const double xd = (double)x;
    const double t = fma(xd, 0.31830988618379067154, 6755399441055744.0);    // 1/pi, 1.5 * 2^52
    const unsigned sign = ((unsigned)__double2loint(t)) << 31;
    const double kd = t - 6755399441055744.0;
    double r = fma(-kd, 3.14159265358979311600e+00, xd);                      // pi (hi)
    r = fma(-kd, 1.22464679914735317723e-16, r);                              // pi (lo)
    const float rf = (float)r;
    const float u = rf * rf;
    float c = fmaf(u, -2.604992346277868e-07f, 2.476006920915097e-05f);
    c = fmaf(u, c, -0.0013888359535485506f);
    c = fmaf(u, c, 0.04166663438081741f);
    c = fmaf(u, c, -0.5f);
    c = fmaf(u, c, 1.0f);
    return __uint_as_float(__float_as_uint(c) ^ sign);
}

__device__ __forceinline__ void pfo_sincosf(float x, float* s, float* c) { pfo_sincosf_f64(x, s, c); }

__device__ __forceinline__ float pfo_cosf_f64(float x) { float s, c; pfo_sincosf_f64(x, &s, &c); return c; }
__device__ __forceinline__ float pfo_cosf_f32(float x) { float s, c; pfo_sincosf_f32(x, &s, &c); return c; }

__device__ __forceinline__ float pfo_cosf(float x) {
    float s, c;
    pfo_sincosf(x, &s, &c);
    return c;
}
