// Full-range sin/cos of an fp32 argument without the Payne-Hanek slow path.
//
// TimeEncode arguments dt*w+b reach 1e13 rad on YYYYMMDDhhmmss timestamps (SURVEY.md hard part 1),
// far beyond the 105615 rad where CUDA's cosf switches to a ~100-instruction, divergent,
// local-memory slow path.  The fp32 argument is exact in fp64, so the quadrant reduction
//     k = rint(x * 2/pi),   r = x - k*pi/2   (two fp64 fmas with a hi/lo split of pi/2)
// is accurate to ~1e-16 for every |x| < 2^44, after which r in [-pi/4, pi/4] is rounded to fp32 and
// fed to the classic single-precision minimax polynomials.  Absolute error <= ~1.5e-7 (about
// 2 ulp of 1.0), the same class as cosf itself, at ~25 uniform instructions.
#pragma once

__device__ __forceinline__ void pfo_sincosf(float x, float* s, float* c) {
    const double xd = (double)x;
    // k = rint(x * 2/pi) by the magic-number trick (|k| < 2^51): the biased sum holds k in its low mantissa bits, so
    // the quadrant comes from a register move instead of an F2I.S64 and the rounding from a DADD instead of an FRND
    // (conversions issue at a quarter of the fp64 FMA rate)
    const double t = fma(xd, 0.63661977236758134308, 6755399441055744.0);   // 2/pi, 1.5 * 2^52
    const int q = __double2loint(t) & 3;
    const double kd = t - 6755399441055744.0;
    double r = fma(-kd, 1.57079632679489655800e+00, xd);            // pi/2 (hi)
    r = fma(-kd, 6.12323399573676603587e-17, r);                    // pi/2 (lo)
    const float rf = (float)r;
    const float r2 = rf * rf;
    const float sn = fmaf(rf * r2, fmaf(r2, fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f), rf);
    const float cs = fmaf(r2 * r2, fmaf(r2, fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f),
                                        4.166664568298827e-2f), fmaf(-0.5f, r2, 1.0f));
    const float s0 = (q & 1) ? cs : sn;
    const float c0 = (q & 1) ? sn : cs;
    *s = (q & 2) ? -s0 : s0;
    *c = ((q + 1) & 2) ? -c0 : c0;
}

__device__ __forceinline__ float pfo_cosf(float x) {
    float s, c;
    pfo_sincosf(x, &s, &c);
    return c;
}
