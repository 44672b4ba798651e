// K4 (neighbour level): fused gather + time encoding + masked softmax over the sampled
// temporal neighbours, in the weight-absorbed form of multi-head attention.
//
// Reference: model/temporal_attention.py:34-90 feeding torch.nn.MultiheadAttention with
// key = value = [h_nbr | e | cos(dt*w+b)] (:52).  Because K and V are linear in the
// neighbour row x_j, the per-neighbour projections are never materialised:
//     score_hj = q_h . (Wk_h x_j + bk_h) = (Wk_h^T q_h) . x_j + const_h   (const cancels in softmax)
//     out_h    = sum_j p_hj (Wv_h x_j + bv_h) = Wv_h (sum_j p_hj x_j) + bv_h * sum_j p_hj
// so this kernel consumes qk_h = Wk_h^T q_h per query and produces xbar_h = sum_j p_hj x_j
// (plus psum_h); the Q-level projections around it are tall-skinny GEMMs (linear_*.cu).
// This cuts the neighbour-level FLOPs by ~8x and leaves a gather-bound kernel:
// one warp per query, lanes own contiguous feature columns, online softmax in registers.
//
// Row layout of QK / XB / dQK / dXB: [Q, H, EKP], segment = [h (d) | e (F) | te (d) | psum | valid | one | 0..],
// EKP >= 2d + F + 3.  `valid` (1 when the query has at least one neighbour, else 0) and `one` (always 1) are
// constant columns the forward kernel emits so that the host can fold the out-projection bias (applied to
// valid rows only, temporal_attention.py:84) and the merge-layer bias into the weight of the GEMM that consumes
// XB.  XB / dXB rows may be strided (ldxb / lddxb floats between queries) so that XB lands directly inside
// the operand row [XB | h_query] of that GEMM.
#include "common.cuh"
#include "philox.cuh"
#include "pfo_math.cuh"

namespace {

constexpr int kMaxHeads = 4;      // instantiated head counts: 1, 2, 4
constexpr int64_t kMaxRowFloats = (int64_t)1 << 28;   // row strides are kept as 32-bit byte counts in the kernels

struct NbrArgs {
    const float* QK; const int32_t* qk_row; const float* T; int64_t ldt;
    const int32_t* idx; const int32_t* eidx; const float* dt;
    const float* efeat; const float* tw; const float* tb;
    int64_t Q; int n; int d; int F; int H; int ekp;
    float p_drop; uint32_t k0, k1, step; const uint32_t* step_dev;
    float* XB; int64_t ldxb; float* P; int32_t* invalid;
    // backward only
    const float* dXB; int64_t lddxb; float* dQK; float* dT; int64_t lddt; float* partial;
};

// the dropout stream is keyed by (query, slot, step): ONE Philox4x32 block per (query, slot), word h of the block is
// head h's draw (kMaxHeads = 4 words).  `step` = p.step + *p.step_dev, the device part being a per-batch counter the
// host bumps on the stream (so a captured CUDA graph replays fresh masks)
template <int NH>
__device__ __forceinline__ void keep_scales(const NbrArgs& p, uint32_t step, int64_t q, int j, float (&keep)[NH]) {
    static_assert(NH <= 4, "one Philox block carries four heads");
#pragma unroll
    for (int h = 0; h < NH; ++h) keep[h] = 1.0f;
    if (p.p_drop <= 0.0f) return;
    const Philox4 r = philox4x32_10((uint32_t)q, (uint32_t)j, step, PFO_PURPOSE_DROPOUT, p.k0, p.k1);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    const float inv = 1.0f / (1.0f - p.p_drop);
#pragma unroll
    for (int h = 0; h < NH; ++h) {
        const float u = (float)(w[h] >> 8) * (1.0f / 16777216.0f);
        keep[h] = u < p.p_drop ? 0.0f : inv;
    }
}

// DPL consecutive floats of a row owned by one lane (rows are 16-byte aligned, checked by the entry points)
// same, through the read-only path of a gathered row whose address came out of pfo_row_ptr (an address built in PTX
// carries no state space: a plain dereference would compile to a generic LD instead of LDG)
template <int DPL>
__device__ __forceinline__ void ldg_cols(const char* p, float (&v)[DPL]) {
    if constexpr (DPL == 2) { const float2 t = __ldg(reinterpret_cast<const float2*>(p)); v[0] = t.x; v[1] = t.y; }
    else if constexpr (DPL == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else {
#pragma unroll
        for (int i = 0; i < DPL; ++i) v[i] = __ldg(reinterpret_cast<const float*>(p) + i);
    }
}
template <int DPL>
__device__ __forceinline__ void ld_cols(const float* __restrict__ p, float (&v)[DPL]) {
    if constexpr (DPL == 2) { const float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y; }
    else if constexpr (DPL == 4) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else {
#pragma unroll
        for (int i = 0; i < DPL; ++i) v[i] = p[i];
    }
}
template <int DPL>
__device__ __forceinline__ void st_cols(float* __restrict__ p, const float (&v)[DPL]) {
    if constexpr (DPL == 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    else if constexpr (DPL == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else {
#pragma unroll
        for (int i = 0; i < DPL; ++i) p[i] = v[i];
    }
}

constexpr int kUnroll = 4;      // neighbour rows in flight per warp (independent gathers issued back to back)

// Sum NV (power of two <= 32) per-lane values across the warp at once: each step of the butterfly halves the
// number of values a lane carries (it sends one half, keeps and accumulates the other), then the usual
// xor-reduction finishes.  NV + log2(32 / NV) - 1 shuffles instead of 5 * NV; the total of value v ends up in
// lanes [v * 32 / NV, (v + 1) * 32 / NV).
template <int NV, int O = 16>
__device__ __forceinline__ float multi_reduce(float (&v)[NV], int lane) {
    if constexpr (NV == 1) {
        float r = v[0];
#pragma unroll
        for (int o = O; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        return r;
    } else {
        const bool upper = (lane & O) != 0;
        float h[NV / 2];
#pragma unroll
        for (int i = 0; i < NV / 2; ++i) {
            const float send = upper ? v[i] : v[i + NV / 2];
            const float keep = upper ? v[i + NV / 2] : v[i];
            h[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
        }
        return multi_reduce<NV / 2, O / 2>(h, lane);
    }
}

// One warp per query; lane l owns feature columns [l*DPL, (l+1)*DPL) of the h and te segments and -- for the
// per-slot scalars (neighbour id, dt, edge id, dropout factor, score, softmax weight) -- neighbour slot l.
//   phase A  slot scalars by one coalesced load each; neighbour rows gathered kUnroll at a time (independent L2
//            requests in flight), x_j = [h_j | e_j | cos(dt_j w + b)] stashed in shared memory, the kUnroll * NH
//            partial scores reduced across the warp together (multi_reduce), score of slot j kept in lane j
//   phase B  softmax over the slots, lane-parallel: one exp per lane and head, two warp reductions per head
//   phase C  xbar_h = sum_j p'_hj x_j from the stash (p' = softmax weight times the dropout factor)
template <int DPL, int NH>
__global__ void __launch_bounds__(128)
attn_nbr_fwd_kernel(const NbrArgs p) {
    pfo_pdl_prologue();
    extern __shared__ float smem[];
    constexpr int d = 32 * DPL;
    constexpr int SW = 2 * d + 32;                      // stash row: [h | cos | e(32)]
    constexpr int NV = kUnroll * NH;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int F = p.F, n = p.n, ekp = p.ekp;
    float* stash = smem + (size_t)wib * n * SW;
    const int c0 = lane * DPL;
    // gather addresses as byte base + index * byte stride with a 32-bit stride (checked by the entry point): one
    // IMAD.WIDE per gathered row instead of a 64 x 32-bit multiply-add chain (the SASS showed 4 + 10 integer
    // instructions per neighbour for the two gathers)
    const char* Tb = reinterpret_cast<const char*>(p.T) + (size_t)c0 * sizeof(float);
    const int ldtb = (int)(p.ldt * (int64_t)sizeof(float));
    const char* efb = reinterpret_cast<const char*>(p.efeat) + (size_t)lane * sizeof(float);
    const int Fb = p.F * (int)sizeof(float);
    const uint32_t step = p.step + (p.step_dev ? *p.step_dev : 0u);
    float tw[DPL], tb[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) { tw[i] = p.tw[c0 + i]; tb[i] = p.tb[c0 + i]; }
    // the per-query header (slot scalars and the query operand) is loaded one query ahead: those loads head every
    // dependency chain of the body, and ncu's source view showed ~12 % of the samples waiting on them
    int id_n = -1, ei_n = 0;
    float dt_n = 0.0f;
    float qa_n[NH][DPL], qg_n[NH][DPL], qe_n[NH];
    auto load_header = [&](int64_t q) {
        id_n = -1; ei_n = 0; dt_n = 0.0f;
        if (q < p.Q) {
            if (lane < n) {
                id_n = __ldg(p.idx + q * n + lane);
                dt_n = __ldg(p.dt + q * n + lane);
                ei_n = __ldg(p.eidx + q * n + lane);
            }
            // queries may share rows of the query operand (pfo_attn_nbr_fwd_rows: one row per distinct query NODE)
            const int64_t qr = p.qk_row ? (int64_t)__ldg(p.qk_row + q) : q;
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const float* qk = p.QK + (qr * NH + h) * ekp;
                ld_cols<DPL>(qk + c0, qa_n[h]);
#pragma unroll
                for (int i = 0; i < DPL; ++i) qg_n[h][i] = qk[d + F + c0 + i];
                qe_n[h] = lane < F ? qk[d + lane] : 0.0f;
            }
        }
    };
    load_header(warp);
    for (int64_t q = warp; q < p.Q; q += nwarps) {
        const int id_l = id_n, ei_l = ei_n;
        const float dt_l = dt_n;
        float qa[NH][DPL], qg[NH][DPL], qe[NH], sj[NH];
#pragma unroll
        for (int h = 0; h < NH; ++h) {
#pragma unroll
            for (int i = 0; i < DPL; ++i) { qa[h][i] = qa_n[h][i]; qg[h][i] = qg_n[h][i]; }
            qe[h] = qe_n[h];
            sj[h] = 0.0f;
        }
        load_header(q + nwarps);
        const bool live_l = id_l >= 0;                  // padded neighbours are masked (embedding_module.py:154)
        const unsigned live_mask = __ballot_sync(0xffffffffu, live_l);
        const bool any = live_mask != 0u;
        // ---- phase A
        for (int j0 = 0; j0 < n; j0 += kUnroll) {
            int id[kUnroll];
            float dtj[kUnroll], xe[kUnroll], xh[kUnroll][DPL];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int j = j0 + u;
                id[u] = __shfl_sync(0xffffffffu, id_l, j & 31);
                dtj[u] = __shfl_sync(0xffffffffu, dt_l, j & 31);
                const int ei = __shfl_sync(0xffffffffu, ei_l, j & 31);
                if (j >= n) id[u] = -1;
                xe[u] = 0.0f;
#pragma unroll
                for (int i = 0; i < DPL; ++i) xh[u][i] = 0.0f;
                if (id[u] >= 0) {
                    ldg_cols<DPL>(pfo_row_ptr(Tb, id[u], ldtb), xh[u]);
                    if (lane < F) xe[u] = __ldg(reinterpret_cast<const float*>(pfo_row_ptr(efb, ei, Fb)));
                }
            }
            float part[NV];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
                for (int h = 0; h < NH; ++h) part[u * NH + h] = 0.0f;
                if (id[u] < 0) continue;
                float xt[DPL];
#pragma unroll
                for (int i = 0; i < DPL; ++i)
                    xt[i] = pfo_cosf_half(fmaf(dtj[u], tw[i], tb[i]));   // full-range, never __cosf (SURVEY hard part 1)
                float* st = stash + (j0 + u) * SW;
                st_cols<DPL>(st + c0, xh[u]);
                st_cols<DPL>(st + d + c0, xt);
                st[2 * d + lane] = xe[u];
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    float acc = qe[h] * xe[u];
#pragma unroll
                    for (int i = 0; i < DPL; ++i) acc = fmaf(qa[h][i], xh[u][i], fmaf(qg[h][i], xt[i], acc));
                    part[u * NH + h] = acc;
                }
            }
            const float r = multi_reduce<NV>(part, lane);          // total of (u, h) in lanes [(u*NH+h) * 32/NV, ..)
            const int u_l = (lane - j0) & (kUnroll - 1);
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const float s = __shfl_sync(0xffffffffu, r, (u_l * NH + h) * (32 / NV));
                if (lane >= j0 && lane < j0 + kUnroll) sj[h] = s;
            }
        }
        __syncwarp();
        // ---- phase B: softmax over the slots (lane = slot), then the dropout factor of torch's attention dropout
        float w[NH], psum[NH], keep[NH];
        keep_scales<NH>(p, step, q, lane, keep);
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const float m = warp_max(live_l ? sj[h] : -INFINITY);
            const float e = live_l ? expf(sj[h] - m) : 0.0f;
            const float l = warp_sum(e);
            const float pw = any ? e / l : 0.0f;
            if (lane < n) p.P[(q * NH + h) * n + lane] = pw;
            w[h] = live_l ? pw * keep[h] : 0.0f;
            psum[h] = warp_sum(w[h]);
        }
        // ---- phase C
        float ah[NH][DPL], at[NH][DPL], ae[NH];
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            ae[h] = 0.0f;
#pragma unroll
            for (int i = 0; i < DPL; ++i) { ah[h][i] = 0.0f; at[h][i] = 0.0f; }
        }
        for (unsigned mleft = live_mask; mleft; mleft &= mleft - 1) {
            const int j = __ffs(mleft) - 1;
            const float* st = stash + j * SW;
            float xh[DPL], xt[DPL];
            ld_cols<DPL>(st + c0, xh);
            ld_cols<DPL>(st + d + c0, xt);
            const float xe = st[2 * d + lane];
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const float wj = __shfl_sync(0xffffffffu, w[h], j);
                ae[h] = fmaf(wj, xe, ae[h]);
#pragma unroll
                for (int i = 0; i < DPL; ++i) { ah[h][i] = fmaf(wj, xh[i], ah[h][i]); at[h][i] = fmaf(wj, xt[i], at[h][i]); }
            }
        }
        __syncwarp();                                   // the stash is rewritten by the next query
        if (lane == 0) p.invalid[q] = any ? 0 : 1;     // rows with no neighbours: output zeroed (temporal_attention.py:84)
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            float* xb = p.XB + q * p.ldxb + (int64_t)h * ekp;
            st_cols<DPL>(xb + c0, ah[h]);
#pragma unroll
            for (int i = 0; i < DPL; ++i) xb[d + F + c0 + i] = at[h][i];
            if (lane < F) xb[d + lane] = ae[h];
            if (lane == 0) xb[2 * d + F] = psum[h];
            if (lane < ekp - (2 * d + F + 1))            // constant tail: [valid, one, 0...]
                xb[2 * d + F + 1 + lane] = lane == 0 ? (any ? 1.0f : 0.0f) : (lane == 1 ? 1.0f : 0.0f);
        }
    }
}

template <int DPL, int NH>
__global__ void __launch_bounds__(128)
attn_nbr_bwd_kernel(const NbrArgs p) {
    pfo_pdl_prologue();
    extern __shared__ float smem[];
    constexpr int d = 32 * DPL;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * wpb + wib;
    const int64_t nwarps = (int64_t)gridDim.x * wpb;
    const int F = p.F, n = p.n, ekp = p.ekp;
    constexpr int sw = 3 * d + 32;                      // stash row: [h | cos | sin | e(32)]
    float* stash = smem + (size_t)wib * n * sw;
    float* red = smem + (size_t)wpb * n * sw;           // [wpb][2][d] for the block reduction
    const int c0 = lane * DPL;
    // byte bases + 32-bit byte strides of the gathers and of the gradient scatter (see the forward kernel)
    const char* Tb = reinterpret_cast<const char*>(p.T) + (size_t)c0 * sizeof(float);
    const int ldtb = (int)(p.ldt * (int64_t)sizeof(float));
    const char* efb = reinterpret_cast<const char*>(p.efeat) + (size_t)lane * sizeof(float);
    const int Fb = p.F * (int)sizeof(float);
    char* dTb = reinterpret_cast<char*>(p.dT) + (size_t)c0 * sizeof(float);
    const int lddtb = (int)(p.lddt * (int64_t)sizeof(float));
    const uint32_t step = p.step + (p.step_dev ? *p.step_dev : 0u);
    float tw[DPL], tb[DPL], dwl[DPL], dbl[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) { tw[i] = p.tw[c0 + i]; tb[i] = p.tb[c0 + i]; dwl[i] = 0.f; dbl[i] = 0.f; }
    // per-query header (slot scalars, query operand, incoming gradient, softmax weights) loaded one query ahead,
    // like the forward kernel
    bool dead_n = true;
    int id_n = -1, ei_n = 0;
    float dt_n = 0.0f;
    float qa_n[NH][DPL], qg_n[NH][DPL], ga_n[NH][DPL], gg_n[NH][DPL], ge_n[NH], gp_n[NH], pj_n[NH];
    auto load_header = [&](int64_t q) {
        dead_n = true; id_n = -1; ei_n = 0; dt_n = 0.0f;
        if (q < p.Q) {
            dead_n = p.invalid[q] != 0;
            if (lane < n && !dead_n) {
                id_n = __ldg(p.idx + q * n + lane);
                dt_n = __ldg(p.dt + q * n + lane);
                ei_n = __ldg(p.eidx + q * n + lane);
            }
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const float* qk = p.QK + (q * NH + h) * ekp;
                const float* gx = p.dXB + q * p.lddxb + (int64_t)h * ekp;
                ld_cols<DPL>(qk + c0, qa_n[h]);
                ld_cols<DPL>(gx + c0, ga_n[h]);
#pragma unroll
                for (int i = 0; i < DPL; ++i) {
                    qg_n[h][i] = qk[d + F + c0 + i];
                    gg_n[h][i] = gx[d + F + c0 + i];
                }
                ge_n[h] = lane < F ? gx[d + lane] : 0.0f;
                gp_n[h] = gx[2 * d + F];
                pj_n[h] = (lane < n) ? p.P[(q * NH + h) * n + lane] : 0.0f;
            }
        }
    };
    load_header(warp);
    for (int64_t q = warp; q < p.Q; q += nwarps) {
        const bool dead = dead_n;
        const int id_l = id_n, ei_l = ei_n;
        const float dt_l = dt_n;
        float qa[NH][DPL], qg[NH][DPL];
        float ga[NH][DPL], gg[NH][DPL], ge[NH], gp[NH];
        float da[NH][DPL], dg[NH][DPL], de[NH];
        float pj[NH], pk[NH], dpj[NH];
#pragma unroll
        for (int h = 0; h < NH; ++h) {
#pragma unroll
            for (int i = 0; i < DPL; ++i) {
                qa[h][i] = qa_n[h][i]; ga[h][i] = ga_n[h][i]; qg[h][i] = qg_n[h][i]; gg[h][i] = gg_n[h][i];
                da[h][i] = 0.f; dg[h][i] = 0.f;
            }
            ge[h] = ge_n[h]; gp[h] = gp_n[h]; pj[h] = pj_n[h];
            de[h] = 0.f;
        }
        load_header(q + nwarps);
        keep_scales<NH>(p, step, q, lane, dpj);         // dpj holds the factor until pass A overwrites it with dp_hj
#pragma unroll
        for (int h = 0; h < NH; ++h) pk[h] = pj[h] * dpj[h];   // softmax weight after dropout (what multiplied x_j forward)
        if (!dead) {
            // pass A: rebuild x_j (kUnroll gathers in flight), stash it, dp_hj = keep_hj * (dXB_h . [x_j | 1])
            for (int j0 = 0; j0 < n; j0 += kUnroll) {
                int id[kUnroll];
                float dtj[kUnroll], xe[kUnroll], xh[kUnroll][DPL];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const int j = j0 + u;
                    id[u] = __shfl_sync(0xffffffffu, id_l, j & 31);
                    dtj[u] = __shfl_sync(0xffffffffu, dt_l, j & 31);
                    const int ei = __shfl_sync(0xffffffffu, ei_l, j & 31);
                    if (j >= n) id[u] = -1;
                    xe[u] = 0.0f;
#pragma unroll
                    for (int i = 0; i < DPL; ++i) xh[u][i] = 0.0f;
                    if (id[u] >= 0) {
                        ldg_cols<DPL>(pfo_row_ptr(Tb, id[u], ldtb), xh[u]);
                        if (lane < F) xe[u] = __ldg(reinterpret_cast<const float*>(pfo_row_ptr(efb, ei, Fb)));
                    }
                }
                float part[kUnroll * NH];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
                    for (int h = 0; h < NH; ++h) part[u * NH + h] = 0.0f;
                    if (id[u] < 0) continue;
                    const int j = j0 + u;
                    float* st = stash + j * sw;
                    float xt[DPL], sn[DPL];
#pragma unroll
                    for (int i = 0; i < DPL; ++i) pfo_sincosf(fmaf(dtj[u], tw[i], tb[i]), &sn[i], &xt[i]);
                    st_cols<DPL>(st + c0, xh[u]);
                    st_cols<DPL>(st + d + c0, xt);
                    st_cols<DPL>(st + 2 * d + c0, sn);
                    st[3 * d + lane] = xe[u];
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        float acc = ge[h] * xe[u];
#pragma unroll
                        for (int i = 0; i < DPL; ++i) acc = fmaf(ga[h][i], xh[u][i], fmaf(gg[h][i], xt[i], acc));
                        part[u * NH + h] = acc;
                    }
                }
                const float r = multi_reduce<kUnroll * NH>(part, lane);
                const int u_l = (lane - j0) & (kUnroll - 1);
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    const float dp = __shfl_sync(0xffffffffu, r, (u_l * NH + h) * (32 / (kUnroll * NH))) + gp[h];
                    if (lane >= j0 && lane < j0 + kUnroll) dpj[h] *= dp;
                }
            }
            __syncwarp();
            // softmax backward: ds_hj = p_hj (dp_hj - sum_j' p_hj' dp_hj')
            float dsj[NH];
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const float dpv = (id_l >= 0) ? dpj[h] : 0.0f;
                const float dot = warp_sum(pj[h] * dpv);
                dsj[h] = pj[h] * (dpv - dot);
            }
            // pass B: dqk_h += ds_hj x_j ; dx_j = sum_h (p'_hj dXB_h + ds_hj qk_h)
            for (int j = 0; j < n; ++j) {
                const int id = __shfl_sync(0xffffffffu, id_l, j);
                if (id < 0) continue;
                const float dtj = __shfl_sync(0xffffffffu, dt_l, j);
                const float* st = stash + j * sw;
                float xh[DPL], cs[DPL], sn[DPL], gxh[DPL], gxt[DPL];
                ld_cols<DPL>(st + c0, xh);
                ld_cols<DPL>(st + d + c0, cs);
                ld_cols<DPL>(st + 2 * d + c0, sn);
#pragma unroll
                for (int i = 0; i < DPL; ++i) { gxh[i] = 0.f; gxt[i] = 0.f; }
                const float xe = st[3 * d + lane];
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    const float ds = __shfl_sync(0xffffffffu, dsj[h], j);
                    const float pw = __shfl_sync(0xffffffffu, pk[h], j);
                    de[h] = fmaf(ds, xe, de[h]);
#pragma unroll
                    for (int i = 0; i < DPL; ++i) {
                        da[h][i] = fmaf(ds, xh[i], da[h][i]);
                        dg[h][i] = fmaf(ds, cs[i], dg[h][i]);
                        gxh[i] += pw * ga[h][i] + ds * qa[h][i];
                        gxt[i] += pw * gg[h][i] + ds * qg[h][i];
                    }
                }
                float* drow = reinterpret_cast<float*>(const_cast<char*>(pfo_row_ptr(dTb, id, lddtb)));
                if constexpr (DPL == 4) red_add_f32x4(drow, gxh[0], gxh[1], gxh[2], gxh[3]);
                else if constexpr (DPL == 2) red_add_f32x2(drow, gxh[0], gxh[1]);
                else {
#pragma unroll
                    for (int i = 0; i < DPL; ++i) atomicAdd(drow + i, gxh[i]);
                }
#pragma unroll
                for (int i = 0; i < DPL; ++i) {          // d cos(dt*w+b) = -sin(.) * (dt dw + db)
                    const float t = -sn[i] * gxt[i];
                    dwl[i] = fmaf(t, dtj, dwl[i]);
                    dbl[i] += t;
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            float* o = p.dQK + (q * NH + h) * ekp;
            st_cols<DPL>(o + c0, da[h]);
#pragma unroll
            for (int i = 0; i < DPL; ++i) o[d + F + c0 + i] = dg[h][i];
            if (lane < F) o[d + lane] = de[h];
            if (lane < ekp - (2 * d + F)) o[2 * d + F + lane] = 0.0f;
        }
    }
    // block reduction of the time-encoder gradients -> partial[block][2][d]
#pragma unroll
    for (int i = 0; i < DPL; ++i) { red[(wib * 2 + 0) * d + c0 + i] = dwl[i]; red[(wib * 2 + 1) * d + c0 + i] = dbl[i]; }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < wpb; ++w) s += red[(w * 2) * d + c];   // red is [wpb][2*d]
        p.partial[(int64_t)blockIdx.x * 2 * d + c] = s;
    }
}

// ---- BPR loss forward + backward (reference main.py:321-337) --------------------------
// one warp per interaction; loss_partial[block] and gradients w.r.t. the three embedding groups
__global__ void __launch_bounds__(256)
bpr_kernel(const float* __restrict__ eu, const float* __restrict__ ep, const float* __restrict__ en,
           int B, int k, int d, float* __restrict__ du, float* __restrict__ dp, float* __restrict__ dn,
           float* __restrict__ loss_partial, float grad_scale) {
    pfo_pdl_prologue();
    __shared__ float wl[8];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float lsum = 0.f;
    for (int64_t b = warp; b < B; b += nwarps) {
        const float* u = eu + b * d;
        const float* pp = ep + b * d;
        float pos = 0.f;
        for (int c = lane; c < d; c += 32) pos = fmaf(u[c], pp[c], pos);
        pos = warp_sum(pos);
        float diff = 0.f;
        for (int j = 0; j < k; ++j) {
            const float* nn = en + (b * k + j) * d;
            float s = 0.f;
            for (int c = lane; c < d; c += 32) s = fmaf(u[c], nn[c], s);
            diff += pos - warp_sum(s);
        }
        const float x = diff / (float)k;
        const float sg = sigmoidf_(x);
        lsum += -logf(sg);
        if (du) {
            const float gx = -(1.0f - sg) * grad_scale / (float)B;     // d(-log sigmoid(x))/dx / B
            for (int c = lane; c < d; c += 32) {
                float nsum = 0.f;
                for (int j = 0; j < k; ++j) {
                    nsum += en[(b * k + j) * d + c];
                    dn[(b * k + j) * d + c] = -gx / (float)k * u[c];
                }
                du[b * d + c] = gx * (pp[c] - nsum / (float)k);
                dp[b * d + c] = gx * u[c];
            }
        }
    }
    if (lane == 0) wl[wib] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += wl[w];
        loss_partial[blockIdx.x] = s / (float)B;
    }
}

// ---- evaluation scoring + ranking (reference evaluation.py:107-115,134-138) -----------
// one CTA per interaction: scores[0] = <src,dst>, scores[1+c] = <src,cand_c>; the rank of the
// positive under argsort(scores)[::-1] with a stable sort, and the top-k candidate positions.
__global__ void __launch_bounds__(256)
eval_score_kernel(const float* __restrict__ es, const float* __restrict__ ed, const float* __restrict__ ec,
                  int n_cand, int d, int topk, float* __restrict__ scores, int32_t* __restrict__ pos_rank,
                  int32_t* __restrict__ top_idx) {
    pfo_pdl_prologue();
    extern __shared__ float sh[];            // [1 + n_cand] scores, then reduction scratch
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const float* u = es + (int64_t)b * d;
    const int total = 1 + n_cand;
    for (int c = wib; c < total; c += wpb) {
        const float* v = c == 0 ? ed + (int64_t)b * d : ec + ((int64_t)b * n_cand + (c - 1)) * d;
        float s = 0.f;
        for (int k = lane; k < d; k += 32) s += u[k] * v[k];    // torch.sum(a * b): mul then add
        s = warp_sum(s);
        if (lane == 0) { sh[c] = s; scores[(int64_t)b * total + c] = s; }
    }
    __syncthreads();
    // reversed stable ascending order: among equal scores the larger index comes first
    __shared__ int cnt;
    __shared__ float best_v[8];
    __shared__ int best_i[8];
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const float s0 = sh[0];
    int local = 0;
    for (int c = 1 + threadIdx.x; c < total; c += blockDim.x) local += (sh[c] >= s0) ? 1 : 0;
    local = warp_sum_i(local);
    if (lane == 0) atomicAdd(&cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) pos_rank[b] = cnt;
    float prev_v = INFINITY;
    int prev_i = 0x7fffffff;
    for (int r = 0; r < topk; ++r) {
        float bv = -INFINITY; int bi = -1;
        for (int c = threadIdx.x; c < total; c += blockDim.x) {
            const float v = sh[c];
            const bool after_prev = v < prev_v || (v == prev_v && c < prev_i);
            if (after_prev && (v > bv || (v == bv && c > bi))) { bv = v; bi = c; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi > bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { best_v[wib] = bv; best_i[wib] = bi; }
        __syncthreads();
        bv = best_v[0]; bi = best_i[0];
        for (int w = 1; w < wpb; ++w)
            if (best_v[w] > bv || (best_v[w] == bv && best_i[w] > bi)) { bv = best_v[w]; bi = best_i[w]; }
        if (threadIdx.x == 0) top_idx[(int64_t)b * topk + r] = bi;
        prev_v = bv; prev_i = bi;
        __syncthreads();
    }
}

template <int DPL, int NH>
int launch_fwd(const NbrArgs& a, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(attn_nbr_fwd_kernel<DPL, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    const size_t smem = (size_t)4 * a.n * (2 * a.d + 32) * sizeof(float);      // 4 warps x n stash rows
    static const int env_waves = [] { const char* e = getenv("PFO_ATTN_FWD_CTAS"); return e ? atoi(e) : 0; }();
    // three resident waves (the occupancy API says 6 CTAs per SM at d = 64, 2 heads).  Measured at bs 8192, CTAs per
    // SM -> step: 5 -> 0.863 ms, 6 -> 0.838, 10 -> 0.837, 12 -> 0.831 / 0.829, 18 -> 0.822 (finer slices of the query
    // list even out the tail; profiles/r2_knob_sweeps.txt); PFO_ATTN_FWD_CTAS overrides the CTAs per SM
    const int per_sm = env_waves > 0 ? env_waves : 3 * pfo_resident(attn_nbr_fwd_kernel<DPL, NH>, 128, smem);
    pfo_launch(attn_nbr_fwd_kernel<DPL, NH>, pfo_grid(a.Q * 32, 128, per_sm), 128, smem, s, a);
    PFO_LAUNCH_CHECK();
}

template <int DPL, int NH>
int launch_bwd(const NbrArgs& a, int& grid, size_t smem, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(attn_nbr_bwd_kernel<DPL, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    static const int env_ctas = [] { const char* e = getenv("PFO_ATTN_BWD_CTAS"); return e ? atoi(e) : 0; }();
    const int per_sm = env_ctas > 0 ? env_ctas : pfo_resident(attn_nbr_bwd_kernel<DPL, NH>, 128, smem);
    const int cap = pfo_num_sms() * per_sm;          // one resident wave; `grid` is the caller's bound (workspace rows)
    if (grid > cap) grid = cap;
    pfo_launch(attn_nbr_bwd_kernel<DPL, NH>, grid, 128, smem, s, a);
    PFO_LAUNCH_CHECK();
}

// (d / 32, heads) -> kernel instance; both are compile-time so the per-head state lives in registers
template <int DPL>
int dispatch_fwd(const NbrArgs& a, cudaStream_t s) {
    switch (a.H) {
        case 1: return launch_fwd<DPL, 1>(a, s);
        case 2: return launch_fwd<DPL, 2>(a, s);
        case 4: return launch_fwd<DPL, 4>(a, s);
        default: return (int)cudaErrorInvalidValue;
    }
}
template <int DPL>
int dispatch_bwd(const NbrArgs& a, int& grid, size_t smem, cudaStream_t s) {
    switch (a.H) {
        case 1: return launch_bwd<DPL, 1>(a, grid, smem, s);
        case 2: return launch_bwd<DPL, 2>(a, grid, smem, s);
        case 4: return launch_bwd<DPL, 4>(a, grid, smem, s);
        default: return (int)cudaErrorInvalidValue;
    }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

PFO_API int pfo_attn_nbr_fwd(const float* QK, const float* T, int64_t ldt, const int32_t* idx, const int32_t* eidx,
                             const float* dt, const float* efeat, const float* tw, const float* tb,
                             int64_t Q, int n, int d, int F, int H, int ekp,
                             float p_drop, uint64_t seed, uint32_t step, const uint32_t* step_dev,
                             float* XB, int64_t ldxb, float* P, int32_t* invalid, void* stream) {
    return pfo_attn_nbr_fwd_rows(QK, nullptr, T, ldt, idx, eidx, dt, efeat, tw, tb, Q, n, d, F, H, ekp, p_drop, seed, step,
                                 step_dev, XB, ldxb, P, invalid, stream);
}

PFO_API int pfo_attn_nbr_fwd_rows(const float* QK, const int32_t* qk_row, const float* T, int64_t ldt, const int32_t* idx,
                                  const int32_t* eidx, const float* dt, const float* efeat, const float* tw, const float* tb,
                                  int64_t Q, int n, int d, int F, int H, int ekp,
                                  float p_drop, uint64_t seed, uint32_t step, const uint32_t* step_dev,
                                  float* XB, int64_t ldxb, float* P, int32_t* invalid, void* stream) {
    if (Q <= 0) return 0;
    if (d % 32 != 0 || d > 128 || F > 32 || H > kMaxHeads || n > 32 || n < 1 || ekp < 2 * d + F + 3 ||
        ldxb < (int64_t)H * ekp)
        return (int)cudaErrorInvalidValue;
    NbrArgs a{};
    a.QK = QK; a.qk_row = qk_row; a.T = T; a.ldt = ldt; a.idx = idx; a.eidx = eidx; a.dt = dt; a.efeat = efeat; a.tw = tw; a.tb = tb;
    a.Q = Q; a.n = n; a.d = d; a.F = F; a.H = H; a.ekp = ekp; a.p_drop = p_drop;
    a.k0 = (uint32_t)(seed & 0xffffffffu); a.k1 = (uint32_t)(seed >> 32); a.step = step; a.step_dev = step_dev;
    a.XB = XB; a.ldxb = ldxb; a.P = P; a.invalid = invalid;
    cudaStream_t s = (cudaStream_t)stream;
    if (ldt % 4 != 0 || ldxb % 4 != 0 || ekp % 4 != 0 || !aligned16(T) || !aligned16(QK) || !aligned16(XB) ||
        ldt > kMaxRowFloats)
        return (int)cudaErrorInvalidValue;             // rows are read / written with 8- and 16-byte accesses
    switch (d / 32) {
        case 1: return dispatch_fwd<1>(a, s);
        case 2: return dispatch_fwd<2>(a, s);
        case 3: return dispatch_fwd<3>(a, s);
        default: return dispatch_fwd<4>(a, s);
    }
}

PFO_API int64_t pfo_attn_nbr_bwd_workspace_floats(int d) { return (int64_t)2 * 148 * 4 * 2 * d; }

PFO_API int pfo_attn_nbr_bwd(const float* QK, const float* dXB, int64_t lddxb, const float* P, const int32_t* invalid,
                             const float* T, int64_t ldt, const int32_t* idx, const int32_t* eidx, const float* dt,
                             const float* efeat, const float* tw, const float* tb,
                             int64_t Q, int n, int d, int F, int H, int ekp,
                             float p_drop, uint64_t seed, uint32_t step, const uint32_t* step_dev,
                             float* dQK, float* dT, int64_t lddt, float* dtw_dtb, int accumulate,
                             float* workspace, void* stream) {
    if (Q <= 0) return 0;
    if (d % 32 != 0 || d > 128 || F > 32 || H > kMaxHeads || n > 32 || n < 1) return (int)cudaErrorInvalidValue;
    NbrArgs a{};
    a.QK = QK; a.T = T; a.ldt = ldt; a.idx = idx; a.eidx = eidx; a.dt = dt; a.efeat = efeat; a.tw = tw; a.tb = tb;
    a.Q = Q; a.n = n; a.d = d; a.F = F; a.H = H; a.ekp = ekp; a.p_drop = p_drop;
    a.k0 = (uint32_t)(seed & 0xffffffffu); a.k1 = (uint32_t)(seed >> 32); a.step = step; a.step_dev = step_dev;
    a.P = const_cast<float*>(P); a.invalid = const_cast<int32_t*>(invalid);
    a.dXB = dXB; a.lddxb = lddxb; a.dQK = dQK; a.dT = dT; a.lddt = lddt; a.partial = workspace;
    cudaStream_t s = (cudaStream_t)stream;
    const int wpb = 4;
    const size_t smem = ((size_t)wpb * n * (3 * d + 32) + (size_t)wpb * 2 * d) * sizeof(float);
    int grid = pfo_grid(Q * 32, 128, 8);             // upper bound: the launcher caps it at one resident wave
    const int max_grid = 2 * 148 * 4;                // rows of the partial workspace
    if (grid > max_grid) grid = max_grid;
    if (ldt % 4 != 0 || lddxb % 4 != 0 || lddt % 4 != 0 || ekp % 4 != 0 || !aligned16(T) || !aligned16(QK) ||
        !aligned16(dXB) || !aligned16(dQK) || !aligned16(dT) || ldt > kMaxRowFloats || lddt > kMaxRowFloats)
        return (int)cudaErrorInvalidValue;
    int rc;
    switch (d / 32) {
        case 1: rc = dispatch_bwd<1>(a, grid, smem, s); break;
        case 2: rc = dispatch_bwd<2>(a, grid, smem, s); break;
        case 3: rc = dispatch_bwd<3>(a, grid, smem, s); break;
        default: rc = dispatch_bwd<4>(a, grid, smem, s); break;
    }
    if (rc) return rc;
    return pfo_reduce_partials(workspace, grid, 2 * d, dtw_dtb, accumulate, stream);
}

PFO_API int pfo_bpr(const float* eu, const float* ep, const float* en, int B, int k, int d,
                    float* du, float* dp, float* dn, float* loss, float grad_scale, float* workspace, void* stream) {
    if (B <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    int grid = pfo_grid((int64_t)B * 32, 256, 8);   // one warp per interaction when they fit (latency-bound, tiny; one
                                                    // resident wave -- 5 CTAs per SM -- measured slower: 27 against 21 us)
    if (grid > 1024) grid = 1024;
    pfo_launch(bpr_kernel, grid, 256, 0, s, eu, ep, en, B, k, d, du, dp, dn, workspace, grad_scale);
    // loss = sum over blocks of per-block means/B contributions: reduce rows=grid, cols=1
    return pfo_reduce_partials(workspace, grid, 1, loss, 0, stream);
}

PFO_API int pfo_eval_score(const float* es, const float* ed, const float* ec, int B, int n_cand, int d, int topk,
                           float* scores, int32_t* pos_rank, int32_t* top_idx, void* stream) {
    if (B <= 0) return 0;
    const size_t smem = (size_t)(1 + n_cand) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(eval_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    pfo_launch(eval_score_kernel, B, 256, smem, (cudaStream_t)stream, es, ed, ec, n_cand, d, topk, scores, pos_rank, top_idx);
    PFO_LAUNCH_CHECK();
}
