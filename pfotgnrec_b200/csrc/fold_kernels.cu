// Parameter-side folding of the temporal-attention layer (reference model/temporal_attention.py:26-90 and
// utils/utils.py:4-17): the ten tensors of nn.MultiheadAttention + MergeLayer become the three GEMM operands
// the per-query kernels consume, and the gradients of those operands are carried back to the ten tensors.
//
//   Wqk  [H*ekp, d], cqk [H*ekp] : qk_h = A_h h_q + cA_h,  A_h = s Wk_h^T Wq_h[:, :d],  cA_h = s Wk_h^T cq_h,
//                                  cq = Wq[:, d:] cos(tb) + bq  (TimeEncode(0), embedding_module.py:92), s = 1/sqrt(hd)
//   Wc1T [H*ekp + d, d]          : rows of head h = (W1a Wo_h [Wv_h | bv_h])^T, row `valid` = W1a bo and row
//                                  `one` = b1 in head 0, zero padding rows, then W1b^T
// Everything here is O(weights) -- a few MFLOP per step -- so the arithmetic is done in fp64 (the folded operands
// are then at least as accurate as the reference's chained fp32 products) with one thread per output element and
// the thread index laid along the operand that makes the inner-loop loads coalesce.  Two launches forward, two
// backward, instead of ~95 tiny framework kernels per step.
#include "common.cuh"

namespace {

struct FoldArgs {
    const float *Wq, *Wk, *Wv, *b_in, *Wo, *bo, *W1, *b1, *tb;
    int d, F, H, ekp, E, Ek, hd;
    double scale;
    double* ws;               // [te0 (d) | cq (E) | T (H, E, Ek+1) | gcq (E) | gT (H, Ek+1, E)]
    float *Wqk, *cqk, *Wc1T;  // forward outputs
    const float *gWqk, *gcqk, *gWc1T;   // backward inputs
    float *gWq, *gWk, *gWv, *gb_in, *gWo, *gbo, *gW1, *gb1, *gtb;
};

__device__ __forceinline__ double* ws_te0(const FoldArgs& p) { return p.ws; }
__device__ __forceinline__ double* ws_cq(const FoldArgs& p) { return p.ws + p.d; }
__device__ __forceinline__ double* ws_T(const FoldArgs& p) { return p.ws + p.d + p.E; }
__device__ __forceinline__ double* ws_gcq(const FoldArgs& p) { return ws_T(p) + (size_t)p.H * p.E * (p.Ek + 1); }
__device__ __forceinline__ double* ws_gT(const FoldArgs& p) { return ws_gcq(p) + p.E; }

// [Wv_h | bv_h](i, r)
__device__ __forceinline__ double wva(const FoldArgs& p, int h, int i, int r) {
    const int row = h * p.hd + i;
    return r < p.Ek ? (double)p.Wv[(size_t)row * p.Ek + r] : (double)p.b_in[2 * p.E + row];
}

// ---- forward stage 1: te0, cq | T_h = Wo_h [Wv_h | bv_h] | Wqk = s Wk_h^T Wq_h[:, :d] (padding rows zero)
__global__ void __launch_bounds__(256) fold_fwd1_kernel(const FoldArgs p) {
    const int d = p.d, E = p.E, Ek = p.Ek, H = p.H, hd = p.hd, ekp = p.ekp, R = Ek + 1;
    const int n_cq = E, n_T = H * E * R, n_A = H * ekp * d;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cq + n_T + n_A; i += gridDim.x * blockDim.x) {
        if (i < n_cq) {
            const int e = i;
            double s = (double)p.b_in[e];
            for (int k = 0; k < d; ++k) s += (double)p.Wq[(size_t)e * E + d + k] * cos((double)p.tb[k]);
            ws_cq(p)[e] = s;
            if (e < d) ws_te0(p)[e] = cos((double)p.tb[e]);
        } else if (i < n_cq + n_T) {
            const int j = i - n_cq;
            const int r = j % R, e = (j / R) % E, h = j / (R * E);
            double s = 0.0;
            for (int t = 0; t < hd; ++t) s += (double)p.Wo[(size_t)e * E + h * hd + t] * wva(p, h, t, r);
            ws_T(p)[j] = s;
        } else {
            const int j = i - n_cq - n_T;
            const int c = j % d, r = (j / d) % ekp, h = j / (d * ekp);
            double s = 0.0;
            if (r < Ek)
                for (int t = 0; t < hd; ++t)
                    s += (double)p.Wk[(size_t)(h * hd + t) * Ek + r] * (double)p.Wq[(size_t)(h * hd + t) * E + c];
            p.Wqk[j] = (float)(p.scale * s);
        }
    }
}

// ---- forward stage 2: cqk = s Wk_h^T cq_h | Wc1T
__global__ void __launch_bounds__(256) fold_fwd2_kernel(const FoldArgs p) {
    const int d = p.d, E = p.E, Ek = p.Ek, H = p.H, hd = p.hd, ekp = p.ekp, R = Ek + 1;
    const int n_c = H * ekp, n_W = (H * ekp + d) * d;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_c + n_W; i += gridDim.x * blockDim.x) {
        if (i < n_c) {
            const int r = i % ekp, h = i / ekp;
            double s = 0.0;
            if (r < Ek)
                for (int t = 0; t < hd; ++t) s += (double)p.Wk[(size_t)(h * hd + t) * Ek + r] * ws_cq(p)[h * hd + t];
            p.cqk[i] = (float)(p.scale * s);
        } else {
            // thread index along the folded row index (rr) so that T_h[e, r] loads coalesce; column c is the slow index
            const int j = i - n_c;
            const int rows = H * ekp + d;
            const int rr = j % rows, c = j / rows;
            double s = 0.0;
            if (rr >= H * ekp) {
                s = (double)p.W1[(size_t)c * (E + d) + E + (rr - H * ekp)];               // W1b^T
            } else {
                const int h = rr / ekp, r = rr % ekp;
                if (r < R) {
                    const double* T = ws_T(p) + (size_t)h * E * R;
                    for (int e = 0; e < E; ++e) s += (double)p.W1[(size_t)c * (E + d) + e] * T[(size_t)e * R + r];
                } else if (h == 0 && r == R) {                                           // `valid` row: W1a bo
                    for (int e = 0; e < E; ++e) s += (double)p.W1[(size_t)c * (E + d) + e] * (double)p.bo[e];
                } else if (h == 0 && r == R + 1) {                                       // `one` row: b1
                    s = (double)p.b1[c];
                }
            }
            p.Wc1T[(size_t)rr * d + c] = (float)s;
        }
    }
}

// ---- backward stage 1 (needs T and cq of the forward):
//   gW1 (a and b parts), gb1 | gT_h = W1a^T gB_h (layout [h][r][e]) | gbo | gWq[:, :d] | gcq | gWk
__global__ void __launch_bounds__(256) fold_bwd1_kernel(const FoldArgs p) {
    const int d = p.d, E = p.E, Ek = p.Ek, H = p.H, hd = p.hd, ekp = p.ekp, R = Ek + 1;
    const int n_W1 = d * (E + d), n_b1 = d, n_gT = H * R * E, n_bo = E, n_Wq = E * d, n_cq = E, n_Wk = E * Ek;
    const int o1 = n_W1, o2 = o1 + n_b1, o3 = o2 + n_gT, o4 = o3 + n_bo, o5 = o4 + n_Wq, o6 = o5 + n_cq, o7 = o6 + n_Wk;
    const float* gc1 = p.gWc1T + (size_t)R * d;                   // head 0, `valid` row
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < o7; i += gridDim.x * blockDim.x) {
        if (i < o1) {                                              // gW1[c, x]; thread index along c
            const int c = i % d, x = i / d;
            double s = 0.0;
            if (x >= E) {
                s = (double)p.gWc1T[(size_t)(H * ekp + (x - E)) * d + c];
            } else {
                for (int h = 0; h < H; ++h) {
                    const double* T = ws_T(p) + ((size_t)h * E + x) * R;
                    const float* g = p.gWc1T + (size_t)h * ekp * d + c;
                    for (int r = 0; r < R; ++r) s += (double)g[(size_t)r * d] * T[r];
                }
                s += (double)gc1[c] * (double)p.bo[x];
            }
            p.gW1[(size_t)c * (E + d) + x] = (float)s;
        } else if (i < o2) {
            const int c = i - o1;
            p.gb1[c] = p.gWc1T[(size_t)(R + 1) * d + c];           // head 0, `one` row
        } else if (i < o3) {                                       // gT[h][r][e]; thread index along e
            const int j = i - o2;
            const int e = j % E, r = (j / E) % R, h = j / (E * R);
            const float* g = p.gWc1T + (size_t)(h * ekp + r) * d;
            double s = 0.0;
            for (int c = 0; c < d; ++c) s += (double)p.W1[(size_t)c * (E + d) + e] * (double)g[c];
            ws_gT(p)[j] = s;
        } else if (i < o4) {
            const int e = i - o3;
            double s = 0.0;
            for (int c = 0; c < d; ++c) s += (double)p.W1[(size_t)c * (E + d) + e] * (double)gc1[c];
            p.gbo[e] = (float)s;
        } else if (i < o5) {                                       // gWq[row, c < d]; thread index along c
            const int j = i - o4;
            const int c = j % d, row = j / d, h = row / hd;
            const float* g = p.gWqk + (size_t)h * ekp * d + c;
            double s = 0.0;
            for (int r = 0; r < Ek; ++r) s += (double)p.Wk[(size_t)row * Ek + r] * (double)g[(size_t)r * d];
            p.gWq[(size_t)row * E + c] = (float)(p.scale * s);
        } else if (i < o6) {
            const int row = i - o5, h = row / hd;
            double s = 0.0;
            for (int r = 0; r < Ek; ++r) s += (double)p.Wk[(size_t)row * Ek + r] * (double)p.gcqk[h * ekp + r];
            ws_gcq(p)[row] = p.scale * s;
        } else {                                                   // gWk[row, r]
            const int j = i - o6;
            const int r = j % Ek, row = j / Ek, h = row / hd;
            const float* g = p.gWqk + (size_t)(h * ekp + r) * d;
            const float* wq = p.Wq + (size_t)row * E;
            double s = 0.0;
            for (int c = 0; c < d; ++c) s += (double)g[c] * (double)wq[c];
            s += (double)p.gcqk[h * ekp + r] * ws_cq(p)[row];
            p.gWk[j] = (float)(p.scale * s);
        }
    }
}

// ---- backward stage 2 (needs gT and gcq): gWo | gWv, gb_in | gWq[:, d:] | gtb
__global__ void __launch_bounds__(256) fold_bwd2_kernel(const FoldArgs p) {
    const int d = p.d, E = p.E, Ek = p.Ek, H = p.H, hd = p.hd, R = Ek + 1;
    const int n_Wo = E * E, n_Wv = E * R, n_bq = 2 * E, n_Wqt = E * d, n_tb = d;
    const int o1 = n_Wo, o2 = o1 + n_Wv, o3 = o2 + n_bq, o4 = o3 + n_Wqt, o5 = o4 + n_tb;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < o5; i += gridDim.x * blockDim.x) {
        if (i < o1) {                                              // gWo[e, col]; thread index along e
            const int e = i % E, col = i / E, h = col / hd, t = col % hd;
            const double* gT = ws_gT(p) + (size_t)h * R * E + e;
            double s = 0.0;
            for (int r = 0; r < R; ++r) s += gT[(size_t)r * E] * wva(p, h, t, r);
            p.gWo[(size_t)e * E + col] = (float)s;
        } else if (i < o2) {                                       // g[Wv | bv][row, r]; thread index along row's t
            const int j = i - o1;
            const int t = j % hd, r = (j / hd) % R, h = j / (hd * R);
            const double* gT = ws_gT(p) + ((size_t)h * R + r) * E;
            double s = 0.0;
            for (int e = 0; e < E; ++e) s += (double)p.Wo[(size_t)e * E + h * hd + t] * gT[e];
            const int row = h * hd + t;
            if (r < Ek) p.gWv[(size_t)row * Ek + r] = (float)s;
            else p.gb_in[2 * E + row] = (float)s;
        } else if (i < o3) {                                       // gb_in: query part = gcq, key part = 0 (cancels in softmax)
            const int e = i - o2;
            p.gb_in[e] = e < E ? (float)ws_gcq(p)[e] : 0.0f;
        } else if (i < o4) {                                       // gWq[e, d + k] = gcq[e] te0[k]
            const int j = i - o3;
            const int k = j % d, e = j / d;
            p.gWq[(size_t)e * E + d + k] = (float)(ws_gcq(p)[e] * ws_te0(p)[k]);
        } else {                                                   // gtb[k] = -sin(tb[k]) sum_e Wq[e, d+k] gcq[e]
            const int k = i - o4;
            double s = 0.0;
            for (int e = 0; e < E; ++e) s += (double)p.Wq[(size_t)e * E + d + k] * ws_gcq(p)[e];
            p.gtb[k] = (float)(-sin((double)p.tb[k]) * s);
        }
    }
}

int fill(FoldArgs& a, const float* Wq, const float* Wk, const float* Wv, const float* b_in, const float* Wo,
         const float* bo, const float* W1, const float* b1, const float* tb, int d, int F, int H, int ekp, double* ws) {
    if (d <= 0 || F < 0 || H <= 0 || (2 * d) % H != 0 || ekp < 2 * d + F + 3 || ws == nullptr) return 1;
    a.Wq = Wq; a.Wk = Wk; a.Wv = Wv; a.b_in = b_in; a.Wo = Wo; a.bo = bo; a.W1 = W1; a.b1 = b1; a.tb = tb;
    a.d = d; a.F = F; a.H = H; a.ekp = ekp; a.E = 2 * d; a.Ek = 2 * d + F; a.hd = 2 * d / H;
    a.scale = 1.0 / sqrt((double)a.hd);
    a.ws = ws;
    return 0;
}

int blocks_for(int64_t n) { int64_t b = (n + 255) / 256; return (int)(b < 1 ? 1 : (b > 4096 ? 4096 : b)); }

}  // namespace

PFO_API int64_t pfo_fold_attention_workspace_doubles(int d, int F, int H) {
    const int64_t E = 2 * d, R = 2 * d + F + 1;
    return d + E + H * E * R + E + H * R * E;
}

PFO_API int pfo_fold_attention_fwd(const float* Wq, const float* Wk, const float* Wv, const float* b_in, const float* Wo,
                                   const float* bo, const float* W1, const float* b1, const float* tb,
                                   int d, int F, int H, int ekp, double* workspace,
                                   float* Wqk, float* cqk, float* Wc1T, void* stream) {
    FoldArgs a{};
    if (fill(a, Wq, Wk, Wv, b_in, Wo, bo, W1, b1, tb, d, F, H, ekp, workspace)) return (int)cudaErrorInvalidValue;
    a.Wqk = Wqk; a.cqk = cqk; a.Wc1T = Wc1T;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n1 = a.E + (int64_t)H * a.E * (a.Ek + 1) + (int64_t)H * ekp * d;
    fold_fwd1_kernel<<<blocks_for(n1), 256, 0, s>>>(a);
    const int64_t n2 = (int64_t)H * ekp + (int64_t)(H * ekp + d) * d;
    fold_fwd2_kernel<<<blocks_for(n2), 256, 0, s>>>(a);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_fold_attention_bwd(const float* Wq, const float* Wk, const float* Wv, const float* b_in, const float* Wo,
                                   const float* bo, const float* W1, const float* b1, const float* tb,
                                   int d, int F, int H, int ekp, double* workspace,
                                   const float* gWqk, const float* gcqk, const float* gWc1T,
                                   float* gWq, float* gWk, float* gWv, float* gb_in, float* gWo, float* gbo,
                                   float* gW1, float* gb1, float* gtb, void* stream) {
    FoldArgs a{};
    if (fill(a, Wq, Wk, Wv, b_in, Wo, bo, W1, b1, tb, d, F, H, ekp, workspace)) return (int)cudaErrorInvalidValue;
    a.gWqk = gWqk; a.gcqk = gcqk; a.gWc1T = gWc1T;
    a.gWq = gWq; a.gWk = gWk; a.gWv = gWv; a.gb_in = gb_in; a.gWo = gWo; a.gbo = gbo; a.gW1 = gW1; a.gb1 = gb1; a.gtb = gtb;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t E = a.E, R = a.Ek + 1;
    const int64_t n1 = (int64_t)d * (E + d) + d + H * R * E + E + E * d + E + E * a.Ek;
    fold_bwd1_kernel<<<blocks_for(n1), 256, 0, s>>>(a);
    const int64_t n2 = E * E + E * R + 2 * E + E * d + d;
    fold_bwd2_kernel<<<blocks_for(n2), 256, 0, s>>>(a);
    PFO_LAUNCH_CHECK();
}
