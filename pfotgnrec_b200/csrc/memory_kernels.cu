// K2/K3 glue: the dense pending-message table, the lazy memory updater gates, persist,
// and message construction with last-wins scatter.
//
// Replaces the reference's dict-of-lists message store and its per-node Python loops:
//   modules/memory.py:35-37,73-75 (store / clear), modules/message_aggregator.py:38-55
//   (`last` aggregator), modules/memory_updater.py:18-53 (GRU/RNN update, functional and
//   in place) and model/tgn.py:357-378 (get_raw_messages).
// State layout in HBM (all fp32 unless noted), N = n_nodes:
//   memory[N, d], last_update[N], pend_msg[N, RAWP], pend_ts[N], pend_valid[N] (u8),
//   last_pos[N] (i32 scratch, -1 when idle).
// RAW = 2d + F + d floats per message, rows padded to RAWP (multiple of 4).
#include "common.cuh"
#include "pfo_math.cuh"

namespace {

// ---- gates: cell outputs for the unique touched nodes -------------------------------
// cell: 0 = GRU (GI/GH hold [r | z | n] pre-activations, 3d wide), 1 = RNN (d wide),
//       2 = no memory (H0 = node features only).
__global__ void __launch_bounds__(256)
cell_forward_kernel(const int32_t* __restrict__ uniq, const int32_t* __restrict__ n_uniq, int64_t u_max, int d,
                    int cell, int merged, const float* __restrict__ GI, const float* __restrict__ GH, int64_t ldg,
                    const float* __restrict__ HG, const uint8_t* __restrict__ valid_u,
                    const float* __restrict__ node_feat, float* __restrict__ Hnew, float* __restrict__ H0) {
    pfo_pdl_prologue();
    int64_t U = *n_uniq; if (U > u_max) U = u_max;
    const int64_t total = U * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = i / d;
        const int c = (int)(i - u * d);
        const int node = uniq[u];
        const float nf = node_feat[(int64_t)node * d + c];
        if (cell == 2) { H0[i] = nf; continue; }
        const float h = HG[i];
        float hn = h;
        if (valid_u[u]) {
            const float* gi = GI + u * ldg;
            if (merged) {
                // one contraction over [message | memory]: G4 = [r | z | n_i | n_h] with the r, z sums already formed
                // (GRU, 4d wide) or the single pre-activation (RNN, d wide)
                if (cell == 0) {
                    const float r = sigmoidf_(gi[c]);
                    const float z = sigmoidf_(gi[d + c]);
                    const float n = tanhf(gi[2 * d + c] + r * gi[3 * d + c]);
                    hn = n + z * (h - n);
                } else {
                    hn = tanhf(gi[c]);
                }
            } else {
                const float* gh = GH + u * ldg;
                if (cell == 0) {
                    // torch GRUCell: r,z = sigmoid(i + h); n = tanh(i_n + r * h_n); h' = n + z * (h - n)
                    const float r = sigmoidf_(gi[c] + gh[c]);
                    const float z = sigmoidf_(gi[d + c] + gh[d + c]);
                    const float n = tanhf(gi[2 * d + c] + r * gh[2 * d + c]);
                    hn = n + z * (h - n);
                } else {
                    hn = tanhf(gi[c] + gh[c]);
                }
            }
        }
        Hnew[i] = hn;
        H0[i] = hn + nf;          // h0 = memory' + node features (embedding_module.py:93-98)
    }
}

// backward of the cell w.r.t. its pre-activations (memory and messages are detached inputs)
__global__ void __launch_bounds__(256)
cell_backward_kernel(const int32_t* __restrict__ uniq, const int32_t* __restrict__ n_uniq, int64_t u_max, int d,
                     int cell, int merged, const float* __restrict__ GI, const float* __restrict__ GH, int64_t ldg,
                     const float* __restrict__ HG, const uint8_t* __restrict__ valid_u,
                     const float* __restrict__ dH, float* __restrict__ dGI, float* __restrict__ dGH) {
    pfo_pdl_prologue();
    int64_t U = *n_uniq; if (U > u_max) U = u_max;
    const int64_t total = U * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = i / d;
        const int c = (int)(i - u * d);
        float* dgi = dGI + u * ldg;
        const bool live = valid_u[u] != 0;
        const float g = dH[i];
        if (merged) {
            if (cell == 0) {
                float dr = 0.f, dz = 0.f, dn = 0.f, dhn = 0.f;
                if (live) {
                    const float* gi = GI + u * ldg;
                    const float h = HG[i];
                    const float r = sigmoidf_(gi[c]);
                    const float z = sigmoidf_(gi[d + c]);
                    const float hn_ = gi[3 * d + c];
                    const float n = tanhf(gi[2 * d + c] + r * hn_);
                    const float dn_ = g * (1.0f - z);
                    dn = dn_ * (1.0f - n * n);
                    dz = g * (h - n) * z * (1.0f - z);
                    dr = dn * hn_ * r * (1.0f - r);
                    dhn = dn * r;
                }
                dgi[c] = dr; dgi[d + c] = dz; dgi[2 * d + c] = dn; dgi[3 * d + c] = dhn;
            } else {
                float dp = 0.f;
                if (live) {
                    const float y = tanhf(GI[u * ldg + c]);
                    dp = g * (1.0f - y * y);
                }
                dgi[c] = dp;
            }
            continue;
        }
        float* dgh = dGH + u * ldg;
        if (cell == 0) {
            float dr = 0.f, dz = 0.f, dn = 0.f, dhn = 0.f;
            if (live) {
                const float* gi = GI + u * ldg;
                const float* gh = GH + u * ldg;
                const float h = HG[i];
                const float r = sigmoidf_(gi[c] + gh[c]);
                const float z = sigmoidf_(gi[d + c] + gh[d + c]);
                const float hn_ = gh[2 * d + c];
                const float n = tanhf(gi[2 * d + c] + r * hn_);
                const float dn_ = g * (1.0f - z);
                dn = dn_ * (1.0f - n * n);            // d pre-activation of n
                dz = g * (h - n) * z * (1.0f - z);
                dr = dn * hn_ * r * (1.0f - r);
                dhn = dn * r;
            }
            dgi[c] = dr; dgi[d + c] = dz; dgi[2 * d + c] = dn;
            dgh[c] = dr; dgh[d + c] = dz; dgh[2 * d + c] = dhn;
        } else {
            float dp = 0.f;
            if (live) {
                const float y = tanhf(GI[u * ldg + c] + GH[u * ldg + c]);
                dp = g * (1.0f - y * y);
            }
            dgi[c] = dp; dgh[c] = dp;
        }
    }
}

// snapshot of the state rows of the unique touched nodes, taken before persist / store overwrite them:
// HG = memory rows, XG = pending raw messages (row stride rawp), valid_u, lu_u = last_update' (message time if pending)
__global__ void __launch_bounds__(256)
gather_state_kernel(const int32_t* __restrict__ uniq, const int32_t* __restrict__ n_uniq, int64_t u_max, int d, int raw,
                    const float* __restrict__ memory, const float* __restrict__ pend_msg, int64_t rawp,
                    const uint8_t* __restrict__ pend_valid, const float* __restrict__ pend_ts,
                    const float* __restrict__ last_update,
                    float* __restrict__ HG, float* __restrict__ XG, int64_t ldx, float* __restrict__ Hcat, int64_t ldh,
                    uint8_t* __restrict__ valid_u, float* __restrict__ lu_u) {
    pfo_pdl_prologue();
    int64_t U = *n_uniq; if (U > u_max) U = u_max;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < U; u += nwarps) {
        const int node = uniq[u];
        const bool v = pend_valid[node] != 0;
        const float* m = memory + (int64_t)node * d;
        for (int c = lane; c < d; c += 32) {
            const float mc = m[c];
            HG[u * d + c] = mc;
            if (Hcat) Hcat[u * ldh + c] = mc;     // second copy behind the cell input: operand row [input | memory]
        }
        if (XG) {
            const float* x = pend_msg + (int64_t)node * rawp;
            // XG rows keep a 16-byte-aligned stride (ldx >= rawp) so that TMA can stream them
            for (int c = lane; c < rawp; c += 32) XG[u * ldx + c] = (v && c < raw) ? x[c] : 0.0f;
        }
        if (lane == 0) {
            valid_u[u] = v ? 1 : 0;
            lu_u[u] = v ? pend_ts[node] : last_update[node];
        }
    }
}

// ---- persist + last-occurrence ranking ----------------------------------------------
// One warp per (side, event).  Positives with a pending message get memory <- cell output
// and last_update <- message time (tgn.py:185, memory_updater.py:18-33).  Every positive
// receives a new message in this batch, so pend_valid is simply re-set by the store kernel.
__global__ void __launch_bounds__(256)
persist_rank_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int B, int d,
                    const int32_t* __restrict__ slot_of_node, const float* __restrict__ Hnew,
                    const uint8_t* __restrict__ pend_valid, const float* __restrict__ pend_ts,
                    float* __restrict__ memory, float* __restrict__ last_update, int32_t* __restrict__ last_pos) {
    pfo_pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t idx = warp; idx < 2 * (int64_t)B; idx += nwarps) {
        const int i = (int)(idx % B);
        const int node = idx < B ? src[i] : dst[i];
        if (lane == 0) atomicMax(last_pos + node, (int)idx);   // dst-view messages are appended last (tgn.py:205-206)
        if (pend_valid[node]) {
            const float* h = Hnew + (int64_t)slot_of_node[node] * d;
            float* m = memory + (int64_t)node * d;
            for (int c = lane; c < d; c += 32) m[c] = h[c];
            if (lane == 0) last_update[node] = pend_ts[node];
        }
    }
}

// One warp per (side, event); only the last occurrence of a node writes (last-wins,
// message_aggregator.py:49-50).  msg = [mem[self] | mem[other] | edge_feat | cos(dt*w+b)],
// dt = fp32(t) - last_update[self] in fp32 (tgn.py:359-371).
__global__ void __launch_bounds__(256)
store_messages_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                      const int32_t* __restrict__ eidx, const double* __restrict__ ts, int B, int d, int F,
                      const float* __restrict__ memory, const float* __restrict__ last_update,
                      const float* __restrict__ edge_feat, const float* __restrict__ tw, const float* __restrict__ tb,
                      const float* __restrict__ other_emb_for_src, const float* __restrict__ other_emb_for_dst,
                      const float* __restrict__ self_emb_for_src, const float* __restrict__ self_emb_for_dst,
                      float* __restrict__ pend_msg, int64_t rawp, float* __restrict__ pend_ts,
                      uint8_t* __restrict__ pend_valid, int32_t* __restrict__ last_pos) {
    pfo_pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t idx = warp; idx < 2 * (int64_t)B; idx += nwarps) {
        const int i = (int)(idx % B);
        const bool is_src = idx < B;
        const int node = is_src ? src[i] : dst[i];
        const int other = is_src ? dst[i] : src[i];
        if (last_pos[node] != (int)idx) continue;
        __syncwarp();
        const float t32 = (float)ts[i];
        const float delta = t32 - last_update[node];
        float* out = pend_msg + (int64_t)node * rawp;
        const float* se = is_src ? self_emb_for_src : self_emb_for_dst;     // use_source_embedding_in_message (tgn.py:360-361)
        const float* ms = se ? se + (int64_t)i * d : memory + (int64_t)node * d;
        const float* oe = is_src ? other_emb_for_src : other_emb_for_dst;   // dyrep: embedding of the other endpoint
        const float* mo = oe ? oe + (int64_t)i * d : memory + (int64_t)other * d;
        const float* ef = edge_feat + (int64_t)eidx[i] * F;
        for (int c = lane; c < d; c += 32) {
            out[c] = ms[c];
            out[d + c] = mo[c];
            out[2 * d + F + c] = pfo_cosf(fmaf(delta, tw[c], tb[c]));   // fmaf == nn.Linear(1, d) (SURVEY hard part 1)
        }
        for (int c = lane; c < F; c += 32) out[2 * d + c] = ef[c];
        __syncwarp();
        if (lane == 0) {
            pend_ts[node] = t32;
            pend_valid[node] = 1;
            last_pos[node] = -1;
        }
    }
}

// ---- jodie time-projection embedding (embedding_module.py:57-61, tgn.py:260-266) ------
// emb = mem'[q] * (1 + W * td + b), td = (float(int64(t) - int64(last_update'[q])) - mean) / std
__global__ void __launch_bounds__(256)
time_embedding_fwd_kernel(const int32_t* __restrict__ q_nodes, const double* __restrict__ q_ts, int64_t Q,
                          int64_t n_src, int d, const int32_t* __restrict__ slot_of_node,
                          const float* __restrict__ Hnew, const float* __restrict__ lu_u,
                          float mean_src, float std_src, float mean_dst, float std_dst,
                          const float* __restrict__ W, const float* __restrict__ b,
                          float* __restrict__ td_out, float* __restrict__ emb) {
    pfo_pdl_prologue();
    const int64_t total = Q * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = i / d;
        const int c = (int)(i - q * d);
        const int node = q_nodes[q];
        const float lu = lu_u[slot_of_node[node]];
        const long long diff = (long long)q_ts[q] - (long long)lu;
        const bool s = q < n_src;
        const float td = ((float)diff - (s ? mean_src : mean_dst)) / (s ? std_src : std_dst);
        if (c == 0) td_out[q] = td;
        emb[i] = Hnew[(int64_t)slot_of_node[node] * d + c] * (1.0f + fmaf(td, W[c], b[c]));
    }
}

// block per d-column-slab is overkill; one warp per query row, lane-strided columns, then
// per-block partial sums for dW/db written to `partial[grid, 2, d]` (deterministic reduce).
__global__ void __launch_bounds__(256)
time_embedding_bwd_kernel(const int32_t* __restrict__ q_nodes, int64_t Q, int d,
                          const int32_t* __restrict__ slot_of_node, const float* __restrict__ Hnew,
                          const float* __restrict__ td, const float* __restrict__ W, const float* __restrict__ b,
                          const float* __restrict__ dEmb, float* __restrict__ dHnew, float* __restrict__ partial) {
    pfo_pdl_prologue();
    extern __shared__ float sm[];          // [2][d]
    for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) sm[c] = 0.0f;
    __syncthreads();
    const int64_t total = Q * d;
    // consecutive threads walk consecutive columns so that the smem atomics spread over banks
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = i / d;
        const int c = (int)(i - q * d);
        const int node = q_nodes[q];
        const int64_t hrow = (int64_t)slot_of_node[node] * d + c;
        const float g = dEmb[i], h = Hnew[hrow], t = td[q];
        atomicAdd(&sm[c], g * h * t);
        atomicAdd(&sm[d + c], g * h);
        atomicAdd(dHnew + hrow, g * (1.0f + fmaf(t, W[c], b[c])));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) partial[(int64_t)blockIdx.x * 2 * d + c] = sm[c];
}

// out[c] (+)= sum_r partial[r, c] in a fixed order (deterministic): a CTA owns 32 columns, its 8 warps stride the
// rows (coalesced 128-byte reads), then the 8 per-warp sums are added in warp order.  Narrow inputs (cols < 32,
// e.g. the BPR loss partials) put the rows across the lanes instead.
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partial, int rows, int cols, float* __restrict__ out, int accumulate) {
    pfo_pdl_prologue();
    __shared__ float sm[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (cols >= 32) {
        const int c = blockIdx.x * 32 + lane;
        float s = 0.0f;
        if (c < cols) {      // four independent chains per warp (fixed association: deterministic)
            float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
            int r = w;
            for (; r + 24 < rows; r += 32) {
                s0 += partial[(int64_t)r * cols + c];
                s1 += partial[(int64_t)(r + 8) * cols + c];
                s2 += partial[(int64_t)(r + 16) * cols + c];
                s3 += partial[(int64_t)(r + 24) * cols + c];
            }
            for (; r < rows; r += 8) s0 += partial[(int64_t)r * cols + c];
            s = (s0 + s1) + (s2 + s3);
        }
        sm[w][lane] = s;
        __syncthreads();
        if (w == 0 && c < cols) {
            float t = 0.0f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += sm[k][lane];
            out[c] = accumulate ? out[c] + t : t;
        }
    } else {
        const int c = blockIdx.x;                      // one CTA per column
        float s = 0.0f;
        for (int r = threadIdx.x; r < rows; r += 256) s += partial[(int64_t)r * cols + c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sm[w][0] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.0f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += sm[k][0];
            out[c] = accumulate ? out[c] + t : t;
        }
    }
}

// scatter-add rows: dst[idx[m], :] += src[m, :]  (gradient of a row gather)
__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(const float* __restrict__ src, int64_t lds, const int32_t* __restrict__ idx, int64_t M, int d,
                        float* __restrict__ dst, int64_t ldd) {
    pfo_pdl_prologue();
    const int64_t total = M * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / d;
        const int c = (int)(i - m * d);
        const int r = idx[m];
        if (r >= 0) atomicAdd(dst + (int64_t)r * ldd + c, src[m * lds + c]);
    }
}

// gather rows: dst[m, :] = idx[m] >= 0 ? src[idx[m], :] : 0
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, int64_t lds, const int32_t* __restrict__ idx, int64_t M, int d,
                   float* __restrict__ dst, int64_t ldd) {
    pfo_pdl_prologue();
    const int64_t total = M * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / d;
        const int c = (int)(i - m * d);
        const int r = idx[m];
        dst[m * ldd + c] = r >= 0 ? src[(int64_t)r * lds + c] : 0.0f;
    }
}

// ---- node-sharded variants (pfotgnrec_b200/dist.py): messages are BUILT where the events live,
// from the feature rows fetched from the owners, and APPLIED where the nodes live.
// build: one warp per (side, event): row = [mem'[self] | mem'[other] | edge_feat | cos((fp32(t) - lu'[self]) w + b)]
// with mem' / lu' = post-persist memory / last_update = the rows of the unique-node table.
__global__ void __launch_bounds__(256)
build_messages_kernel(const int32_t* __restrict__ src_slot, const int32_t* __restrict__ dst_slot,
                      const int32_t* __restrict__ eidx, const double* __restrict__ ts, int B, int d, int F,
                      const float* __restrict__ Hnew, const float* __restrict__ lu_u,
                      const float* __restrict__ edge_feat, const float* __restrict__ tw, const float* __restrict__ tb,
                      const float* __restrict__ other_emb_for_src, const float* __restrict__ other_emb_for_dst,
                      float* __restrict__ rows, int64_t ldr, float* __restrict__ t32_out,
                      const int32_t* __restrict__ src_node, const int32_t* __restrict__ dst_node, int n_ranks,
                      int key_base, int key_side) {
    pfo_pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t idx = warp; idx < 2 * (int64_t)B; idx += nwarps) {
        const int i = (int)(idx % B);
        const bool is_src = idx < B;
        const int self = is_src ? src_slot[i] : dst_slot[i];
        const int other = is_src ? dst_slot[i] : src_slot[i];
        const float t32 = (float)ts[i];
        const float delta = t32 - lu_u[self];
        float* out = rows + idx * ldr;
        const float* ms = Hnew + (int64_t)self * d;
        const float* oe = is_src ? other_emb_for_src : other_emb_for_dst;
        const float* mo = oe ? oe + (int64_t)i * d : Hnew + (int64_t)other * d;
        const float* ef = edge_feat + (int64_t)eidx[i] * F;
        for (int c = lane; c < d; c += 32) {
            out[c] = ms[c];
            out[d + c] = mo[c];
            out[2 * d + F + c] = pfo_cosf(fmaf(delta, tw[c], tb[c]));
        }
        for (int c = lane; c < F; c += 32) out[2 * d + c] = ef[c];
        if (lane == 0) {
            if (t32_out) t32_out[idx] = t32;
            if (src_node) {
                // routed row: [message (raw) | owner-local node id | global batch position | fp32 time] -- the three
                // words the owner's last-wins pass reads (key = side * B_global + rank * B + event)
                const int raw = 3 * d + F;
                const int node = is_src ? src_node[i] : dst_node[i];
                int* meta = reinterpret_cast<int*>(out + raw);
                meta[0] = node / n_ranks;
                meta[1] = key_base + i + (is_src ? 0 : key_side);
                meta[2] = __float_as_int(t32);
            }
        }
    }
}

// apply (owner side), pass 1: last-wins key + persist of the positives that had a pending message
__global__ void __launch_bounds__(256)
apply_rank_kernel(const int32_t* __restrict__ node, const int32_t* __restrict__ key, int64_t meta_stride, int64_t R, int d,
                  const int32_t* __restrict__ slot_of_node, const float* __restrict__ Hnew,
                  const uint8_t* __restrict__ pend_valid, const float* __restrict__ pend_ts,
                  float* __restrict__ memory, float* __restrict__ last_update, int32_t* __restrict__ last_pos) {
    pfo_pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < R; r += nwarps) {
        const int v = node[r * meta_stride];
        if (v < 0) continue;                              // empty routing slot
        if (lane == 0) atomicMax(last_pos + v, key[r * meta_stride]);
        if (pend_valid[v]) {
            const float* h = Hnew + (int64_t)slot_of_node[v] * d;
            float* m = memory + (int64_t)v * d;
            for (int c = lane; c < d; c += 32) m[c] = h[c];
            if (lane == 0) last_update[v] = pend_ts[v];
        }
    }
}

// pass 2: the row with the largest global position per node becomes the pending message
__global__ void __launch_bounds__(256)
apply_store_kernel(const int32_t* __restrict__ node, const int32_t* __restrict__ key, int64_t meta_stride, int64_t R, int raw,
                   const float* __restrict__ rows, int64_t ldr, const float* __restrict__ t32,
                   float* __restrict__ pend_msg, int64_t rawp, float* __restrict__ pend_ts,
                   uint8_t* __restrict__ pend_valid, int32_t* __restrict__ last_pos) {
    pfo_pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < R; r += nwarps) {
        const int v = node[r * meta_stride];
        if (v < 0 || last_pos[v] != key[r * meta_stride]) continue;
        __syncwarp();
        const float* in = rows + r * ldr;
        float* out = pend_msg + (int64_t)v * rawp;
        for (int c = lane; c < raw; c += 32) out[c] = in[c];
        __syncwarp();
        if (lane == 0) { pend_ts[v] = t32[r * meta_stride]; pend_valid[v] = 1; last_pos[v] = -1; }
    }
}

// `mean` aggregator (reference modules/message_aggregator.py:62-81): the pending message of a node is the MEAN of the
// raw messages its interactions of this batch appended, the pending time that of the last one.  A node's list only
// ever holds the messages of the most recent batch it was a positive of (cleared at tgn.py:191, refilled at :205-206),
// so the dense slot is exact here too.  `sorted_node` / `order` are the (node, position) pairs of the batch sorted by
// node with a STABLE sort (position = side * B + event, ascending inside a node = append order): the warp at the
// head of a node's run walks it in order and sums sequentially -- deterministic, no float atomics.
__global__ void __launch_bounds__(256)
store_messages_mean_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                           const int32_t* __restrict__ eidx, const double* __restrict__ ts, int B, int d, int F,
                           const int32_t* __restrict__ sorted_node, const int32_t* __restrict__ order,
                           const float* __restrict__ memory, const float* __restrict__ last_update,
                           const float* __restrict__ edge_feat, const float* __restrict__ tw, const float* __restrict__ tb,
                           const float* __restrict__ other_emb_for_src, const float* __restrict__ other_emb_for_dst,
                           const float* __restrict__ self_emb_for_src, const float* __restrict__ self_emb_for_dst,
                           float* __restrict__ pend_msg, int64_t rawp, float* __restrict__ pend_ts,
                           uint8_t* __restrict__ pend_valid, int32_t* __restrict__ last_pos) {
    pfo_pdl_prologue();
    constexpr int MAXC = 4;                                    // d <= 128: columns per lane and segment
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n2 = 2 * (int64_t)B;
    for (int64_t k = warp; k < n2; k += nwarps) {
        const int node = sorted_node[k];
        if (k > 0 && sorted_node[k - 1] == node) continue;     // not the head of its run
        float a_self[MAXC], a_other[MAXC], a_te[MAXC], a_e = 0.0f;
#pragma unroll
        for (int j = 0; j < MAXC; ++j) a_self[j] = a_other[j] = a_te[j] = 0.0f;
        const float lu = last_update[node];
        int count = 0;
        float t_last = 0.0f;
        for (int64_t kk = k; kk < n2 && sorted_node[kk] == node; ++kk) {
            const int pos = order[kk];
            const int i = pos % B;
            const bool is_src = pos < B;
            const int other = is_src ? dst[i] : src[i];
            const float t32 = (float)ts[i];
            const float delta = t32 - lu;
            const float* se = is_src ? self_emb_for_src : self_emb_for_dst;
            const float* ms = se ? se + (int64_t)i * d : memory + (int64_t)node * d;
            const float* oe = is_src ? other_emb_for_src : other_emb_for_dst;
            const float* mo = oe ? oe + (int64_t)i * d : memory + (int64_t)other * d;
#pragma unroll
            for (int j = 0; j < MAXC; ++j) {
                const int c = lane + 32 * j;
                if (c < d) {
                    a_self[j] += ms[c];
                    a_other[j] += mo[c];
                    a_te[j] += pfo_cosf(fmaf(delta, tw[c], tb[c]));
                }
            }
            if (lane < F) a_e += edge_feat[(int64_t)eidx[i] * F + lane];
            t_last = t32;
            ++count;
        }
        const float n = (float)count;
        float* out = pend_msg + (int64_t)node * rawp;
#pragma unroll
        for (int j = 0; j < MAXC; ++j) {
            const int c = lane + 32 * j;
            if (c < d) {
                out[c] = a_self[j] / n;
                out[d + c] = a_other[j] / n;
                out[2 * d + F + c] = a_te[j] / n;
            }
        }
        if (lane < F) out[2 * d + lane] = a_e / n;
        if (lane == 0) {
            pend_ts[node] = t_last;
            pend_valid[node] = 1;
            last_pos[node] = -1;
        }
    }
}


}  // namespace

PFO_API int pfo_cell_forward(const int32_t* uniq, const int32_t* n_uniq, int64_t u_max, int d, int cell, int merged,
                             const float* GI, const float* GH, int64_t ldg, const float* HG,
                             const uint8_t* valid_u, const float* node_feat, float* Hnew, float* H0,
                             void* stream) {
    if (u_max <= 0) return 0;
    pfo_launch(cell_forward_kernel, pfo_grid(u_max * d, 256, 8), 256, 0, (cudaStream_t)stream, 
        uniq, n_uniq, u_max, d, cell, merged, GI, GH, ldg, HG, valid_u, node_feat, Hnew, H0);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_cell_backward(const int32_t* uniq, const int32_t* n_uniq, int64_t u_max, int d, int cell, int merged,
                              const float* GI, const float* GH, int64_t ldg, const float* HG,
                              const uint8_t* valid_u, const float* dH, float* dGI, float* dGH, void* stream) {
    if (u_max <= 0) return 0;
    pfo_launch(cell_backward_kernel, pfo_grid(u_max * d, 256, 8), 256, 0, (cudaStream_t)stream, 
        uniq, n_uniq, u_max, d, cell, merged, GI, GH, ldg, HG, valid_u, dH, dGI, dGH);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_gather_state(const int32_t* uniq, const int32_t* n_uniq, int64_t u_max, int d, int raw,
                             const float* memory, const float* pend_msg, int64_t rawp, const uint8_t* pend_valid,
                             const float* pend_ts, const float* last_update,
                             float* HG, float* XG, int64_t ldx, float* Hcat, int64_t ldh, uint8_t* valid_u, float* lu_u,
                             void* stream) {
    if (u_max <= 0) return 0;
    if (XG != nullptr && ldx < rawp) return (int)cudaErrorInvalidValue;
    pfo_launch(gather_state_kernel, pfo_grid(u_max * 32, 256, 8), 256, 0, (cudaStream_t)stream, 
        uniq, n_uniq, u_max, d, raw, memory, pend_msg, rawp, pend_valid, pend_ts, last_update, HG, XG, ldx, Hcat, ldh,
        valid_u, lu_u);
    PFO_LAUNCH_CHECK();
}

// ---- one contraction for the memory updater.  torch's GRUCell / RNNCell (modules/memory_updater.py:60,68) apply
// W_ih [g*d, kx] to the message and W_hh [g*d, d] to the memory in two GEMMs; over the concatenated operand
// [message (kxp columns, zero padded) | memory (d)] they are ONE GEMM with the block weight
//   GRU  rows [r | z] = [W_i{r,z} | W_h{r,z}] (the sums r, z need),  n_i = [W_in | 0],  n_h = [0 | W_hn]   -> 4d x (kxp + d)
//   RNN  rows         = [W_ih | W_hh]                                                                    ->  d x (kxp + d)
// which reads the operand once and saves a launch each way.  pack builds the block weight / bias from the reference's
// four tensors, unpack is its adjoint (accumulating into their gradients).
static __global__ void __launch_bounds__(256)
pack_cell_kernel(const float* __restrict__ W_ih, const float* __restrict__ W_hh, const float* __restrict__ b_ih,
                 const float* __restrict__ b_hh, int d, int kx, int kxp, int cell, float* __restrict__ Wc,
                 float* __restrict__ bc) {
    pfo_pdl_prologue();
    const int K = kxp + d, rows = cell == 0 ? 4 * d : d;
    const int64_t total = (int64_t)rows * K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total + rows; i += (int64_t)gridDim.x * blockDim.x) {
        if (i >= total) {                                   // bias
            const int r = (int)(i - total);
            float v;
            if (cell != 0 || r < 2 * d) v = b_ih[r] + b_hh[r];
            else if (r < 3 * d) v = b_ih[r];
            else v = b_hh[r - d];
            bc[r] = v;
            continue;
        }
        const int r = (int)(i / K), k = (int)(i - (int64_t)r * K);
        float v = 0.0f;
        const bool msg = k < kxp;
        if (cell != 0 || r < 2 * d) v = msg ? (k < kx ? W_ih[(int64_t)r * kx + k] : 0.0f) : W_hh[(int64_t)r * d + (k - kxp)];
        else if (r < 3 * d) v = msg && k < kx ? W_ih[(int64_t)r * kx + k] : 0.0f;
        else v = msg ? 0.0f : W_hh[(int64_t)(r - d) * d + (k - kxp)];
        Wc[i] = v;
    }
}

static __global__ void __launch_bounds__(256)
unpack_cell_grads_kernel(const float* __restrict__ gWc, const float* __restrict__ gbc, int d, int kx, int kxp, int cell,
                         float* __restrict__ gW_ih, float* __restrict__ gW_hh, float* __restrict__ gb_ih,
                         float* __restrict__ gb_hh) {
    pfo_pdl_prologue();
    const int K = kxp + d, g = cell == 0 ? 3 : 1;
    const int64_t n_ih = (int64_t)g * d * kx, n_hh = (int64_t)g * d * d;
    const int64_t total = n_ih + n_hh + 2 * (int64_t)g * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n_ih) {                                     // W_ih[r, k]: block row r of the message columns
            const int r = (int)(i / kx), k = (int)(i - (int64_t)r * kx);
            gW_ih[i] += gWc[(int64_t)r * K + k];
        } else if (i < n_ih + n_hh) {                       // W_hh[r, k]: the n rows live one block further down
            const int64_t j = i - n_ih;
            const int r = (int)(j / d), k = (int)(j - (int64_t)r * d);
            const int rc = (cell == 0 && r >= 2 * d) ? r + d : r;
            gW_hh[j] += gWc[(int64_t)rc * K + kxp + k];
        } else {
            const int64_t j = i - n_ih - n_hh;
            const bool hh = j >= (int64_t)g * d;
            const int r = (int)(hh ? j - (int64_t)g * d : j);
            const int rc = (cell == 0 && r >= 2 * d && hh) ? r + d : r;
            (hh ? gb_hh : gb_ih)[r] += gbc[rc];
        }
    }
}

PFO_API int pfo_pack_cell(const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, int d, int kx,
                          int kxp, int cell, float* Wc, float* bc, void* stream) {
    if (d <= 0 || kx <= 0 || kxp < kx || (cell != 0 && cell != 1)) return (int)cudaErrorInvalidValue;
    const int64_t total = (int64_t)(cell == 0 ? 4 * d : d) * (kxp + d + 1);
    pfo_launch(pack_cell_kernel, pfo_grid(total, 256, 2), 256, 0, (cudaStream_t)stream, W_ih, W_hh, b_ih, b_hh, d, kx,
               kxp, cell, Wc, bc);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_unpack_cell_grads(const float* gWc, const float* gbc, int d, int kx, int kxp, int cell, float* gW_ih,
                                  float* gW_hh, float* gb_ih, float* gb_hh, void* stream) {
    if (d <= 0 || kx <= 0 || kxp < kx || (cell != 0 && cell != 1)) return (int)cudaErrorInvalidValue;
    const int64_t total = (int64_t)(cell == 0 ? 3 : 1) * d * (kx + d + 2);
    pfo_launch(unpack_cell_grads_kernel, pfo_grid(total, 256, 2), 256, 0, (cudaStream_t)stream, gWc, gbc, d, kx, kxp,
               cell, gW_ih, gW_hh, gb_ih, gb_hh);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_persist_rank(const int32_t* src, const int32_t* dst, int B, int d, const int32_t* slot_of_node,
                             const float* Hnew, const uint8_t* pend_valid, const float* pend_ts,
                             float* memory, float* last_update, int32_t* last_pos, void* stream) {
    if (B <= 0) return 0;
    pfo_launch(persist_rank_kernel, pfo_grid(2 * (int64_t)B * 32, 256, 8), 256, 0, (cudaStream_t)stream, 
        src, dst, B, d, slot_of_node, Hnew, pend_valid, pend_ts, memory, last_update, last_pos);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_store_messages(const int32_t* src, const int32_t* dst, const int32_t* eidx, const double* ts,
                               int B, int d, int F, const float* memory, const float* last_update,
                               const float* edge_feat, const float* tw, const float* tb,
                               const float* other_emb_for_src, const float* other_emb_for_dst,
                               const float* self_emb_for_src, const float* self_emb_for_dst,
                               float* pend_msg, int64_t rawp, float* pend_ts, uint8_t* pend_valid,
                               int32_t* last_pos, void* stream) {
    if (B <= 0) return 0;
    pfo_launch(store_messages_kernel, pfo_grid(2 * (int64_t)B * 32, 256, 8), 256, 0, (cudaStream_t)stream, 
        src, dst, eidx, ts, B, d, F, memory, last_update, edge_feat, tw, tb, other_emb_for_src,
        other_emb_for_dst, self_emb_for_src, self_emb_for_dst, pend_msg, rawp, pend_ts, pend_valid, last_pos);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_store_messages_mean(const int32_t* src, const int32_t* dst, const int32_t* eidx, const double* ts,
                                    int B, int d, int F, const int32_t* sorted_node, const int32_t* order,
                                    const float* memory, const float* last_update,
                                    const float* edge_feat, const float* tw, const float* tb,
                                    const float* other_emb_for_src, const float* other_emb_for_dst,
                                    const float* self_emb_for_src, const float* self_emb_for_dst,
                                    float* pend_msg, int64_t rawp, float* pend_ts, uint8_t* pend_valid,
                                    int32_t* last_pos, void* stream) {
    if (B <= 0) return 0;
    if (d > 128 || F > 32) return (int)cudaErrorInvalidValue;
    pfo_launch(store_messages_mean_kernel, pfo_grid(2 * (int64_t)B * 32, 256, 8), 256, 0, (cudaStream_t)stream, 
        src, dst, eidx, ts, B, d, F, sorted_node, order, memory, last_update, edge_feat, tw, tb, other_emb_for_src,
        other_emb_for_dst, self_emb_for_src, self_emb_for_dst, pend_msg, rawp, pend_ts, pend_valid, last_pos);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_time_embedding_fwd(const int32_t* q_nodes, const double* q_ts, int64_t Q, int64_t n_src, int d,
                                   const int32_t* slot_of_node, const float* Hnew, const float* lu_u,
                                   float mean_src, float std_src, float mean_dst, float std_dst,
                                   const float* W, const float* b, float* td_out, float* emb, void* stream) {
    if (Q <= 0) return 0;
    pfo_launch(time_embedding_fwd_kernel, pfo_grid(Q * d, 256, 8), 256, 0, (cudaStream_t)stream, 
        q_nodes, q_ts, Q, n_src, d, slot_of_node, Hnew, lu_u,
        mean_src, std_src, mean_dst, std_dst, W, b, td_out, emb);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_time_embedding_bwd(const int32_t* q_nodes, int64_t Q, int d, const int32_t* slot_of_node,
                                   const float* Hnew, const float* td, const float* W, const float* b,
                                   const float* dEmb, float* dHnew, float* dWdb,
                                   float* workspace, int64_t workspace_floats, void* stream) {
    if (Q <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    int grid = pfo_grid(Q * d, 256, 2);
    if ((int64_t)grid * 2 * d > workspace_floats) grid = (int)(workspace_floats / (2 * d));
    if (grid < 1) return (int)cudaErrorInvalidValue;
    pfo_launch(time_embedding_bwd_kernel, grid, 256, 2 * d * sizeof(float), s, q_nodes, Q, d, slot_of_node, Hnew, td, W, b,
                                                                      dEmb, dHnew, workspace);
    // partial rows are [dW(d) | db(d)]; dWdb receives the same layout
    pfo_launch(reduce_partials_kernel, (2 * d + 31) / 32, 256, 0, s, workspace, grid, 2 * d, dWdb, 0);
    PFO_LAUNCH_CHECK();
}

// TimeEncode.forward on its own (model/time_encoding.py:17-25): out[m, c] = cos(fmaf(t[m], w[c], b[c])), optionally the
// sine too.  mode 0 / 1 = the fp64 quadrant reduction of the fused kernels (pfo_math.cuh), 2 = fp32 Cody-Waite reduction
// (only valid for |argument| < 2^17), 3 = the cosine-only half-turn form of the neighbour forward kernel (sine as in
// mode 0): the parity tests compare the reductions through this entry point.
static __global__ void time_encode_kernel(const float* __restrict__ t, const float* __restrict__ w,
                                          const float* __restrict__ b, int64_t M, int d, int mode,
                                          float* __restrict__ out_cos, float* __restrict__ out_sin) {
    pfo_pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t m = warp; m < M; m += nwarps) {
        const float tm = t[m];
        for (int c0 = 0; c0 < d; c0 += 32) {          // warp-uniform trip count; lanes past d idle on a zero argument
            const int c = c0 + lane;
            const float x = c < d ? fmaf(tm, w[c], b[c]) : 0.0f;
            float sn, cs;
            if (mode == 1) pfo_sincosf_f64(x, &sn, &cs);
            else if (mode == 2) pfo_sincosf_f32(x, &sn, &cs);
            else pfo_sincosf(x, &sn, &cs);
            if (mode == 3) cs = pfo_cosf_half(x);      // the cosine-only form of the neighbour forward kernel
            if (c < d) {
                out_cos[m * d + c] = cs;
                if (out_sin) out_sin[m * d + c] = sn;
            }
        }
    }
}

PFO_API int pfo_time_encode(const float* t, const float* w, const float* b, int64_t M, int d, int mode,
                            float* out_cos, float* out_sin, void* stream) {
    if (M <= 0) return 0;
    if (d <= 0 || mode < 0 || mode > 3) return (int)cudaErrorInvalidValue;
    pfo_launch(time_encode_kernel, pfo_grid(M * 32, 256, 8), 256, 0, (cudaStream_t)stream, t, w, b, M, d, mode, out_cos, out_sin);
    PFO_LAUNCH_CHECK();
}

// ---- Adam over the flat parameter buffer (torch.optim.Adam, main.py:123: no weight decay, no amsgrad):
//   m <- m + (1 - b1)(g - m);  v <- b2 v + (1 - b2) g^2;  p <- p - lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// t is read from device memory (the caller bumps it on the stream first), so a captured graph keeps counting.  One launch
// over ~133 k floats instead of torch's multi-tensor pass over ~22 small tensors (25 us of launch-bound tail per step).
static __global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, float lr, float b1, float b2, float eps, const int32_t* __restrict__ step) {
    pfo_pdl_prologue();
    const float t = (float)*step;
    const float bc1 = 1.0f - powf(b1, t), bc2 = 1.0f - powf(b2, t);
    const float step_size = lr / bc1, rs = rsqrtf(bc2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = fmaf(1.0f - b1, gi - m[i], m[i]);
        const float vi = fmaf(1.0f - b2, gi * gi, b2 * v[i]);
        m[i] = mi; v[i] = vi;
        p[i] -= step_size * (mi / (sqrtf(vi) * rs + eps));
    }
}

PFO_API int pfo_adam_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                          float beta1, float beta2, float eps, const int32_t* step, void* stream) {
    if (n <= 0) return 0;
    pfo_launch(adam_flat_kernel, pfo_grid(n, 256, 4), 256, 0, (cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq,
               n, lr, beta1, beta2, eps, step);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_reduce_partials(const float* partial, int rows, int cols, float* out, int accumulate, void* stream) {
    if (cols <= 0) return 0;
    pfo_launch(reduce_partials_kernel, cols >= 32 ? (cols + 31) / 32 : cols, 256, 0, (cudaStream_t)stream, partial, rows, cols, out, accumulate);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_scatter_add_rows(const float* src, int64_t lds, const int32_t* idx, int64_t M, int d,
                                 float* dst, int64_t ldd, void* stream) {
    if (M <= 0) return 0;
    pfo_launch(scatter_add_rows_kernel, pfo_grid(M * d, 256, 8), 256, 0, (cudaStream_t)stream, src, lds, idx, M, d, dst, ldd);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_gather_rows(const float* src, int64_t lds, const int32_t* idx, int64_t M, int d,
                            float* dst, int64_t ldd, void* stream) {
    if (M <= 0) return 0;
    pfo_launch(gather_rows_kernel, pfo_grid(M * d, 256, 8), 256, 0, (cudaStream_t)stream, src, lds, idx, M, d, dst, ldd);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_build_messages(const int32_t* src_slot, const int32_t* dst_slot, const int32_t* eidx, const double* ts,
                               int B, int d, int F, const float* Hnew, const float* lu_u, const float* edge_feat,
                               const float* tw, const float* tb, const float* other_emb_for_src,
                               const float* other_emb_for_dst, float* rows, int64_t ldr, float* t32_out, void* stream) {
    if (B <= 0) return 0;
    pfo_launch(build_messages_kernel, pfo_grid(2 * (int64_t)B * 32, 256, 8), 256, 0, (cudaStream_t)stream, 
        src_slot, dst_slot, eidx, ts, B, d, F, Hnew, lu_u, edge_feat, tw, tb, other_emb_for_src, other_emb_for_dst,
        rows, ldr, t32_out, nullptr, nullptr, 1, 0, 0);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_build_routed_messages(const int32_t* src_slot, const int32_t* dst_slot, const int32_t* src_node,
                                      const int32_t* dst_node, const int32_t* eidx, const double* ts, int B, int d,
                                      int F, const float* Hnew, const float* lu_u, const float* edge_feat,
                                      const float* tw, const float* tb, const float* other_emb_for_src,
                                      const float* other_emb_for_dst, int n_ranks, int key_base, int key_side,
                                      float* rows, int64_t ldr, void* stream) {
    if (B <= 0) return 0;
    if (ldr < 3 * d + F + 3 || n_ranks <= 0) return (int)cudaErrorInvalidValue;
    pfo_launch(build_messages_kernel, pfo_grid(2 * (int64_t)B * 32, 256, 8), 256, 0, (cudaStream_t)stream, 
        src_slot, dst_slot, eidx, ts, B, d, F, Hnew, lu_u, edge_feat, tw, tb, other_emb_for_src, other_emb_for_dst,
        rows, ldr, nullptr, src_node, dst_node, n_ranks, key_base, key_side);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_apply_routed_messages(const float* rows, int64_t ldr, int64_t R, int d, int raw,
                                      const int32_t* slot_of_node, const float* Hnew, float* memory,
                                      float* last_update, float* pend_msg, int64_t rawp, float* pend_ts,
                                      uint8_t* pend_valid, int32_t* last_pos, void* stream) {
    if (R <= 0) return 0;
    if (ldr < raw + 3) return (int)cudaErrorInvalidValue;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = pfo_grid(R * 32, 256, 8);
    const int32_t* node = reinterpret_cast<const int32_t*>(rows) + raw;      // meta words behind the message
    const int32_t* key = node + 1;
    const float* t32 = rows + raw + 2;
    pfo_launch(apply_rank_kernel, grid, 256, 0, s, node, key, ldr, R, d, slot_of_node, Hnew, pend_valid, pend_ts, memory,
                                           last_update, last_pos);
    pfo_launch(apply_store_kernel, grid, 256, 0, s, node, key, ldr, R, raw, rows, ldr, t32, pend_msg, rawp, pend_ts, pend_valid,
                                            last_pos);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_apply_messages(const int32_t* node, const int32_t* key, int64_t R, int d, int raw,
                               const int32_t* slot_of_node, const float* Hnew, const float* rows, int64_t ldr,
                               const float* t32, float* memory, float* last_update, float* pend_msg, int64_t rawp,
                               float* pend_ts, uint8_t* pend_valid, int32_t* last_pos, void* stream) {
    if (R <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = pfo_grid(R * 32, 256, 8);
    pfo_launch(apply_rank_kernel, grid, 256, 0, s, node, key, 1, R, d, slot_of_node, Hnew, pend_valid, pend_ts, memory,
                                           last_update, last_pos);
    pfo_launch(apply_store_kernel, grid, 256, 0, s, node, key, 1, R, raw, rows, ldr, t32, pend_msg, rawp, pend_ts, pend_valid,
                                            last_pos);
    PFO_LAUNCH_CHECK();
}
