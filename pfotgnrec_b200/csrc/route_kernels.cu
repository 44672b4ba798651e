// Device-side exchange plans of the node-sharded multi-GPU mode (pfotgnrec_b200/dist.py).
//
// The reference is single-process; sharding the node state over G GPUs by owner(x) = x mod G adds one step before
// every exchange: rows (queries, unique node ids, gradient rows, message rows) have to be grouped by destination rank.
// Round 1 did that with a sort + bincount and copied the split sizes to the HOST (two syncs per exchange).  Here the
// plan never leaves the device and every shape is static: each (source, destination) pair owns a bucket of `cap` rows
// in the send buffer, a row's place is slot = dest * cap + (its arrival order in that bucket), handed out by
// warp-aggregated atomics; rows that do not fit raise an overflow flag the host reads lazily.  The buffers then go
// through ONE equal-split all-to-all (static sizes: capturable in the step's CUDA graph), empty slots carry id -1 and
// every consumer kernel skips them, and replies travel back in the same slots, so the requester finds the answer to
// row i at slot[i] without any bookkeeping on the owner.
#include "common.cuh"

namespace {

// slot[i] = (ids[i] mod G) * cap + arrival order, or -1 when the row is dropped (ids[i] < 0, i >= *n_valid, or the
// bucket is full).  counts[g] ends as the number of rows destined to rank g (including the ones that did not fit).
__global__ void __launch_bounds__(256)
route_plan_kernel(const int32_t* __restrict__ ids, int64_t R, const int32_t* __restrict__ n_valid, int G, int cap,
                  int32_t* __restrict__ counts, int32_t* __restrict__ slot, int32_t* __restrict__ local_id,
                  int32_t* __restrict__ overflow) {
    pfo_pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int64_t limit = n_valid ? (int64_t)*n_valid : R;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // whole warps walk the array together (the trip count is rounded up to a multiple of 32)
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < R; i0 += stride) {
        const int64_t i = i0 + lane;
        int id = -1;
        if (i < R && i < limit) id = ids[i];
        const int dest = id >= 0 ? id % G : -1;
        int pos = -1;
        unsigned todo = __ballot_sync(0xffffffffu, dest >= 0);
        while (todo) {                                    // one round per distinct destination in the warp (<= G)
            const int leader = __ffs(todo) - 1;
            const int g = __shfl_sync(0xffffffffu, dest, leader);
            const unsigned same = __ballot_sync(0xffffffffu, dest == g);
            int base = 0;
            if (lane == leader) base = atomicAdd(counts + g, __popc(same));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (dest == g) pos = base + __popc(same & ((1u << lane) - 1u));
            todo &= ~same;
        }
        if (i < R) {
            int s = -1;
            if (dest >= 0) {
                if (pos < cap) s = dest * cap + pos;
                else atomicOr(overflow, 1);
            }
            slot[i] = s;
            if (local_id) local_id[i] = id >= 0 ? id / G : -1;
        }
    }
}

// dst[slot[m], 0:w] = src[m, 0:w] for slot[m] >= 0 (32-bit words, any payload type)
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const uint32_t* __restrict__ src, int64_t lds, const int32_t* __restrict__ slot, int64_t M,
                    const int32_t* __restrict__ n_valid, int w, uint32_t* __restrict__ dst, int64_t ldd) {
    pfo_pdl_prologue();
    if (n_valid && *n_valid < M) M = *n_valid;      // the tables are sized for the worst case: walk the live rows only
    const int64_t total = M * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / w;
        const int c = (int)(i - m * w);
        const int s = slot[m];
        if (s >= 0) dst[(int64_t)s * ldd + c] = src[m * lds + c];
    }
}

// dst[m, 0:w] = slot[m] >= 0 ? src[slot[m], 0:w] : fill   (32-bit words)
__global__ void __launch_bounds__(256)
gather_words_kernel(const uint32_t* __restrict__ src, int64_t lds, const int32_t* __restrict__ slot, int64_t M, int w,
                    uint32_t* __restrict__ dst, int64_t ldd, uint32_t fill) {
    pfo_pdl_prologue();
    const int64_t total = M * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / w;
        const int c = (int)(i - m * w);
        const int s = slot[m];
        dst[m * ldd + c] = s >= 0 ? src[(int64_t)s * lds + c] : fill;
    }
}

// request rows of exchange R1 in one pass: row = [local node id, timestamp (2 words), query id] at the row's slot
__global__ void __launch_bounds__(256)
pack_queries_kernel(const int32_t* __restrict__ local_id, const double* __restrict__ q_ts,
                    const int32_t* __restrict__ q_ids, const int32_t* __restrict__ slot, int64_t Q,
                    int32_t* __restrict__ out) {
    pfo_pdl_prologue();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Q; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = slot[i];
        if (s < 0) continue;
        const long long tb = __double_as_longlong(q_ts[i]);
        int4 row;
        row.x = local_id[i];
        row.y = (int)(tb & 0xffffffffll);
        row.z = (int)(tb >> 32);
        row.w = q_ids ? q_ids[i] : (int)i;
        reinterpret_cast<int4*>(out)[s] = row;
    }
}

// owner side of R1: request rows -> the dense query arrays K1 takes (empty slots stay node = -1)
__global__ void __launch_bounds__(256)
unpack_queries_kernel(const int32_t* __restrict__ in, int64_t R, int32_t* __restrict__ q_nodes,
                      double* __restrict__ q_ts, int32_t* __restrict__ q_ids) {
    pfo_pdl_prologue();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < R; i += (int64_t)gridDim.x * blockDim.x) {
        const int4 row = reinterpret_cast<const int4*>(in)[i];
        q_nodes[i] = row.x;
        const long long tb = ((long long)row.z << 32) | (unsigned int)row.y;
        q_ts[i] = row.x >= 0 ? __longlong_as_double(tb) : 0.0;
        q_ids[i] = row.w;
    }
}

// requester side of R1: the reply rows [nbr (n) | eidx (n) | dt (n)] come back in the slots the queries left in; one pass
// puts them into the three dense [Q, n] arrays of the un-sharded finder (dropped rows: zeros)
__global__ void __launch_bounds__(256)
unroute_neighbors_kernel(const int32_t* __restrict__ back, const int32_t* __restrict__ slot, int64_t Q, int n,
                         int32_t* __restrict__ nbr, int32_t* __restrict__ eidx, float* __restrict__ dt) {
    pfo_pdl_prologue();
    const int64_t total = Q * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = i / n;
        const int j = (int)(i - q * n);
        const int s = slot[q];
        int a = 0, b = 0, c = 0;
        if (s >= 0) {
            const int32_t* row = back + (int64_t)s * 3 * n;
            a = row[j]; b = row[n + j]; c = row[2 * n + j];
        }
        nbr[i] = a; eidx[i] = b; dt[i] = __int_as_float(c);
    }
}

// owner side of R2: reply row = [updated memory row (d) | last_update'] of the node each received id names (holes: zeros)
__global__ void __launch_bounds__(256)
route_reply_rows_kernel(const float* __restrict__ Hnew_own, const float* __restrict__ lu_own,
                        const int32_t* __restrict__ slots_own, int64_t R, int d, float* __restrict__ reply) {
    pfo_pdl_prologue();
    const int w = d + 1;
    const int64_t total = R * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / w;
        const int c = (int)(i - r * w);
        const int s = slots_own[r];
        float v = 0.0f;
        if (s >= 0) v = c < d ? Hnew_own[(int64_t)s * d + c] : lu_own[s];
        reply[i] = v;
    }
}

// requester side of R2: rows of the unique-node table from the reply slots, and H0 = memory' + node features
__global__ void __launch_bounds__(256)
unroute_rows_kernel(const float* __restrict__ back, const int32_t* __restrict__ slot, const int32_t* __restrict__ uniq,
                    const int32_t* __restrict__ n_valid, const float* __restrict__ node_feat, int64_t U, int d,
                    float* __restrict__ Hnew, float* __restrict__ lu_u, float* __restrict__ H0) {
    pfo_pdl_prologue();
    if (n_valid && *n_valid < U) U = *n_valid;      // rows past the unique count are never read downstream
    const int w = d + 1;
    const int64_t total = U * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = i / w;
        const int c = (int)(i - u * w);
        const int s = slot[u];
        const float v = s >= 0 ? back[(int64_t)s * w + c] : 0.0f;
        if (c < d) {
            Hnew[u * d + c] = v;
            H0[u * d + c] = v + node_feat[(int64_t)uniq[u] * d + c];
        } else {
            lu_u[u] = v;
        }
    }
}

}  // namespace

PFO_API int pfo_unroute_neighbors(const int32_t* back, const int32_t* slot, int64_t n_queries, int n, int32_t* nbr,
                                  int32_t* eidx, float* dt, void* stream) {
    if (n_queries <= 0) return 0;
    pfo_launch(unroute_neighbors_kernel, pfo_grid(n_queries * n, 256, 8), 256, 0, (cudaStream_t)stream, back, slot,
               n_queries, n, nbr, eidx, dt);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_route_reply_rows(const float* Hnew_own, const float* lu_own, const int32_t* slots_own, int64_t n_rows,
                                 int d, float* reply, void* stream) {
    if (n_rows <= 0) return 0;
    pfo_launch(route_reply_rows_kernel, pfo_grid(n_rows * (d + 1), 256, 8), 256, 0, (cudaStream_t)stream, Hnew_own,
               lu_own, slots_own, n_rows, d, reply);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_unroute_rows(const float* back, const int32_t* slot, const int32_t* uniq, const int32_t* n_valid,
                             const float* node_feat, int64_t n_rows, int d, float* Hnew, float* lu_u, float* H0,
                             void* stream) {
    if (n_rows <= 0) return 0;
    pfo_launch(unroute_rows_kernel, pfo_grid(n_rows * (d + 1), 256, 8), 256, 0, (cudaStream_t)stream, back, slot, uniq,
               n_valid, node_feat, n_rows, d, Hnew, lu_u, H0);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_route_plan(const int32_t* ids, int64_t n_rows, const int32_t* n_valid, int n_ranks, int cap,
                           int32_t* counts, int32_t* slot, int32_t* local_id, int32_t* overflow, void* stream) {
    if (n_ranks <= 0 || cap <= 0) return (int)cudaErrorInvalidValue;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int32_t) * n_ranks, s);
    if (e != cudaSuccess) return (int)e;
    if (n_rows <= 0) return 0;
    pfo_launch(route_plan_kernel, pfo_grid(n_rows, 256, 4), 256, 0, s, ids, n_rows, n_valid, n_ranks, cap, counts, slot,
                                                              local_id, overflow);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_scatter_rows(const void* src, int64_t lds, const int32_t* slot, int64_t M, const int32_t* n_valid,
                             int w, void* dst, int64_t ldd, void* stream) {
    if (M <= 0 || w <= 0) return 0;
    pfo_launch(scatter_rows_kernel, pfo_grid(M * w, 256, 8), 256, 0, (cudaStream_t)stream, 
        (const uint32_t*)src, lds, slot, M, n_valid, w, (uint32_t*)dst, ldd);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_gather_words(const void* src, int64_t lds, const int32_t* slot, int64_t M, int w, void* dst,
                             int64_t ldd, uint32_t fill, void* stream) {
    if (M <= 0 || w <= 0) return 0;
    pfo_launch(gather_words_kernel, pfo_grid(M * w, 256, 8), 256, 0, (cudaStream_t)stream, 
        (const uint32_t*)src, lds, slot, M, w, (uint32_t*)dst, ldd, fill);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_pack_queries(const int32_t* local_id, const double* q_ts, const int32_t* q_ids, const int32_t* slot,
                             int64_t n_queries, int32_t* out_rows, void* stream) {
    if (n_queries <= 0) return 0;
    pfo_launch(pack_queries_kernel, pfo_grid(n_queries, 256, 8), 256, 0, (cudaStream_t)stream, local_id, q_ts, q_ids, slot,
                                                                                     n_queries, out_rows);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_unpack_queries(const int32_t* in_rows, int64_t n_rows, int32_t* q_nodes, double* q_ts, int32_t* q_ids,
                               void* stream) {
    if (n_rows <= 0) return 0;
    pfo_launch(unpack_queries_kernel, pfo_grid(n_rows, 256, 8), 256, 0, (cudaStream_t)stream, in_rows, n_rows, q_nodes, q_ts, q_ids);
    PFO_LAUNCH_CHECK();
}
