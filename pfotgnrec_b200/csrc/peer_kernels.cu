// Peer-memory transport of the node-sharded exchanges (pfotgnrec_b200/dist.py): the all-to-all as direct stores into the
// other GPUs' memory over NVLink / NVSwitch, plus a flag barrier -- in place of one NCCL all-to-all per exchange.
//
// Every rank allocates one arena with cudaMalloc, exports it as a CUDA IPC handle and maps the arenas of its peers
// (`pfo_peer_alloc / open`).  A receive buffer is the SAME offset in every arena (the ranks run the same allocation
// sequence).  An exchange is then two launches on the step's stream, both capturable in the step's CUDA graph:
//   push     block g of the local send buffer -> peer g's receive buffer, block `rank`   (16-byte stores over NVLink;
//            the local block is an ordinary copy)
//   barrier  __threadfence_system, then flag[id][rank] = epoch on every peer (release), then wait until every
//            flag[id][g] of the local arena reached the epoch (acquire).  Epochs count the invocations of barrier `id`
//            in device memory, so a graph replay keeps counting; a rank that runs ahead can only raise a flag, and the
//            comparison is >=.  A wait that exceeds `timeout_ns` raises an error word instead of hanging the GPU.
// A barrier with a dedicated id at the start of every step keeps a fast rank from writing into buffers a slow rank is
// still reading from the previous step.  Small messages make these exchanges latency-bound: the two launches cost
// ~2 x 3 us + one NVLink round trip against ~20-30 us for an NCCL all-to-all inside a graph.
#include "common.cuh"

namespace {

constexpr int kMaxRanks = 8;
constexpr int kMaxBarriers = 64;

struct PeerBases { unsigned long long base[kMaxRanks]; };

__global__ void __launch_bounds__(256)
peer_push_kernel(const uint32_t* __restrict__ send, PeerBases peers, int64_t recv_off_bytes, int G, int rank,
                 int64_t block_words) {
    pfo_pdl_prologue();
    // grid.y = destination rank; 16-byte vectors when the block allows it
    const int g = blockIdx.y;
    const uint32_t* src = send + (int64_t)g * block_words;
    uint32_t* dst = reinterpret_cast<uint32_t*>(peers.base[g] + recv_off_bytes) + (int64_t)rank * block_words;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    if ((block_words & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (int64_t i = tid; i < (block_words >> 2); i += nth) d4[i] = s4[i];
    } else {
        for (int64_t i = tid; i < block_words; i += nth) dst[i] = src[i];
    }
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(32)
peer_barrier_kernel(PeerBases peers, int64_t flags_off_bytes, int64_t epoch_off_bytes, int G, int rank, int id,
                    uint32_t* __restrict__ error_word, unsigned long long timeout_ns) {
    pfo_pdl_prologue();
    __shared__ uint32_t epoch_s;
    uint32_t* epoch = reinterpret_cast<uint32_t*>(peers.base[rank] + epoch_off_bytes) + id;
    if (threadIdx.x == 0) { epoch_s = *epoch + 1u; *epoch = epoch_s; }
    __syncthreads();
    const uint32_t e = epoch_s;
    __threadfence_system();                               // this rank's stores into the peers' buffers come first
    const int t = threadIdx.x;
    if (t < G) {
        uint32_t* remote = reinterpret_cast<uint32_t*>(peers.base[t] + flags_off_bytes) + (int64_t)id * kMaxRanks + rank;
        st_release_sys(remote, e);
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(peers.base[rank] + flags_off_bytes) + (int64_t)id * kMaxRanks + t;
        const unsigned long long t0 = globaltimer_ns();
        while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
            if (globaltimer_ns() - t0 > timeout_ns) { atomicOr(error_word, 2u); break; }
        }
    }
    __threadfence_system();
}

}  // namespace

PFO_API int pfo_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle64) {
    if (bytes <= 0 || ptr == nullptr || handle64 == nullptr) return (int)cudaErrorInvalidValue;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*ptr, 0, (size_t)bytes);
    if (e != cudaSuccess) return (int)e;
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, *ptr);
    if (e != cudaSuccess) return (int)e;
    memcpy(handle64, &h, 64);
    return 0;
}

PFO_API int pfo_peer_open(const unsigned char* handle64, void** ptr) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

PFO_API int pfo_peer_close(void* ptr) { return (int)cudaIpcCloseMemHandle(ptr); }
PFO_API int pfo_peer_free(void* ptr) { return (int)cudaFree(ptr); }
PFO_API int pfo_peer_max_ranks(void) { return kMaxRanks; }
PFO_API int pfo_peer_max_barriers(void) { return kMaxBarriers; }
PFO_API int64_t pfo_peer_header_bytes(void) { return (int64_t)kMaxBarriers * kMaxRanks * 4 + kMaxBarriers * 4; }

static int fill_bases(PeerBases& p, const uint64_t* bases, int G) {
    if (G <= 0 || G > kMaxRanks) return 1;
    for (int g = 0; g < kMaxRanks; ++g) p.base[g] = g < G ? bases[g] : 0ull;
    return 0;
}

// bases: HOST array of the G arena base addresses as seen from this process (own arena at index `rank`)
PFO_API int pfo_peer_push(const void* send, const uint64_t* bases, int64_t recv_off_bytes, int n_ranks, int rank,
                          int64_t block_words, void* stream) {
    PeerBases p;
    if (fill_bases(p, bases, n_ranks) || block_words <= 0 || (recv_off_bytes & 15) != 0) return (int)cudaErrorInvalidValue;
    int64_t per = (block_words / 4 + 255) / 256;
    int gx = (int)(per < 1 ? 1 : (per > 4 * pfo_num_sms() / n_ranks + 1 ? 4 * pfo_num_sms() / n_ranks + 1 : per));
    pfo_launch(peer_push_kernel, dim3(gx, n_ranks), 256, 0, (cudaStream_t)stream, (const uint32_t*)send, p,
               recv_off_bytes, n_ranks, rank, block_words);
    PFO_LAUNCH_CHECK();
}

PFO_API int pfo_peer_barrier(const uint64_t* bases, int n_ranks, int rank, int id, uint32_t* error_word,
                             double timeout_s, void* stream) {
    PeerBases p;
    if (fill_bases(p, bases, n_ranks) || id < 0 || id >= kMaxBarriers) return (int)cudaErrorInvalidValue;
    const int64_t flags_off = 0, epoch_off = (int64_t)kMaxBarriers * kMaxRanks * 4;
    pfo_launch(peer_barrier_kernel, 1, 32, 0, (cudaStream_t)stream, p, flags_off, epoch_off, n_ranks, rank, id,
               error_word, (unsigned long long)(timeout_s * 1e9));
    PFO_LAUNCH_CHECK();
}
