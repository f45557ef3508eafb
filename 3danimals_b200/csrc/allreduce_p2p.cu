// allreduce_p2p.cu - gradient all-reduce (average) over NVLink / NVSwitch PEER MEMORY in one kernel launch.
//
// The path's only exchange is the DDP all-reduce of parameter gradients (reference: accelerate / DistributedDataParallel,
// Trainer.py:170-180).  The ranks of this path are host-bound, so what a collective costs a step is mostly the HOST
// time of issuing it: a c10d / NCCL call is ~40 us of host work per bucket; this kernel is one ~5 us launch per bucket set.
// Every rank maps the other ranks' gradient buffers (CUDA IPC, set up once by 3danimals_b200/parallel.py) and runs the same kernel:
//   1. barrier    - a flag word per (rank, peer) in peer memory: "my buffer holds this step's gradients"
//   2. reduce     - rank r owns slice r of the buffer: it reads that slice from all N buffers (its own from HBM, the others over
//                   NVLink, 16-byte loads), averages in a fixed rank order (every rank computes bit-identical results), and
//                   writes the result into ALL N buffers (remote 16-byte stores): reduce-scatter and all-gather in one pass
//   3. barrier    - "my writes into your buffer are done"
// No slice is read by one rank while another writes it: slice r is read and written by rank r only.  Traffic per rank:
// (N-1)/N of the buffer in, the same out - what a ring all-reduce moves, at one hop through NVSwitch.
#include "common.cuh"

namespace {

constexpr int MAX_RANKS = 8;

struct P2PArgs {
    float* buf[MAX_RANKS];          // every rank's buffer, mapped into this process
    unsigned* flags[MAX_RANKS];     // every rank's flag words [2 * MAX_RANKS]: (phase, writer rank)
    int rank, world;
    int64_t n4;                     // float4 elements
    unsigned epoch;                 // strictly increasing per launch of one channel
};

__device__ __forceinline__ void flag_store(unsigned* p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned flag_load(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// grid barrier across the blocks of THIS rank's launch (the kernel is launched with at most one resident wave)
__device__ __forceinline__ void grid_sync(unsigned* counter, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (atomicAdd(counter, 0u) < target) {}
        __threadfence();
    }
    __syncthreads();
}

// cross-rank barrier: block 0 announces `epoch` in every peer's flag row and waits for every peer's announcement in its own.
// A peer that never arrives (its process died) must not hang this GPU: after SPIN_LIMIT cycles (~10 s) the wait gives up and
// raises the channel's error word (local_counter[1]), which the host reads when it closes the peer memory.
constexpr long long SPIN_LIMIT = 20000000000ll;
__device__ __forceinline__ void rank_barrier(const P2PArgs& a, int phase, unsigned* error_word)
{
    if (blockIdx.x == 0 && threadIdx.x < a.world) {
        const int peer = threadIdx.x;
        __threadfence_system();
        flag_store(a.flags[peer] + phase * MAX_RANKS + a.rank, a.epoch);
        const long long t0 = clock64();
        while (flag_load(a.flags[a.rank] + phase * MAX_RANKS + peer) < a.epoch) {
            if (clock64() - t0 > SPIN_LIMIT) { atomicExch(error_word, 1u); break; }
        }
    }
}

__global__ void __launch_bounds__(512) allreduce_p2p_kernel(P2PArgs a, unsigned* local_counter, unsigned counter_base)
{
    // phase 0: everybody's gradients are in place (block 0 talks to the peers, then releases this rank's other blocks)
    rank_barrier(a, 0, local_counter + 1);
    grid_sync(local_counter, counter_base + gridDim.x);
    const int64_t per = (a.n4 + a.world - 1) / a.world;
    const int64_t lo = per * a.rank, hi = lo + per < a.n4 ? lo + per : a.n4;
    const float inv = 1.f / (float)a.world;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < MAX_RANKS; p++)
            if (p < a.world) {
                const float4 v = reinterpret_cast<const float4*>(a.buf[p])[i];
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
#pragma unroll
        for (int p = 0; p < MAX_RANKS; p++)
            if (p < a.world) reinterpret_cast<float4*>(a.buf[p])[i] = s;
    }
    // phase 1: my slice has been written everywhere; wait until every peer's slice has landed here
    __threadfence_system();
    grid_sync(local_counter, counter_base + 2 * gridDim.x);
    rank_barrier(a, 1, local_counter + 1);
}

}  // namespace

// A buffer other ranks can map: cudaMalloc'ed (zero-filled) with its 64-byte IPC handle.  Peer-memory buffers are allocated here, not
// by the caller's allocator, so that the handle names exactly this buffer (offset 0).
B2A_API int b2a_p2p_alloc(size_t bytes, void** ptr, void* handle64)
{
    B2A_CHECK_ARG(ptr && handle64 && bytes > 0, "arguments");
    void* p = nullptr;
    B2A_CUDA_OK(cudaMalloc(&p, bytes));
    B2A_CUDA_OK(cudaMemset(p, 0, bytes));
    B2A_CUDA_OK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    B2A_CUDA_OK(cudaIpcGetMemHandle(&h, p));
    static_assert(sizeof(h) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
    *ptr = p;
    return 0;
}

// Map another rank's buffer (same node) into this process, on the CURRENT device: peer access is enabled as part of the mapping.
B2A_API int b2a_p2p_open(const void* handle64, void** ptr)
{
    B2A_CHECK_ARG(ptr && handle64, "arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    B2A_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

B2A_API int b2a_p2p_close(void* ptr)
{
    if (ptr) B2A_CUDA_OK(cudaIpcCloseMemHandle(ptr));
    return 0;
}

B2A_API int b2a_p2p_free(void* ptr)
{
    if (ptr) B2A_CUDA_OK(cudaFree(ptr));
    return 0;
}

// bufs / flags: `world` device pointers each (this rank's own at index `rank`): every rank's buffer of n floats (n % 4 == 0, 16-byte
// aligned) and its 2*8 zero-initialised flag words for this channel; local_counter: this rank's two zero-initialised device words for this
// channel (grid-barrier counter, error word - nonzero after a peer failed to arrive within ~10 s); epoch: 1, 2, 3, ... per channel (the same on every rank for the same collective; a channel always reduces the same n).
// Averages in place on every rank.  Every rank of the group must launch it; the kernel spins until the peers arrive.
B2A_API int b2a_allreduce_p2p(const void* const* bufs, const void* const* flags, int rank, int world, int64_t n, int epoch, void* local_counter,
                              b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(bufs && flags && local_counter && world >= 1 && world <= MAX_RANKS && rank >= 0 && rank < world && n >= 0 && (n & 3) == 0 && epoch > 0,
                  "arguments");
    if (n == 0) return 0;
    P2PArgs a;
    for (int p = 0; p < MAX_RANKS; p++) {
        a.buf[p] = p < world ? (float*)bufs[p] : nullptr;
        a.flags[p] = p < world ? (unsigned*)flags[p] : nullptr;
        B2A_CHECK_ARG(p >= world || (a.buf[p] && a.flags[p] && ((uintptr_t)a.buf[p] & 15) == 0), "peer pointers");
    }
    a.rank = rank; a.world = world; a.n4 = n / 4; a.epoch = (unsigned)epoch;
    const int64_t per = (a.n4 + world - 1) / world;
    int blocks = (int)((per + 511) / 512);
    if (blocks > 64) blocks = 64;       // a fraction of the SMs: the copy engines are not involved, the kernel shares the GPU with the backward
    if (blocks < 1) blocks = 1;
    const unsigned base = (unsigned)(epoch - 1) * 2u * (unsigned)blocks;      // the local counter only ever grows (same block count every launch of a set)
    allreduce_p2p_kernel<<<blocks, 512, 0, stream>>>(a, (unsigned*)local_counter, base);
    B2A_LAUNCH_OK();
    return 0;
}
