// mt_tables.cu - static per-grid tables of the marching-tetrahedra extraction, built once when a tet grid is loaded.
// Replaces DMTetGeometry.generate_edges (reference model/geometry/dmtet.py:283-288: the six edges of every tet, sorted per edge,
// torch.unique(dim=0)) and adds the tile skip table of csrc/marching_tets.cu.  One-off work (not on the training step): built
// on CUB's device-wide radix sort / unique / scan and block-level sort.
//   edges : key = min * (Vg + 1) + max over the 6 T tet edges -> radix sort -> unique -> CSR by the smaller endpoint:
//           edge_start [Vg + 1], edge_b [E] (lexicographic (min, max) order = the order torch.unique(dim=0) gives the reference).
//   tiles : for every tile of MT_BLOCK consecutive tets the (<= 32) distinct occupancy words (vertex >> 5) its vertices live in,
//           ascending, padded with the smallest; slot 0 = -1 when a tile touches more than 32 words (never skipped).
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int TT = 512;     // tets per tile  (= MT_BLOCK of marching_tets.cu; checked against b2a_mt_tile_shape at run time)
constexpr int TW = 32;      // words per tile row (= MT_TILE_WORDS)

template <typename IdxT>
__global__ void __launch_bounds__(256) edge_keys_kernel(const IdxT* __restrict__ tets, int64_t T, uint64_t n, unsigned long long* __restrict__ keys)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // one thread per (tet, edge)
    if (i >= 6 * T) return;
    const int64_t t = i / 6;
    const int e = (int)(i - t * 6);
    const int ea[6] = {0, 0, 0, 1, 1, 2}, eb[6] = {1, 2, 3, 2, 3, 3};      // base_tet_edges (dmtet.py:46)
    const uint64_t a = (uint64_t)tets[t * 4 + ea[e]], b = (uint64_t)tets[t * 4 + eb[e]];
    keys[i] = (a < b ? a : b) * n + (a < b ? b : a);
}

__global__ void __launch_bounds__(256) edge_count_kernel(const unsigned long long* __restrict__ uniq, const int64_t* __restrict__ num, uint64_t n,
                                                         int* __restrict__ counts)
{
    const int64_t E = *num;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(counts + (uniq[i] / n), 1);
}

__global__ void __launch_bounds__(256) edge_emit_kernel(const unsigned long long* __restrict__ uniq, int64_t E, uint64_t n, int* __restrict__ edge_b)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) edge_b[i] = (int)(uniq[i] % n);
}

struct EdgeWs {
    unsigned long long* keys_a;
    unsigned long long* keys_b;
    int* counts;        // [Vg + 2]
    int64_t* num;       // [1] device: E
    void* cub_tmp;
    size_t cub_bytes;
};

size_t edge_ws_layout(int64_t Vg, int64_t T, void* base, EdgeWs* ws)
{
    const size_t kb = b2a_align((size_t)6 * T * sizeof(unsigned long long));
    const size_t cb = b2a_align((size_t)(Vg + 2) * sizeof(int));
    size_t sort_b = 0, uniq_b = 0, scan_b = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, sort_b, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (size_t)6 * T);
    cub::DeviceSelect::Unique(nullptr, uniq_b, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int64_t*)nullptr, (size_t)6 * T);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_b, (int*)nullptr, (int*)nullptr, (size_t)(Vg + 1));
    size_t tb = sort_b > uniq_b ? sort_b : uniq_b;
    if (scan_b > tb) tb = scan_b;
    tb = b2a_align(tb);
    if (ws) {
        char* p = (char*)base;
        ws->keys_a = (unsigned long long*)p;
        ws->keys_b = (unsigned long long*)(p + kb);
        ws->counts = (int*)(p + 2 * kb);
        ws->num = (int64_t*)(p + 2 * kb + cb);
        ws->cub_tmp = p + 2 * kb + cb + 256;
        ws->cub_bytes = tb;
    }
    return 2 * kb + cb + 256 + tb;
}

// one block per tile: 2048 words -> block radix sort -> distinct ones ranked by a block scan
__global__ void __launch_bounds__(256) tile_words_kernel(const int* __restrict__ tets, int64_t T, int* __restrict__ table)
{
    using Sort = cub::BlockRadixSort<int, 256, 8>;
    using Scan = cub::BlockScan<int, 256>;
    __shared__ union { typename Sort::TempStorage sort; typename Scan::TempStorage scan; } tmp;
    __shared__ int sorted[TT * 4 + 1];
    __shared__ int total;
    const int64_t tile = blockIdx.x;
    int keys[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        int64_t idx = tile * TT * 4 + (int64_t)threadIdx.x * 8 + j;            // flat index into tets [T,4]
        if (idx >= T * 4) idx = (T - 1) * 4 + (idx & 3);                       // last tile: repeat the final tet's words
        keys[j] = __ldg(tets + idx) >> 5;
    }
    Sort(tmp.sort).Sort(keys);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; j++) sorted[threadIdx.x * 8 + j] = keys[j];
    __syncthreads();
    int first[8], cnt = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int i = threadIdx.x * 8 + j;
        first[j] = (i == 0 || sorted[i] != sorted[i - 1]) ? 1 : 0;
        cnt += first[j];
    }
    int rank0, tot;
    Scan(tmp.scan).ExclusiveSum(cnt, rank0, tot);
    if (threadIdx.x == 0) total = tot;
    int* row = table + tile * TW;
    if (threadIdx.x < TW) row[threadIdx.x] = sorted[0];                            // padding = the smallest word
    __syncthreads();
    int r = rank0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (first[j]) {
            if (r < TW) row[r] = keys[j];
            r++;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && total > TW) row[0] = -1;                              // touches more words than the row holds: never skipped
}

}  // namespace

B2A_API int b2a_mt_tables_workspace_bytes(int64_t Vg, int64_t T, size_t* bytes)
{
    B2A_CHECK_ARG(bytes && Vg > 0 && T > 0, "shape");
    *bytes = edge_ws_layout(Vg, T, nullptr, nullptr);
    return 0;
}

// Phase 1: unique sorted edges inside the workspace; edge_start [Vg + 1] written; *num_edges (DEVICE int64, also readable as pinned
// host memory) = E.  Phase 2 (b2a_mt_emit_edges, after the caller has read E and allocated edge_b [E]) copies the larger endpoints.
B2A_API int b2a_mt_build_edges(const void* tets, int tets_are_i64, int64_t Vg, int64_t T, void* workspace, size_t workspace_bytes, int32_t* edge_start,
                               int64_t* num_edges, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(tets && workspace && edge_start && num_edges && Vg > 0 && T > 0 && Vg < (1ll << 31), "arguments");
    EdgeWs ws;
    B2A_CHECK_ARG(workspace_bytes >= edge_ws_layout(Vg, T, workspace, &ws) && ((uintptr_t)workspace & 255) == 0, "workspace");
    const uint64_t n = (uint64_t)Vg + 1;
    const size_t N6 = (size_t)6 * T;
    if (tets_are_i64) edge_keys_kernel<int64_t><<<b2a_blocks(6 * T, 256), 256, 0, stream>>>((const int64_t*)tets, T, n, ws.keys_a);
    else edge_keys_kernel<int><<<b2a_blocks(6 * T, 256), 256, 0, stream>>>((const int*)tets, T, n, ws.keys_a);
    int bits = 1;
    while (bits < 64 && ((n * n - 1) >> bits) != 0) bits++;
    size_t tb = ws.cub_bytes;
    B2A_CUDA_OK(cub::DeviceRadixSort::SortKeys(ws.cub_tmp, tb, ws.keys_a, ws.keys_b, N6, 0, bits, stream));
    tb = ws.cub_bytes;
    B2A_CUDA_OK(cub::DeviceSelect::Unique(ws.cub_tmp, tb, ws.keys_b, ws.keys_a, ws.num, N6, stream));
    B2A_CUDA_OK(cudaMemsetAsync(ws.counts, 0, (size_t)(Vg + 2) * sizeof(int), stream));
    edge_count_kernel<<<148 * 8, 256, 0, stream>>>(ws.keys_a, ws.num, n, ws.counts);
    tb = ws.cub_bytes;
    B2A_CUDA_OK(cub::DeviceScan::ExclusiveSum(ws.cub_tmp, tb, ws.counts, edge_start, (size_t)(Vg + 1), stream));
    B2A_CUDA_OK(cudaMemcpyAsync(num_edges, ws.num, sizeof(int64_t), cudaMemcpyDefault, stream));
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_mt_emit_edges(const void* workspace, size_t workspace_bytes, int64_t Vg, int64_t T, int64_t E, int32_t* edge_b, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(workspace && edge_b && E >= 0 && E <= 6 * T, "arguments");
    EdgeWs ws;
    B2A_CHECK_ARG(workspace_bytes >= edge_ws_layout(Vg, T, const_cast<void*>(workspace), &ws), "workspace");
    if (E) edge_emit_kernel<<<148 * 8, 256, 0, stream>>>(ws.keys_a, E, (uint64_t)Vg + 1, edge_b);
    B2A_LAUNCH_OK();
    return 0;
}

// tile_words [ceil(T / tile_tets), tile_words] int32 (b2a_mt_tile_shape) from int32 tets [T,4]
B2A_API int b2a_mt_build_tile_words(const int32_t* tets, int64_t T, int32_t* tile_words, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(tets && tile_words && T > 0, "arguments");
    int tt = 0, tw = 0;
    b2a_mt_tile_shape(&tt, &tw);
    B2A_CHECK_ARG(tt == TT && tw == TW, "tile shape of marching_tets.cu changed: rebuild mt_tables.cu");
    tile_words_kernel<<<(unsigned)((T + TT - 1) / TT), 256, 0, stream>>>(tets, T, tile_words);
    B2A_LAUNCH_OK();
    return 0;
}
