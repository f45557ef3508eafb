// marching_tets.cu - DMTet SDF -> mesh extraction on sm_100a.
// Replaces DMTet.__call__ (reference model/geometry/dmtet.py:104-155): same vertex order (lexicographic order of
// the unique sorted crossing edges, i.e. torch.unique(dim=0) order) and same face order ([1-triangle tets in tet
// order][2-triangle tets in tet order]) without any runtime sort: a static per-grid CSR of unique edges, crossing
// counts per min-vertex, and device-wide exclusive scans.  HBM-bound: algorithmic bytes per extraction are
// 4*Vg (sdf) + 4*(Vg+E) (edge CSR) + 16*T (tets) + small.
#include "common.cuh"

namespace {

constexpr int MT_BLOCK = 512;

__constant__ int8_t c_tri_table[16][6] = {
    {-1, -1, -1, -1, -1, -1}, {1, 0, 2, -1, -1, -1}, {4, 0, 3, -1, -1, -1}, {1, 4, 2, 1, 3, 4},
    {3, 1, 5, -1, -1, -1},    {2, 3, 0, 2, 5, 3},    {1, 4, 0, 1, 5, 4},    {4, 2, 5, -1, -1, -1},
    {4, 5, 2, -1, -1, -1},    {4, 1, 0, 4, 5, 1},    {3, 2, 0, 3, 5, 2},    {1, 3, 5, -1, -1, -1},
    {4, 1, 2, 4, 3, 1},       {3, 0, 4, -1, -1, -1}, {2, 0, 1, -1, -1, -1}, {-1, -1, -1, -1, -1, -1}};
__constant__ int8_t c_num_tri[16] = {0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0};
__constant__ int8_t c_base_edges[12] = {0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3};

struct MtWorkspace {
    uint32_t* occ_bits;   // [ceil(Vg/32)]
    uint32_t* cand_bits;  // [ceil(Vg/32)] vertices of tets that straddle the surface (the only ones that can own a crossing edge)
    uint8_t* vcnt;        // [Vg]
    int* vtile;           // [nVT]
    uint8_t* tetidx;      // [T]
    int* t1tile;          // [nTT]
    int* t2tile;          // [nTT]
    int* edge_vidx;       // [E]
    int* err;             // [1]
    int64_t nVT, nTT;
};

size_t mt_layout(int64_t Vg, int64_t E, int64_t T, void* base, MtWorkspace* ws)
{
    int64_t nVT = (Vg + MT_BLOCK - 1) / MT_BLOCK, nTT = (T + MT_BLOCK - 1) / MT_BLOCK;
    size_t off = 0;
    char* p = (char*)base;
    auto take = [&](size_t bytes) { size_t o = off; off += b2a_align(bytes); return p ? (void*)(p + o) : nullptr; };
    void* occ = take((size_t)((Vg + 31) / 32) * 4);
    void* cand = take((size_t)((Vg + 31) / 32) * 4);
    void* vcnt = take((size_t)Vg);
    void* vtile = take((size_t)(nVT + 1) * 4);
    void* tetidx = take((size_t)T);
    void* t1 = take((size_t)(nTT + 1) * 4);
    void* t2 = take((size_t)(nTT + 1) * 4);
    void* ev = take((size_t)E * 4);
    void* err = take(256);
    if (ws) {
        ws->occ_bits = (uint32_t*)occ; ws->cand_bits = (uint32_t*)cand; ws->vcnt = (uint8_t*)vcnt; ws->vtile = (int*)vtile;
        ws->tetidx = (uint8_t*)tetidx; ws->t1tile = (int*)t1; ws->t2tile = (int*)t2;
        ws->edge_vidx = (int*)ev; ws->err = (int*)err; ws->nVT = nVT; ws->nTT = nTT;
    }
    return off;
}

__device__ __forceinline__ bool occ_at(const uint32_t* __restrict__ bits, int v) { return (__ldg(bits + (v >> 5)) >> (v & 31)) & 1u; }

// occupancy bitmask: occ = sdf > 0  (dmtet.py:106).  Each warp packs four 32-vertex words (four loads in flight per lane).
// Also clears the candidate-vertex bits (same word grid) and the error flag for the kernels that follow: no memsets.
__global__ void __launch_bounds__(MT_BLOCK) mt_occ_kernel(const float* __restrict__ sdf, int64_t Vg, uint32_t* __restrict__ bits,
                                                          uint32_t* __restrict__ cand, int* __restrict__ err)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t v0 = warp * 128 + lane;
    if (blockIdx.x == 0 && threadIdx.x == 0) *err = 0;
    float x[4];
#pragma unroll
    for (int k = 0; k < 4; k++) x[k] = v0 + k * 32 < Vg ? __ldg(sdf + v0 + k * 32) : 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t m = __ballot_sync(0xffffffffu, x[k] > 0.f);
        if (lane == 0 && warp * 128 + k * 32 < Vg) { bits[warp * 4 + k] = m; cand[warp * 4 + k] = 0u; }
    }
}

// crossing edges per min-vertex + per-tile totals.  Runs after mt_tcount: only vertices of surface-straddling tets
// (cand bits) can own a crossing edge, so a tile without candidates leaves after one 64-byte read and the 60 MB edge CSR
// is only touched along the surface.
__global__ void __launch_bounds__(MT_BLOCK) mt_vcount_kernel(const int* __restrict__ edge_start, const int* __restrict__ edge_b,
                                                             const uint32_t* __restrict__ bits, const uint32_t* __restrict__ cand, int64_t Vg,
                                                             uint8_t* __restrict__ vcnt, int* __restrict__ vtile, int* __restrict__ err)
{
    __shared__ int sm[34];
    int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t cw = a < Vg ? __ldg(cand + (a >> 5)) : 0u;
    if (!__syncthreads_or(cw != 0u)) {
        if (threadIdx.x == 0) vtile[blockIdx.x] = 0;
        return;
    }
    int cnt = 0;
    if (a < Vg) {
        if ((cw >> (a & 31)) & 1u) {
            int s = __ldg(edge_start + a), e = __ldg(edge_start + a + 1);
            bool oa = occ_at(bits, (int)a);
            for (int i = s; i < e; i++) cnt += (occ_at(bits, __ldg(edge_b + i)) != oa);
            if (cnt > 255) { atomicExch(err, 1); cnt = 255; }
        }
        vcnt[a] = (uint8_t)cnt;
    }
    // only the tile total is needed here
    int wsum = warp_sum_i(cnt);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = wsum;
    __syncthreads();
    if (threadIdx.x < 32) {
        int x = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0;
        x = warp_sum_i(x);
        if (threadIdx.x == 0) vtile[blockIdx.x] = x;
    }
}

// tet case index + per-tile totals of 1-triangle and 2-triangle tets  (dmtet.py:107-109,135-137).  A tile is MT_BLOCK
// consecutive tets; a block streams MT_TPB tiles, four 16-byte tet loads in flight per thread; only the tile TOTALS are
// needed here (warp ballots + one shared-memory atomic per warp), the ordered positions are recomputed by the emit
// kernel for the few tiles that hold triangles.
// tile_words (nullable): static per-grid table [nTT, MT_TILE_WORDS] of the occupancy words a tile's tets touch (built once
// when the grid is loaded; -1 in slot 0: too many to list).  If all of them are all-zero or all-one the tile cannot
// straddle the surface and its 8 KB of tet indices are never read: the extraction streams the surface band, not the grid.
constexpr int MT_TPB = 4;
constexpr int MT_TILE_WORDS = 32;
__global__ void __launch_bounds__(MT_BLOCK) mt_tcount_kernel(const int4* __restrict__ tets, const uint32_t* __restrict__ bits,
                                                             const int* __restrict__ tile_words, int64_t T, int64_t nTT,
                                                             uint8_t* __restrict__ tetidx, int* __restrict__ t1tile, int* __restrict__ t2tile,
                                                             uint32_t* __restrict__ cand)
{
    __shared__ int s_cnt[MT_TPB][2];
    __shared__ int s_live[MT_TPB];
    if (threadIdx.x < MT_TPB * 2) (&s_cnt[0][0])[threadIdx.x] = 0;
    const int64_t tile0 = (int64_t)blockIdx.x * MT_TPB;
    if (threadIdx.x < MT_TPB * 32) {
        const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
        bool live = true;
        if (tile_words && tile0 + k < nTT) {
            const int widx = __ldg(tile_words + (tile0 + k) * MT_TILE_WORDS + lane);
            const bool always = __shfl_sync(0xffffffffu, widx, 0) < 0;
            const uint32_t w = (always || widx < 0) ? 1u : __ldg(bits + widx);
            const bool z = __all_sync(0xffffffffu, w == 0u), o = __all_sync(0xffffffffu, w == 0xffffffffu);
            live = always || !(z || o);
        }
        if (lane == 0) s_live[k] = live;
    }
    __syncthreads();
    int4 q[MT_TPB];
#pragma unroll
    for (int k = 0; k < MT_TPB; k++) {
        const int64_t t = (tile0 + k) * MT_BLOCK + threadIdx.x;
        q[k] = (s_live[k] && t < T) ? __ldg(tets + t) : make_int4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < MT_TPB; k++) {
        if (!s_live[k]) continue;
        const int64_t t = (tile0 + k) * MT_BLOCK + threadIdx.x;
        int n = 0;
        if (t < T) {
            int ti = (int)occ_at(bits, q[k].x) | ((int)occ_at(bits, q[k].y) << 1) | ((int)occ_at(bits, q[k].z) << 2) | ((int)occ_at(bits, q[k].w) << 3);
            tetidx[t] = (uint8_t)ti;
            n = c_num_tri[ti];
            if (n) {   // straddles the surface: its vertices are the candidates for crossing edges
                atomicOr(cand + (q[k].x >> 5), 1u << (q[k].x & 31)); atomicOr(cand + (q[k].y >> 5), 1u << (q[k].y & 31));
                atomicOr(cand + (q[k].z >> 5), 1u << (q[k].z & 31)); atomicOr(cand + (q[k].w >> 5), 1u << (q[k].w & 31));
            }
        }
        const unsigned m1 = __ballot_sync(0xffffffffu, n == 1), m2 = __ballot_sync(0xffffffffu, n == 2);
        if ((threadIdx.x & 31) == 0) {
            if (m1) atomicAdd(&s_cnt[k][0], __popc(m1));
            if (m2) atomicAdd(&s_cnt[k][1], __popc(m2));
        }
    }
    __syncthreads();
    if (threadIdx.x < MT_TPB && tile0 + threadIdx.x < nTT) {
        t1tile[tile0 + threadIdx.x] = s_cnt[threadIdx.x][0];
        t2tile[tile0 + threadIdx.x] = s_cnt[threadIdx.x][1];
    }
}

// One block per array: in-place exclusive scan of up to three tile-sum arrays; totals -> counts[which].  Each warp owns
// a contiguous segment and walks it in coalesced 32-element rows (pass 1: segment total; block scan of the 32 warp
// totals; pass 2: shuffle scan per row with a running carry).  The array keeps one extra slot [n] = total so that tile
// b's own count is data[b+1] - data[b].
struct ScanJob { int* data; int64_t n; };
__global__ void __launch_bounds__(1024) mt_scan_tiles_kernel(ScanJob j0, ScanJob j1, ScanJob j2, const int* __restrict__ err, int* __restrict__ counts)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[3] = *err;   // written by mt_vcount, which has completed
    __shared__ int s_warp[32];
    ScanJob j = blockIdx.x == 0 ? j0 : (blockIdx.x == 1 ? j1 : j2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t rows = (j.n + 31) / 32;
    const int64_t rows_per_warp = (rows + 31) / 32;
    const int64_t r0 = (int64_t)warp * rows_per_warp, r1 = min(r0 + rows_per_warp, rows);
    int local = 0;
#pragma unroll 4
    for (int64_t r = r0; r < r1; r++) {
        int64_t i = r * 32 + lane;
        local += i < j.n ? j.data[i] : 0;
    }
    local = warp_sum_i(local);
    if (lane == 0) s_warp[warp] = local;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane], inc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        s_warp[lane] = inc - w;
        if (lane == 31) { j.data[j.n] = inc; counts[blockIdx.x] = inc; }
    }
    __syncthreads();
    int carry = s_warp[warp];
#pragma unroll 4
    for (int64_t r = r0; r < r1; r++) {
        int64_t i = r * 32 + lane;
        int v = i < j.n ? j.data[i] : 0, inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (i < j.n) j.data[i] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// emit vertices: v = p_a*((-s_b)/den) + p_b*(s_a/den), den = s_a - s_b  (dmtet.py:124-131), unfused fp32 ops
__global__ void __launch_bounds__(MT_BLOCK) mt_vemit_kernel(const float* __restrict__ pos, const float* __restrict__ sdf,
                                                            const int* __restrict__ edge_start, const int* __restrict__ edge_b,
                                                            const uint32_t* __restrict__ bits, const uint8_t* __restrict__ vcnt,
                                                            const int* __restrict__ vtile, int64_t Vg, int64_t nVT, float* __restrict__ verts,
                                                            int* __restrict__ vert_edge, int* __restrict__ edge_vidx)
{
    __shared__ int sm[34];
    __shared__ int s_list[MT_BLOCK];
    __shared__ int s_n;
    // Phase A: one thread per tile finds the tiles where a crossing edge
    // starts (few: the surface) - one round of loads instead of a serial walk; phase B: emit those tiles.
    // (tiles are dealt round-robin: the surface occupies a contiguous band of tiles, contiguous ranges would pile it
    // onto a few blocks)
    for (int64_t chunk = blockIdx.x; chunk < nVT; chunk += (int64_t)gridDim.x * MT_BLOCK) {
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    {
        const int64_t tile = chunk + (int64_t)threadIdx.x * gridDim.x;
        if (tile < nVT && vtile[tile + 1] != vtile[tile]) s_list[atomicAdd(&s_n, 1)] = (int)threadIdx.x;
    }
    __syncthreads();
    const int n_live = s_n;
    for (int li = 0; li < n_live; li++) {
        const int64_t tile = chunk + (int64_t)s_list[li] * gridDim.x;
        const int tile_base = vtile[tile];
        int64_t a = tile * blockDim.x + threadIdx.x;
        int cnt = a < Vg ? (int)vcnt[a] : 0;
        int total;
        int ex = block_exclusive_scan(cnt, sm, &total);
        if (cnt == 0) continue;
        int out = tile_base + ex;
        int s = __ldg(edge_start + a), e = __ldg(edge_start + a + 1);
        bool oa = occ_at(bits, (int)a);
        float sa = __ldg(sdf + a);
        float ax = __ldg(pos + a * 3), ay = __ldg(pos + a * 3 + 1), az = __ldg(pos + a * 3 + 2);
        for (int i = s; i < e; i++) {
            int b = __ldg(edge_b + i);
            if (occ_at(bits, b) == oa) continue;
            float sb = -__ldg(sdf + b);
            float den = sa + sb;
            float wa = sb / den, wb = sa / den;
            float bx = __ldg(pos + (int64_t)b * 3), by = __ldg(pos + (int64_t)b * 3 + 1), bz = __ldg(pos + (int64_t)b * 3 + 2);
            verts[(int64_t)out * 3 + 0] = ax * wa + bx * wb;
            verts[(int64_t)out * 3 + 1] = ay * wa + by * wb;
            verts[(int64_t)out * 3 + 2] = az * wa + bz * wb;
            vert_edge[(int64_t)out * 2 + 0] = (int)a;
            vert_edge[(int64_t)out * 2 + 1] = b;
            edge_vidx[i] = out;
            out++;
        }
    }
    __syncthreads();
    }
}

// output-vertex id of grid edge (va,vb): position of max(va,vb) in the CSR row of min(va,vb).  The first 16 row entries
// are fetched with independent predicated loads (one latency instead of a serial compare-and-branch walk; Kuhn rows hold
// <= 14 entries), longer rows continue with a loop.
__device__ __forceinline__ int find_edge_vertex(const int* __restrict__ edge_start, const int* __restrict__ edge_b,
                                                const int* __restrict__ edge_vidx, int va, int vb)
{
    int lo = min(va, vb), hi = max(va, vb);
    int s = __ldg(edge_start + lo), e = __ldg(edge_start + lo + 1);
    int hit = -1;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        int b = s + j < e ? __ldg(edge_b + s + j) : -1;
        if (b == hi) hit = s + j;
    }
    for (int i = s + 16; i < e && hit < 0; i++)
        if (__ldg(edge_b + i) == hi) hit = i;
    return hit >= 0 ? edge_vidx[hit] : -1;
}

// emit faces (dmtet.py:139-151) and uv indices (map_uv :86-96)
__global__ void __launch_bounds__(MT_BLOCK) mt_temit_kernel(const int4* __restrict__ tets, const uint8_t* __restrict__ tetidx,
                                                            const int* __restrict__ t1tile, const int* __restrict__ t2tile,
                                                            const int* __restrict__ edge_start, const int* __restrict__ edge_b,
                                                            const int* __restrict__ edge_vidx, int64_t T, int64_t nTT, int64_t N1,
                                                            int* __restrict__ faces32, long long* __restrict__ faces64,
                                                            long long* __restrict__ uv64)
{
    __shared__ int sm[34];
    __shared__ int s_list[MT_BLOCK];
    __shared__ int s_n, s_m;
    __shared__ int s_rec[MT_BLOCK][7];
    // Phase A: one thread per tile finds the tiles that hold surface
    // (few) - one round of loads instead of a serial walk; phase B: emit those tiles.
    // (tiles are dealt round-robin: the surface occupies a contiguous band of tiles, contiguous ranges would pile it
    // onto a few blocks)
    for (int64_t chunk = blockIdx.x; chunk < nTT; chunk += (int64_t)gridDim.x * MT_BLOCK) {
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    {
        const int64_t tile = chunk + (int64_t)threadIdx.x * gridDim.x;
        if (tile < nTT && (t1tile[tile + 1] != t1tile[tile] || t2tile[tile + 1] != t2tile[tile])) s_list[atomicAdd(&s_n, 1)] = (int)threadIdx.x;
    }
    __syncthreads();
    const int n_live = s_n;
    for (int li = 0; li < n_live; li++) {
        const int64_t tile = chunk + (int64_t)s_list[li] * gridDim.x;
        const int base1 = t1tile[tile], base2 = t2tile[tile];
        int64_t t = tile * blockDim.x + threadIdx.x;
        int ti = t < T ? (int)tetidx[t] : 0;
        int n = c_num_tri[ti];
        int tot;
        int ex = block_exclusive_scan((int)(n == 1) | ((int)(n == 2) << 16), sm, &tot);   // both counts in one scan (MT_BLOCK < 2^16)
        int e1 = ex & 0xffff, e2 = ex >> 16;
        // surface tets of the tile -> shared list; then ONE (tet, triangle corner) lookup per thread, all in flight together
        // (a thread walking its own 3-6 corners serialises ~20 dependent L2 round trips)
        if (threadIdx.x == 0) s_m = 0;
        __syncthreads();
        if (n) {
            int64_t fbase = n == 1 ? (int64_t)base1 + e1 : N1 + 2 * ((int64_t)base2 + e2);
            int slot = atomicAdd(&s_m, 1);
            int4 q = __ldg(tets + t);
            s_rec[slot][0] = (int)fbase; s_rec[slot][1] = ti | (n << 8); s_rec[slot][2] = threadIdx.x;
            s_rec[slot][3] = q.x; s_rec[slot][4] = q.y; s_rec[slot][5] = q.z; s_rec[slot][6] = q.w;
        }
        __syncthreads();
        const int items = s_m * 6;
        for (int it = threadIdx.x; it < items; it += blockDim.x) {
            const int r = it / 6, slot6 = it - r * 6, k = slot6 / 3, c = slot6 - k * 3;
            const int tin = s_rec[r][1], ti2 = tin & 0xff, n2 = tin >> 8;
            if (k >= n2) continue;
            const int64_t f = (int64_t)s_rec[r][0] + k;
            const int le = c_tri_table[ti2][k * 3 + c];
            const int vid = find_edge_vertex(edge_start, edge_b, edge_vidx, s_rec[r][3 + c_base_edges[le * 2]], s_rec[r][3 + c_base_edges[le * 2 + 1]]);
            if (faces32) faces32[f * 3 + c] = vid;
            if (faces64) faces64[f * 3 + c] = vid;
            if (uv64) {
                const long long g = (long long)(tile * blockDim.x + s_rec[r][2]) * 4;
                uv64[f * 3 + c] = c == 0 ? g : g + k + c;     // (4t, 4t+k+1, 4t+k+2)
            }
        }
    }
    __syncthreads();
    }
}

// backward of the vertex interpolation w.r.t. sdf (and optionally pos)
__global__ void mt_bwd_kernel(const float* __restrict__ pos, const float* __restrict__ sdf, const int* __restrict__ vert_edge,
                              const float* __restrict__ d_verts, int64_t V, float* __restrict__ d_sdf, float* __restrict__ d_pos)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    int a = vert_edge[i * 2], b = vert_edge[i * 2 + 1];
    float sa = __ldg(sdf + a), sb = __ldg(sdf + b);
    float den = sa - sb, id2 = 1.f / (den * den);
    float gx = d_verts[i * 3], gy = d_verts[i * 3 + 1], gz = d_verts[i * 3 + 2];
    float dx = __ldg(pos + (int64_t)a * 3) - __ldg(pos + (int64_t)b * 3);
    float dy = __ldg(pos + (int64_t)a * 3 + 1) - __ldg(pos + (int64_t)b * 3 + 1);
    float dz = __ldg(pos + (int64_t)a * 3 + 2) - __ldg(pos + (int64_t)b * 3 + 2);
    float g = gx * dx + gy * dy + gz * dz;
    atomicAdd(d_sdf + a, g * sb * id2);
    atomicAdd(d_sdf + b, -g * sa * id2);
    if (d_pos) {
        float wa = -sb / den, wb = sa / den;
        atomicAdd(d_pos + (int64_t)a * 3 + 0, gx * wa); atomicAdd(d_pos + (int64_t)a * 3 + 1, gy * wa); atomicAdd(d_pos + (int64_t)a * 3 + 2, gz * wa);
        atomicAdd(d_pos + (int64_t)b * 3 + 0, gx * wb); atomicAdd(d_pos + (int64_t)b * 3 + 1, gy * wb); atomicAdd(d_pos + (int64_t)b * 3 + 2, gz * wb);
    }
}

}  // namespace

B2A_API int b2a_mt_workspace_bytes(int64_t Vg, int64_t E, int64_t T, size_t* bytes)
{
    B2A_CHECK_ARG(bytes && Vg >= 0 && E >= 0 && T >= 0, "sizes");
    *bytes = mt_layout(Vg, E, T, nullptr, nullptr);
    return 0;
}

B2A_API int b2a_mt_tile_shape(int* tile_tets, int* tile_words)
{
    B2A_CHECK_ARG(tile_tets && tile_words, "null pointer");
    *tile_tets = MT_BLOCK;
    *tile_words = MT_TILE_WORDS;
    return 0;
}

B2A_API int b2a_mt_count(const float* sdf, const int32_t* tets, const int32_t* edge_start, const int32_t* edge_b,
                         const int32_t* tile_words, int64_t Vg, int64_t E, int64_t T, void* workspace, size_t workspace_bytes,
                         int32_t* counts, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(sdf && tets && edge_start && edge_b && workspace && counts, "null pointer");
    B2A_CHECK_ARG(Vg > 0 && T > 0 && Vg < (1ll << 31) && E < (1ll << 31) && T < (1ll << 29), "grid size");
    B2A_CHECK_ARG(((uintptr_t)tets & 15) == 0, "tets must be 16-byte aligned");
    MtWorkspace ws;
    B2A_CHECK_ARG(mt_layout(Vg, E, T, workspace, &ws) <= workspace_bytes, "workspace too small");
    mt_occ_kernel<<<b2a_blocks(Vg, MT_BLOCK * 4), MT_BLOCK, 0, stream>>>(sdf, Vg, ws.occ_bits, ws.cand_bits, ws.err);
    mt_tcount_kernel<<<(unsigned)((ws.nTT + MT_TPB - 1) / MT_TPB), MT_BLOCK, 0, stream>>>((const int4*)tets, ws.occ_bits, tile_words, T, ws.nTT,
                                                                                          ws.tetidx, ws.t1tile, ws.t2tile, ws.cand_bits);
    mt_vcount_kernel<<<(unsigned)ws.nVT, MT_BLOCK, 0, stream>>>(edge_start, edge_b, ws.occ_bits, ws.cand_bits, Vg, ws.vcnt, ws.vtile, ws.err);
    ScanJob j0{ws.vtile, ws.nVT}, j1{ws.t1tile, ws.nTT}, j2{ws.t2tile, ws.nTT};
    mt_scan_tiles_kernel<<<3, 1024, 0, stream>>>(j0, j1, j2, ws.err, counts);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_mt_emit(const float* pos, const float* sdf, const int32_t* tets, const int32_t* edge_start,
                        const int32_t* edge_b, int64_t Vg, int64_t E, int64_t T, void* workspace, size_t workspace_bytes,
                        int64_t V, int64_t N1, int64_t N2, float* verts, int32_t* vert_edge, int32_t* faces_i32,
                        int64_t* faces_i64, int64_t* uv_idx_i64, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(pos && sdf && tets && edge_start && edge_b && workspace, "null pointer");
    MtWorkspace ws;
    B2A_CHECK_ARG(mt_layout(Vg, E, T, workspace, &ws) <= workspace_bytes, "workspace too small");
    if (V > 0) {
        B2A_CHECK_ARG(verts && vert_edge, "null vertex outputs");
        mt_vemit_kernel<<<(unsigned)min((int64_t)148 * 4, ws.nVT), MT_BLOCK, 0, stream>>>(pos, sdf, edge_start, edge_b, ws.occ_bits, ws.vcnt,
                                                                                          ws.vtile, Vg, ws.nVT, verts, vert_edge, ws.edge_vidx);
    }
    if (N1 + N2 > 0 && (faces_i32 || faces_i64 || uv_idx_i64))
        mt_temit_kernel<<<(unsigned)min((int64_t)148 * 4, ws.nTT), MT_BLOCK, 0, stream>>>((const int4*)tets, ws.tetidx, ws.t1tile, ws.t2tile,
                                                                                          edge_start, edge_b, ws.edge_vidx, T, ws.nTT, N1, faces_i32,
                                                                                          (long long*)faces_i64, (long long*)uv_idx_i64);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_mt_bwd(const float* pos, const float* sdf, const int32_t* vert_edge, const float* d_verts, int64_t V,
                       float* d_sdf, float* d_pos, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(pos && sdf && vert_edge && d_verts && d_sdf, "null pointer");
    if (V > 0) mt_bwd_kernel<<<b2a_blocks(V, 256), 256, 0, stream>>>(pos, sdf, vert_edge, d_verts, V, d_sdf, d_pos);
    B2A_LAUNCH_OK();
    return 0;
}
