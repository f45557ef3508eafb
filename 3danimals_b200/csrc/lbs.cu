// lbs.cu - linear blend skinning on sm_100a.
// Replaces skinning() (reference model/geometry/skinning.py:369-439).  Closed form:
//   T_i = Rest_i * Rot_xyz(theta_i) * Rest_i^-1,   G_k = T_root ... T_parent(k) T_k,
//   w[k,v] = softmax_k(-sqrt(d^2(v, segment_k) + 1e-6) / temperature)   (weights see detached vertices),
//   out[b,v] = sum_k w[k,v] * (G[b,k] [v,1]).
// One fused kernel per direction over all (image, vertex) pairs: weights are recomputed, never stored.
// HBM-bound: 12 B read + 12 B write per (image, vertex) forward; same again backward.
#include "common.cuh"

namespace {

constexpr int LBS_BLOCK = 256;
constexpr int LBS_MAX_K = 64;
constexpr int LBS_MAX_CHAIN = 16;

struct Aff {  // 3x4 affine, row-major
    float m[12];
};

__device__ __forceinline__ Aff aff_identity()
{
    Aff a;
#pragma unroll
    for (int i = 0; i < 12; i++) a.m[i] = (i % 5 == 0) ? 1.f : 0.f;
    return a;
}
__device__ __forceinline__ Aff aff_mul(const Aff& A, const Aff& B)
{
    Aff C;
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float s = A.m[r * 4 + 0] * B.m[0 * 4 + c] + A.m[r * 4 + 1] * B.m[1 * 4 + c] + A.m[r * 4 + 2] * B.m[2 * 4 + c];
            if (c == 3) s += A.m[r * 4 + 3];
            C.m[r * 4 + c] = s;
        }
    }
    return C;
}
__device__ __forceinline__ Aff aff_load(const float* p)
{
    Aff a;
#pragma unroll
    for (int i = 0; i < 12; i++) a.m[i] = p[i];
    return a;
}

// rest frame of a bone: columns [right | up | forward]  (skinning.py:251-270)
__device__ __forceinline__ void bone_frame(const float* bone, float R[9], float j[3])
{
    j[0] = bone[0]; j[1] = bone[1]; j[2] = bone[2];
    float fx = bone[3] - bone[0], fy = bone[4] - bone[1], fz = bone[5] - bone[2];
    float n = fmaxf(sqrtf(fx * fx + fy * fy + fz * fz), 1e-12f);
    fx /= n; fy /= n; fz /= n;
    // up = normalize(forward x (1,0,0)) = normalize((0, fz, -fy))
    float ux = 0.f, uy = fz, uz = -fy;
    float un = fmaxf(sqrtf(uy * uy + uz * uz), 1e-12f);
    uy /= un; uz /= un;
    // right = up x forward
    float rx = uy * fz - uz * fy, ry = uz * fx - ux * fz, rz = ux * fy - uy * fx;
    float un2 = fmaxf(sqrtf(ux * ux + uy * uy + uz * uz), 1e-12f);  // second normalize of up (skinning.py:264)
    ux /= un2; uy /= un2; uz /= un2;
    R[0] = rx; R[1] = ux; R[2] = fx;
    R[3] = ry; R[4] = uy; R[5] = fy;
    R[6] = rz; R[7] = uz; R[8] = fz;
}

__device__ __forceinline__ void mat3_mul(const float* A, const float* B, float* C)
{
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}
__device__ __forceinline__ void mat3_mul_bt(const float* A, const float* B, float* C)  // A * B^T
{
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) C[r * 3 + c] = A[r * 3] * B[c * 3] + A[r * 3 + 1] * B[c * 3 + 1] + A[r * 3 + 2] * B[c * 3 + 2];
}
__device__ __forceinline__ void mat3_mul_at(const float* A, const float* B, float* C)  // A^T * B
{
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) C[r * 3 + c] = A[r] * B[c] + A[3 + r] * B[3 + c] + A[6 + r] * B[6 + c];
}

__device__ __forceinline__ void euler_mats(const float* th, float* Rx, float* Ry, float* Rz)
{
    float sx, cx, sy, cy, sz, cz;
    sincosf(th[0], &sx, &cx); sincosf(th[1], &sy, &cy); sincosf(th[2], &sz, &cz);
    Rx[0] = 1; Rx[1] = 0; Rx[2] = 0; Rx[3] = 0; Rx[4] = cx; Rx[5] = -sx; Rx[6] = 0; Rx[7] = sx; Rx[8] = cx;
    Ry[0] = cy; Ry[1] = 0; Ry[2] = sy; Ry[3] = 0; Ry[4] = 1; Ry[5] = 0; Ry[6] = -sy; Ry[7] = 0; Ry[8] = cy;
    Rz[0] = cz; Rz[1] = -sz; Rz[2] = 0; Rz[3] = sz; Rz[4] = cz; Rz[5] = 0; Rz[6] = 0; Rz[7] = 0; Rz[8] = 1;
}

// T_local[b,i] = Rest_i Rot(theta) Rest_i^-1 :  Q = R Rot R^T,  t = j - Q j
__global__ void lbs_local_kernel(const float* __restrict__ bones, const float* __restrict__ angles, int B, int Bb, int K,
                                 float* __restrict__ T_local)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * K) return;
    int b = idx / K, i = idx % K;
    float R[9], j[3], Rx[9], Ry[9], Rz[9], A[9], Rot[9], Q[9];
    bone_frame(bones + ((size_t)(Bb == 1 ? 0 : b) * K + i) * 6, R, j);
    euler_mats(angles + (size_t)idx * 3, Rx, Ry, Rz);
    mat3_mul(Rx, Ry, A);
    mat3_mul(A, Rz, Rot);
    mat3_mul(R, Rot, A);
    mat3_mul_bt(A, R, Q);
    float* o = T_local + (size_t)idx * 12;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        o[r * 4] = Q[r * 3]; o[r * 4 + 1] = Q[r * 3 + 1]; o[r * 4 + 2] = Q[r * 3 + 2];
        o[r * 4 + 3] = j[r] - (Q[r * 3] * j[0] + Q[r * 3 + 1] * j[1] + Q[r * 3 + 2] * j[2]);
    }
}

// G[b,k] = prod over chain (root first) of T_local; posed bones = G [bone,1]  (skinning.py:398-426)
__global__ void lbs_chain_kernel(const float* __restrict__ T_local, const float* __restrict__ bones, const int* __restrict__ chain_ptr,
                                 const int* __restrict__ chain_ids, int B, int Bb, int K, float* __restrict__ G,
                                 float* __restrict__ posed)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * K) return;
    int b = idx / K, k = idx % K;
    int s = chain_ptr[k], e = chain_ptr[k + 1];
    Aff M = aff_load(T_local + ((size_t)b * K + chain_ids[s]) * 12);
    for (int c = s + 1; c < e; c++) M = aff_mul(M, aff_load(T_local + ((size_t)b * K + chain_ids[c]) * 12));
#pragma unroll
    for (int i = 0; i < 12; i++) G[(size_t)idx * 12 + i] = M.m[i];
    if (posed) {
        const float* bn = bones + ((size_t)(Bb == 1 ? 0 : b) * K + k) * 6;
        for (int ep = 0; ep < 2; ep++) {
            float x = bn[ep * 3], y = bn[ep * 3 + 1], z = bn[ep * 3 + 2];
            for (int r = 0; r < 3; r++)
                posed[((size_t)idx * 2 + ep) * 3 + r] = M.m[r * 4] * x + M.m[r * 4 + 1] * y + M.m[r * 4 + 2] * z + M.m[r * 4 + 3];
        }
    }
}

// distance of p to segment (a,b)  (geometry/util.py:30-53)
__device__ __forceinline__ float seg_dist(const float* bn, float px, float py, float pz)
{
    float ax = bn[0], ay = bn[1], az = bn[2];
    float abx = bn[3] - ax, aby = bn[4] - ay, abz = bn[5] - az;
    float t = ((px - ax) * abx + (py - ay) * aby + (pz - az) * abz) / fmaxf(abx * abx + aby * aby + abz * abz, 1e-6f);
    t = fminf(fmaxf(t, 0.f), 1.f);
    float sx = ax + t * abx - px, sy = ay + t * aby - py, sz = az + t * abz - pz;
    return sqrtf(sx * sx + sy * sy + sz * sz + 1e-6f);
}

// Soft bone weights of one vertex: w_k = softmax_k(-dist_k / T).  Distances/exponentials are evaluated once and kept in
// a per-thread column of shared memory (K <= LBS_CACHE_K), so the blend / gradient loops only read them back.
constexpr int LBS_CACHE_K = 32;

template <bool CACHE>
__device__ __forceinline__ float lbs_weights(const float* __restrict__ s_bones, float (*s_w)[LBS_BLOCK], int K, float px, float py, float pz,
                                             float inv_temp, float& xmax)
{
    xmax = -3.4e38f;
    for (int k = 0; k < K; k++) {
        float x = -seg_dist(s_bones + k * 6, px, py, pz) * inv_temp;
        if (CACHE) s_w[k][threadIdx.x] = x;
        xmax = fmaxf(xmax, x);
    }
    float sum = 0.f;
    for (int k = 0; k < K; k++) {
        float x = CACHE ? s_w[k][threadIdx.x] : -seg_dist(s_bones + k * 6, px, py, pz) * inv_temp;
        float e = expf(x - xmax);
        if (CACHE) s_w[k][threadIdx.x] = e;
        sum += e;
    }
    return 1.f / sum;
}

template <bool CACHE>
__global__ void __launch_bounds__(LBS_BLOCK) lbs_fwd_kernel(const float* __restrict__ v_pos, const float* __restrict__ bones,
                                                            const float* __restrict__ G, int B, int Bv, int Bb, int K, int64_t V,
                                                            float inv_temp, float* __restrict__ out, float* __restrict__ weights)
{
    __shared__ float s_bones[LBS_MAX_K * 6];
    __shared__ float s_G[LBS_MAX_K * 12];
    __shared__ float s_w[CACHE ? LBS_CACHE_K : 1][LBS_BLOCK];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < K * 6; i += blockDim.x) s_bones[i] = bones[(size_t)(Bb == 1 ? 0 : b) * K * 6 + i];
    for (int i = threadIdx.x; i < K * 12; i += blockDim.x) s_G[i] = G[(size_t)b * K * 12 + i];
    __syncthreads();
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const float* p = v_pos + ((size_t)(Bv == 1 ? 0 : b) * V + v) * 3;
    float px = p[0], py = p[1], pz = p[2];
    float xmax;
    float inv = lbs_weights<CACHE>(s_bones, s_w, K, px, py, pz, inv_temp, xmax);
    float ox = 0.f, oy = 0.f, oz = 0.f;
    for (int k = 0; k < K; k++) {
        float e = CACHE ? s_w[k][threadIdx.x] : expf(-seg_dist(s_bones + k * 6, px, py, pz) * inv_temp - xmax);
        const float* g = s_G + k * 12;
        ox += e * (g[0] * px + g[1] * py + g[2] * pz + g[3]);
        oy += e * (g[4] * px + g[5] * py + g[6] * pz + g[7]);
        oz += e * (g[8] * px + g[9] * py + g[10] * pz + g[11]);
    }
    float* o = out + ((size_t)b * V + v) * 3;
    o[0] = ox * inv; o[1] = oy * inv; o[2] = oz * inv;
    if (weights) {
        // weights [K,Bw,V]; written once per weight-batch entry (image 0 covers the broadcast case)
        int Bw = max(Bv, Bb);
        if (Bw == B || b == 0) {
            int bw = Bw == 1 ? 0 : b;
            for (int k = 0; k < K; k++) {
                float e = CACHE ? s_w[k][threadIdx.x] : expf(-seg_dist(s_bones + k * 6, px, py, pz) * inv_temp - xmax);
                weights[((size_t)k * Bw + bw) * V + v] = e * inv;
            }
        }
    }
}

// Sum 16 per-lane values over the 32 lanes of a warp with a transposing butterfly: 8+4+2+1+1 = 16 shuffles (a plain
// per-value butterfly needs 80).  Returns, in every lane, the warp total of value index (lane >> 1).
__device__ __forceinline__ float warp_reduce16(const float (&v)[16], int lane)
{
    float a[8], b4[4], c2[2];
    bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float keep = hi ? v[i + 8] : v[i], send = hi ? v[i] : v[i + 8];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float keep = hi ? a[i + 4] : a[i], send = hi ? a[i] : a[i + 4];
        b4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        float keep = hi ? b4[i + 2] : b4[i], send = hi ? b4[i] : b4[i + 2];
        c2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    hi = lane & 2;
    float keep = hi ? c2[1] : c2[0], send = hi ? c2[0] : c2[1];
    float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    return d + __shfl_xor_sync(0xffffffffu, d, 1);
}

// backward: d_v = sum_k w_k R_k^T g ; d_G[b,k] += w_k g (x) [v,1]  (warp-ballot skips bones with no weight in the warp)
template <bool CACHE>
__global__ void __launch_bounds__(LBS_BLOCK) lbs_bwd_kernel(const float* __restrict__ v_pos, const float* __restrict__ bones,
                                                            const float* __restrict__ G, const float* __restrict__ d_out, int B, int Bv,
                                                            int Bb, int K, int64_t V, float inv_temp, float* __restrict__ d_v_pos,
                                                            float* __restrict__ d_G)
{
    __shared__ float s_bones[LBS_MAX_K * 6];
    __shared__ float s_G[LBS_MAX_K * 12];
    __shared__ float s_dG[LBS_MAX_K * 12];
    __shared__ float s_w[CACHE ? LBS_CACHE_K : 1][LBS_BLOCK];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < K * 6; i += blockDim.x) s_bones[i] = bones[(size_t)(Bb == 1 ? 0 : b) * K * 6 + i];
    for (int i = threadIdx.x; i < K * 12; i += blockDim.x) { s_G[i] = G[(size_t)b * K * 12 + i]; s_dG[i] = 0.f; }
    __syncthreads();
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = v < V;
    float px = 0.f, py = 0.f, pz = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
    if (valid) {
        const float* p = v_pos + ((size_t)(Bv == 1 ? 0 : b) * V + v) * 3;
        px = p[0]; py = p[1]; pz = p[2];
        const float* g = d_out + ((size_t)b * V + v) * 3;
        gx = g[0]; gy = g[1]; gz = g[2];
    }
    float xmax;
    float inv = lbs_weights<CACHE>(s_bones, s_w, K, px, py, pz, inv_temp, xmax);
    if (!valid) inv = 0.f;
    float dvx = 0.f, dvy = 0.f, dvz = 0.f;
    const int lane = threadIdx.x & 31;
    for (int k = 0; k < K; k++) {
        float w = (CACHE ? s_w[k][threadIdx.x] : expf(-seg_dist(s_bones + k * 6, px, py, pz) * inv_temp - xmax)) * inv;
        const float* g = s_G + k * 12;
        dvx += w * (g[0] * gx + g[4] * gy + g[8] * gz);
        dvy += w * (g[1] * gx + g[5] * gy + g[9] * gz);
        dvz += w * (g[2] * gx + g[6] * gy + g[10] * gz);
        if (__ballot_sync(0xffffffffu, w > 1e-10f) == 0u) continue;  // < fp32 eps of the dominant terms
        float wx = w * gx, wy = w * gy, wz = w * gz;
        const float r[16] = {wx * px, wx * py, wx * pz, wx, wy * px, wy * py, wy * pz, wy, wz * px, wz * py, wz * pz, wz, 0.f, 0.f, 0.f, 0.f};
        float tot = warp_reduce16(r, lane);
        if (!(lane & 1) && (lane >> 1) < 12) atomicAdd(&s_dG[k * 12 + (lane >> 1)], tot);
    }
    if (valid && d_v_pos) {
        if (Bv == 1 && B > 1) {
            float* o = d_v_pos + (size_t)v * 3;
            atomicAdd(o, dvx); atomicAdd(o + 1, dvy); atomicAdd(o + 2, dvz);
        } else {
            float* o = d_v_pos + ((size_t)b * V + v) * 3;
            o[0] = dvx; o[1] = dvy; o[2] = dvz;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * 12; i += blockDim.x) {
        float s = s_dG[i];
        if (s != 0.f) atomicAdd(&d_G[(size_t)b * K * 12 + i], s);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Shared-weight fast path (Bv == 1 and Bb == 1: one prior shape and one bone set for the whole batch - the reference's
// regime while per-instance deformation is off, InstancePredictorBase.py:514-518).  The soft weights depend only on
// (vertex, bones), so a thread owns a vertex, evaluates its K distances / exponentials ONCE and loops over a chunk of
// LBS_BC images; the generic kernels above redo that transcendental work per (image, vertex).
// ------------------------------------------------------------------------------------------------------------
constexpr int LBS_BC = 4;
constexpr int LBS_SBLOCK = 128;

__device__ __forceinline__ float lbs_weights_shared(const float* __restrict__ s_bones, float (*s_w)[LBS_SBLOCK], int K, float px, float py,
                                                    float pz, float inv_temp)
{
    float xmax = -3.4e38f;
    for (int k = 0; k < K; k++) {
        float x = -seg_dist(s_bones + k * 6, px, py, pz) * inv_temp;
        s_w[k][threadIdx.x] = x;
        xmax = fmaxf(xmax, x);
    }
    float sum = 0.f;
    for (int k = 0; k < K; k++) {
        float e = expf(s_w[k][threadIdx.x] - xmax);
        s_w[k][threadIdx.x] = e;
        sum += e;
    }
    return 1.f / sum;
}

__global__ void __launch_bounds__(LBS_SBLOCK) lbs_fwd_shared_kernel(const float* __restrict__ v_pos, const float* __restrict__ bones,
                                                                    const float* __restrict__ G, int B, int K, int64_t V, float inv_temp,
                                                                    float* __restrict__ out, float* __restrict__ weights)
{
    __shared__ float s_bones[LBS_CACHE_K * 6];
    __shared__ float s_G[LBS_BC * LBS_CACHE_K * 12];
    __shared__ float s_w[LBS_CACHE_K][LBS_SBLOCK];
    const int b0 = blockIdx.y * LBS_BC, nb = min(LBS_BC, B - b0);
    for (int i = threadIdx.x; i < K * 6; i += blockDim.x) s_bones[i] = bones[i];
    for (int i = threadIdx.x; i < nb * K * 12; i += blockDim.x) s_G[i] = G[(size_t)b0 * K * 12 + i];
    __syncthreads();
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const float px = v_pos[v * 3], py = v_pos[v * 3 + 1], pz = v_pos[v * 3 + 2];
    const float inv = lbs_weights_shared(s_bones, s_w, K, px, py, pz, inv_temp);
    for (int bb = 0; bb < nb; bb++) {
        float ox = 0.f, oy = 0.f, oz = 0.f;
        const float* gb = s_G + bb * K * 12;
        for (int k = 0; k < K; k++) {
            const float e = s_w[k][threadIdx.x];
            const float* g = gb + k * 12;
            ox += e * (g[0] * px + g[1] * py + g[2] * pz + g[3]);
            oy += e * (g[4] * px + g[5] * py + g[6] * pz + g[7]);
            oz += e * (g[8] * px + g[9] * py + g[10] * pz + g[11]);
        }
        float* o = out + ((size_t)(b0 + bb) * V + v) * 3;
        o[0] = ox * inv; o[1] = oy * inv; o[2] = oz * inv;
    }
    if (weights && blockIdx.y == 0)
        for (int k = 0; k < K; k++) weights[(size_t)k * V + v] = s_w[k][threadIdx.x] * inv;
}

// Backward of the shared-weight path.  d_G[b,k] = sum_v w[k,v] * g[b,v] (x) [p_v,1] is a small contraction over the
// vertices (K x V times V x 12B): instead of reducing every vertex's outer product across the warp (shuffles + shared
// atomics per (vertex, image, bone) - instruction-bound), each block stages a chunk of LBS_VC vertices in shared
// memory (weights, [p,1], upstream gradients) and thread (k, c) accumulates its 3 x LBS_BC2 outputs in REGISTERS over
// the chunk; one global atomic per output per block at the very end.  d_v = sum_b (sum_k w_k R[b,k])^T g[b] is
// per-vertex work done in the staging phase.  d_v_pos [1,V,3] is accumulated (zero-initialised by the caller).
constexpr int LBS_VC = 128;    // vertices per chunk == threads per block
constexpr int LBS_BC2 = 4;     // images per block

__global__ void __launch_bounds__(LBS_VC, 5) lbs_bwd_shared_kernel(const float* __restrict__ v_pos, const float* __restrict__ bones,
                                                                const float* __restrict__ G, const float* __restrict__ d_out, int B, int K,
                                                                int64_t V, float inv_temp, float* __restrict__ d_v_pos, float* __restrict__ d_G)
{
    __shared__ float s_bones[LBS_CACHE_K * 6];
    __shared__ __align__(16) float s_G[LBS_BC2 * LBS_CACHE_K * 12];
    __shared__ float s_w[LBS_CACHE_K][LBS_VC + 1];
    __shared__ __align__(16) float4 s_g[LBS_BC2][LBS_VC];
    __shared__ __align__(16) float4 s_ph[LBS_VC];
    const int t = threadIdx.x;
    const int b0 = blockIdx.y * LBS_BC2, nb = min(LBS_BC2, B - b0);
    for (int i = t; i < K * 6; i += blockDim.x) s_bones[i] = bones[i];
    for (int i = t; i < nb * K * 12; i += blockDim.x) s_G[i] = G[(size_t)b0 * K * 12 + i];
    const int k2 = t >> 2, c2 = t & 3;           // contraction role: output column c2 of bone k2, all 3 rows, all images
    const bool role = k2 < K;
    float acc[LBS_BC2][3];
#pragma unroll
    for (int bb = 0; bb < LBS_BC2; bb++) acc[bb][0] = acc[bb][1] = acc[bb][2] = 0.f;
    __syncthreads();
    for (int64_t chunk = blockIdx.x; chunk * LBS_VC < V; chunk += gridDim.x) {
        // ---- staging: this thread's vertex ----
        const int64_t v = chunk * LBS_VC + t;
        const bool valid = v < V;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (valid) { px = v_pos[v * 3]; py = v_pos[v * 3 + 1]; pz = v_pos[v * 3 + 2]; }
        float xmax = -3.4e38f;
        for (int k = 0; k < K; k++) {
            float x = -seg_dist(s_bones + k * 6, px, py, pz) * inv_temp;
            s_w[k][t] = x;
            xmax = fmaxf(xmax, x);
        }
        float sum = 0.f;
        for (int k = 0; k < K; k++) {
            float e = expf(s_w[k][t] - xmax);
            s_w[k][t] = e;
            sum += e;
        }
        const float inv = valid ? 1.f / sum : 0.f;
        for (int k = 0; k < K; k++) s_w[k][t] *= inv;
        s_ph[t] = make_float4(px, py, pz, 1.f);
        float dvx = 0.f, dvy = 0.f, dvz = 0.f;
        for (int bb = 0; bb < nb; bb++) {
            float gx = 0.f, gy = 0.f, gz = 0.f;
            if (valid) {
                const float* g = d_out + ((size_t)(b0 + bb) * V + v) * 3;
                gx = g[0]; gy = g[1]; gz = g[2];
            }
            s_g[bb][t] = make_float4(gx, gy, gz, 0.f);
            // blended rotation of this vertex in image bb, then its transpose applied to g
            float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f, m5 = 0.f, m6 = 0.f, m7 = 0.f, m8 = 0.f;
            const float4* Gb = reinterpret_cast<const float4*>(s_G + (size_t)bb * K * 12);
            for (int k = 0; k < K; k++) {
                const float w = s_w[k][t];
                const float4 r0 = Gb[k * 3], r1 = Gb[k * 3 + 1], r2 = Gb[k * 3 + 2];
                m0 += w * r0.x; m1 += w * r0.y; m2 += w * r0.z;
                m3 += w * r1.x; m4 += w * r1.y; m5 += w * r1.z;
                m6 += w * r2.x; m7 += w * r2.y; m8 += w * r2.z;
            }
            dvx += m0 * gx + m3 * gy + m6 * gz;
            dvy += m1 * gx + m4 * gy + m7 * gz;
            dvz += m2 * gx + m5 * gy + m8 * gz;
        }
        if (valid && d_v_pos) {
            float* o = d_v_pos + (size_t)v * 3;
            atomicAdd(o, dvx); atomicAdd(o + 1, dvy); atomicAdd(o + 2, dvz);
        }
        __syncthreads();
        // ---- contraction over the chunk ----
        if (role) {
            const float* phc = reinterpret_cast<const float*>(s_ph) + c2;
#pragma unroll 4
            for (int vv = 0; vv < LBS_VC; vv++) {
                const float wp = s_w[k2][vv] * phc[vv * 4];
#pragma unroll
                for (int bb = 0; bb < LBS_BC2; bb++) {
                    if (bb < nb) {
                        const float4 g4 = s_g[bb][vv];
                        acc[bb][0] += wp * g4.x; acc[bb][1] += wp * g4.y; acc[bb][2] += wp * g4.z;
                    }
                }
            }
        }
        __syncthreads();
    }
    if (role) {
#pragma unroll
        for (int bb = 0; bb < LBS_BC2; bb++) {
            if (bb < nb) {
                float* o = d_G + ((size_t)(b0 + bb) * K + k2) * 12 + c2;
                if (acc[bb][0] != 0.f) atomicAdd(o, acc[bb][0]);
                if (acc[bb][1] != 0.f) atomicAdd(o + 4, acc[bb][1]);
                if (acc[bb][2] != 0.f) atomicAdd(o + 8, acc[bb][2]);
            }
        }
    }
}

// chain backward: d_T_local[b, c_j] += P_j^T dG S_j^T  (affine algebra), after folding d_posed into dG
__global__ void lbs_chain_bwd_kernel(const float* __restrict__ T_local, const float* __restrict__ bones, const int* __restrict__ chain_ptr,
                                     const int* __restrict__ chain_ids, float* __restrict__ d_G, const float* __restrict__ d_posed,
                                     int B, int Bb, int K, float* __restrict__ d_T_local)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * K) return;
    int b = idx / K, k = idx % K;
    float dG[12];
#pragma unroll
    for (int i = 0; i < 12; i++) dG[i] = d_G[(size_t)idx * 12 + i];
    if (d_posed) {
        const float* bn = bones + ((size_t)(Bb == 1 ? 0 : b) * K + k) * 6;
        for (int ep = 0; ep < 2; ep++)
            for (int r = 0; r < 3; r++) {
                float g = d_posed[((size_t)idx * 2 + ep) * 3 + r];
                dG[r * 4] += g * bn[ep * 3]; dG[r * 4 + 1] += g * bn[ep * 3 + 1]; dG[r * 4 + 2] += g * bn[ep * 3 + 2]; dG[r * 4 + 3] += g;
            }
    }
    int s = chain_ptr[k], n = chain_ptr[k + 1] - s;
    if (n > LBS_MAX_CHAIN) n = LBS_MAX_CHAIN;  // guarded on the host
    Aff P[LBS_MAX_CHAIN];
    P[0] = aff_identity();
    for (int j = 1; j < n; j++) P[j] = aff_mul(P[j - 1], aff_load(T_local + ((size_t)b * K + chain_ids[s + j - 1]) * 12));
    Aff S = aff_identity();
    for (int j = n - 1; j >= 0; j--) {
        int cj = chain_ids[s + j];
        // G = X T Y with X = P[j] (Rp,tp), Y = S (Rs,ts):  dR = Rp^T dG.R Rs^T + (Rp^T dG.t) ts^T ; dt = Rp^T dG.t
        const float* X = P[j].m;
        float A[9], dt[3], dR[9];
        float dGR[9] = {dG[0], dG[1], dG[2], dG[4], dG[5], dG[6], dG[8], dG[9], dG[10]};
        float XR[9] = {X[0], X[1], X[2], X[4], X[5], X[6], X[8], X[9], X[10]};
        float SR[9] = {S.m[0], S.m[1], S.m[2], S.m[4], S.m[5], S.m[6], S.m[8], S.m[9], S.m[10]};
        mat3_mul_at(XR, dGR, A);
        mat3_mul_bt(A, SR, dR);
        for (int r = 0; r < 3; r++) dt[r] = XR[r] * dG[3] + XR[3 + r] * dG[7] + XR[6 + r] * dG[11];
        float* o = d_T_local + ((size_t)b * K + cj) * 12;
        for (int r = 0; r < 3; r++) {
            atomicAdd(o + r * 4 + 0, dR[r * 3 + 0] + dt[r] * S.m[3]);
            atomicAdd(o + r * 4 + 1, dR[r * 3 + 1] + dt[r] * S.m[7]);
            atomicAdd(o + r * 4 + 2, dR[r * 3 + 2] + dt[r] * S.m[11]);
            atomicAdd(o + r * 4 + 3, dt[r]);
        }
        S = aff_mul(aff_load(T_local + ((size_t)b * K + cj) * 12), S);
    }
}

// local backward: d_T_local -> d_angles   (Q = R Rot R^T, t = j - Q j, Rot = Rx Ry Rz)
__global__ void lbs_local_bwd_kernel(const float* __restrict__ bones, const float* __restrict__ angles, const float* __restrict__ d_T_local,
                                     int B, int Bb, int K, float* __restrict__ d_angles)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * K) return;
    int b = idx / K, i = idx % K;
    float R[9], j[3], Rx[9], Ry[9], Rz[9];
    bone_frame(bones + ((size_t)(Bb == 1 ? 0 : b) * K + i) * 6, R, j);
    const float* th = angles + (size_t)idx * 3;
    euler_mats(th, Rx, Ry, Rz);
    const float* g = d_T_local + (size_t)idx * 12;
    float dQ[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) dQ[r * 3 + c] = g[r * 4 + c] - g[r * 4 + 3] * j[c];
    float A[9], dRot[9];
    mat3_mul_at(R, dQ, A);
    mat3_mul(A, R, dRot);
    float YZ[9], XY[9], dA[9], dB[9], dC[9], tmp[9];
    mat3_mul(Ry, Rz, YZ);
    mat3_mul(Rx, Ry, XY);
    mat3_mul_bt(dRot, YZ, dA);      // dRx = dRot (Ry Rz)^T
    mat3_mul_at(Rx, dRot, tmp);     // Rx^T dRot
    mat3_mul_bt(tmp, Rz, dB);       // dRy = Rx^T dRot Rz^T
    mat3_mul_at(XY, dRot, dC);      // dRz = (Rx Ry)^T dRot
    float sx = Rx[7], cx = Rx[4], sy = Ry[2], cy = Ry[0], sz = Rz[3], cz = Rz[0];
    float* o = d_angles + (size_t)idx * 3;
    o[0] = dA[4] * (-sx) + dA[5] * (-cx) + dA[7] * cx + dA[8] * (-sx);
    o[1] = dB[0] * (-sy) + dB[2] * cy + dB[6] * (-cy) + dB[8] * (-sy);
    o[2] = dC[0] * (-sz) + dC[1] * (-cz) + dC[3] * cz + dC[4] * (-sz);
}

}  // namespace

B2A_API int b2a_lbs_bone_transforms(const float* bones, const float* angles, const int32_t* chain_ptr, const int32_t* chain_ids,
                                    int B, int Bb, int K, float* T_local, float* G, float* posed_bones, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(bones && angles && chain_ptr && chain_ids && T_local && G, "null pointer");
    B2A_CHECK_ARG(B > 0 && K > 0 && K <= LBS_MAX_K && (Bb == 1 || Bb == B), "shape");
    int n = B * K;
    lbs_local_kernel<<<b2a_blocks(n, 128), 128, 0, stream>>>(bones, angles, B, Bb, K, T_local);
    lbs_chain_kernel<<<b2a_blocks(n, 128), 128, 0, stream>>>(T_local, bones, chain_ptr, chain_ids, B, Bb, K, G, posed_bones);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_lbs_fwd(const float* v_pos, const float* bones, const float* G, int B, int Bv, int Bb, int K, int64_t V,
                        float inv_temperature, float* out, float* weights, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(v_pos && bones && G && out, "null pointer");
    B2A_CHECK_ARG(B > 0 && B <= 65535 && K > 0 && K <= LBS_MAX_K && (Bb == 1 || Bb == B) && (Bv == 1 || Bv == B), "shape");
    if (V == 0) return 0;
    if (Bv == 1 && Bb == 1 && B > 1 && K <= LBS_CACHE_K) {
        lbs_fwd_shared_kernel<<<dim3(b2a_blocks(V, LBS_SBLOCK), (B + LBS_BC - 1) / LBS_BC), LBS_SBLOCK, 0, stream>>>(v_pos, bones, G, B, K, V,
                                                                                                                 inv_temperature, out, weights);
        B2A_LAUNCH_OK();
        return 0;
    }
    dim3 grid(b2a_blocks(V, LBS_BLOCK), B);
    if (K <= LBS_CACHE_K) lbs_fwd_kernel<true><<<grid, LBS_BLOCK, 0, stream>>>(v_pos, bones, G, B, Bv, Bb, K, V, inv_temperature, out, weights);
    else lbs_fwd_kernel<false><<<grid, LBS_BLOCK, 0, stream>>>(v_pos, bones, G, B, Bv, Bb, K, V, inv_temperature, out, weights);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_lbs_bwd(const float* v_pos, const float* bones, const float* G, const float* d_out, int B, int Bv, int Bb, int K,
                        int64_t V, float inv_temperature, float* d_v_pos, float* d_G, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(v_pos && bones && G && d_out && d_G, "null pointer");
    B2A_CHECK_ARG(B > 0 && B <= 65535 && K > 0 && K <= LBS_MAX_K && (Bb == 1 || Bb == B) && (Bv == 1 || Bv == B), "shape");
    if (V == 0) return 0;
    if (Bv == 1 && Bb == 1 && B > 1 && K <= LBS_CACHE_K) {
        unsigned chunks = b2a_blocks(V, LBS_VC);
        unsigned by = (B + LBS_BC2 - 1) / LBS_BC2;
        unsigned cap = 148u * 6u / by;      // ~6 resident blocks per SM
        dim3 sgrid(chunks < cap ? chunks : (cap ? cap : 1u), by);
        lbs_bwd_shared_kernel<<<sgrid, LBS_VC, 0, stream>>>(v_pos, bones, G, d_out, B, K, V, inv_temperature, d_v_pos, d_G);
        B2A_LAUNCH_OK();
        return 0;
    }
    dim3 grid(b2a_blocks(V, LBS_BLOCK), B);
    if (K <= LBS_CACHE_K) lbs_bwd_kernel<true><<<grid, LBS_BLOCK, 0, stream>>>(v_pos, bones, G, d_out, B, Bv, Bb, K, V, inv_temperature, d_v_pos, d_G);
    else lbs_bwd_kernel<false><<<grid, LBS_BLOCK, 0, stream>>>(v_pos, bones, G, d_out, B, Bv, Bb, K, V, inv_temperature, d_v_pos, d_G);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_lbs_bone_transforms_bwd(const float* bones, const float* angles, const int32_t* chain_ptr, const int32_t* chain_ids,
                                        const float* T_local, float* d_G, const float* d_posed_bones, int B, int Bb, int K,
                                        float* d_T_local, float* d_angles, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(bones && angles && chain_ptr && chain_ids && T_local && d_G && d_T_local && d_angles, "null pointer");
    B2A_CHECK_ARG(B > 0 && K > 0 && K <= LBS_MAX_K && (Bb == 1 || Bb == B), "shape");
    int n = B * K;
    lbs_chain_bwd_kernel<<<b2a_blocks(n, 64), 64, 0, stream>>>(T_local, bones, chain_ptr, chain_ids, d_G, d_posed_bones, B, Bb, K, d_T_local);
    lbs_local_bwd_kernel<<<b2a_blocks(n, 128), 128, 0, stream>>>(bones, angles, d_T_local, B, Bb, K, d_angles);
    B2A_LAUNCH_OK();
    return 0;
}
