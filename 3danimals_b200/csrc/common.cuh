// common.cuh - shared helpers for libb2a.so (sm_100a).  Compiled with -fmad=false: every fp32 op in decision-
// making code is individually rounded, so index buffers (triangle ids, faces) are bit-reproducible against the
// CPU oracle; plain `a*b+c` below therefore never contracts to FMA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b2a.h"

#define B2A_API extern "C" __attribute__((visibility("default")))

void b2a_set_error(const char* fmt, ...);

#define B2A_CHECK_ARG(cond, msg)                                        \
    do {                                                                \
        if (!(cond)) {                                                  \
            b2a_set_error("%s: invalid argument: %s", __func__, msg);   \
            return 2;                                                   \
        }                                                               \
    } while (0)

#define B2A_CUDA_OK(expr)                                                                  \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            b2a_set_error("%s: CUDA error %s (%s)", __func__, cudaGetErrorName(e__), #expr); \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

#define B2A_LAUNCH_OK() B2A_CUDA_OK(cudaGetLastError())

static inline unsigned b2a_blocks(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }
static inline size_t b2a_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Exclusive prefix of `v` over a block of up to 1024 threads; returns prefix, writes block total to *total.
// smem: 33 ints.
__device__ __forceinline__ int block_exclusive_scan(int v, int* smem, int* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nwarp ? smem[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        smem[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    int res = smem[warp] + inc - v;
    *total = smem[32];
    __syncthreads();
    return res;
}

// ---- z-buffer key helpers (rasterizer) -------------------------------------------------------------------
__device__ __forceinline__ uint32_t depth_key(float f)
{
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct TriEval {
    float a0, a1, a2, S, zw, u, v;
};

// Bit-identical restatement target: oracle/raster_ref.c tri_eval (see its header for the fill rule).
__device__ __forceinline__ bool tri_eval(const float4 p0, const float4 p1, const float4 p2, float fx, float fy,
                                         TriEval& e)
{
    float q0x = p0.x - fx * p0.w, q0y = p0.y - fy * p0.w;
    float q1x = p1.x - fx * p1.w, q1y = p1.y - fy * p1.w;
    float q2x = p2.x - fx * p2.w, q2y = p2.y - fy * p2.w;
    float a0 = q1x * q2y - q1y * q2x;
    float a1 = q2x * q0y - q2y * q0x;
    float a2 = q0x * q1y - q0y * q1x;
    float S = (a0 + a1) + a2;
    if (S > 0.f) {
        if (a0 < 0.f || a1 < 0.f || a2 < 0.f) return false;
    } else if (S < 0.f) {
        if (a0 > 0.f || a1 > 0.f || a2 > 0.f) return false;
    } else
        return false;
    float z = (p0.z * a0 + p1.z * a1) + p2.z * a2;
    float w = (p0.w * a0 + p1.w * a1) + p2.w * a2;
    if (S > 0.f ? !(w > 0.f) : !(w < 0.f)) return false;
    float zw = __fdiv_rn(z, w);
    if (!(zw >= -1.f && zw <= 1.f)) return false;
    float iw = __fdiv_rn(1.f, S);
    float u = a0 * iw, v = a1 * iw;
    e.a0 = a0; e.a1 = a1; e.a2 = a2; e.S = S; e.zw = zw;
    e.u = u < 0.f ? 0.f : (u > 1.f ? 1.f : u);
    e.v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
    return true;
}

__device__ __forceinline__ void pixel_ndc(int px, int py, int H, int W, float& fx, float& fy)
{
    float xs = __fdiv_rn(2.f, (float)W), ys = __fdiv_rn(2.f, (float)H);
    fx = (float)px * xs + (xs * 0.5f - 1.f);
    fy = (float)py * ys + (ys * 0.5f - 1.f);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Workspace layout shared by b2a_rasterize_fwd / b2a_rasterize_zbuffer / b2a_gbuffer_fwd.
struct RasterWorkspace {
    unsigned long long* zbuf;  // [B*H*W] keys
    int* queue_count;          // [1] (+ padding)
    int2* queue;               // [queue_cap] (b, f) of large triangles
    int64_t queue_cap;
};
static inline size_t raster_workspace_layout(int B, int64_t F, int H, int W, void* base, RasterWorkspace* ws)
{
    size_t off = 0;
    size_t zb = b2a_align((size_t)B * H * W * sizeof(unsigned long long));
    int64_t cap = (int64_t)B * F;
    if (cap > (1 << 22)) cap = (1 << 22);
    if (ws) {
        char* p = (char*)base;
        ws->zbuf = (unsigned long long*)(p + off);
        ws->queue_count = (int*)(p + off + zb);
        ws->queue = (int2*)(p + off + zb + 256);
        ws->queue_cap = cap;
    }
    off += zb + 256 + b2a_align((size_t)cap * sizeof(int2));
    return off;
}
