// raster.cu - clip transform and z-buffer triangle rasterizer on sm_100a.
// Replaces ru.xfm_points(use_python=True) (reference model/render/renderutils/ops.py:524-525) and
// nvdiffrast.torch.rasterize / DepthPeeler first layer (call sites model/render/render.py:292-294, :351).
//
// Design (B200-first): DMTet meshes at training resolution have MORE triangles than covered pixels (res-128 grid:
// ~71k faces vs ~20k covered pixels of a 256^2 image), so the rasterizer is triangle-bound, not pixel-bound.  One
// thread per (image, triangle) evaluates the homogeneous edge functions at the handful of pixel centres inside the
// triangle's bounding box and resolves visibility with a 64-bit atomicMin on (depth key << 32 | triangle id); the
// whole-batch z-buffer (8 B/pixel) lives in the 126 MB L2.  Triangles with large bounding boxes are queued and
// re-done by whole warps, lanes striding the box (edge set-up broadcast by shuffle).  A resolve pass turns the
// keys into (u, v, z/w, id+1).  Fill rule and arithmetic: oracle/raster_ref.c header - bit-identical ids.
#include "common.cuh"

namespace {

constexpr int SMALL_BBOX_PIXELS = 32;

// ------------------------------------------------------------------------------------------------------------
// clip transform
// ------------------------------------------------------------------------------------------------------------
__global__ void xfm_fwd_kernel(const float* __restrict__ pts, const float* __restrict__ mtx, int Bp, int64_t V, float* __restrict__ out)
{
    __shared__ float m[16];
    const int b = blockIdx.y;
    if (threadIdx.x < 16) m[threadIdx.x] = mtx[(size_t)b * 16 + threadIdx.x];
    __syncthreads();
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const float* p = pts + ((size_t)(Bp == 1 ? 0 : b) * V + v) * 3;
    float x = p[0], y = p[1], z = p[2];
    float4 o;
    o.x = ((m[0] * x + m[1] * y) + m[2] * z) + m[3];
    o.y = ((m[4] * x + m[5] * y) + m[6] * z) + m[7];
    o.z = ((m[8] * x + m[9] * y) + m[10] * z) + m[11];
    o.w = ((m[12] * x + m[13] * y) + m[14] * z) + m[15];
    reinterpret_cast<float4*>(out)[(size_t)b * V + v] = o;
}

// d_out2 (nullable): a second upstream gradient added to d_out (the fused render node sums the antialias and g-buffer
// contributions here instead of in a separate add launch); accumulate: d_pts += instead of = (Bp == B case)
__global__ void xfm_bwd_kernel(const float* __restrict__ pts, const float* __restrict__ mtx, const float* __restrict__ d_out,
                               const float* __restrict__ d_out2, int accumulate, int B, int Bp, int64_t V, float* __restrict__ d_pts,
                               float* __restrict__ d_mtx)
{
    __shared__ float m[16];
    __shared__ float acc[16];
    const int b = blockIdx.y;
    if (threadIdx.x < 16) { m[threadIdx.x] = mtx[(size_t)b * 16 + threadIdx.x]; acc[threadIdx.x] = 0.f; }
    __syncthreads();
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    float x = 0.f, y = 0.f, z = 0.f, one = 0.f;
    if (v < V) {
        g = reinterpret_cast<const float4*>(d_out)[(size_t)b * V + v];
        if (d_out2) {
            const float4 g2 = reinterpret_cast<const float4*>(d_out2)[(size_t)b * V + v];
            g.x += g2.x; g.y += g2.y; g.z += g2.z; g.w += g2.w;
        }
        if (d_mtx) {
            const float* p = pts + ((size_t)(Bp == 1 ? 0 : b) * V + v) * 3;
            x = p[0]; y = p[1]; z = p[2]; one = 1.f;
        }
        if (d_pts) {
            float dx = m[0] * g.x + m[4] * g.y + m[8] * g.z + m[12] * g.w;
            float dy = m[1] * g.x + m[5] * g.y + m[9] * g.z + m[13] * g.w;
            float dz = m[2] * g.x + m[6] * g.y + m[10] * g.z + m[14] * g.w;
            if (Bp == 1 && B > 1) {
                float* o = d_pts + (size_t)v * 3;
                atomicAdd(o, dx); atomicAdd(o + 1, dy); atomicAdd(o + 2, dz);
            } else {
                float* o = d_pts + ((size_t)b * V + v) * 3;
                if (accumulate) { o[0] += dx; o[1] += dy; o[2] += dz; }
                else { o[0] = dx; o[1] = dy; o[2] = dz; }
            }
        }
    }
    if (d_mtx) {
        if (__ballot_sync(0xffffffffu, g.x != 0.f || g.y != 0.f || g.z != 0.f || g.w != 0.f)) {
            float gr[4] = {g.x, g.y, g.z, g.w}, h[4] = {x, y, z, one};
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    float s = warp_sum(gr[r] * h[c]);
                    if ((threadIdx.x & 31) == 0) atomicAdd(&acc[r * 4 + c], s);
                }
        }
        __syncthreads();
        if (threadIdx.x < 16 && acc[threadIdx.x] != 0.f) atomicAdd(&d_mtx[(size_t)b * 16 + threadIdx.x], acc[threadIdx.x]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// rasterizer
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool tri_bbox(const float4 p0, const float4 p1, const float4 p2, int H, int W, int& x0, int& x1, int& y0, int& y1)
{
    if (!(p0.w > 0.f) && !(p1.w > 0.f) && !(p2.w > 0.f)) return false;
    if (p0.w > 1e-6f && p1.w > 1e-6f && p2.w > 1e-6f) {
        float fW = (float)W, fH = (float)H;
        float sx0 = (p0.x / p0.w * 0.5f + 0.5f) * fW, sy0 = (p0.y / p0.w * 0.5f + 0.5f) * fH;
        float sx1 = (p1.x / p1.w * 0.5f + 0.5f) * fW, sy1 = (p1.y / p1.w * 0.5f + 0.5f) * fH;
        float sx2 = (p2.x / p2.w * 0.5f + 0.5f) * fW, sy2 = (p2.y / p2.w * 0.5f + 0.5f) * fH;
        float mnx = fminf(sx0, fminf(sx1, sx2)), mxx = fmaxf(sx0, fmaxf(sx1, sx2));
        float mny = fminf(sy0, fminf(sy1, sy2)), mxy = fmaxf(sy0, fmaxf(sy1, sy2));
        if (!(mxx >= 0.f && mnx <= fW && mxy >= 0.f && mny <= fH)) return false;
        // pixel centre px+.5 in [mn,mx] -> px in [mn-.5, mx-.5], widened by 1/16 px (>> fp32 projection error, ~1e-4 px;
        // the oracle uses a full pixel; the box only has to be conservative, coverage itself is decided by tri_eval).
        // Rounding INWARD (ceil / floor): most DMTet triangles are sub-pixel and contain no pixel centre at all - an
        // outward-rounded box made them evaluate 4 pixels each.
        x0 = (int)fmaxf(ceilf(mnx - 0.5625f), 0.f);
        x1 = (int)fminf(floorf(mxx - 0.4375f), fW - 1.f);
        y0 = (int)fmaxf(ceilf(mny - 0.5625f), 0.f);
        y1 = (int)fminf(floorf(mxy - 0.4375f), fH - 1.f);
        return x0 <= x1 && y0 <= y1;
    }
    x0 = 0; x1 = W - 1; y0 = 0; y1 = H - 1;
    return true;
}

__device__ __forceinline__ void zbuf_test(unsigned long long* __restrict__ zb, const float4 p0, const float4 p1, const float4 p2,
                                          int px, int py, int H, int W, int f)
{
    float fx, fy;
    pixel_ndc(px, py, H, W, fx, fy);
    TriEval e;
    if (!tri_eval(p0, p1, p2, fx, fy, e)) return;
    unsigned long long key = ((unsigned long long)depth_key(e.zw) << 32) | (unsigned)f;
    atomicMin(zb + (size_t)py * W + px, key);
}

__global__ void __launch_bounds__(256) raster_scatter_kernel(const float* __restrict__ pos, const int* __restrict__ tri, int64_t V, int64_t F,
                                                             int H, int W, RasterWorkspace ws)
{
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int b = blockIdx.y;
    int i0 = __ldg(tri + f * 3), i1 = __ldg(tri + f * 3 + 1), i2 = __ldg(tri + f * 3 + 2);
    if ((unsigned)i0 >= (unsigned)V || (unsigned)i1 >= (unsigned)V || (unsigned)i2 >= (unsigned)V) return;
    const float* pb = pos + (size_t)b * V * 4;
    float4 p0 = ldg4(pb + (size_t)i0 * 4), p1 = ldg4(pb + (size_t)i1 * 4), p2 = ldg4(pb + (size_t)i2 * 4);
    int x0, x1, y0, y1;
    if (!tri_bbox(p0, p1, p2, H, W, x0, x1, y0, y1)) return;
    int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
    if (bw * bh > SMALL_BBOX_PIXELS) {
        int slot = atomicAdd(ws.queue_count, 1);
        if (slot < ws.queue_cap) { ws.queue[slot] = make_int2(b, (int)f); return; }
        // queue full: fall through and do it here (correct, just slow)
    }
    unsigned long long* zb = ws.zbuf + (size_t)b * H * W;
    for (int py = y0; py <= y1; py++)
        for (int px = x0; px <= x1; px++) zbuf_test(zb, p0, p1, p2, px, py, H, W, (int)f);
}

// one warp per queued large triangle; lane 0 fetches the triangle, set-up is broadcast by shuffle, lanes stride pixels
__global__ void __launch_bounds__(256) raster_large_kernel(const float* __restrict__ pos, const int* __restrict__ tri, int64_t V, int H, int W,
                                                           RasterWorkspace ws)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    int count = *ws.queue_count;
    if (count > ws.queue_cap) count = (int)ws.queue_cap;
    for (int item = blockIdx.x * warps_per_block + (threadIdx.x >> 5); item < count; item += gridDim.x * warps_per_block) {
        float v[12];
        int bf0 = 0, bf1 = 0;
        if (lane == 0) {
            int2 q = ws.queue[item];
            bf0 = q.x; bf1 = q.y;
            const float* pb = pos + (size_t)q.x * V * 4;
            float4 a = ldg4(pb + (size_t)__ldg(tri + (size_t)q.y * 3) * 4);
            float4 c = ldg4(pb + (size_t)__ldg(tri + (size_t)q.y * 3 + 1) * 4);
            float4 d = ldg4(pb + (size_t)__ldg(tri + (size_t)q.y * 3 + 2) * 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
            v[8] = d.x; v[9] = d.y; v[10] = d.z; v[11] = d.w;
        }
        bf0 = __shfl_sync(0xffffffffu, bf0, 0);
        bf1 = __shfl_sync(0xffffffffu, bf1, 0);
#pragma unroll
        for (int i = 0; i < 12; i++) v[i] = __shfl_sync(0xffffffffu, v[i], 0);
        float4 p0 = make_float4(v[0], v[1], v[2], v[3]), p1 = make_float4(v[4], v[5], v[6], v[7]), p2 = make_float4(v[8], v[9], v[10], v[11]);
        int x0, x1, y0, y1;
        if (!tri_bbox(p0, p1, p2, H, W, x0, x1, y0, y1)) continue;
        int bw = x1 - x0 + 1, n = bw * (y1 - y0 + 1);
        unsigned long long* zb = ws.zbuf + (size_t)bf0 * H * W;
        for (int i = lane; i < n; i += 32) zbuf_test(zb, p0, p1, p2, x0 + i % bw, y0 + i / bw, H, W, bf1);
    }
}

// keys -> (u, v, z/w, id+1); optionally appends the covered pixels to a compact list of 16-byte entries
// (flat pixel index b*H*W + p, vertex ids i0, i1, i2) that lets the g-buffer backward run dense warps and start its
// vertex gathers without the rast -> triangle -> index chain (warp-aggregated append: one atomic per warp)
__global__ void __launch_bounds__(256) raster_resolve_kernel(const unsigned long long* __restrict__ zbuf, const float* __restrict__ pos,
                                                             const int* __restrict__ tri, int64_t V, int H, int W, float* __restrict__ rast,
                                                             int4* __restrict__ cov_list, int* __restrict__ cov_count)
{
    int ip = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const bool in = ip < H * W;
    bool covered = false;
    int i0 = 0, i1 = 0, i2 = 0;
    size_t pi = (size_t)b * H * W + ip;
    if (in) {
        int px = ip % W, py = ip / W;
        unsigned long long key = zbuf[pi];
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (key != 0xffffffffffffffffull) {
            int f = (int)(unsigned)key;
            const float* pb = pos + (size_t)b * V * 4;
            i0 = __ldg(tri + (size_t)f * 3); i1 = __ldg(tri + (size_t)f * 3 + 1); i2 = __ldg(tri + (size_t)f * 3 + 2);
            float4 p0 = ldg4(pb + (size_t)i0 * 4);
            float4 p1 = ldg4(pb + (size_t)i1 * 4);
            float4 p2 = ldg4(pb + (size_t)i2 * 4);
            float fx, fy;
            pixel_ndc(px, py, H, W, fx, fy);
            TriEval e;
            if (tri_eval(p0, p1, p2, fx, fy, e)) { o = make_float4(e.u, e.v, e.zw, (float)(f + 1)); covered = true; }
        }
        reinterpret_cast<float4*>(rast)[pi] = o;
    }
    if (cov_list) {
        const unsigned m = __ballot_sync(0xffffffffu, covered);
        if (m) {
            const int lane = threadIdx.x & 31;
            int base = 0;
            if (lane == 0) base = atomicAdd(cov_count, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (covered) cov_list[base + __popc(m & ((1u << lane) - 1u))] = make_int4((int)pi, i0, i1, i2);
        }
    }
}

// d(u,v) -> d(x,y,w) of the three vertices
__global__ void __launch_bounds__(256) raster_bwd_kernel(const float* __restrict__ pos, const int* __restrict__ tri, const float* __restrict__ rast,
                                                         const float* __restrict__ d_rast, int64_t V, int64_t F, int H, int W,
                                                         float* __restrict__ d_pos)
{
    int ip = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (ip >= H * W) return;
    int px = ip % W, py = ip / W;
    size_t pi = (size_t)b * H * W + ip;
    float4 r = ldg4(rast + pi * 4);
    int f = (int)r.w - 1;
    if (f < 0 || f >= F) return;
    float4 g = ldg4(d_rast + pi * 4);
    if (g.x == 0.f && g.y == 0.f) return;
    int vi0 = __ldg(tri + (size_t)f * 3), vi1 = __ldg(tri + (size_t)f * 3 + 1), vi2 = __ldg(tri + (size_t)f * 3 + 2);
    const float* pb = pos + (size_t)b * V * 4;
    float4 p0 = ldg4(pb + (size_t)vi0 * 4), p1 = ldg4(pb + (size_t)vi1 * 4), p2 = ldg4(pb + (size_t)vi2 * 4);
    float fx, fy;
    pixel_ndc(px, py, H, W, fx, fy);
    float q0x = p0.x - fx * p0.w, q0y = p0.y - fy * p0.w;
    float q1x = p1.x - fx * p1.w, q1y = p1.y - fy * p1.w;
    float q2x = p2.x - fx * p2.w, q2y = p2.y - fy * p2.w;
    float a0 = q1x * q2y - q1y * q2x, a1 = q2x * q0y - q2y * q0x, a2 = q0x * q1y - q0y * q1x;
    float iw = 1.f / ((a0 + a1) + a2);
    float u = a0 * iw, v = a1 * iw;
    float gs = u * g.x + v * g.y;
    float ga0 = (g.x - gs) * iw, ga1 = (g.y - gs) * iw, ga2 = -gs * iw;
    float gq0x = ga2 * q1y - ga1 * q2y, gq0y = ga1 * q2x - ga2 * q1x;
    float gq1x = ga0 * q2y - ga2 * q0y, gq1y = ga2 * q0x - ga0 * q2x;
    float gq2x = ga1 * q0y - ga0 * q1y, gq2y = ga0 * q1x - ga1 * q0x;
    float* gb = d_pos + (size_t)b * V * 4;
    atomicAdd(gb + (size_t)vi0 * 4, gq0x); atomicAdd(gb + (size_t)vi0 * 4 + 1, gq0y); atomicAdd(gb + (size_t)vi0 * 4 + 3, -(fx * gq0x + fy * gq0y));
    atomicAdd(gb + (size_t)vi1 * 4, gq1x); atomicAdd(gb + (size_t)vi1 * 4 + 1, gq1y); atomicAdd(gb + (size_t)vi1 * 4 + 3, -(fx * gq1x + fy * gq1y));
    atomicAdd(gb + (size_t)vi2 * 4, gq2x); atomicAdd(gb + (size_t)vi2 * 4 + 1, gq2y); atomicAdd(gb + (size_t)vi2 * 4 + 3, -(fx * gq2x + fy * gq2y));
}

int raster_zbuffer_impl(const float* pos, const int32_t* tri, int B, int64_t V, int64_t F, int H, int W, void* workspace,
                        size_t workspace_bytes, RasterWorkspace* ws, cudaStream_t stream)
{
    B2A_CHECK_ARG(pos && tri && workspace, "null pointer");
    B2A_CHECK_ARG(B > 0 && B <= 65535 && V > 0 && F >= 0 && H > 0 && W > 0 && (int64_t)H * W < (1ll << 31) && F < (1ll << 31), "shape");
    B2A_CHECK_ARG(((uintptr_t)pos & 15) == 0, "pos must be 16-byte aligned");
    B2A_CHECK_ARG(raster_workspace_layout(B, F, H, W, workspace, ws) <= workspace_bytes, "workspace too small");
    B2A_CUDA_OK(cudaMemsetAsync(ws->zbuf, 0xff, (size_t)B * H * W * sizeof(unsigned long long), stream));
    B2A_CUDA_OK(cudaMemsetAsync(ws->queue_count, 0, sizeof(int), stream));
    if (F > 0) {
        raster_scatter_kernel<<<dim3(b2a_blocks(F, 256), B), 256, 0, stream>>>(pos, tri, V, F, H, W, *ws);
        raster_large_kernel<<<148 * 4, 256, 0, stream>>>(pos, tri, V, H, W, *ws);
    }
    B2A_LAUNCH_OK();
    return 0;
}

}  // namespace

B2A_API int b2a_xfm_points_fwd(const float* pts, const float* mtx, int B, int Bp, int64_t V, float* out, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(pts && mtx && out, "null pointer");
    B2A_CHECK_ARG(B > 0 && B <= 65535 && (Bp == 1 || Bp == B) && V >= 0, "shape");
    B2A_CHECK_ARG(((uintptr_t)out & 15) == 0, "out must be 16-byte aligned");
    if (V > 0) xfm_fwd_kernel<<<dim3(b2a_blocks(V, 256), B), 256, 0, stream>>>(pts, mtx, Bp, V, out);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_xfm_points_bwd(const float* pts, const float* mtx, const float* d_out, const float* d_out2, int accumulate, int B, int Bp,
                               int64_t V, float* d_pts, float* d_mtx, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(pts && mtx && d_out, "null pointer");
    B2A_CHECK_ARG(B > 0 && B <= 65535 && (Bp == 1 || Bp == B) && V >= 0, "shape");
    B2A_CHECK_ARG(((uintptr_t)d_out & 15) == 0 && ((uintptr_t)d_out2 & 15) == 0, "gradients must be 16-byte aligned");
    if (V > 0 && (d_pts || d_mtx))
        xfm_bwd_kernel<<<dim3(b2a_blocks(V, 256), B), 256, 0, stream>>>(pts, mtx, d_out, d_out2, accumulate, B, Bp, V, d_pts, d_mtx);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_rasterize_workspace_bytes(int B, int64_t F, int H, int W, size_t* bytes)
{
    B2A_CHECK_ARG(bytes && B > 0 && F >= 0 && H > 0 && W > 0, "shape");
    *bytes = raster_workspace_layout(B, F, H, W, nullptr, nullptr);
    return 0;
}

B2A_API int b2a_rasterize_fwd(const float* pos, const int32_t* tri, int B, int64_t V, int64_t F, int H, int W, void* workspace,
                              size_t workspace_bytes, float* rast, int32_t* cov_list, int32_t* cov_count, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(rast, "null pointer");
    B2A_CHECK_ARG((cov_list == nullptr) == (cov_count == nullptr) && (!cov_list || (int64_t)B * H * W < (1ll << 31)), "covered-pixel list");
    RasterWorkspace ws;
    int rc = raster_zbuffer_impl(pos, tri, B, V, F, H, W, workspace, workspace_bytes, &ws, stream);
    if (rc) return rc;
    if (cov_count) B2A_CUDA_OK(cudaMemsetAsync(cov_count, 0, sizeof(int), stream));
    B2A_CHECK_ARG(!cov_list || ((uintptr_t)cov_list & 15) == 0, "cov_list must be 16-byte aligned");
    raster_resolve_kernel<<<dim3(b2a_blocks((int64_t)H * W, 256), B), 256, 0, stream>>>(ws.zbuf, pos, tri, V, H, W, rast, (int4*)cov_list, cov_count);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_rasterize_bwd(const float* pos, const int32_t* tri, const float* rast, const float* d_rast, int B, int64_t V, int64_t F,
                              int H, int W, float* d_pos, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(pos && tri && rast && d_rast && d_pos, "null pointer");
    B2A_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && (int64_t)H * W < (1ll << 31), "shape");
    raster_bwd_kernel<<<dim3(b2a_blocks((int64_t)H * W, 256), B), 256, 0, stream>>>(pos, tri, rast, d_rast, V, F, H, W, d_pos);
    B2A_LAUNCH_OK();
    return 0;
}
