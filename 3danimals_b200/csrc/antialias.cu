// antialias.cu - silhouette antialiasing (+ fused alpha composite) and triangle edge adjacency on sm_100a.
// Replaces nvdiffrast.torch.antialias and the lerp composite of render_mesh.composite_buffer (reference call sites
// model/render/render.py:258-268).  Semantics and arithmetic: oracle/raster_ref.c (aa_analyze) - bit-identical.
//
// Design (B200-first): nvdiffrast scatters blends with atomics after a work-queue pass.  Here both directions are
// GATHERS: each output element looks at its four pixel pairs (up, left, right, down - the oracle's accumulation
// order), re-runs the pair analysis only where triangle ids differ (the silhouette: O(perimeter) pixels) and writes
// its result once.  No atomics on image data, deterministic, coalesced one-thread-per-element streaming; the
// composite lerp(bg, [color,1], id>0) is folded in so the composited image never exists in HBM.  Only the vertex
// position gradient of silhouette pairs uses atomics (a few thousand per image).
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

// ------------------------------------------------------------------------------------------------------------
// edge adjacency: open-addressing hash on the undirected edge, two lowest triangle ids per edge
// ------------------------------------------------------------------------------------------------------------
struct AdjWorkspace {
    unsigned long long* keys;
    int* t0;
    int* t1;
    uint32_t mask;
};

size_t adj_layout(int64_t F, void* base, AdjWorkspace* ws)
{
    uint64_t cap = 1024;
    while (cap < (uint64_t)F * 6) cap <<= 1;
    size_t kb = b2a_align(cap * 8), tb = b2a_align(cap * 4);
    if (ws) {
        char* p = (char*)base;
        ws->keys = (unsigned long long*)p;
        ws->t0 = (int*)(p + kb);
        ws->t1 = (int*)(p + kb + tb);
        ws->mask = (uint32_t)(cap - 1);
    }
    return kb + 2 * tb;
}

__device__ __forceinline__ uint32_t hash64(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (uint32_t)k;
}

__device__ __forceinline__ bool edge_of(const int* __restrict__ tri, int64_t i, int64_t V, int& f, int& lo, int& hi)
{
    f = (int)(i / 3);
    int e = (int)(i % 3);
    int a = __ldg(tri + (size_t)f * 3 + (e + 1) % 3), b = __ldg(tri + (size_t)f * 3 + (e + 2) % 3);
    if ((unsigned)a >= (unsigned)V || (unsigned)b >= (unsigned)V) return false;
    lo = min(a, b); hi = max(a, b);
    return true;
}

__global__ void adj_insert_kernel(const int* __restrict__ tri, int64_t F, int64_t V, AdjWorkspace ws)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * 3) return;
    int f, lo, hi;
    if (!edge_of(tri, i, V, f, lo, hi)) return;
    unsigned long long key = (unsigned long long)lo * (unsigned long long)(V + 1) + (unsigned long long)hi + 1ull;
    uint32_t slot = hash64(key) & ws.mask;
    while (true) {
        unsigned long long prev = atomicCAS(ws.keys + slot, 0ull, key);
        if (prev == 0ull || prev == key) break;
        slot = (slot + 1) & ws.mask;
    }
    atomicMin(ws.t0 + slot, f);
}

__device__ __forceinline__ uint32_t adj_find(const AdjWorkspace& ws, unsigned long long key)
{
    uint32_t slot = hash64(key) & ws.mask;
    while (ws.keys[slot] != key) slot = (slot + 1) & ws.mask;
    return slot;
}

__global__ void adj_second_kernel(const int* __restrict__ tri, int64_t F, int64_t V, AdjWorkspace ws)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * 3) return;
    int f, lo, hi;
    if (!edge_of(tri, i, V, f, lo, hi)) return;
    uint32_t slot = adj_find(ws, (unsigned long long)lo * (unsigned long long)(V + 1) + (unsigned long long)hi + 1ull);
    if (f != ws.t0[slot]) atomicMin(ws.t1 + slot, f);
}

__global__ void adj_emit_kernel(const int* __restrict__ tri, int64_t F, int64_t V, AdjWorkspace ws, int* __restrict__ opp)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * 3) return;
    int f, lo, hi;
    int ov = -1;
    if (edge_of(tri, i, V, f, lo, hi)) {
        uint32_t slot = adj_find(ws, (unsigned long long)lo * (unsigned long long)(V + 1) + (unsigned long long)hi + 1ull);
        int a = ws.t0[slot], b = ws.t1[slot];
        int partner = (f == a) ? b : a;
        if (partner >= 0 && partner < F) {
#pragma unroll
            for (int c = 2; c >= 0; c--) {
                int vv = __ldg(tri + (size_t)partner * 3 + c);
                if (vv != lo && vv != hi) ov = vv;  // descending loop: the first match in ascending order wins
            }
        }
    }
    opp[i] = ov;
}

// ------------------------------------------------------------------------------------------------------------
// pair analysis (oracle/raster_ref.c aa_analyze)
// ------------------------------------------------------------------------------------------------------------
struct AAParams {
    const float* color;
    const float* bg;
    const float* rast;
    const float* pos;
    const int* tri;
    const int* opp;
    int Bg, composite, B, H, W, C;
    int64_t V, F;
};

struct AAPair {
    float alpha;
    int tri, di, px, py;
};

__device__ __forceinline__ bool same_sign(float a, float b) { return (__float_as_int(a) ^ __float_as_int(b)) >= 0; }

#define B2A_F32_MAX 3.402823466e+38f

// PROVENANCE.  The reference calls nvdiffrast.torch.antialias (NVlabs nvdiffrast, NVIDIA Source Code License; an un-pinned git
// dependency that is NOT part of the reference tree, INSTALL.md:22).  Results identical to it require its pair analysis: which
// surface of a pixel pair is nearer, which of that triangle's edges are silhouette edges (no neighbour, or the neighbour's
// opposite vertex on the same screen side), where the edge crosses the segment between the pixel centres, the 1/16 guard on
// near-parallel edges, the blend weight 0.5 - distance.  aa_analyze restates that published algorithm (AntialiasFwdAnalysisKernel
// in nvdiffrast/common/antialias.cu) from memory - no nvdiffrast source is present in this container or this repository -
// so its structure and several local names follow the original; everything around it (analyse once per render, gather-based
// streaming kernels, pair fusion, coefficient folding for the position gradient) is this repository's own design.  Whether the
// restatement matches nvdiffrast bit for bit cannot be checked here: PARITY UNPINNED.
// (px,py) = first pixel of the pair, d = 0: neighbour to the right, 1: neighbour below.  r0/r1 = rast of the two pixels.
__device__ bool aa_analyze(const AAParams& P, const float* __restrict__ pos_b, float4 r0, float4 r1, int px, int py, int d, AAPair& r)
{
    int tri0 = (int)r0.w - 1, tri1 = (int)r1.w - 1;
    if (tri0 == tri1) return false;
    int t = (tri0 >= 0) ? tri0 : tri1;
    if (tri0 >= 0 && tri1 >= 0) t = (r0.z < r1.z) ? tri0 : tri1;
    if (t == tri1) { px += 1 - d; py += d; }
    if (t < 0 || t >= P.F) return false;
    int vi0 = __ldg(P.tri + (size_t)t * 3), vi1 = __ldg(P.tri + (size_t)t * 3 + 1), vi2 = __ldg(P.tri + (size_t)t * 3 + 2);
    if ((unsigned)vi0 >= (unsigned)P.V || (unsigned)vi1 >= (unsigned)P.V || (unsigned)vi2 >= (unsigned)P.V) return false;
    int op0 = __ldg(P.opp + (size_t)t * 3), op1 = __ldg(P.opp + (size_t)t * 3 + 1), op2 = __ldg(P.opp + (size_t)t * 3 + 2);
    if (op0 < 0) op0 = vi0;
    if (op1 < 0) op1 = vi1;
    if (op2 < 0) op2 = vi2;
    float4 p0 = ldg4(pos_b + (size_t)vi0 * 4), p1 = ldg4(pos_b + (size_t)vi1 * 4), p2 = ldg4(pos_b + (size_t)vi2 * 4);
    float4 o0 = ldg4(pos_b + (size_t)op0 * 4), o1 = ldg4(pos_b + (size_t)op1 * 4), o2 = ldg4(pos_b + (size_t)op2 * 4);
    float xh = 0.5f * (float)P.W, yh = 0.5f * (float)P.H;
    float fx = (float)px + 0.5f - xh, fy = (float)py + 0.5f - yh;
    float w0 = 1.f / p0.w, w1 = 1.f / p1.w, w2 = 1.f / p2.w;
    float ow0 = 1.f / o0.w, ow1 = 1.f / o1.w, ow2 = 1.f / o2.w;
    float x0 = p0.x * w0 * xh - fx, y0 = p0.y * w0 * yh - fy;
    float x1 = p1.x * w1 * xh - fx, y1 = p1.y * w1 * yh - fy;
    float x2 = p2.x * w2 * xh - fx, y2 = p2.y * w2 * yh - fy;
    float ox0 = o0.x * ow0 * xh - fx, oy0 = o0.y * ow0 * yh - fy;
    float ox1 = o1.x * ow1 * xh - fx, oy1 = o1.y * ow1 * yh - fy;
    float ox2 = o2.x * ow2 * xh - fx, oy2 = o2.y * ow2 * yh - fy;
    float bb = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
    float a0 = (x1 - ox0) * (y2 - oy0) - (x2 - ox0) * (y1 - oy0);
    float a1 = (x2 - ox1) * (y0 - oy1) - (x0 - ox1) * (y2 - oy1);
    float a2 = (x0 - ox2) * (y1 - oy2) - (x1 - ox2) * (y0 - oy2);
    bool s0 = same_sign(a0, bb), s1 = same_sign(a1, bb), s2 = same_sign(a2, bb);
    if (!(s0 || s1 || s2)) return false;
    if (d) { float tmp; tmp = x0; x0 = y0; y0 = tmp; tmp = x1; x1 = y1; y1 = tmp; tmp = x2; x2 = y2; y2 = tmp; }
    float dx0 = x2 - x1, dx1 = x0 - x2, dx2 = x1 - x0;
    float dy0 = y2 - y1, dy1 = y0 - y2, dy2 = y1 - y0;
    float ds = (t == tri0) ? 1.f : -1.f;
    float c0 = -B2A_F32_MAX, c1 = -B2A_F32_MAX, c2 = -B2A_F32_MAX;
    if (!same_sign(y1, y2)) c0 = ds * (x1 * dy0 - y1 * dx0) / dy0;
    if (!same_sign(y2, y0)) c1 = ds * (x2 * dy1 - y2 * dx1) / dy1;
    if (!same_sign(y0, y1)) c2 = ds * (x0 * dy2 - y0 * dx2) / dy2;
    int di = 0;
    float cm = c0;
    if (c1 > cm) { di = 1; cm = c1; }
    if (c2 > cm) { di = 2; cm = c2; }
    float dc = -B2A_F32_MAX;
    if (di == 0 && s0 && fabsf(dy0) >= fabsf(dx0)) dc = c0;
    if (di == 1 && s1 && fabsf(dy1) >= fabsf(dx1)) dc = c1;
    if (di == 2 && s2 && fabsf(dy2) >= fabsf(dx2)) dc = c2;
    const float eps = 0.0625f;
    if (dc > -eps && dc < 1.f + eps) {
        dc = fminf(fmaxf(dc, 0.f), 1.f);
        r.alpha = ds * (0.5f - dc);
        r.tri = t; r.di = di; r.px = px; r.py = py;
        return true;
    }
    return false;
}

// value of channel c of the (optionally composited) input image at pixel `pix` of image b
__device__ __forceinline__ float comp_color(const AAParams& P, int b, int pix, int c, bool covered)
{
    const size_t HW = (size_t)P.H * P.W;
    if (!P.composite) return __ldg(P.color + ((size_t)b * HW + pix) * P.C + c);
    if (covered) return c < P.C - 1 ? __ldg(P.color + ((size_t)b * HW + pix) * (P.C - 1) + c) : 1.f;
    return P.bg ? __ldg(P.bg + ((size_t)(P.Bg == 1 ? 0 : b) * HW + pix) * P.C + c) : 0.f;
}

// the four pixel pairs of pixel (px,py) in the oracle's accumulation order: (first pixel offset, direction)
__constant__ int c_pair_dx[4] = {0, -1, 0, 0};
__constant__ int c_pair_dy[4] = {-1, 0, 0, 0};
__constant__ int c_pair_d[4] = {1, 0, 0, 1};

__global__ void __launch_bounds__(256) aa_fwd_kernel(AAParams P, float* __restrict__ out)
{
    const int HW = P.H * P.W;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)HW * P.C) return;
    const int b = blockIdx.y;
    int p = (int)(idx / P.C), c = (int)(idx % P.C);
    int px = p % P.W, py = p / P.W;
    const float* rast_b = P.rast + (size_t)b * HW * 4;
    float idc = __ldg(rast_b + (size_t)p * 4 + 3);
    float acc = comp_color(P, b, p, c, idc > 0.f);
    // quick reject: all four neighbours carry the same id
    float idu = py > 0 ? __ldg(rast_b + (size_t)(p - P.W) * 4 + 3) : idc;
    float idl = px > 0 ? __ldg(rast_b + (size_t)(p - 1) * 4 + 3) : idc;
    float idr = px + 1 < P.W ? __ldg(rast_b + (size_t)(p + 1) * 4 + 3) : idc;
    float idd = py + 1 < P.H ? __ldg(rast_b + (size_t)(p + P.W) * 4 + 3) : idc;
    if (idu != idc || idl != idc || idr != idc || idd != idc) {
        const float* pos_b = P.pos + (size_t)b * P.V * 4;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int qx = px + c_pair_dx[k], qy = py + c_pair_dy[k], d = c_pair_d[k];
            if (qx < 0 || qy < 0) continue;
            if (d == 0 ? qx + 1 >= P.W : qy + 1 >= P.H) continue;
            int q0 = qy * P.W + qx, q1 = q0 + (d ? P.W : 1);
            float4 r0 = ldg4(rast_b + (size_t)q0 * 4), r1 = ldg4(rast_b + (size_t)q1 * 4);
            AAPair r;
            if (!aa_analyze(P, pos_b, r0, r1, qx, qy, d, r)) continue;
            int target = r.alpha > 0.f ? q0 : q1;
            if (target != p) continue;
            acc += r.alpha * (comp_color(P, b, q1, c, r1.w > 0.f) - comp_color(P, b, q0, c, r0.w > 0.f));
        }
    }
    out[((size_t)b * HW) * P.C + idx] = acc;
}

// gradient of one blended pair w.r.t. the clip-space positions of the silhouette edge's two vertices
// (oracle/raster_ref.c orc_antialias_bwd) is linear in dd = sum_c d_out[target,c] * (color[p1,c] - color[p0,c]):
// d_pos[e1] += dd * (g[0], g[1], 0, g[2]),  d_pos[e2] += dd * (g[3], g[4], 0, g[5]).
struct AAPosCoef {
    int e1, e2;
    float g[6];
};

__device__ void aa_pos_coef(const AAParams& P, const float* __restrict__ pos_b, const AAPair& r, int d, AAPosCoef& c)
{
    int e1 = __ldg(P.tri + (size_t)r.tri * 3 + (r.di + 1) % 3), e2 = __ldg(P.tri + (size_t)r.tri * 3 + (r.di + 2) % 3);
    float4 q1v = ldg4(pos_b + (size_t)e1 * 4), q2v = ldg4(pos_b + (size_t)e2 * 4);
    float pxh = 0.5f * (float)P.W, pyh = 0.5f * (float)P.H;
    float fx = (float)r.px + 0.5f - pxh, fy = (float)r.py + 0.5f - pyh;
    if (d) {
        float t_;
        t_ = q1v.x; q1v.x = q1v.y; q1v.y = t_;
        t_ = q2v.x; q2v.x = q2v.y; q2v.y = t_;
        t_ = pxh; pxh = pyh; pyh = t_;
        t_ = fx; fx = fy; fy = t_;
    }
    float w1 = 1.f / q1v.w, w2 = 1.f / q2v.w;
    float x1 = q1v.x * w1 * pxh - fx, y1 = q1v.y * w1 * pyh - fy;
    float x2 = q2v.x * w2 * pxh - fx, y2 = q2v.y * w2 * pyh - fy;
    float dx = x2 - x1, dy = y2 - y1;
    float db = x1 * dy - y1 * dx;
    float ep = copysignf(1e-3f, dy);
    float iy = 1.f / (dy + ep);
    float dby = db * iy;
    float iw1 = -w1 * iy, iw2 = w2 * iy;
    float gp1x = iw1 * pxh * y2, gp2x = iw2 * pxh * y1;
    float gp1y = iw1 * pyh * (dby - x2), gp2y = iw2 * pyh * (dby - x1);
    float gp1w = -(q1v.x * gp1x + q1v.y * gp1y) * w1;
    float gp2w = -(q2v.x * gp2x + q2v.y * gp2y) * w2;
    if (d) { float t_; t_ = gp1x; gp1x = gp1y; gp1y = t_; t_ = gp2x; gp2x = gp2y; gp2y = t_; }
    c.e1 = e1; c.e2 = e2;
    c.g[0] = gp1x; c.g[1] = gp1y; c.g[2] = gp1w; c.g[3] = gp2x; c.g[4] = gp2y; c.g[5] = gp2w;
}

__device__ __forceinline__ void aa_pos_apply(const AAPosCoef& c, float dd, float* __restrict__ d_pos_b)
{
    float* g1 = d_pos_b + (size_t)c.e1 * 4;
    float* g2 = d_pos_b + (size_t)c.e2 * 4;
    atomicAdd(g1, dd * c.g[0]); atomicAdd(g1 + 1, dd * c.g[1]); atomicAdd(g1 + 3, dd * c.g[2]);
    atomicAdd(g2, dd * c.g[3]); atomicAdd(g2 + 1, dd * c.g[4]); atomicAdd(g2 + 3, dd * c.g[5]);
}

__device__ void aa_pos_grad(const AAParams& P, const float* __restrict__ pos_b, const AAPair& r, int d, float dd, float* __restrict__ d_pos_b)
{
    AAPosCoef c;
    aa_pos_coef(P, pos_b, r, d, c);
    aa_pos_apply(c, dd, d_pos_b);
}

struct AAGrad {
    const float* d_out;
    int64_t sb, sy, sx, sc;
    int Cg;
};
__device__ __forceinline__ float grad_at(const AAGrad& G, int W, int b, int pix, int c)
{
    if (c >= G.Cg) return 0.f;
    return __ldg(G.d_out + (int64_t)b * G.sb + (int64_t)(pix / W) * G.sy + (int64_t)(pix % W) * G.sx + (int64_t)c * G.sc);
}

__global__ void __launch_bounds__(256) aa_bwd_kernel(AAParams P, AAGrad G, float* __restrict__ d_color, float* __restrict__ d_pos)
{
    const int HW = P.H * P.W;
    const int Cc = P.composite ? P.C - 1 : P.C;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)HW * Cc) return;
    const int b = blockIdx.y;
    int p = (int)(idx / Cc), c = (int)(idx % Cc);
    int px = p % P.W, py = p / P.W;
    const float* rast_b = P.rast + (size_t)b * HW * 4;
    float idc = __ldg(rast_b + (size_t)p * 4 + 3);
    float acc = grad_at(G, P.W, b, p, c);
    float idu = py > 0 ? __ldg(rast_b + (size_t)(p - P.W) * 4 + 3) : idc;
    float idl = px > 0 ? __ldg(rast_b + (size_t)(p - 1) * 4 + 3) : idc;
    float idr = px + 1 < P.W ? __ldg(rast_b + (size_t)(p + 1) * 4 + 3) : idc;
    float idd = py + 1 < P.H ? __ldg(rast_b + (size_t)(p + P.W) * 4 + 3) : idc;
    if (idu != idc || idl != idc || idr != idc || idd != idc) {
        const float* pos_b = P.pos + (size_t)b * P.V * 4;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int qx = px + c_pair_dx[k], qy = py + c_pair_dy[k], d = c_pair_d[k];
            if (qx < 0 || qy < 0) continue;
            if (d == 0 ? qx + 1 >= P.W : qy + 1 >= P.H) continue;
            int q0 = qy * P.W + qx, q1 = q0 + (d ? P.W : 1);
            float4 r0 = ldg4(rast_b + (size_t)q0 * 4), r1 = ldg4(rast_b + (size_t)q1 * 4);
            AAPair r;
            if (!aa_analyze(P, pos_b, r0, r1, qx, qy, d, r)) continue;
            int target = r.alpha > 0.f ? q0 : q1;
            float gy = grad_at(G, P.W, b, target, c);
            if (p == q0) acc -= r.alpha * gy; else acc += r.alpha * gy;
            // vertex-position gradient: once per pair, by channel-0 thread of the pair's first pixel
            if (c != 0 || p != q0 || !d_pos) continue;
            float dd = 0.f;
            for (int cc = 0; cc < G.Cg; cc++)
                dd += grad_at(G, P.W, b, target, cc) * (comp_color(P, b, q1, cc, r1.w > 0.f) - comp_color(P, b, q0, cc, r0.w > 0.f));
            if (dd == 0.f || fabsf(r.alpha) >= 0.5f) continue;
            aa_pos_grad(P, pos_b, r, d, dd, d_pos + (size_t)b * P.V * 4);
        }
    }
    if (d_color) d_color[((size_t)b * HW) * Cc + idx] = (P.composite && !(idc > 0.f)) ? 0.f : acc;
}

// ------------------------------------------------------------------------------------------------------------
// Fast path (composite mode = the training path): analyse once per render, then every launch is ONE streaming pass.
//   aa_prepare   : one pass over rast -> coverage bitmask + silhouette bitmask (1 bit/pixel each; silhouette = a
//                  4-neighbour carries a different triangle id) + compact list of silhouette pixels.
//   aa_pairs     : one thread per (silhouette pixel, owned pair: right / down) runs the pair analysis ONCE and stores
//                  (alpha, triangle, edge) in a dense per-pixel record touched only at silhouette pixels.  The 2 keys x
//                  (fwd, bwd) launches of a render share it, so none of them reads rast / pos / tri again.
//   aa_fwd_tile  : warp-autonomous 32-pixel tiles: float4 loads of NHWC(C-1) colour -> smem -> float4 stores of the
//                  composited NHWC(C) image; elements of silhouette pixels add their <=4 blend terms in the generic
//                  kernel's order (bit-identical results).
//   aa_bwd_tile  : same shape for the gradient: NCHW rows or NHWC float4s -> smem -> masked NHWC(C-1) float4 stores,
//                  silhouette elements gather their pair terms; the first AA_POS_BLOCKS blocks of the same launch
//                  scatter the edge-vertex position gradients (16 lanes per pair, channels across lanes).
// ------------------------------------------------------------------------------------------------------------
struct AAContext {
    uint32_t* cover;   // [B*HW/32] coverage bits
    uint32_t* act;     // [B*HW/32] pixels with at least one active (blending) pair - the true silhouette
    int* count;        // [0] pixels whose 4-neighbourhood carries another triangle id (candidates), [1] active pixels
    int* list;         // [B*HW] candidate pixels (flat index b*HW + p)
    int* alist;        // [B*HW] active pixels
    AAPosCoef* acoef;  // [2*B*HW] per active-list entry: position-gradient coefficients of the owned pairs (p,right), (p,down)
    float4* rec;       // [B*HW] dense, valid where the act bit is set: blend weights of the pixel's four pairs in the
                       //        generic kernel's order (up,p) (left,p) (p,right) (p,down); 0 = inactive
};

size_t aa_ctx_layout(int B, int H, int W, void* base, AAContext* ctx)
{
    size_t npix = (size_t)B * H * W;
    size_t cb = b2a_align(((npix + 31) / 32) * 4);
    size_t lb = b2a_align(npix * 4);
    if (ctx) {
        char* p = (char*)base;
        ctx->cover = (uint32_t*)p;
        ctx->act = (uint32_t*)(p + cb);
        ctx->count = (int*)(p + 2 * cb);
        ctx->list = (int*)(p + 2 * cb + 256);
        ctx->alist = (int*)(p + 2 * cb + 256 + lb);
        ctx->rec = (float4*)(p + 2 * cb + 256 + 2 * lb);
        ctx->acoef = (AAPosCoef*)(p + 2 * cb + 256 + 2 * lb + b2a_align(npix * sizeof(float4)));
    }
    return 2 * cb + 256 + 2 * lb + b2a_align(npix * sizeof(float4)) + b2a_align(npix * 2 * sizeof(AAPosCoef));
}

// grid-stride over all B*HW pixels (HW % 32 == 0, so a warp never straddles two images)
__global__ void __launch_bounds__(256) aa_prepare_kernel(const float* __restrict__ rast, int B, int H, int W, AAContext ctx)
{
    const int HW = H * W;
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)B * HW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int p = (int)(i % HW);
        int px = p % W, py = p / W;
        float idc = __ldg(rast + i * 4 + 3);
        float idl = __shfl_up_sync(0xffffffffu, idc, 1), idr = __shfl_down_sync(0xffffffffu, idc, 1);
        if (lane == 0 && px > 0) idl = __ldg(rast + (i - 1) * 4 + 3);
        if (lane == 31 && px + 1 < W) idr = __ldg(rast + (i + 1) * 4 + 3);
        if (px == 0) idl = idc;
        if (px + 1 >= W) idr = idc;
        float idu = py > 0 ? __ldg(rast + (i - W) * 4 + 3) : idc;
        float idd = py + 1 < H ? __ldg(rast + (i + W) * 4 + 3) : idc;
        bool cand = idu != idc || idl != idc || idr != idc || idd != idc;
        uint32_t cov = __ballot_sync(0xffffffffu, idc > 0.f);
        uint32_t cm = __ballot_sync(0xffffffffu, cand);
        if (lane == 0) ctx.cover[i >> 5] = cov;
        if (cm) {
            int base = 0;
            if (lane == 0) base = atomicAdd(ctx.count, __popc(cm));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (cand) ctx.list[base + __popc(cm & ((1u << lane) - 1u))] = (int)i;
        }
    }
}

// four lanes per candidate pixel, one per pair k = 0 (up,p)  1 (left,p)  2 (p,right)  3 (p,down).  Each pair is analysed
// from both of its pixels (identical arithmetic), so a pixel's record holds everything its fix-up needs.  Pixels with
// a blending pair (few: the silhouette proper) get a record, an act bit and an entry in the active list.
__global__ void __launch_bounds__(128) aa_pairs_kernel(AAParams P, AAContext ctx)
{
    const int HW = P.H * P.W;
    const int count = ctx.count[0];
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < 4 * count; base += gridDim.x * blockDim.x) {
        const int i = base + lane, k = i & 3;
        float alpha = 0.f;
        int flat = 0;
        AAPosCoef coef;
        coef.e1 = coef.e2 = 0;
#pragma unroll
        for (int q = 0; q < 6; q++) coef.g[q] = 0.f;
        if (i < 4 * count) {
            flat = ctx.list[i >> 2];
            const int b = flat / HW, p = flat - b * HW;
            const int qx = p % P.W + c_pair_dx[k], qy = p / P.W + c_pair_dy[k], d = c_pair_d[k];
            if (qx >= 0 && qy >= 0 && (d == 0 ? qx + 1 < P.W : qy + 1 < P.H)) {
                const float* rast_b = P.rast + (size_t)b * HW * 4;
                const float* pos_b = P.pos + (size_t)b * P.V * 4;
                const int q0 = qy * P.W + qx;
                float4 r0 = ldg4(rast_b + (size_t)q0 * 4), r1 = ldg4(rast_b + (size_t)(q0 + (d ? P.W : 1)) * 4);
                AAPair r;
                if (aa_analyze(P, pos_b, r0, r1, qx, qy, d, r)) {
                    alpha = r.alpha;
                    if (k >= 2 && fabsf(alpha) < 0.5f) aa_pos_coef(P, pos_b, r, d, coef);   // owned pair: position-gradient coefficients
                }
            }
        }
        const int g0 = lane & ~3;
        float4 a;
        a.x = __shfl_sync(0xffffffffu, alpha, g0);
        a.y = __shfl_sync(0xffffffffu, alpha, g0 + 1);
        a.z = __shfl_sync(0xffffffffu, alpha, g0 + 2);
        a.w = __shfl_sync(0xffffffffu, alpha, g0 + 3);
        const bool live = a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f;
        int slot = 0;
        if (k == 0 && live) {
            ctx.rec[flat] = a;
            atomicOr(ctx.act + (flat >> 5), 1u << (flat & 31));
            slot = atomicAdd(ctx.count + 1, 1);
            ctx.alist[slot] = flat;
        }
        slot = __shfl_sync(0xffffffffu, slot, g0);
        if (k >= 2 && live) ctx.acoef[(size_t)slot * 2 + (k - 2)] = coef;
    }
}

__device__ __forceinline__ bool aa_bit(const uint32_t* __restrict__ bits, size_t flat) { return (__ldg(bits + (flat >> 5)) >> (flat & 31)) & 1u; }

// composited colour of channel c at flat pixel q (image b): colour [.,C-1] | 1 where covered, bg (or 0) elsewhere
template <int C>
__device__ __forceinline__ float aa_comp(const float* __restrict__ color, const float* __restrict__ bg, int Bg, const uint32_t* __restrict__ cover,
                                         int b, int HW, size_t q, int c)
{
    if (aa_bit(cover, q)) return c < C - 1 ? __ldg(color + q * (C - 1) + c) : 1.f;
    return bg ? __ldg(bg + (Bg == 1 ? q - (size_t)b * HW : q) * C + c) : 0.f;
}

// j-th (0-based) set bit of m
__device__ __forceinline__ int nth_bit(uint32_t m, int j) { return (int)__fns(m, 0, j + 1); }

// Forward.  The smem tile holds the composited image of 32 pixels in OUTPUT layout [pp*C + c]; elements of active
// pixels are fixed in place (work items = active pixel x channel, spread over the lanes) before the float4 stores.
// A warp owns TPW consecutive tiles and issues the colour loads of all of them up front (bytes in flight).
template <int C, int TPW>
__device__ __forceinline__ void aa_fwd_tile_body(const float* __restrict__ color, const float* __restrict__ bg, int Bg, const AAContext& ctx,
                                                 int B, int H, int W, float* __restrict__ out, int64_t vblock, float* __restrict__ smem)
{
    constexpr int CI = C - 1;
    constexpr int NV = (8 * CI + 31) / 32;   // float4 colour loads per lane per tile
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int HW = H * W;
    const int64_t ntiles = ((int64_t)B * HW) / 32;
    const int64_t t0 = (vblock * 8 + w) * TPW;
    if (t0 >= ntiles) return;
    float* tile = smem + w * (32 * C);
    uint32_t cov[TPW], act[TPW];
    float4 cv[TPW][NV];
#pragma unroll
    for (int tt = 0; tt < TPW; tt++) {
        const int64_t t = t0 + tt;
        cov[tt] = t < ntiles ? __ldg(ctx.cover + t) : 0u;
        act[tt] = t < ntiles ? __ldg(ctx.act + t) : 0u;
    }
#pragma unroll
    for (int tt = 0; tt < TPW; tt++) {
        if (cov[tt] == 0u) continue;
        const float4* src = reinterpret_cast<const float4*>(color + (size_t)(t0 + tt) * 32 * CI);
#pragma unroll
        for (int k = 0; k < NV; k++)
            if (k * 32 + lane < 8 * CI) cv[tt][k] = __ldg(src + k * 32 + lane);
    }
#pragma unroll
    for (int tt = 0; tt < TPW; tt++) {
        if (t0 + tt >= ntiles) break;
        const size_t P0 = (size_t)(t0 + tt) * 32;
        const int b = (int)(P0 / HW);
        const uint32_t cm = cov[tt], am = act[tt];
        float4* dst = reinterpret_cast<float4*>(out + P0 * C);
        const float4* bsrc = bg ? reinterpret_cast<const float4*>(bg + (Bg == 1 ? P0 - (size_t)b * HW : P0) * C) : nullptr;
        if (cm == 0u && am == 0u) {      // pure background tile: straight copy
#pragma unroll
            for (int i = lane; i < 8 * C; i += 32) dst[i] = bsrc ? __ldg(bsrc + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        __syncwarp();
        if (cm != 0xffffffffu) {         // background (or zeros) in output layout
#pragma unroll
            for (int i = lane; i < 8 * C; i += 32) reinterpret_cast<float4*>(tile)[i] = bsrc ? __ldg(bsrc + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
        }
        if (cm) {
#pragma unroll
            for (int k = 0; k < NV; k++) {
                const int i = k * 32 + lane;
                if (i < 8 * CI) {
                    const float4 x = cv[tt][k];
                    const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int e = 4 * i + j;
                        const int pp = e / CI, c = e - pp * CI;
                        if ((cm >> pp) & 1u) tile[pp * C + c] = xv[j];
                    }
                }
            }
            if ((cm >> lane) & 1u) tile[lane * C + CI] = 1.f;
        }
        __syncwarp();
        if (am) {
            const int items = __popc(am) * C;
            for (int item = lane; item < items; item += 32) {
                const int j = item / C, c = item - j * C;
                const int pp = nth_bit(am, j);
                const size_t flat = P0 + pp;
                const float4 a = __ldg(ctx.rec + flat);
                const float own = tile[pp * C + c];
                // this pixel is the blend target of (up,p)/(left,p) when alpha < 0 and of (p,right)/(p,down) when alpha > 0.
                // Neighbour colours come from global memory (the unblended input), loaded up front so the loads overlap.
                const bool k0 = a.x < 0.f, k1 = a.y < 0.f, k2 = a.z > 0.f, k3 = a.w > 0.f;
                const float cu = k0 ? aa_comp<C>(color, bg, Bg, ctx.cover, b, HW, flat - W, c) : own;
                const float cl = k1 ? aa_comp<C>(color, bg, Bg, ctx.cover, b, HW, flat - 1, c) : own;
                const float cr = k2 ? aa_comp<C>(color, bg, Bg, ctx.cover, b, HW, flat + 1, c) : own;
                const float cd = k3 ? aa_comp<C>(color, bg, Bg, ctx.cover, b, HW, flat + W, c) : own;
                float acc = own;
                if (k0) acc += a.x * (own - cu);
                if (k1) acc += a.y * (own - cl);
                if (k2) acc += a.z * (cr - own);
                if (k3) acc += a.w * (cd - own);
                tile[pp * C + c] = acc;     // each item touches only its own slot
            }
            __syncwarp();
        }
#pragma unroll
        for (int i = lane; i < 8 * C; i += 32) dst[i] = reinterpret_cast<const float4*>(tile)[i];
    }
}

template <int C, int TPW>
__global__ void __launch_bounds__(256, TPW == 1 ? 5 : 4) aa_fwd_tile_kernel(const float* __restrict__ color, const float* __restrict__ bg, int Bg, AAContext ctx,
                                                          int B, int H, int W, float* __restrict__ out)
{
    __shared__ __align__(16) float s_tile[8 * 32 * C];
    aa_fwd_tile_body<C, TPW>(color, bg, Bg, ctx, B, H, W, out, (int64_t)blockIdx.x, s_tile);
}

constexpr int AA_POS_BLOCKS = 148 * 4;

// vertex-position gradient role: one half-warp per owned pair of an active pixel, channels across the 16 lanes.  All
// geometry was folded into per-pair coefficients by aa_pairs_kernel, so this is a short gather + six atomics and costs
// the streaming kernel no registers; it runs in the FIRST blocks of the same launch.
template <int C>
__device__ __forceinline__ void aa_bwd_pos_role(const AAParams& P, const AAGrad& G, const AAContext& ctx, float* __restrict__ d_pos, int role_block,
                                                int role_blocks = AA_POS_BLOCKS)
{
    const int HW = P.H * P.W;
    const int lane = threadIdx.x & 31, sub = lane & 15, d = lane >> 4;
    const int count = ctx.count[1];
    const int nwarp = role_blocks * (blockDim.x >> 5);
    for (int i = role_block * (blockDim.x >> 5) + (threadIdx.x >> 5); i < count; i += nwarp) {
        const int flat = ctx.alist[i];
        const int b = flat / HW, p = flat - b * HW;
        const float4 av = __ldg(ctx.rec + flat);
        const float a = d ? av.w : av.z;
        const bool on = a != 0.f && fabsf(a) < 0.5f;
        float dd = 0.f;
        const size_t q0 = (size_t)flat, q1 = q0 + (d ? P.W : 1);
        if (on) {
            const int target = a > 0.f ? p : p + (d ? P.W : 1);
            for (int c = sub; c < G.Cg; c += 16)
                dd += grad_at(G, P.W, b, target, c) * (aa_comp<C>(P.color, P.bg, P.Bg, ctx.cover, b, HW, q1, c) -
                                                       aa_comp<C>(P.color, P.bg, P.Bg, ctx.cover, b, HW, q0, c));
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, o);
        if (on && sub == 0 && dd != 0.f) {
            const AAPosCoef* cp = ctx.acoef + (size_t)i * 2 + d;
            const int4 c0 = __ldg(reinterpret_cast<const int4*>(cp));
            const float4 c1 = __ldg(reinterpret_cast<const float4*>(cp) + 1);
            float* g1 = d_pos + ((size_t)b * P.V + c0.x) * 4;
            float* g2 = d_pos + ((size_t)b * P.V + c0.y) * 4;
            atomicAdd(g1, dd * __int_as_float(c0.z)); atomicAdd(g1 + 1, dd * __int_as_float(c0.w)); atomicAdd(g1 + 3, dd * c1.x);
            atomicAdd(g2, dd * c1.y); atomicAdd(g2 + 1, dd * c1.z); atomicAdd(g2 + 3, dd * c1.w);
        }
    }
}

// Backward.  CC = colour channels written, CG = gradient channels read (CG >= CC; CG == CC + 1 when the alpha channel is
// kept).  NCHW: the gradient's x stride is 1 (each warp loads CC coalesced 128-byte rows); else NHWC-contiguous (float4s).
// A warp owns TPW consecutive 32-pixel tiles and issues the loads of all of them before touching any (bytes in flight:
// the narrow keys are latency-bound otherwise).  Per tile the smem tile [pp*ST + c] holds g; covered silhouette elements
// gather their pair terms in place, then masked float4 stores.
template <int C, int CC, int CG, bool NCHW, int TPW>
__device__ __forceinline__ void aa_bwd_tile_body(const AAParams& P, const AAGrad& G, const AAContext& ctx, float* __restrict__ d_color,
                                                 int64_t vblock, float* __restrict__ smem)
{
    constexpr int ST = CG + 1;   // padded pixel stride: conflict-free for both access patterns
    constexpr int NV = NCHW ? CC : (8 * CG + 31) / 32;   // registers per lane per tile: scalars (NCHW) or float4s (NHWC)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int HW = P.H * P.W;
    const int64_t ntiles = ((int64_t)P.B * HW) / 32;
    const int64_t t0 = (vblock * 8 + w) * TPW;
    if (t0 >= ntiles) return;
    float* tile = smem + w * (32 * ST);
    uint32_t cov[TPW], act[TPW];
    float vs[NCHW ? TPW : 1][NCHW ? NV : 1];
    float4 vv[NCHW ? 1 : TPW][NCHW ? 1 : NV];
#pragma unroll
    for (int tt = 0; tt < TPW; tt++) {
        const int64_t t = t0 + tt;
        cov[tt] = t < ntiles ? __ldg(ctx.cover + t) : 0u;
        act[tt] = t < ntiles ? __ldg(ctx.act + t) & cov[tt] : 0u;
    }
#pragma unroll
    for (int tt = 0; tt < TPW; tt++) {
        if (cov[tt] == 0u) continue;
        const size_t P0 = (size_t)(t0 + tt) * 32;
        const int b = (int)(P0 / HW), p0 = (int)(P0 - (size_t)b * HW);
        if (NCHW) {
            const int py = p0 / P.W, px0 = p0 - py * P.W;   // W % 32 == 0: the tile lies in one row
            const float* gp = G.d_out + (int64_t)b * G.sb + (int64_t)py * G.sy + px0 + lane;
#pragma unroll
            for (int c = 0; c < NV; c++) vs[NCHW ? tt : 0][NCHW ? c : 0] = __ldg(gp + (int64_t)c * G.sc);
        } else {
            const float4* src = reinterpret_cast<const float4*>(G.d_out + (int64_t)b * G.sb + (int64_t)p0 * CG);
#pragma unroll
            for (int k = 0; k < NV; k++)
                if (k * 32 + lane < 8 * CG) vv[NCHW ? 0 : tt][NCHW ? 0 : k] = __ldg(src + k * 32 + lane);
        }
    }
#pragma unroll
    for (int tt = 0; tt < TPW; tt++) {
        if (t0 + tt >= ntiles) break;
        const size_t P0 = (size_t)(t0 + tt) * 32;
        const int b = (int)(P0 / HW), p0 = (int)(P0 - (size_t)b * HW);
        float4* dst = reinterpret_cast<float4*>(d_color + P0 * CC);
        if (cov[tt] == 0u) {   // nothing covered in this tile: the colour gradient is zero (background never reaches `color`)
#pragma unroll
            for (int i = lane; i < 8 * CC; i += 32) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        __syncwarp();
        if (NCHW) {
#pragma unroll
            for (int c = 0; c < NV; c++) tile[lane * ST + c] = vs[NCHW ? tt : 0][NCHW ? c : 0];
        } else {
#pragma unroll
            for (int k = 0; k < NV; k++) {
                const int i = k * 32 + lane;
                if (i < 8 * CG) {
                    const float4 x = vv[NCHW ? 0 : tt][NCHW ? 0 : k];
                    const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int e = 4 * i + j;
                        const int pp = e / CG, c = e - pp * CG;
                        tile[pp * ST + c] = xv[j];
                    }
                }
            }
        }
        __syncwarp();
        if (act[tt]) {
            const uint32_t am = act[tt];
            const int items = __popc(am) * CC;
            for (int item = lane; item < items; item += 32) {
                const int j = item / CC, c = item - j * CC;
                const int pp = nth_bit(am, j);
                const int p = p0 + pp;
                const float4 a = __ldg(ctx.rec + P0 + pp);
                const float own = tile[pp * ST + c];
                // d_color[p] = g[p] + a_up g[t] + a_left g[t] - a_right g[t] - a_down g[t], t = the pair's blend target
                const float gu = a.x != 0.f ? (a.x > 0.f ? grad_at(G, P.W, b, p - P.W, c) : own) : 0.f;
                const float gl = a.y != 0.f ? (a.y > 0.f ? grad_at(G, P.W, b, p - 1, c) : own) : 0.f;
                const float gr = a.z != 0.f ? (a.z > 0.f ? own : grad_at(G, P.W, b, p + 1, c)) : 0.f;
                const float gd = a.w != 0.f ? (a.w > 0.f ? own : grad_at(G, P.W, b, p + P.W, c)) : 0.f;
                float acc = own;
                if (a.x != 0.f) acc += a.x * gu;
                if (a.y != 0.f) acc += a.y * gl;
                if (a.z != 0.f) acc -= a.z * gr;
                if (a.w != 0.f) acc -= a.w * gd;
                tile[pp * ST + c] = acc;    // each item touches only its own slot
            }
            __syncwarp();
        }
        const uint32_t cm = cov[tt];
#pragma unroll
        for (int i = lane; i < 8 * CC; i += 32) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int e = 4 * i + j;
                const int pp = e / CC, c = e - pp * CC;
                v[j] = ((cm >> pp) & 1u) ? tile[pp * ST + c] : 0.f;
            }
            dst[i] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

template <int C, int CC, int CG, bool NCHW, int TPW>
__global__ void __launch_bounds__(256, TPW == 1 ? 5 : 4) aa_bwd_tile_kernel(AAParams P, AAGrad G, AAContext ctx, float* __restrict__ d_color,
                                                                            float* __restrict__ d_pos, int pos_blocks)
{
    __shared__ float s_tile[8 * 32 * (CG + 1)];
    if ((int)blockIdx.x < pos_blocks) {   // first in the grid: their latency chains overlap the streaming blocks
        aa_bwd_pos_role<C>(P, G, ctx, d_pos, (int)blockIdx.x);
        return;
    }
    aa_bwd_tile_body<C, CC, CG, NCHW, TPW>(P, G, ctx, d_color, (int64_t)blockIdx.x - pos_blocks, s_tile);
}

// ------------------------------------------------------------------------------------------------------------
// Narrow keys (C <= 4: shaded RGBA, shading, flow, depth): one thread per PIXEL, no shared memory, ~24 registers ->
// full occupancy.  The warp-tile kernels above pay a fixed per-tile latency chain that only amortises over wide rows.
// ------------------------------------------------------------------------------------------------------------
constexpr int AA_PPT = 1;   // pixels per thread (measured: 4 pixels per thread with batched loads was 40 % SLOWER - fewer, fatter threads)

template <int C>
__device__ __forceinline__ void aa_fwd_pix_body(const float* __restrict__ color, const float* __restrict__ bg, int Bg, const AAContext& ctx,
                                                int B, int H, int W, float* __restrict__ out, int64_t vblock)
{
    constexpr int CI = C - 1;
    const int HW = H * W;
    const int64_t n = (int64_t)B * HW;
    const int64_t base = vblock * (blockDim.x * AA_PPT) + threadIdx.x;
    uint32_t cw[AA_PPT], aw[AA_PPT];
    float v[AA_PPT][C];
#pragma unroll
    for (int j = 0; j < AA_PPT; j++) {
        const int64_t flat = base + j * blockDim.x;
        cw[j] = flat < n ? __ldg(ctx.cover + (flat >> 5)) : 0u;
        aw[j] = flat < n ? __ldg(ctx.act + (flat >> 5)) : 0u;
    }
#pragma unroll
    for (int j = 0; j < AA_PPT; j++) {
        const int64_t flat = base + j * blockDim.x;
        if (flat >= n) continue;
        if ((cw[j] >> (flat & 31)) & 1u) {
#pragma unroll
            for (int c = 0; c < CI; c++) v[j][c] = __ldg(color + flat * CI + c);
            v[j][CI] = 1.f;
        } else {
            const int b = (int)(flat / HW);
            const float* bp = bg ? bg + (Bg == 1 ? flat - (int64_t)b * HW : flat) * C : nullptr;
#pragma unroll
            for (int c = 0; c < C; c++) v[j][c] = bp ? __ldg(bp + c) : 0.f;
        }
    }
#pragma unroll
    for (int j = 0; j < AA_PPT; j++) {
        const int64_t flat = base + j * blockDim.x;
        if (flat >= n) continue;
        if ((aw[j] >> (flat & 31)) & 1u) {
            const int b = (int)(flat / HW);
            const float4 a = __ldg(ctx.rec + flat);
            const bool k0 = a.x < 0.f, k1 = a.y < 0.f, k2 = a.z > 0.f, k3 = a.w > 0.f;
#pragma unroll
            for (int c = 0; c < C; c++) {
                const float own = v[j][c];
                const float cu = k0 ? aa_comp<C>(color, bg, Bg, ctx.cover, b, HW, flat - W, c) : own;
                const float cl = k1 ? aa_comp<C>(color, bg, Bg, ctx.cover, b, HW, flat - 1, c) : own;
                const float cr = k2 ? aa_comp<C>(color, bg, Bg, ctx.cover, b, HW, flat + 1, c) : own;
                const float cd = k3 ? aa_comp<C>(color, bg, Bg, ctx.cover, b, HW, flat + W, c) : own;
                float acc = own;
                if (k0) acc += a.x * (own - cu);
                if (k1) acc += a.y * (own - cl);
                if (k2) acc += a.z * (cr - own);
                if (k3) acc += a.w * (cd - own);
                v[j][c] = acc;
            }
        }
        float* o = out + flat * C;
        if (C == 4) reinterpret_cast<float4*>(o)[0] = make_float4(v[j][0], v[j][1], v[j][2], v[j][C - 1]);
        else if (C == 2) reinterpret_cast<float2*>(o)[0] = make_float2(v[j][0], v[j][C - 1]);
        else {
#pragma unroll
            for (int c = 0; c < C; c++) o[c] = v[j][c];
        }
    }
}

template <int C>
__global__ void __launch_bounds__(256) aa_fwd_pix_kernel(const float* __restrict__ color, const float* __restrict__ bg, int Bg, AAContext ctx,
                                                         int B, int H, int W, float* __restrict__ out)
{
    aa_fwd_pix_body<C>(color, bg, Bg, ctx, B, H, W, out, (int64_t)blockIdx.x);
}

template <int C, int CC, int CG>
__device__ __forceinline__ void aa_bwd_pix_body(const AAParams& P, const AAGrad& G, const AAContext& ctx, float* __restrict__ d_color, int64_t vblock)
{
    const int HW = P.H * P.W;
    const int64_t n = (int64_t)P.B * HW;
    const int64_t base = vblock * (blockDim.x * AA_PPT) + threadIdx.x;
    uint32_t cw[AA_PPT], aw[AA_PPT];
    float v[AA_PPT][CC];
#pragma unroll
    for (int j = 0; j < AA_PPT; j++) {
        const int64_t flat = base + j * blockDim.x;
        cw[j] = flat < n ? __ldg(ctx.cover + (flat >> 5)) : 0u;
        aw[j] = flat < n ? __ldg(ctx.act + (flat >> 5)) : 0u;
    }
#pragma unroll
    for (int j = 0; j < AA_PPT; j++) {
        const int64_t flat = base + j * blockDim.x;
#pragma unroll
        for (int c = 0; c < CC; c++) v[j][c] = 0.f;
        if (flat < n && ((cw[j] >> (flat & 31)) & 1u)) {
            const int b = (int)(flat / HW), p = (int)(flat - (int64_t)b * HW);
            const int py = p / P.W, px = p - py * P.W;
            const float* gp = G.d_out + (int64_t)b * G.sb + (int64_t)py * G.sy + (int64_t)px * G.sx;
#pragma unroll
            for (int c = 0; c < CC; c++) v[j][c] = __ldg(gp + (int64_t)c * G.sc);
        }
    }
#pragma unroll
    for (int j = 0; j < AA_PPT; j++) {
        const int64_t flat = base + j * blockDim.x;
        if (flat >= n) continue;
        if (((cw[j] & aw[j]) >> (flat & 31)) & 1u) {
            const int b = (int)(flat / HW), p = (int)(flat - (int64_t)b * HW);
            const float4 a = __ldg(ctx.rec + flat);
#pragma unroll
            for (int c = 0; c < CC; c++) {
                const float own = v[j][c];
                const float gu = a.x != 0.f ? (a.x > 0.f ? grad_at(G, P.W, b, p - P.W, c) : own) : 0.f;
                const float gl = a.y != 0.f ? (a.y > 0.f ? grad_at(G, P.W, b, p - 1, c) : own) : 0.f;
                const float gr = a.z != 0.f ? (a.z > 0.f ? own : grad_at(G, P.W, b, p + 1, c)) : 0.f;
                const float gd = a.w != 0.f ? (a.w > 0.f ? own : grad_at(G, P.W, b, p + P.W, c)) : 0.f;
                float acc = own;
                if (a.x != 0.f) acc += a.x * gu;
                if (a.y != 0.f) acc += a.y * gl;
                if (a.z != 0.f) acc -= a.z * gr;
                if (a.w != 0.f) acc -= a.w * gd;
                v[j][c] = acc;
            }
        }
        float* o = d_color + flat * CC;
        if (CC == 2) reinterpret_cast<float2*>(o)[0] = make_float2(v[j][0], v[j][CC - 1]);
        else {
#pragma unroll
            for (int c = 0; c < CC; c++) o[c] = v[j][c];
        }
    }
}

template <int C, int CC, int CG>
__global__ void __launch_bounds__(256, 6) aa_bwd_pix_kernel(AAParams P, AAGrad G, AAContext ctx, float* __restrict__ d_color,
                                                            float* __restrict__ d_pos, int pos_blocks)
{
    if ((int)blockIdx.x < pos_blocks) {
        aa_bwd_pos_role<C>(P, G, ctx, d_pos, (int)blockIdx.x);
        return;
    }
    aa_bwd_pix_body<C, CC, CG>(P, G, ctx, d_color, (int64_t)blockIdx.x - pos_blocks);
}

// ------------------------------------------------------------------------------------------------------------
// Two keys of one render in ONE launch (the training pair: a wide key, e.g. dino_pred [16+1], and a narrow key, e.g.
// shaded [3+1]).  Both share the render's prepared context; the wide key's tiles go first (longest blocks), the narrow
// key's pixels fill in behind them, and in the backward the position-gradient roles of both keys lead the grid and add
// into one d_pos.  Saves a launch ramp + tail per direction and lets the narrow key's latency-bound blocks overlap
// the wide key's bandwidth-bound ones.
// ------------------------------------------------------------------------------------------------------------
// Block -> role of the fused launch: all wide blocks first, the narrow key's blocks behind them.  (Measured: interleaving
// the two kinds 1:1 is SLOWER - 43.0 vs 39.3 us backward, 34.8 vs 32.8 us forward - the wide key's stream loses
// residency to the latency-bound narrow blocks for the whole launch instead of only at its end.)
__device__ __forceinline__ bool pair_role(int bx, int wide_blocks, int& wb, int& nb)
{
    wb = bx; nb = bx - wide_blocks;
    return bx < wide_blocks;
}

template <int CW, int CN>
__global__ void __launch_bounds__(256, 5) aa_fwd_pair_kernel(const float* __restrict__ color_w, const float* __restrict__ bg_w, int Bg_w,
                                                             float* __restrict__ out_w, const float* __restrict__ color_n,
                                                             const float* __restrict__ bg_n, int Bg_n, float* __restrict__ out_n, AAContext ctx,
                                                             int B, int H, int W, int wide_blocks)
{
    __shared__ __align__(16) float s_tile[8 * 32 * CW];
    int wb, nb;
    const bool wide = pair_role((int)blockIdx.x, wide_blocks, wb, nb);
    if (wide) aa_fwd_tile_body<CW, 1>(color_w, bg_w, Bg_w, ctx, B, H, W, out_w, (int64_t)wb, s_tile);
    else aa_fwd_pix_body<CN>(color_n, bg_n, Bg_n, ctx, B, H, W, out_n, (int64_t)nb);
}

template <int CW, int CGW, bool NCHW, int CN, int CGN>
__global__ void __launch_bounds__(256, 5) aa_bwd_pair_kernel(AAParams Pw, AAGrad Gw, float* __restrict__ d_color_w, AAParams Pn, AAGrad Gn,
                                                             float* __restrict__ d_color_n, AAContext ctx, float* __restrict__ d_pos,
                                                             int pos_blocks, int wide_blocks, int pos_first)
{
    __shared__ float s_tile[8 * 32 * (CGW + 1)];
    int bx = (int)blockIdx.x;
    if (pos_first < 0) {         // experiment switch (B2A_AA_POS): position-gradient roles at the END of the grid
        const int stream_blocks = (int)gridDim.x - 2 * pos_blocks;
        if (bx >= stream_blocks) {
            bx -= stream_blocks;
            if (bx < pos_blocks) aa_bwd_pos_role<CW>(Pw, Gw, ctx, d_pos, bx, pos_blocks);
            else aa_bwd_pos_role<CN>(Pn, Gn, ctx, d_pos, bx - pos_blocks, pos_blocks);
            return;
        }
    } else {
        if (bx < pos_blocks) { aa_bwd_pos_role<CW>(Pw, Gw, ctx, d_pos, bx, pos_blocks); return; }
        bx -= pos_blocks;
        if (bx < pos_blocks) { aa_bwd_pos_role<CN>(Pn, Gn, ctx, d_pos, bx, pos_blocks); return; }
        bx -= pos_blocks;
    }
    int wb, nb;
    const bool wide = pair_role(bx, wide_blocks, wb, nb);
    if (wide) aa_bwd_tile_body<CW, CW - 1, CGW, NCHW, 1>(Pw, Gw, ctx, d_color_w, (int64_t)wb, s_tile);
    else aa_bwd_pix_body<CN, CN - 1, CGN>(Pn, Gn, ctx, d_color_n, (int64_t)nb);
}

// ------------------------------------------------------------------------------------------------------------
// msaa / logging path (spp > 1, or keys that are composited but not antialiased: kd, normal, geo_normal).
// render_mesh shades the g-buffer at [H/up, W/up] and the reference nearest-UPSAMPLES every shaded buffer to the raster
// resolution before compositing it (render.py:217-219 util.scale_img_nhwc, :258-268) - at 2048^2 internal that is a
// 64 MB copy per key forward and a 4x4 block sum per key backward, in separate PyTorch kernels.  Here the low-resolution
// colour is read IN PLACE: hi-res pixel (y,x) takes colour row (y/up, x/up); the backward accumulates the up*up hi-res
// contributions of a low-res pixel in registers and writes d_color at low resolution.  aa = 0: composite only.
// Narrow keys only (C <= 4), prepared context required; same blend arithmetic and accumulation order as aa_fwd_pix.
// ------------------------------------------------------------------------------------------------------------
struct AAUp {
    int up, W, HW, lw, lhw;   // hi-res width / pixels per image, low-res width / pixels per image
};

__device__ __forceinline__ size_t up_index(const AAUp& U, int b, size_t q)
{
    if (U.up == 1) return q;
    const int p = (int)(q - (size_t)b * U.HW);
    const int y = p / U.W, x = p - y * U.W;
    return (size_t)b * U.lhw + (size_t)(y / U.up) * U.lw + (x / U.up);
}

template <int C>
__device__ __forceinline__ float aa_comp_up(const float* __restrict__ color, const AAUp& U, const float* __restrict__ bg, int Bg,
                                            const uint32_t* __restrict__ cover, int b, size_t q, int c)
{
    if (aa_bit(cover, q)) return c < C - 1 ? __ldg(color + up_index(U, b, q) * (C - 1) + c) : 1.f;
    return bg ? __ldg(bg + (Bg == 1 ? q - (size_t)b * U.HW : q) * C + c) : 0.f;
}

template <int C>
__global__ void __launch_bounds__(256) aa_up_fwd_kernel(const float* __restrict__ color, AAUp U, const float* __restrict__ bg, int Bg, AAContext ctx,
                                                        int B, int aa, float* __restrict__ out)
{
    constexpr int CI = C - 1;
    const int64_t n = (int64_t)B * U.HW;
    const int64_t flat = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (flat >= n) return;
    const int b = (int)(flat / U.HW);
    const uint32_t cw = __ldg(ctx.cover + (flat >> 5));
    const uint32_t aw = aa ? __ldg(ctx.act + (flat >> 5)) : 0u;
    float v[C];
    if ((cw >> (flat & 31)) & 1u) {
        const float* cp = color + up_index(U, b, (size_t)flat) * CI;
#pragma unroll
        for (int c = 0; c < CI; c++) v[c] = __ldg(cp + c);
        v[CI] = 1.f;
    } else {
        const float* bp = bg ? bg + (Bg == 1 ? flat - (int64_t)b * U.HW : flat) * C : nullptr;
#pragma unroll
        for (int c = 0; c < C; c++) v[c] = bp ? __ldg(bp + c) : 0.f;
    }
    if ((aw >> (flat & 31)) & 1u) {
        const float4 a = __ldg(ctx.rec + flat);
        const bool k0 = a.x < 0.f, k1 = a.y < 0.f, k2 = a.z > 0.f, k3 = a.w > 0.f;
#pragma unroll
        for (int c = 0; c < C; c++) {
            const float own = v[c];
            const float cu = k0 ? aa_comp_up<C>(color, U, bg, Bg, ctx.cover, b, (size_t)flat - U.W, c) : own;
            const float cl = k1 ? aa_comp_up<C>(color, U, bg, Bg, ctx.cover, b, (size_t)flat - 1, c) : own;
            const float cr = k2 ? aa_comp_up<C>(color, U, bg, Bg, ctx.cover, b, (size_t)flat + 1, c) : own;
            const float cd = k3 ? aa_comp_up<C>(color, U, bg, Bg, ctx.cover, b, (size_t)flat + U.W, c) : own;
            float acc = own;
            if (k0) acc += a.x * (own - cu);
            if (k1) acc += a.y * (own - cl);
            if (k2) acc += a.z * (cr - own);
            if (k3) acc += a.w * (cd - own);
            v[c] = acc;
        }
    }
    float* o = out + flat * C;
    if (C == 4) reinterpret_cast<float4*>(o)[0] = make_float4(v[0], v[1], v[2], v[C - 1]);
    else if (C == 2) reinterpret_cast<float2*>(o)[0] = make_float2(v[0], v[C - 1]);
    else {
#pragma unroll
        for (int c = 0; c < C; c++) o[c] = v[c];
    }
}

// Upstream gradient of hi-res pixel `pix`.  POOL: the caller's gradient is that of the spp x spp AVERAGE of the antialiased image
// (render.py:322-323 util.avg_pool_nhwc) at [H/up, W/up]: every hi-res pixel of a block receives g / (up*up) - avg_pool2d's backward,
// never materialised at the raster resolution.
template <bool POOL>
__device__ __forceinline__ float grad_up(const AAGrad& G, const AAUp& U, int b, int pix, int c)
{
    if (c >= G.Cg) return 0.f;
    int y = pix / U.W, x = pix - y * U.W;
    if (POOL) { y /= U.up; x /= U.up; }
    const float g = __ldg(G.d_out + (int64_t)b * G.sb + (int64_t)y * G.sy + (int64_t)x * G.sx + (int64_t)c * G.sc);
    return POOL ? g / (float)(U.up * U.up) : g;
}

// Composite (+ antialias) at the raster resolution and AVERAGE over the up x up block in one pass: one thread per low-resolution
// output pixel.  The hi-res image (64 MB per 4-channel key at 2048^2) is never written or re-read; the per-pixel arithmetic and the
// row-major summation order are those of aa_up_fwd_kernel followed by avg_pool2d, so the result is bit-identical to that pair.
template <int C>
__global__ void __launch_bounds__(256) aa_up_pool_fwd_kernel(const float* __restrict__ color, AAUp U, const float* __restrict__ bg, int Bg, AAContext ctx,
                                                             int B, int aa, int keep, float* __restrict__ out)
{
    constexpr int CI = C - 1;
    const int64_t n = (int64_t)B * U.lhw;
    const int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n) return;
    const int b = (int)(li / U.lhw);
    const int lp = (int)(li - (int64_t)b * U.lhw);
    const int yl = lp / U.lw, xl = lp - yl * U.lw;
    float col[CI], acc[C];
#pragma unroll
    for (int c = 0; c < CI; c++) col[c] = __ldg(color + li * CI + c);
#pragma unroll
    for (int c = 0; c < C; c++) acc[c] = 0.f;
    for (int dy = 0; dy < U.up; dy++) {
        for (int dx = 0; dx < U.up; dx++) {
            const int p = (yl * U.up + dy) * U.W + xl * U.up + dx;
            const size_t flat = (size_t)b * U.HW + p;
            float v[C];
            if (aa_bit(ctx.cover, flat)) {
#pragma unroll
                for (int c = 0; c < CI; c++) v[c] = col[c];
                v[CI] = 1.f;
            } else {
                const float* bp = bg ? bg + (Bg == 1 ? (size_t)p : flat) * C : nullptr;
#pragma unroll
                for (int c = 0; c < C; c++) v[c] = bp ? __ldg(bp + c) : 0.f;
            }
            if (aa && aa_bit(ctx.act, flat)) {
                const float4 a = __ldg(ctx.rec + flat);
                const bool k0 = a.x < 0.f, k1 = a.y < 0.f, k2 = a.z > 0.f, k3 = a.w > 0.f;
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const float own = v[c];
                    const float cu = k0 ? aa_comp_up<C>(color, U, bg, Bg, ctx.cover, b, flat - U.W, c) : own;
                    const float cl = k1 ? aa_comp_up<C>(color, U, bg, Bg, ctx.cover, b, flat - 1, c) : own;
                    const float cr = k2 ? aa_comp_up<C>(color, U, bg, Bg, ctx.cover, b, flat + 1, c) : own;
                    const float cd = k3 ? aa_comp_up<C>(color, U, bg, Bg, ctx.cover, b, flat + U.W, c) : own;
                    float t = own;
                    if (k0) t += a.x * (own - cu);
                    if (k1) t += a.y * (own - cl);
                    if (k2) t += a.z * (cr - own);
                    if (k3) t += a.w * (cd - own);
                    v[c] = t;
                }
            }
#pragma unroll
            for (int c = 0; c < C; c++) acc[c] += v[c];
        }
    }
    const float area = (float)(U.up * U.up);
    float* o = out + li * keep;
#pragma unroll
    for (int c = 0; c < C; c++)
        if (c < keep) o[c] = acc[c] / area;
}

// position-gradient role with up-sampled colour (see aa_bwd_pos_role)
template <int C, bool POOL>
__device__ __forceinline__ void aa_up_pos_role(const float* __restrict__ color, const AAUp& U, const float* __restrict__ bg, int Bg, int64_t V,
                                               const AAGrad& G, const AAContext& ctx, float* __restrict__ d_pos, int role_block)
{
    const int lane = threadIdx.x & 31, sub = lane & 15, d = lane >> 4;
    const int count = ctx.count[1];
    const int nwarp = AA_POS_BLOCKS * (blockDim.x >> 5);
    for (int i = role_block * (blockDim.x >> 5) + (threadIdx.x >> 5); i < count; i += nwarp) {
        const int flat = ctx.alist[i];
        const int b = flat / U.HW, p = flat - b * U.HW;
        const float4 av = __ldg(ctx.rec + flat);
        const float a = d ? av.w : av.z;
        const bool on = a != 0.f && fabsf(a) < 0.5f;
        float dd = 0.f;
        const size_t q0 = (size_t)flat, q1 = q0 + (d ? U.W : 1);
        if (on) {
            const int target = a > 0.f ? p : p + (d ? U.W : 1);
            for (int c = sub; c < G.Cg; c += 16)
                dd += grad_up<POOL>(G, U, b, target, c) * (aa_comp_up<C>(color, U, bg, Bg, ctx.cover, b, q1, c) -
                                                           aa_comp_up<C>(color, U, bg, Bg, ctx.cover, b, q0, c));
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, o);
        if (on && sub == 0 && dd != 0.f) {
            const AAPosCoef* cp = ctx.acoef + (size_t)i * 2 + d;
            const int4 c0 = __ldg(reinterpret_cast<const int4*>(cp));
            const float4 c1 = __ldg(reinterpret_cast<const float4*>(cp) + 1);
            float* g1 = d_pos + ((size_t)b * V + c0.x) * 4;
            float* g2 = d_pos + ((size_t)b * V + c0.y) * 4;
            atomicAdd(g1, dd * __int_as_float(c0.z)); atomicAdd(g1 + 1, dd * __int_as_float(c0.w)); atomicAdd(g1 + 3, dd * c1.x);
            atomicAdd(g2, dd * c1.y); atomicAdd(g2 + 1, dd * c1.z); atomicAdd(g2 + 3, dd * c1.w);
        }
    }
}

// one thread per LOW-resolution pixel: sums the colour gradient of its up*up hi-res pixels
template <int C, int CC, bool POOL>
__global__ void __launch_bounds__(256) aa_up_bwd_kernel(const float* __restrict__ color, AAUp U, const float* __restrict__ bg, int Bg, int64_t V,
                                                        AAGrad G, AAContext ctx, int B, int aa, float* __restrict__ d_color,
                                                        float* __restrict__ d_pos, int pos_blocks)
{
    if ((int)blockIdx.x < pos_blocks) {
        aa_up_pos_role<C, POOL>(color, U, bg, Bg, V, G, ctx, d_pos, (int)blockIdx.x);
        return;
    }
    const int64_t n = (int64_t)B * U.lhw;
    const int64_t li = (int64_t)(blockIdx.x - pos_blocks) * blockDim.x + threadIdx.x;
    if (li >= n) return;
    const int b = (int)(li / U.lhw);
    const int lp = (int)(li - (int64_t)b * U.lhw);
    const int yl = lp / U.lw, xl = lp - yl * U.lw;
    float acc[CC];
#pragma unroll
    for (int c = 0; c < CC; c++) acc[c] = 0.f;
    for (int dy = 0; dy < U.up; dy++) {
        for (int dx = 0; dx < U.up; dx++) {
            const int p = (yl * U.up + dy) * U.W + xl * U.up + dx;
            const size_t flat = (size_t)b * U.HW + p;
            if (!aa_bit(ctx.cover, flat)) continue;
            float v[CC];
#pragma unroll
            for (int c = 0; c < CC; c++) v[c] = grad_up<POOL>(G, U, b, p, c);
            if (aa && aa_bit(ctx.act, flat)) {
                const float4 a = __ldg(ctx.rec + flat);
#pragma unroll
                for (int c = 0; c < CC; c++) {
                    const float own = v[c];
                    const float gu = a.x != 0.f ? (a.x > 0.f ? grad_up<POOL>(G, U, b, p - U.W, c) : own) : 0.f;
                    const float gl = a.y != 0.f ? (a.y > 0.f ? grad_up<POOL>(G, U, b, p - 1, c) : own) : 0.f;
                    const float gr = a.z != 0.f ? (a.z > 0.f ? own : grad_up<POOL>(G, U, b, p + 1, c)) : 0.f;
                    const float gd = a.w != 0.f ? (a.w > 0.f ? own : grad_up<POOL>(G, U, b, p + U.W, c)) : 0.f;
                    float t = own;
                    if (a.x != 0.f) t += a.x * gu;
                    if (a.y != 0.f) t += a.y * gl;
                    if (a.z != 0.f) t -= a.z * gr;
                    if (a.w != 0.f) t -= a.w * gd;
                    v[c] = t;
                }
            }
#pragma unroll
            for (int c = 0; c < CC; c++) acc[c] += v[c];
        }
    }
    float* o = d_color + li * CC;
#pragma unroll
    for (int c = 0; c < CC; c++) o[c] = acc[c];
}

int aa_check(const float* color, const float* rast, const float* pos, const int32_t* tri, const int32_t* opp, int Bg, int composite, int B,
             int64_t V, int64_t F, int H, int W, int C)
{
    B2A_CHECK_ARG(rast && pos && tri && opp, "null pointer");
    B2A_CHECK_ARG(color || (composite && C == 1), "null color");
    B2A_CHECK_ARG(B > 0 && B <= 65535 && V > 0 && F >= 0 && H > 0 && W > 0 && (int64_t)H * W < (1ll << 31) && C > 0 && C <= 1024, "shape");
    B2A_CHECK_ARG(Bg == 1 || Bg == B, "bg batch");
    B2A_CHECK_ARG(((uintptr_t)pos & 15) == 0 && ((uintptr_t)rast & 15) == 0, "pos/rast must be 16-byte aligned");
    return 0;
}

}  // namespace

B2A_API int b2a_edge_adjacency_workspace_bytes(int64_t F, size_t* bytes)
{
    B2A_CHECK_ARG(bytes && F >= 0, "sizes");
    *bytes = adj_layout(F, nullptr, nullptr);
    return 0;
}

B2A_API int b2a_edge_adjacency(const int32_t* tri, int64_t F, int64_t V, void* workspace, size_t workspace_bytes, int32_t* opp,
                               b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(tri && workspace && opp, "null pointer");
    B2A_CHECK_ARG(F >= 0 && F < (1ll << 29) && V > 0 && V < (1ll << 31), "sizes");
    AdjWorkspace ws;
    size_t need = adj_layout(F, workspace, &ws);
    B2A_CHECK_ARG(need <= workspace_bytes, "workspace too small");
    if (F == 0) return 0;
    size_t kb = (size_t)((char*)ws.t0 - (char*)ws.keys);
    B2A_CUDA_OK(cudaMemsetAsync(ws.keys, 0, kb, stream));
    B2A_CUDA_OK(cudaMemsetAsync(ws.t0, 0x7f, need - kb, stream));
    unsigned nb = b2a_blocks(F * 3, 256);
    adj_insert_kernel<<<nb, 256, 0, stream>>>(tri, F, V, ws);
    adj_second_kernel<<<nb, 256, 0, stream>>>(tri, F, V, ws);
    adj_emit_kernel<<<nb, 256, 0, stream>>>(tri, F, V, ws, opp);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_antialias_workspace_bytes(int B, int H, int W, size_t* bytes)
{
    B2A_CHECK_ARG(bytes && B > 0 && H > 0 && W > 0, "shape");
    *bytes = aa_ctx_layout(B, H, W, nullptr, nullptr);
    return 0;
}

B2A_API int b2a_antialias_prepare(const float* rast, const float* pos, const int32_t* tri, const int32_t* opp, int B, int64_t V, int64_t F,
                                  int H, int W, void* aa_ctx, size_t aa_ctx_bytes, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(rast && pos && tri && opp && aa_ctx, "null pointer");
    B2A_CHECK_ARG(B > 0 && H > 0 && W > 0 && (int64_t)B * H * W < (1ll << 31) && V > 0 && F >= 0 && F < (1ll << 28), "shape");
    B2A_CHECK_ARG(((int64_t)H * W) % 32 == 0, "H*W must be a multiple of 32 for the prepared fast path");
    B2A_CHECK_ARG(((uintptr_t)pos & 15) == 0 && ((uintptr_t)rast & 15) == 0 && ((uintptr_t)aa_ctx & 15) == 0, "pos/rast/aa_ctx must be 16-byte aligned");
    AAContext ctx;
    B2A_CHECK_ARG(aa_ctx_layout(B, H, W, aa_ctx, &ctx) <= aa_ctx_bytes, "context workspace too small");
    int64_t n = (int64_t)B * H * W;
    B2A_CUDA_OK(cudaMemsetAsync(ctx.act, 0, (size_t)((char*)ctx.count - (char*)ctx.act) + 2 * sizeof(int), stream));
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    aa_prepare_kernel<<<blocks, 256, 0, stream>>>(rast, B, H, W, ctx);
    AAParams P{nullptr, nullptr, rast, pos, tri, opp, 1, 1, B, H, W, 0, V, F};
    unsigned pblocks = (unsigned)((4 * n + 127) / 128);
    if (pblocks > 148u * 16u) pblocks = 148u * 16u;
    aa_pairs_kernel<<<pblocks, 128, 0, stream>>>(P, ctx);
    B2A_LAUNCH_OK();
    return 0;
}

namespace {
bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

bool aa_fast_ok(int composite, const void* aa_ctx, size_t aa_ctx_bytes, int B, int H, int W, int C, AAContext* ctx)
{
    if (!composite || !aa_ctx) return false;
    if (((int64_t)H * W) % 32 != 0 || (int64_t)B * H * W * C >= (1ll << 31)) return false;
    return aa_ctx_layout(B, H, W, const_cast<void*>(aa_ctx), ctx) <= aa_ctx_bytes;
}

template <int C>
void aa_fwd_tile(const float* color, const float* bg, int Bg, const AAContext& ctx, int B, int H, int W, float* out, cudaStream_t stream)
{
    if constexpr (C <= 4) {   // narrow keys: one thread per pixel
        aa_fwd_pix_kernel<C><<<b2a_blocks((int64_t)B * H * W, 256 * AA_PPT), 256, 0, stream>>>(color, bg, Bg, ctx, B, H, W, out);
        return;
    }
    constexpr int TPW = C >= 9 ? 1 : 4;
    unsigned tiles = (unsigned)(((int64_t)B * H * W) / 32);
    aa_fwd_tile_kernel<C, TPW><<<b2a_blocks(tiles, 8 * TPW), 256, 0, stream>>>(color, bg, Bg, ctx, B, H, W, out);
}
}  // namespace

B2A_API int b2a_antialias_fwd(const float* color, const float* bg, int Bg, int composite, const float* rast, const float* pos,
                              const int32_t* tri, const int32_t* opp, int B, int64_t V, int64_t F, int H, int W, int C, float* out,
                              const void* aa_ctx, size_t aa_ctx_bytes, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = aa_check(color, rast, pos, tri, opp, Bg, composite, B, V, F, H, W, C);
    if (rc) return rc;
    B2A_CHECK_ARG(out, "null pointer");
    AAParams P{color, bg, rast, pos, tri, opp, Bg, composite, B, H, W, C, V, F};
    AAContext ctx;
    bool fast = aa_fast_ok(composite, aa_ctx, aa_ctx_bytes, B, H, W, C, &ctx) && aligned16(color) && aligned16(bg) && aligned16(out);
    if (fast) {
        switch (C) {
            case 2: aa_fwd_tile<2>(color, bg, Bg, ctx, B, H, W, out, stream); break;
            case 3: aa_fwd_tile<3>(color, bg, Bg, ctx, B, H, W, out, stream); break;
            case 4: aa_fwd_tile<4>(color, bg, Bg, ctx, B, H, W, out, stream); break;
            case 17: aa_fwd_tile<17>(color, bg, Bg, ctx, B, H, W, out, stream); break;
            default: fast = false;
        }
    }
    if (!fast) aa_fwd_kernel<<<dim3(b2a_blocks((int64_t)H * W * C, 256), B), 256, 0, stream>>>(P, out);
    B2A_LAUNCH_OK();
    return 0;
}

namespace {
template <int C, int CC, int CG>
bool aa_bwd_tile(const AAParams& P, const AAGrad& G, const AAContext& ctx, float* d_color, float* d_pos, cudaStream_t stream)
{
    const int pos_blocks = d_pos ? AA_POS_BLOCKS : 0;
    if constexpr (CC <= 4) {   // narrow keys: one thread per pixel, any gradient strides
        unsigned grid = b2a_blocks((int64_t)P.B * P.H * P.W, 256 * AA_PPT) + pos_blocks;
        aa_bwd_pix_kernel<C, CC, CG><<<grid, 256, 0, stream>>>(P, G, ctx, d_color, d_pos, pos_blocks);
        return true;
    }
    constexpr int TPW = CC >= 8 ? 1 : 4;   // tiles per warp: keep >= ~12 loads in flight per lane
    unsigned tiles = (unsigned)(((int64_t)P.B * P.H * P.W) / 32);
    unsigned grid = b2a_blocks(tiles, 8 * TPW);
    const bool nhwc = G.sc == 1 && G.sx == CG && G.sy == (int64_t)P.W * CG && G.sb % 4 == 0 && aligned16(G.d_out);
    const bool nchw = G.sx == 1 && P.W % 32 == 0;
    if (!nhwc && !nchw) return false;
    const unsigned total = grid + pos_blocks;
    if (nhwc) aa_bwd_tile_kernel<C, CC, CG, false, TPW><<<total, 256, 0, stream>>>(P, G, ctx, d_color, d_pos, pos_blocks);
    else aa_bwd_tile_kernel<C, CC, CG, true, TPW><<<total, 256, 0, stream>>>(P, G, ctx, d_color, d_pos, pos_blocks);
    return true;
}
}  // namespace

B2A_API int b2a_antialias_bwd(const float* color, const float* bg, int Bg, int composite, const float* rast, const float* pos,
                              const int32_t* tri, const int32_t* opp, const float* d_out, int64_t d_sb, int64_t d_sy, int64_t d_sx,
                              int64_t d_sc, int Cg, int B, int64_t V, int64_t F, int H, int W, int C, float* d_color, float* d_pos,
                              const void* aa_ctx, size_t aa_ctx_bytes, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = aa_check(color, rast, pos, tri, opp, Bg, composite, B, V, F, H, W, C);
    if (rc) return rc;
    B2A_CHECK_ARG(d_out && Cg >= 0 && Cg <= C, "d_out");
    int Cc = composite ? C - 1 : C;
    if (Cc == 0) return 0;
    AAParams P{color, bg, rast, pos, tri, opp, Bg, composite, B, H, W, C, V, F};
    AAGrad G{d_out, d_sb, d_sy, d_sx, d_sc, Cg};
    AAContext ctx;
    bool fast = d_color && aa_fast_ok(composite, aa_ctx, aa_ctx_bytes, B, H, W, C, &ctx) && aligned16(d_color);
    if (fast) {
        if (C == 4 && Cg == 4) fast = aa_bwd_tile<4, 3, 4>(P, G, ctx, d_color, d_pos, stream);
        else if (C == 4 && Cg == 3) fast = aa_bwd_tile<4, 3, 3>(P, G, ctx, d_color, d_pos, stream);
        else if (C == 17 && Cg == 16) fast = aa_bwd_tile<17, 16, 16>(P, G, ctx, d_color, d_pos, stream);
        else if (C == 2 && Cg == 1) fast = aa_bwd_tile<2, 1, 1>(P, G, ctx, d_color, d_pos, stream);
        else if (C == 3 && Cg == 2) fast = aa_bwd_tile<3, 2, 2>(P, G, ctx, d_color, d_pos, stream);
        else fast = false;
    }
    if (!fast) aa_bwd_kernel<<<dim3(b2a_blocks((int64_t)H * W * Cc, 256), B), 256, 0, stream>>>(P, G, d_color, d_pos);
    B2A_LAUNCH_OK();
    return 0;
}

/* Two keys of one render in one launch per direction (the training pair: wide key dino_pred [16+1], narrow key shaded
 * [3+1]; composite mode with a prepared context only).  Returns B2A_ERR_UNSUPPORTED-style non-zero when the combination
 * has no fused instantiation: the caller then issues the two single-key calls (same kernels' bodies, same results). */
B2A_API int b2a_antialias_pair_fwd(const float* color_w, const float* bg_w, int Bg_w, int Cw, float* out_w, const float* color_n,
                                   const float* bg_n, int Bg_n, int Cn, float* out_n, int B, int H, int W, const void* aa_ctx,
                                   size_t aa_ctx_bytes, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(color_w && color_n && out_w && out_n && aa_ctx, "null pointer");
    B2A_CHECK_ARG(B > 0 && H > 0 && W > 0 && (Bg_w == 1 || Bg_w == B) && (Bg_n == 1 || Bg_n == B), "shape");
    B2A_CHECK_ARG(Cw == 17 && Cn == 4, "fused pair: only (wide 17, narrow 4) channels are instantiated");
    AAContext ctx;
    B2A_CHECK_ARG(aa_fast_ok(1, aa_ctx, aa_ctx_bytes, B, H, W, Cw, &ctx), "prepared context required (H*W % 32 == 0)");
    B2A_CHECK_ARG(aligned16(color_w) && aligned16(bg_w) && aligned16(out_w) && aligned16(out_n), "16-byte alignment");
    const unsigned tiles = (unsigned)(((int64_t)B * H * W) / 32);
    const unsigned wide_blocks = b2a_blocks(tiles, 8);
    const unsigned narrow_blocks = b2a_blocks((int64_t)B * H * W, 256 * AA_PPT);
    aa_fwd_pair_kernel<17, 4><<<wide_blocks + narrow_blocks, 256, 0, stream>>>(color_w, bg_w, Bg_w, out_w, color_n, bg_n, Bg_n, out_n, ctx, B, H, W,
                                                                               (int)wide_blocks);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_antialias_pair_bwd(const float* color_w, const float* bg_w, int Bg_w, int Cw, const float* d_out_w, int64_t w_sb, int64_t w_sy,
                                   int64_t w_sx, int64_t w_sc, int Cgw, float* d_color_w, const float* color_n, const float* bg_n, int Bg_n,
                                   int Cn, const float* d_out_n, int64_t n_sb, int64_t n_sy, int64_t n_sx, int64_t n_sc, int Cgn,
                                   float* d_color_n, int B, int64_t V, int H, int W, float* d_pos, const void* aa_ctx, size_t aa_ctx_bytes,
                                   b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(color_w && color_n && d_out_w && d_out_n && d_color_w && d_color_n && aa_ctx, "null pointer");
    B2A_CHECK_ARG(B > 0 && V > 0 && H > 0 && W > 0 && (Bg_w == 1 || Bg_w == B) && (Bg_n == 1 || Bg_n == B), "shape");
    B2A_CHECK_ARG(Cw == 17 && Cgw == 16 && Cn == 4 && (Cgn == 4 || Cgn == 3), "fused pair: only (wide 17/16, narrow 4/{3,4}) channels are instantiated");
    AAContext ctx;
    B2A_CHECK_ARG(aa_fast_ok(1, aa_ctx, aa_ctx_bytes, B, H, W, Cw, &ctx), "prepared context required (H*W % 32 == 0)");
    B2A_CHECK_ARG(aligned16(d_color_w) && aligned16(d_color_n), "16-byte alignment");
    AAParams Pw{color_w, bg_w, nullptr, nullptr, nullptr, nullptr, Bg_w, 1, B, H, W, Cw, V, 0};
    AAParams Pn{color_n, bg_n, nullptr, nullptr, nullptr, nullptr, Bg_n, 1, B, H, W, Cn, V, 0};
    AAGrad Gw{d_out_w, w_sb, w_sy, w_sx, w_sc, Cgw};
    AAGrad Gn{d_out_n, n_sb, n_sy, n_sx, n_sc, Cgn};
    const bool nhwc = w_sc == 1 && w_sx == Cgw && w_sy == (int64_t)W * Cgw && w_sb % 4 == 0 && aligned16(d_out_w);
    const bool nchw = w_sx == 1 && W % 32 == 0;
    B2A_CHECK_ARG(nhwc || nchw, "fused pair: the wide gradient must be NCHW- or NHWC-contiguous");
    // Position-gradient blocks per key, leading the grid.  Measured at C1 (1 M pixels, ~5 k active pixels): 148 blocks 38.4 us,
    // 592 41.0 us, 1184 41.1 us; at the END of the grid 592 40.7 us, 148 47.1 us, 37 80.3 us (their latency chains must overlap
    // the streaming blocks).  Scaled with the image area so that a 2048^2 render (8x the silhouette) keeps ~600.
    // B2A_AA_POS="<n>[,last]" overrides (tuning switch).
    static int s_pos_override = -1, s_pos_first = 1;
    if (s_pos_override < 0) {
        s_pos_override = 0;
        const char* e = getenv("B2A_AA_POS");
        if (e) {
            int n = atoi(e);
            if (n > 0 && n <= 4096) s_pos_override = n;
            if (strstr(e, "last")) s_pos_first = -1;
        }
    }
    int64_t scaled = ((int64_t)B * H * W) >> 13;
    if (scaled < 148) scaled = 148;
    if (scaled > 2 * AA_POS_BLOCKS) scaled = 2 * AA_POS_BLOCKS;
    const int pos_blocks = d_pos ? (s_pos_override ? s_pos_override : (int)scaled) : 0;
    // Measured alternatives at C1 (B2A_AA_VAR sweep, round 1): (256,4) launch bounds / 64 registers, no spills: 42.3 us;
    // two tiles per wide warp (32 loads in flight per lane): 41.5 us; this configuration ((256,5), one tile per warp): 38.6 us.
    const unsigned tiles = (unsigned)(((int64_t)B * H * W) / 32);
    const unsigned wide_blocks = b2a_blocks(tiles, 8);
    const unsigned narrow_blocks = b2a_blocks((int64_t)B * H * W, 256 * AA_PPT);
    const unsigned grid = 2 * pos_blocks + wide_blocks + narrow_blocks;
#define B2A_PAIR_BWD(NCHW_, CGN_) \
    aa_bwd_pair_kernel<17, 16, NCHW_, 4, CGN_><<<grid, 256, 0, stream>>>(Pw, Gw, d_color_w, Pn, Gn, d_color_n, ctx, d_pos, pos_blocks, (int)wide_blocks, s_pos_first)
    if (nhwc) { if (Cgn == 4) B2A_PAIR_BWD(false, 4); else B2A_PAIR_BWD(false, 3); }
    else      { if (Cgn == 4) B2A_PAIR_BWD(true, 4);  else B2A_PAIR_BWD(true, 3); }
#undef B2A_PAIR_BWD
    B2A_LAUNCH_OK();
    return 0;
}

namespace {
bool up_args_ok(int up, int B, int H, int W, int C, const void* aa_ctx, size_t aa_ctx_bytes, AAContext* ctx, AAUp* U)
{
    if (up < 1 || H % up || W % up || C < 2 || C > 4 || !aa_fast_ok(1, aa_ctx, aa_ctx_bytes, B, H, W, C, ctx)) return false;
    *U = AAUp{up, W, H * W, W / up, (H / up) * (W / up)};
    return true;
}
}  // namespace

B2A_API int b2a_composite_up_fwd(const float* color, int up, const float* bg, int Bg, int antialias, int B, int H, int W, int C,
                                 float* out, const void* aa_ctx, size_t aa_ctx_bytes, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(color && out && aa_ctx, "null pointer");
    B2A_CHECK_ARG(B > 0 && H > 0 && W > 0 && (Bg == 1 || Bg == B), "shape");
    AAContext ctx;
    AAUp U;
    B2A_CHECK_ARG(up_args_ok(up, B, H, W, C, aa_ctx, aa_ctx_bytes, &ctx, &U), "needs C in 2..4, H and W multiples of `up`, and a prepared context");
    B2A_CHECK_ARG(aligned16(out) || C == 3, "out must be 16-byte aligned");
    const unsigned blocks = b2a_blocks((int64_t)B * H * W, 256);
    switch (C) {
        case 2: aa_up_fwd_kernel<2><<<blocks, 256, 0, stream>>>(color, U, bg, Bg, ctx, B, antialias, out); break;
        case 3: aa_up_fwd_kernel<3><<<blocks, 256, 0, stream>>>(color, U, bg, Bg, ctx, B, antialias, out); break;
        default: aa_up_fwd_kernel<4><<<blocks, 256, 0, stream>>>(color, U, bg, Bg, ctx, B, antialias, out); break;
    }
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_composite_up_bwd(const float* color, int up, const float* bg, int Bg, int antialias, const float* d_out, int64_t d_sb,
                                 int64_t d_sy, int64_t d_sx, int64_t d_sc, int Cg, int B, int64_t V, int H, int W, int C, float* d_color,
                                 float* d_pos, const void* aa_ctx, size_t aa_ctx_bytes, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(color && d_out && d_color && aa_ctx, "null pointer");
    B2A_CHECK_ARG(B > 0 && V > 0 && H > 0 && W > 0 && (Bg == 1 || Bg == B) && Cg >= 0 && Cg <= C, "shape");
    AAContext ctx;
    AAUp U;
    B2A_CHECK_ARG(up_args_ok(up, B, H, W, C, aa_ctx, aa_ctx_bytes, &ctx, &U), "needs C in 2..4, H and W multiples of `up`, and a prepared context");
    AAGrad G{d_out, d_sb, d_sy, d_sx, d_sc, Cg};
    const int pos_blocks = (d_pos && antialias) ? AA_POS_BLOCKS : 0;
    const unsigned grid = b2a_blocks((int64_t)B * U.lhw, 256) + pos_blocks;
    switch (C) {
        case 2: aa_up_bwd_kernel<2, 1, false><<<grid, 256, 0, stream>>>(color, U, bg, Bg, V, G, ctx, B, antialias, d_color, d_pos, pos_blocks); break;
        case 3: aa_up_bwd_kernel<3, 2, false><<<grid, 256, 0, stream>>>(color, U, bg, Bg, V, G, ctx, B, antialias, d_color, d_pos, pos_blocks); break;
        default: aa_up_bwd_kernel<4, 3, false><<<grid, 256, 0, stream>>>(color, U, bg, Bg, V, G, ctx, B, antialias, d_color, d_pos, pos_blocks); break;
    }
    B2A_LAUNCH_OK();
    return 0;
}

// msaa renders (reference render.py:217-219 nearest up-sampling, :258-268 composite + antialias at [H,W] = the raster resolution,
// :322-323 util.avg_pool_nhwc(., spp)) in ONE kernel per direction: out [B, H/up, W/up, keep] is the up x up average of the composited
// (+ antialiased) image, which is never materialised.  d_out: the gradient of that average, any strides (sb, sy, sx, sc over
// [B, H/up, W/up, Cg]).  Bit-identical to b2a_composite_up_fwd followed by avg_pool2d.
B2A_API int b2a_composite_up_pool_fwd(const float* color, int up, const float* bg, int Bg, int antialias, int B, int H, int W, int C, int keep,
                                      float* out, const void* aa_ctx, size_t aa_ctx_bytes, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(color && out && aa_ctx, "null pointer");
    B2A_CHECK_ARG(B > 0 && H > 0 && W > 0 && (Bg == 1 || Bg == B) && keep >= 1 && keep <= C && up >= 1, "shape");
    AAContext ctx;
    AAUp U;
    B2A_CHECK_ARG(up_args_ok(up, B, H, W, C, aa_ctx, aa_ctx_bytes, &ctx, &U), "needs C in 2..4, H and W multiples of `up`, and a prepared context");
    const unsigned blocks = b2a_blocks((int64_t)B * U.lhw, 256);
    switch (C) {
        case 2: aa_up_pool_fwd_kernel<2><<<blocks, 256, 0, stream>>>(color, U, bg, Bg, ctx, B, antialias, keep, out); break;
        case 3: aa_up_pool_fwd_kernel<3><<<blocks, 256, 0, stream>>>(color, U, bg, Bg, ctx, B, antialias, keep, out); break;
        default: aa_up_pool_fwd_kernel<4><<<blocks, 256, 0, stream>>>(color, U, bg, Bg, ctx, B, antialias, keep, out); break;
    }
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_composite_up_pool_bwd(const float* color, int up, const float* bg, int Bg, int antialias, const float* d_out, int64_t d_sb,
                                      int64_t d_sy, int64_t d_sx, int64_t d_sc, int Cg, int B, int64_t V, int H, int W, int C, float* d_color,
                                      float* d_pos, const void* aa_ctx, size_t aa_ctx_bytes, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(color && d_out && d_color && aa_ctx, "null pointer");
    B2A_CHECK_ARG(B > 0 && V > 0 && H > 0 && W > 0 && (Bg == 1 || Bg == B) && Cg >= 0 && Cg <= C, "shape");
    AAContext ctx;
    AAUp U;
    B2A_CHECK_ARG(up_args_ok(up, B, H, W, C, aa_ctx, aa_ctx_bytes, &ctx, &U), "needs C in 2..4, H and W multiples of `up`, and a prepared context");
    AAGrad G{d_out, d_sb, d_sy, d_sx, d_sc, Cg};
    const int pos_blocks = (d_pos && antialias) ? AA_POS_BLOCKS : 0;
    const unsigned grid = b2a_blocks((int64_t)B * U.lhw, 256) + pos_blocks;
    switch (C) {
        case 2: aa_up_bwd_kernel<2, 1, true><<<grid, 256, 0, stream>>>(color, U, bg, Bg, V, G, ctx, B, antialias, d_color, d_pos, pos_blocks); break;
        case 3: aa_up_bwd_kernel<3, 2, true><<<grid, 256, 0, stream>>>(color, U, bg, Bg, V, G, ctx, B, antialias, d_color, d_pos, pos_blocks); break;
        default: aa_up_bwd_kernel<4, 3, true><<<grid, 256, 0, stream>>>(color, U, bg, Bg, V, G, ctx, B, antialias, d_color, d_pos, pos_blocks); break;
    }
    B2A_LAUNCH_OK();
    return 0;
}
