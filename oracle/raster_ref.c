/*
 * oracle/raster_ref.c - CPU restatement of the rasterizer-side arithmetic of the 3DAnimals hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Never linked into, imported by, or called from the
 * product library; used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
 *
 * What it restates
 * ----------------
 * The reference renders through `nvdiffrast.torch` (NVlabs; installed un-pinned from git, INSTALL.md:22; NOT
 * under /root/reference).  Call sites: model/render/render.py:24 (interpolate), :264 (antialias), :292-294
 * (DepthPeeler.rasterize_next_layer, first layer only since num_layers=1, models/AnimalModel.py:247).
 * There is no golden vector or test for this boundary in the reference: PARITY UNPINNED.
 * The semantics below follow nvdiffrast's published algorithm (common/rasterize.cu, interpolate.cu,
 * antialias.cu, recalled) with the fill rule fixed and documented here:
 *
 *   rasterize  - pixel centre (px+.5, py+.5); image row 0 is clip y=-1 (GL convention, so the reference's
 *                negated-y projection render/util.py:189-194 comes out upright).  Coverage and barycentrics by
 *                2-D homogeneous edge functions on clip-space (x,y,w):  q_i = (x_i - fx*w_i, y_i - fy*w_i),
 *                a0 = q1 x q2, a1 = q2 x q0, a2 = q0 x q1, S = a0+a1+a2.  A pixel is inside when every a_i has
 *                the sign of S or is zero (inclusive edges; both orientations rendered, no culling), the
 *                interpolated w has the sign of S (in front of the eye; this replaces polygon clipping), and
 *                -1 <= z/w <= 1.  Nearest z/w wins; equal z/w -> lowest triangle index (GL_LESS draw order).
 *                Every operation is an individually rounded fp32 op (compile with -ffp-contract=off; the CUDA
 *                product compiles with -fmad=false), so the triangle-id buffer is bit-reproducible.
 *                Output (u, v, z/w, id+1): u,v = a0/S, a1/S saturated to [0,1]; id+1 = 0 for empty pixels.
 *                Backward: d(u,v) -> d(x,y,w) of the three vertices; nothing through z/w or id.
 *   interpolate- out = u*A[i0] + v*A[i1] + (1-u-v)*A[i2]; zeros on empty pixels; attr batch 1 broadcasts.
 *                Backward: scatter-add to attr, and d(u,v).
 *   antialias  - for each horizontally / vertically adjacent pixel pair with different ids: take the nearer
 *                surface's triangle, keep only its silhouette edges (no neighbour across the edge, or the
 *                neighbour's opposite vertex on the same screen side), intersect with the segment between the
 *                pixel centres, blend the two colours by (0.5 - crossing distance).  Backward: colour and the
 *                two edge vertices' clip positions.  Edge adjacency: per undirected edge the two lowest
 *                triangle indices are kept; a triangle's neighbour is the lower one that is not itself.
 *
 * Also here (used for the CPU baseline so that the whole render leg is native code): clip transform
 * (renderutils/ops.py:524-525), vertex normals (render/mesh.py:276-304) forward and backward.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------ */
/* clip transform: out[b,v,:] = [p,1] . M[b]^T   (ops.py:524-525)                                    */
/* ------------------------------------------------------------------------------------------------ */
ORC_API void orc_xfm_points_fwd(const float* pts, const float* mtx, int B, int Bp, int V, float* out)
{
#pragma omp parallel for
    for (int b = 0; b < B; b++) {
        const float* m = mtx + (size_t)b * 16;
        const float* p = pts + (size_t)(Bp == 1 ? 0 : b) * V * 3;
        float* o = out + (size_t)b * V * 4;
        for (int v = 0; v < V; v++) {
            float x = p[v * 3], y = p[v * 3 + 1], z = p[v * 3 + 2];
            for (int r = 0; r < 4; r++)
                o[v * 4 + r] = ((m[r * 4] * x + m[r * 4 + 1] * y) + m[r * 4 + 2] * z) + m[r * 4 + 3];
        }
    }
}

/* d_pts[b or 0,v,:] += M[b][:, :3]^T d_out ; d_mtx[b][r][c] += sum_v d_out[r] * [p,1][c] */
ORC_API void orc_xfm_points_bwd(const float* pts, const float* mtx, const float* d_out, int B, int Bp, int V,
                                float* d_pts, float* d_mtx)
{
    for (int b = 0; b < B; b++) {
        const float* m = mtx + (size_t)b * 16;
        const float* p = pts + (size_t)(Bp == 1 ? 0 : b) * V * 3;
        float* dp = d_pts ? d_pts + (size_t)(Bp == 1 ? 0 : b) * V * 3 : 0;
        const float* g = d_out + (size_t)b * V * 4;
        double acc[16];
        for (int i = 0; i < 16; i++) acc[i] = 0;
        for (int v = 0; v < V; v++) {
            float h[4] = {p[v * 3], p[v * 3 + 1], p[v * 3 + 2], 1.f};
            for (int r = 0; r < 4; r++) {
                float gr = g[v * 4 + r];
                if (dp) for (int c = 0; c < 3; c++) dp[v * 3 + c] += m[r * 4 + c] * gr;
                for (int c = 0; c < 4; c++) acc[r * 4 + c] += (double)gr * h[c];
            }
        }
        if (d_mtx) for (int i = 0; i < 16; i++) d_mtx[(size_t)b * 16 + i] += (float)acc[i];
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* rasterize                                                                                        */
/* ------------------------------------------------------------------------------------------------ */
static inline uint32_t depth_key(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

typedef struct { float a0, a1, a2, S, zw, u, v; } TriEval;

/* Shared by the z-buffer scatter and the resolve so both see identical bits. */
static inline int tri_eval(const float* p0, const float* p1, const float* p2, float fx, float fy, TriEval* e)
{
    float q0x = p0[0] - fx * p0[3], q0y = p0[1] - fy * p0[3];
    float q1x = p1[0] - fx * p1[3], q1y = p1[1] - fy * p1[3];
    float q2x = p2[0] - fx * p2[3], q2y = p2[1] - fy * p2[3];
    float a0 = q1x * q2y - q1y * q2x;
    float a1 = q2x * q0y - q2y * q0x;
    float a2 = q0x * q1y - q0y * q1x;
    float S = (a0 + a1) + a2;
    if (S > 0.f) { if (a0 < 0.f || a1 < 0.f || a2 < 0.f) return 0; }
    else if (S < 0.f) { if (a0 > 0.f || a1 > 0.f || a2 > 0.f) return 0; }
    else return 0; /* zero area or NaN */
    float z = (p0[2] * a0 + p1[2] * a1) + p2[2] * a2;
    float w = (p0[3] * a0 + p1[3] * a1) + p2[3] * a2;
    if (S > 0.f ? !(w > 0.f) : !(w < 0.f)) return 0;
    float zw = z / w;
    if (!(zw >= -1.f && zw <= 1.f)) return 0;
    float iw = 1.f / S;
    float u = a0 * iw, v = a1 * iw;
    e->a0 = a0; e->a1 = a1; e->a2 = a2; e->S = S; e->zw = zw;
    e->u = u < 0.f ? 0.f : (u > 1.f ? 1.f : u);
    e->v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
    return 1;
}

static inline void pixel_ndc(int px, int py, int H, int W, float* fx, float* fy)
{
    float xs = 2.f / (float)W, ys = 2.f / (float)H;
    *fx = (float)px * xs + (xs * 0.5f - 1.f);
    *fy = (float)py * ys + (ys * 0.5f - 1.f);
}

/* conservative pixel bounding box of a triangle; full screen when any w <= 0 */
static inline int tri_bbox(const float* p0, const float* p1, const float* p2, int H, int W,
                           int* x0, int* x1, int* y0, int* y1)
{
    if (!(p0[3] > 0.f) && !(p1[3] > 0.f) && !(p2[3] > 0.f)) return 0;
    if (p0[3] > 1e-6f && p1[3] > 1e-6f && p2[3] > 1e-6f) {
        float sx0 = (p0[0] / p0[3] * 0.5f + 0.5f) * W, sy0 = (p0[1] / p0[3] * 0.5f + 0.5f) * H;
        float sx1 = (p1[0] / p1[3] * 0.5f + 0.5f) * W, sy1 = (p1[1] / p1[3] * 0.5f + 0.5f) * H;
        float sx2 = (p2[0] / p2[3] * 0.5f + 0.5f) * W, sy2 = (p2[1] / p2[3] * 0.5f + 0.5f) * H;
        float mnx = fminf(sx0, fminf(sx1, sx2)), mxx = fmaxf(sx0, fmaxf(sx1, sx2));
        float mny = fminf(sy0, fminf(sy1, sy2)), mxy = fmaxf(sy0, fmaxf(sy1, sy2));
        if (!(mxx >= 0.f && mnx <= (float)W && mxy >= 0.f && mny <= (float)H)) return 0; /* also NaN */
        /* pixel centre px+.5 in [mn,mx]  ->  px in [mn-.5, mx-.5]; widen by one pixel for rounding */
        *x0 = (int)fmaxf(floorf(mnx - 0.5f) - 1.f, 0.f);
        *x1 = (int)fminf(ceilf(mxx - 0.5f) + 1.f, (float)(W - 1));
        *y0 = (int)fmaxf(floorf(mny - 0.5f) - 1.f, 0.f);
        *y1 = (int)fminf(ceilf(mxy - 0.5f) + 1.f, (float)(H - 1));
        return 1;
    }
    *x0 = 0; *x1 = W - 1; *y0 = 0; *y1 = H - 1;
    return 1;
}

ORC_API void orc_rasterize_fwd(const float* pos, const int* tri, int B, int V, int F, int H, int W, float* rast)
{
    size_t npix = (size_t)B * H * W;
    uint64_t* zbuf = (uint64_t*)malloc(npix * sizeof(uint64_t));
    memset(zbuf, 0xff, npix * sizeof(uint64_t));
    long long total = (long long)B * F;
#pragma omp parallel for schedule(dynamic, 1024)
    for (long long i = 0; i < total; i++) {
        int b = (int)(i / F), f = (int)(i % F);
        int i0 = tri[f * 3], i1 = tri[f * 3 + 1], i2 = tri[f * 3 + 2];
        if (i0 < 0 || i0 >= V || i1 < 0 || i1 >= V || i2 < 0 || i2 >= V) continue;
        const float* pb = pos + (size_t)b * V * 4;
        const float *p0 = pb + (size_t)i0 * 4, *p1 = pb + (size_t)i1 * 4, *p2 = pb + (size_t)i2 * 4;
        int x0, x1, y0, y1;
        if (!tri_bbox(p0, p1, p2, H, W, &x0, &x1, &y0, &y1)) continue;
        for (int py = y0; py <= y1; py++)
            for (int px = x0; px <= x1; px++) {
                float fx, fy;
                pixel_ndc(px, py, H, W, &fx, &fy);
                TriEval e;
                if (!tri_eval(p0, p1, p2, fx, fy, &e)) continue;
                uint64_t key = ((uint64_t)depth_key(e.zw) << 32) | (uint32_t)f;
                uint64_t* slot = zbuf + ((size_t)b * H + py) * W + px;
                uint64_t old = __atomic_load_n(slot, __ATOMIC_RELAXED);
                while (key < old &&
                       !__atomic_compare_exchange_n(slot, &old, key, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
            }
    }
#pragma omp parallel for
    for (long long r = 0; r < (long long)B * H; r++) {
        int b = (int)(r / H), py = (int)(r % H);
        const float* pb = pos + (size_t)b * V * 4;
        for (int px = 0; px < W; px++) {
            size_t pi = ((size_t)b * H + py) * W + px;
            float* o = rast + pi * 4;
            uint64_t key = zbuf[pi];
            o[0] = o[1] = o[2] = o[3] = 0.f;
            if (key == UINT64_MAX) continue;
            int f = (int)(uint32_t)key;
            float fx, fy;
            pixel_ndc(px, py, H, W, &fx, &fy);
            TriEval e;
            if (!tri_eval(pb + (size_t)tri[f * 3] * 4, pb + (size_t)tri[f * 3 + 1] * 4, pb + (size_t)tri[f * 3 + 2] * 4,
                          fx, fy, &e)) continue;
            o[0] = e.u; o[1] = e.v; o[2] = e.zw; o[3] = (float)(f + 1);
        }
    }
    free(zbuf);
}

/* d_pos[b,v,(x,y,w)] += ... from d_rast[...,0:2]  (nothing through z/w or id) */
ORC_API void orc_rasterize_bwd(const float* pos, const int* tri, const float* rast, const float* d_rast,
                               int B, int V, int F, int H, int W, float* d_pos)
{
#pragma omp parallel for
    for (int b = 0; b < B; b++) {
        const float* pb = pos + (size_t)b * V * 4;
        float* gb = d_pos + (size_t)b * V * 4;
        for (int py = 0; py < H; py++)
            for (int px = 0; px < W; px++) {
                size_t pi = ((size_t)b * H + py) * W + px;
                int f = (int)rast[pi * 4 + 3] - 1;
                if (f < 0 || f >= F) continue;
                float du = d_rast[pi * 4], dv = d_rast[pi * 4 + 1];
                if (du == 0.f && dv == 0.f) continue;
                int vi[3] = {tri[f * 3], tri[f * 3 + 1], tri[f * 3 + 2]};
                const float *p0 = pb + (size_t)vi[0] * 4, *p1 = pb + (size_t)vi[1] * 4, *p2 = pb + (size_t)vi[2] * 4;
                float fx, fy;
                pixel_ndc(px, py, H, W, &fx, &fy);
                float q0x = p0[0] - fx * p0[3], q0y = p0[1] - fy * p0[3];
                float q1x = p1[0] - fx * p1[3], q1y = p1[1] - fy * p1[3];
                float q2x = p2[0] - fx * p2[3], q2y = p2[1] - fy * p2[3];
                float a0 = q1x * q2y - q1y * q2x, a1 = q2x * q0y - q2y * q0x, a2 = q0x * q1y - q0y * q1x;
                float S = (a0 + a1) + a2;
                float iw = 1.f / S;
                float u = a0 * iw, v = a1 * iw;
                float gs = u * du + v * dv;
                float ga0 = (du - gs) * iw, ga1 = (dv - gs) * iw, ga2 = -gs * iw;
                float gq0x = ga2 * q1y - ga1 * q2y, gq0y = ga1 * q2x - ga2 * q1x;
                float gq1x = ga0 * q2y - ga2 * q0y, gq1y = ga2 * q0x - ga0 * q2x;
                float gq2x = ga1 * q0y - ga0 * q1y, gq2y = ga0 * q1x - ga1 * q0x;
                float* g0 = gb + (size_t)vi[0] * 4; float* g1 = gb + (size_t)vi[1] * 4; float* g2 = gb + (size_t)vi[2] * 4;
                g0[0] += gq0x; g0[1] += gq0y; g0[3] += -(fx * gq0x + fy * gq0y);
                g1[0] += gq1x; g1[1] += gq1y; g1[3] += -(fx * gq1x + fy * gq1y);
                g2[0] += gq2x; g2[1] += gq2y; g2[3] += -(fx * gq2x + fy * gq2y);
            }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* interpolate                                                                                      */
/* ------------------------------------------------------------------------------------------------ */
ORC_API void orc_interpolate_fwd(const float* attr, const float* rast, const int* tri, int B, int Ba, int V, int F,
                                 int H, int W, int C, float* out)
{
#pragma omp parallel for
    for (long long r = 0; r < (long long)B * H; r++) {
        int b = (int)(r / H);
        const float* ab = attr + (size_t)(Ba == 1 ? 0 : b) * V * C;
        for (int px = 0; px < W; px++) {
            size_t pi = (size_t)r * W + px;
            float* o = out + pi * C;
            int f = (int)rast[pi * 4 + 3] - 1;
            if (f < 0 || f >= F) { for (int c = 0; c < C; c++) o[c] = 0.f; continue; }
            float u = rast[pi * 4], v = rast[pi * 4 + 1], w = 1.f - u - v;
            const float *A0 = ab + (size_t)tri[f * 3] * C, *A1 = ab + (size_t)tri[f * 3 + 1] * C, *A2 = ab + (size_t)tri[f * 3 + 2] * C;
            for (int c = 0; c < C; c++) o[c] = (u * A0[c] + v * A1[c]) + w * A2[c];
        }
    }
}

ORC_API void orc_interpolate_bwd(const float* attr, const float* rast, const int* tri, const float* d_out,
                                 int B, int Ba, int V, int F, int H, int W, int C, float* d_attr, float* d_rast)
{
    /* sequential over images when the attribute is broadcast (shared accumulator) */
#pragma omp parallel for if (Ba != 1)
    for (int b = 0; b < B; b++) {
        const float* ab = attr + (size_t)(Ba == 1 ? 0 : b) * V * C;
        float* gab = d_attr ? d_attr + (size_t)(Ba == 1 ? 0 : b) * V * C : 0;
        for (int py = 0; py < H; py++)
            for (int px = 0; px < W; px++) {
                size_t pi = ((size_t)b * H + py) * W + px;
                float* gr = d_rast ? d_rast + pi * 4 : 0;
                if (gr) gr[0] = gr[1] = gr[2] = gr[3] = 0.f;
                int f = (int)rast[pi * 4 + 3] - 1;
                if (f < 0 || f >= F) continue;
                float u = rast[pi * 4], v = rast[pi * 4 + 1], w = 1.f - u - v;
                size_t o0 = (size_t)tri[f * 3] * C, o1 = (size_t)tri[f * 3 + 1] * C, o2 = (size_t)tri[f * 3 + 2] * C;
                const float* g = d_out + pi * C;
                float du = 0.f, dv = 0.f;
                for (int c = 0; c < C; c++) {
                    float gc = g[c];
                    if (gab) { gab[o0 + c] += u * gc; gab[o1 + c] += v * gc; gab[o2 + c] += w * gc; }
                    du += gc * (ab[o0 + c] - ab[o2 + c]);
                    dv += gc * (ab[o1 + c] - ab[o2 + c]);
                }
                if (gr) { gr[0] = du; gr[1] = dv; }
            }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* edge adjacency: opp[f][e] = third vertex of the neighbour across edge e (e=0:(v1,v2) 1:(v2,v0) 2:(v0,v1)) */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { int64_t key; int f; } EdgeRec;
static int edge_cmp(const void* a, const void* b)
{
    const EdgeRec* x = (const EdgeRec*)a; const EdgeRec* y = (const EdgeRec*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->f < y->f ? -1 : (x->f > y->f ? 1 : 0);
}

ORC_API void orc_edge_adjacency(const int* tri, int F, int V, int* opp)
{
    size_t n = (size_t)F * 3;
    EdgeRec* recs = (EdgeRec*)malloc(n * sizeof(EdgeRec));
    for (int f = 0; f < F; f++)
        for (int e = 0; e < 3; e++) {
            int a = tri[f * 3 + (e + 1) % 3], b = tri[f * 3 + (e + 2) % 3];
            int lo = a < b ? a : b, hi = a < b ? b : a;
            recs[(size_t)f * 3 + e].key = (int64_t)lo * (int64_t)(V + 1) + hi;
            recs[(size_t)f * 3 + e].f = f;
        }
    qsort(recs, n, sizeof(EdgeRec), edge_cmp);
    for (size_t i = 0; i < n; i++) opp[i] = -1;
    size_t i = 0;
    while (i < n) {
        size_t j = i;
        while (j < n && recs[j].key == recs[i].key) j++;
        /* distinct triangle ids in ascending order: t0 = lowest, t1 = second lowest */
        int t0 = recs[i].f, t1 = -1;
        for (size_t k = i + 1; k < j; k++) if (recs[k].f != t0) { t1 = recs[k].f; break; }
        int lo = (int)(recs[i].key / (V + 1)), hi = (int)(recs[i].key % (V + 1));
        for (size_t k = i; k < j; k++) {
            int f = recs[k].f;
            int partner = (f == t0) ? t1 : t0;
            if (partner < 0) continue;
            int ov = -1;
            for (int c = 0; c < 3; c++) { int vv = tri[partner * 3 + c]; if (vv != lo && vv != hi) { ov = vv; break; } }
            for (int e = 0; e < 3; e++) {
                int a = tri[f * 3 + (e + 1) % 3], b = tri[f * 3 + (e + 2) % 3];
                if ((a == lo && b == hi) || (a == hi && b == lo)) opp[f * 3 + e] = ov;
            }
        }
        i = j;
    }
    free(recs);
}

/* ------------------------------------------------------------------------------------------------ */
/* antialias                                                                                        */
/* ------------------------------------------------------------------------------------------------ */
static inline int same_sign(float a, float b)
{
    int32_t x, y;
    memcpy(&x, &a, 4); memcpy(&y, &b, 4);
    return (x ^ y) >= 0;
}

typedef struct {
    int ok;        /* pair produces a blend */
    float alpha;   /* blend weight, target pixel = alpha > 0 ? p0 : p1 */
    int tri, di;   /* owning triangle and which of its edges (0:(v1,v2) 1:(v2,v0) 2:(v0,v1)) */
    int px, py;    /* pixel the analysis is relative to (the owner's pixel) */
} AAPair;

#define ORC_F32_MAX 3.402823466e+38f

static void aa_analyze(const float* rast_b, const float* pos_b, const int* tri, const int* opp, int F, int V, int H,
                       int W, int px, int py, int d, AAPair* r)
{
    r->ok = 0;
    size_t pidx0 = (size_t)py * W + px, pidx1 = pidx0 + (d ? W : 1);
    float z0 = rast_b[pidx0 * 4 + 2], z1 = rast_b[pidx1 * 4 + 2];
    int tri0 = (int)rast_b[pidx0 * 4 + 3] - 1, tri1 = (int)rast_b[pidx1 * 4 + 3] - 1;
    if (tri0 == tri1) return;
    int t = (tri0 >= 0) ? tri0 : tri1;
    if (tri0 >= 0 && tri1 >= 0) t = (z0 < z1) ? tri0 : tri1;
    if (t == tri1) { px += 1 - d; py += d; }
    if (t < 0 || t >= F) return;
    int vi0 = tri[t * 3], vi1 = tri[t * 3 + 1], vi2 = tri[t * 3 + 2];
    if (vi0 < 0 || vi0 >= V || vi1 < 0 || vi1 >= V || vi2 < 0 || vi2 >= V) return;
    int op0 = opp[t * 3], op1 = opp[t * 3 + 1], op2 = opp[t * 3 + 2];
    if (op0 < 0) op0 = vi0;
    if (op1 < 0) op1 = vi1;
    if (op2 < 0) op2 = vi2;
    const float *p0 = pos_b + (size_t)vi0 * 4, *p1 = pos_b + (size_t)vi1 * 4, *p2 = pos_b + (size_t)vi2 * 4;
    const float *o0 = pos_b + (size_t)op0 * 4, *o1 = pos_b + (size_t)op1 * 4, *o2 = pos_b + (size_t)op2 * 4;
    float xh = 0.5f * (float)W, yh = 0.5f * (float)H;
    float fx = (float)px + 0.5f - xh, fy = (float)py + 0.5f - yh;
    float w0 = 1.f / p0[3], w1 = 1.f / p1[3], w2 = 1.f / p2[3];
    float ow0 = 1.f / o0[3], ow1 = 1.f / o1[3], ow2 = 1.f / o2[3];
    float x0 = p0[0] * w0 * xh - fx, y0 = p0[1] * w0 * yh - fy;
    float x1 = p1[0] * w1 * xh - fx, y1 = p1[1] * w1 * yh - fy;
    float x2 = p2[0] * w2 * xh - fx, y2 = p2[1] * w2 * yh - fy;
    float ox0 = o0[0] * ow0 * xh - fx, oy0 = o0[1] * ow0 * yh - fy;
    float ox1 = o1[0] * ow1 * xh - fx, oy1 = o1[1] * ow1 * yh - fy;
    float ox2 = o2[0] * ow2 * xh - fx, oy2 = o2[1] * ow2 * yh - fy;
    float bb = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
    float a0 = (x1 - ox0) * (y2 - oy0) - (x2 - ox0) * (y1 - oy0);
    float a1 = (x2 - ox1) * (y0 - oy1) - (x0 - ox1) * (y2 - oy1);
    float a2 = (x0 - ox2) * (y1 - oy2) - (x1 - ox2) * (y0 - oy2);
    int s0 = same_sign(a0, bb), s1 = same_sign(a1, bb), s2 = same_sign(a2, bb);
    if (!(s0 || s1 || s2)) return;
    if (d) { float tmp; tmp = x0; x0 = y0; y0 = tmp; tmp = x1; x1 = y1; y1 = tmp; tmp = x2; x2 = y2; y2 = tmp; }
    float dx0 = x2 - x1, dx1 = x0 - x2, dx2 = x1 - x0;
    float dy0 = y2 - y1, dy1 = y0 - y2, dy2 = y1 - y0;
    float ds = (t == tri0) ? 1.f : -1.f;
    /* crossing distance of each edge's line with the row through the pixel centre, towards the neighbour */
    float c0 = -ORC_F32_MAX, c1 = -ORC_F32_MAX, c2 = -ORC_F32_MAX;
    if (!same_sign(y1, y2)) c0 = ds * (x1 * dy0 - y1 * dx0) / dy0;
    if (!same_sign(y2, y0)) c1 = ds * (x2 * dy1 - y2 * dx1) / dy1;
    if (!same_sign(y0, y1)) c2 = ds * (x0 * dy2 - y0 * dx2) / dy2;
    int di = 0; float cm = c0;
    if (c1 > cm) { di = 1; cm = c1; }
    if (c2 > cm) { di = 2; cm = c2; }
    float dc = -ORC_F32_MAX;
    if (di == 0 && s0 && fabsf(dy0) >= fabsf(dx0)) dc = c0;
    if (di == 1 && s1 && fabsf(dy1) >= fabsf(dx1)) dc = c1;
    if (di == 2 && s2 && fabsf(dy2) >= fabsf(dx2)) dc = c2;
    const float eps = 0.0625f;
    if (dc > -eps && dc < 1.f + eps) {
        dc = fminf(fmaxf(dc, 0.f), 1.f);
        r->alpha = ds * (0.5f - dc);
        r->ok = 1; r->tri = t; r->di = di; r->px = px; r->py = py;
    }
}

ORC_API void orc_antialias_fwd(const float* color, const float* rast, const float* pos, const int* tri, const int* opp,
                               int B, int V, int F, int H, int W, int C, float* out)
{
    memcpy(out, color, (size_t)B * H * W * C * sizeof(float));
#pragma omp parallel for
    for (int b = 0; b < B; b++) {
        const float* rb = rast + (size_t)b * H * W * 4;
        const float* pb = pos + (size_t)b * V * 4;
        const float* cb = color + (size_t)b * H * W * C;
        float* ob = out + (size_t)b * H * W * C;
        for (int py = 0; py < H; py++)
            for (int px = 0; px < W; px++)
                for (int d = 0; d < 2; d++) {
                    if (d == 0 ? px + 1 >= W : py + 1 >= H) continue;
                    AAPair r;
                    aa_analyze(rb, pb, tri, opp, F, V, H, W, px, py, d, &r);
                    if (!r.ok) continue;
                    size_t p0 = (size_t)py * W + px, p1 = p0 + (d ? W : 1);
                    float* o = ob + (r.alpha > 0.f ? p0 : p1) * C;
                    for (int c = 0; c < C; c++) o[c] += r.alpha * (cb[p1 * C + c] - cb[p0 * C + c]);
                }
    }
}

ORC_API void orc_antialias_bwd(const float* color, const float* rast, const float* pos, const int* tri, const int* opp,
                               const float* d_out, int B, int V, int F, int H, int W, int C, float* d_color, float* d_pos)
{
    memcpy(d_color, d_out, (size_t)B * H * W * C * sizeof(float));
#pragma omp parallel for
    for (int b = 0; b < B; b++) {
        const float* rb = rast + (size_t)b * H * W * 4;
        const float* pb = pos + (size_t)b * V * 4;
        const float* cb = color + (size_t)b * H * W * C;
        const float* gob = d_out + (size_t)b * H * W * C;
        float* gcb = d_color + (size_t)b * H * W * C;
        float* gpb = d_pos ? d_pos + (size_t)b * V * 4 : 0;
        for (int py = 0; py < H; py++)
            for (int px = 0; px < W; px++)
                for (int d = 0; d < 2; d++) {
                    if (d == 0 ? px + 1 >= W : py + 1 >= H) continue;
                    AAPair r;
                    aa_analyze(rb, pb, tri, opp, F, V, H, W, px, py, d, &r);
                    if (!r.ok) continue;
                    size_t p0 = (size_t)py * W + px, p1 = p0 + (d ? W : 1);
                    const float* g = gob + (r.alpha > 0.f ? p0 : p1) * C;
                    float dd = 0.f;
                    for (int c = 0; c < C; c++) {
                        float gy = g[c];
                        dd += gy * (cb[p1 * C + c] - cb[p0 * C + c]);
                        gcb[p0 * C + c] -= r.alpha * gy;
                        gcb[p1 * C + c] += r.alpha * gy;
                    }
                    if (!gpb || dd == 0.f || fabsf(r.alpha) >= 0.5f) continue;
                    /* the edge's two vertices */
                    int e1 = tri[r.tri * 3 + (r.di + 1) % 3], e2 = tri[r.tri * 3 + (r.di + 2) % 3];
                    float q1[4], q2[4];
                    memcpy(q1, pb + (size_t)e1 * 4, 16); memcpy(q2, pb + (size_t)e2 * 4, 16);
                    float pxh = 0.5f * (float)W, pyh = 0.5f * (float)H;
                    float fx = (float)r.px + 0.5f - pxh, fy = (float)r.py + 0.5f - pyh;
                    if (d) { float t_; t_ = q1[0]; q1[0] = q1[1]; q1[1] = t_; t_ = q2[0]; q2[0] = q2[1]; q2[1] = t_;
                             t_ = pxh; pxh = pyh; pyh = t_; t_ = fx; fx = fy; fy = t_; }
                    float w1 = 1.f / q1[3], w2 = 1.f / q2[3];
                    float x1 = q1[0] * w1 * pxh - fx, y1 = q1[1] * w1 * pyh - fy;
                    float x2 = q2[0] * w2 * pxh - fx, y2 = q2[1] * w2 * pyh - fy;
                    float dx = x2 - x1, dy = y2 - y1;
                    float db = x1 * dy - y1 * dx;
                    float ep = copysignf(1e-3f, dy);
                    float iy = 1.f / (dy + ep);
                    float dby = db * iy;
                    float iw1 = -w1 * iy * dd, iw2 = w2 * iy * dd;
                    float gp1x = iw1 * pxh * y2, gp2x = iw2 * pxh * y1;
                    float gp1y = iw1 * pyh * (dby - x2), gp2y = iw2 * pyh * (dby - x1);
                    float gp1w = -(q1[0] * gp1x + q1[1] * gp1y) * w1;
                    float gp2w = -(q2[0] * gp2x + q2[1] * gp2y) * w2;
                    if (d) { float t_; t_ = gp1x; gp1x = gp1y; gp1y = t_; t_ = gp2x; gp2x = gp2y; gp2y = t_; }
                    float* g1 = gpb + (size_t)e1 * 4; float* g2 = gpb + (size_t)e2 * 4;
                    g1[0] += gp1x; g1[1] += gp1y; g1[3] += gp1w;
                    g2[0] += gp2x; g2[1] += gp2y; g2[3] += gp2w;
                }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* vertex normals (render/mesh.py:276-304)                                                          */
/* ------------------------------------------------------------------------------------------------ */
ORC_API void orc_vertex_normals_fwd(const float* v_pos, const int* tri, int B, int V, int F, float* nsum, float* v_nrm)
{
#pragma omp parallel for
    for (int b = 0; b < B; b++) {
        const float* p = v_pos + (size_t)b * V * 3;
        float* s = nsum + (size_t)b * V * 3;
        float* n = v_nrm + (size_t)b * V * 3;
        memset(s, 0, (size_t)V * 3 * sizeof(float));
        for (int f = 0; f < F; f++) {
            int i0 = tri[f * 3], i1 = tri[f * 3 + 1], i2 = tri[f * 3 + 2];
            float e1[3], e2[3];
            for (int c = 0; c < 3; c++) { e1[c] = p[i1 * 3 + c] - p[i0 * 3 + c]; e2[c] = p[i2 * 3 + c] - p[i0 * 3 + c]; }
            float fn[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
            for (int c = 0; c < 3; c++) { s[i0 * 3 + c] += fn[c]; s[i1 * 3 + c] += fn[c]; s[i2 * 3 + c] += fn[c]; }
        }
        for (int v = 0; v < V; v++) {
            float x = s[v * 3], y = s[v * 3 + 1], z = s[v * 3 + 2];
            float d = (x * x + y * y) + z * z;
            if (!(d > 1e-20f)) { x = 0.f; y = 0.f; z = 1.f; d = 1.f; }
            float inv = 1.f / sqrtf(fmaxf(d, 1e-20f));
            n[v * 3] = x * inv; n[v * 3 + 1] = y * inv; n[v * 3 + 2] = z * inv;
        }
    }
}

ORC_API void orc_vertex_normals_bwd(const float* v_pos, const int* tri, const float* nsum, const float* d_nrm,
                                    int B, int V, int F, float* d_pos)
{
#pragma omp parallel for
    for (int b = 0; b < B; b++) {
        const float* p = v_pos + (size_t)b * V * 3;
        const float* s = nsum + (size_t)b * V * 3;
        const float* g = d_nrm + (size_t)b * V * 3;
        float* gp = d_pos + (size_t)b * V * 3;
        float* gs = (float*)calloc((size_t)V * 3, sizeof(float));
        for (int v = 0; v < V; v++) {
            float x = s[v * 3], y = s[v * 3 + 1], z = s[v * 3 + 2];
            float d = (x * x + y * y) + z * z;
            if (!(d > 1e-20f)) continue; /* constant fallback normal: no gradient */
            float inv = 1.f / sqrtf(d);
            float nx = x * inv, ny = y * inv, nz = z * inv;
            float dot = nx * g[v * 3] + ny * g[v * 3 + 1] + nz * g[v * 3 + 2];
            gs[v * 3] = (g[v * 3] - nx * dot) * inv;
            gs[v * 3 + 1] = (g[v * 3 + 1] - ny * dot) * inv;
            gs[v * 3 + 2] = (g[v * 3 + 2] - nz * dot) * inv;
        }
        for (int f = 0; f < F; f++) {
            int i0 = tri[f * 3], i1 = tri[f * 3 + 1], i2 = tri[f * 3 + 2];
            float gf[3], e1[3], e2[3];
            for (int c = 0; c < 3; c++) {
                gf[c] = gs[i0 * 3 + c] + gs[i1 * 3 + c] + gs[i2 * 3 + c];
                e1[c] = p[i1 * 3 + c] - p[i0 * 3 + c]; e2[c] = p[i2 * 3 + c] - p[i0 * 3 + c];
            }
            /* fn = e1 x e2 ; d e1 = e2 x gf ; d e2 = gf x e1 */
            float ge1[3] = {e2[1] * gf[2] - e2[2] * gf[1], e2[2] * gf[0] - e2[0] * gf[2], e2[0] * gf[1] - e2[1] * gf[0]};
            float ge2[3] = {gf[1] * e1[2] - gf[2] * e1[1], gf[2] * e1[0] - gf[0] * e1[2], gf[0] * e1[1] - gf[1] * e1[0]};
            for (int c = 0; c < 3; c++) {
                gp[i1 * 3 + c] += ge1[c]; gp[i2 * 3 + c] += ge2[c]; gp[i0 * 3 + c] -= ge1[c] + ge2[c];
            }
        }
        free(gs);
    }
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
