// gbuffer.cu - fused g-buffer pass on sm_100a: the geometry half of render_layer + shade in ONE kernel per direction.
// Replaces (reference model/render/render.py): 4x dr.interpolate (:182 v_pos, :185-191 per-face normals, :195 v_nrm,
// :209 prior v_pos), ru.prepare_shading_normal(use_python=True) (:72 -> renderutils/bsdf.py:46-51 with the constant
// perturbed normal (0,0,1), renderutils/ops.py:217-218) and the camera-space normal safe_normalize(n . R_w2c^T) (:75).
// The vertex tangent interpolation (:196) is numerically dead and is not computed.
//
// Why fused: the reference materialises gb_pos / gb_geometric_normal / gb_normal / gb_tangent (4 x 12 B/pixel, written
// then re-read by ~20 elementwise kernels).  Only what the PyTorch field MLPs and the light consume has to cross the
// kernel boundary: gb_tex_pos and the camera-space normal (12 B each) - every other buffer is optional (NULL).
// Backward (Phase B of the graded raster backward) reads d_cam_nrm + d_tex_pos + rast and scatters
// vertex gradients (d_v_pos, d_v_nrm, d_prior_pos, d_clip) with fp32 atomics that stay in the 126 MB L2.
#include "common.cuh"

namespace {

struct GbParams {
    const float* rast;
    const int* tri;
    const float* v_pos;
    const float* v_nrm;
    const float* prior;
    const float* w2c;
    const float* campos;
    int spp, Bq, two_sided, B, H, W;
    int64_t V, F;
    const float4* pn;   // packed (P.xyz, N.x | N.y, N.z, -, -) [B,V,2] or NULL: vertex attributes as 16-byte gathers
    const float4* q4;   // packed prior positions [Bq,V] (xyz, -) or NULL
    const float* mtx;   // [B,4,4] clip transform or NULL: when set the backward recomputes the clip-space positions of the pixel's
                        // three vertices from P (12 FMAs each) instead of gathering them from pos_clip (3 scattered 16-byte loads)
};

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 ld3(const float* p) { return V3{__ldg(p), __ldg(p + 1), __ldg(p + 2)}; }
__device__ __forceinline__ void st3(float* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 neg(V3 a) { return V3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ V3 bary(V3 a, V3 b, V3 c, float u, float v, float w)
{
    return V3{(u * a.x + v * b.x) + w * c.x, (u * a.y + v * b.y) + w * c.y, (u * a.z + v * b.z) + w * c.z};
}
// y = x / max(|x|, eps)  (torch.nn.functional.normalize);  returns the clamped length
__device__ __forceinline__ V3 fnormalize(V3 x, float eps, float& len)
{
    len = fmaxf(sqrtf(dot3(x, x)), eps);
    return V3{x.x / len, x.y / len, x.z / len};
}
// adjoint of fnormalize: d_x = (g - y (y.g)) / len when unclamped, g / eps when clamped
__device__ __forceinline__ V3 fnormalize_bwd(V3 x, V3 y, float len, float eps, V3 g)
{
    if (sqrtf(dot3(x, x)) > eps) {
        float d = dot3(y, g);
        return V3{(g.x - y.x * d) / len, (g.y - y.y * d) / len, (g.z - y.z * d) / len};
    }
    return V3{g.x / eps, g.y / eps, g.z / eps};
}
// y = x / sqrt(max(x.x, 1e-20))  (render/util.py:28-32 safe_normalize)
__device__ __forceinline__ V3 snormalize(V3 x, float& len)
{
    len = sqrtf(fmaxf(dot3(x, x), 1e-20f));
    return V3{x.x / len, x.y / len, x.z / len};
}
__device__ __forceinline__ V3 snormalize_bwd(V3 x, V3 y, float len, V3 g)
{
    if (dot3(x, x) > 1e-20f) {
        float d = dot3(y, g);
        return V3{(g.x - y.x * d) / len, (g.y - y.y * d) / len, (g.z - y.z * d) / len};
    }
    return V3{g.x / len, g.y / len, g.z / len};
}

struct GbPixel {  // forward intermediates of one covered pixel
    int i0, i1, i2;
    float u, v, w;
    V3 P0, P1, P2, N0, N1, N2, Q0, Q1, Q2;
    V3 n, fn, gpos, ggeo, gnrm, gtex;
    float nlen;
    V3 s1, sh, vv, view, shf, geof, out, cam, cn;
    float ln, ls, lv, lc, tpre, t;
    bool front;
};

// One 32-byte vertex record (P.xyz, N.xyz, -, -) with ONE 256-bit load (LDG.E.256, sm_100+): the per-pixel gathers are bound
// by L1 sector lookups per instruction, and a 32-byte aligned record is exactly one sector.
struct PN8 { float px, py, pz, nx, ny, nz, s0, s1; };
__device__ __forceinline__ PN8 ldg_pn8(const float4* p)
{
    PN8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.px), "=f"(r.py), "=f"(r.pz), "=f"(r.nx), "=f"(r.ny), "=f"(r.nz), "=f"(r.s0), "=f"(r.s1)
                 : "l"(p));
    return r;
}

// f >= 0: vertex ids come from the triangle table; f < 0: the caller already set g.i0..i2 (covered-pixel list entry)
template <bool PACKED = false>
__device__ __forceinline__ void gb_forward(const GbParams& P, int b, int f, float u, float v, GbPixel& g)
{
    g.u = u; g.v = v; g.w = 1.f - u - v;
    if (f >= 0) { g.i0 = __ldg(P.tri + (size_t)f * 3); g.i1 = __ldg(P.tri + (size_t)f * 3 + 1); g.i2 = __ldg(P.tri + (size_t)f * 3 + 2); }
    const float* vp = P.v_pos + (size_t)b * P.V * 3;
    const float* vn = P.v_nrm + (size_t)b * P.V * 3;
    const float* vq = P.prior + (size_t)(P.Bq == 1 ? 0 : b) * P.V * 3;
    if (PACKED || P.pn) {
        // the gather is bound by L1 wavefronts per instruction, not bytes: 9 x 16-byte loads instead of 27 scalar ones
        const float4* pn = P.pn + (size_t)b * P.V * 2;
        const float4* q4 = P.q4 + (size_t)(P.Bq == 1 ? 0 : b) * P.V;
        const PN8 a0 = ldg_pn8(pn + (size_t)g.i0 * 2), a1 = ldg_pn8(pn + (size_t)g.i1 * 2), a2 = ldg_pn8(pn + (size_t)g.i2 * 2);
        const float4 c0 = __ldg(q4 + g.i0), c1 = __ldg(q4 + g.i1), c2 = __ldg(q4 + g.i2);
        g.P0 = V3{a0.px, a0.py, a0.pz}; g.N0 = V3{a0.nx, a0.ny, a0.nz}; g.Q0 = V3{c0.x, c0.y, c0.z};
        g.P1 = V3{a1.px, a1.py, a1.pz}; g.N1 = V3{a1.nx, a1.ny, a1.nz}; g.Q1 = V3{c1.x, c1.y, c1.z};
        g.P2 = V3{a2.px, a2.py, a2.pz}; g.N2 = V3{a2.nx, a2.ny, a2.nz}; g.Q2 = V3{c2.x, c2.y, c2.z};
    } else {
        g.P0 = ld3(vp + (size_t)g.i0 * 3); g.P1 = ld3(vp + (size_t)g.i1 * 3); g.P2 = ld3(vp + (size_t)g.i2 * 3);
        g.N0 = ld3(vn + (size_t)g.i0 * 3); g.N1 = ld3(vn + (size_t)g.i1 * 3); g.N2 = ld3(vn + (size_t)g.i2 * 3);
        g.Q0 = ld3(vq + (size_t)g.i0 * 3); g.Q1 = ld3(vq + (size_t)g.i1 * 3); g.Q2 = ld3(vq + (size_t)g.i2 * 3);
    }
    g.gpos = bary(g.P0, g.P1, g.P2, g.u, g.v, g.w);
    g.n = cross3(g.P1 - g.P0, g.P2 - g.P0);
    g.fn = snormalize(g.n, g.nlen);
    g.ggeo = bary(g.fn, g.fn, g.fn, g.u, g.v, g.w);
    g.gnrm = bary(g.N0, g.N1, g.N2, g.u, g.v, g.w);
    g.gtex = bary(g.Q0, g.Q1, g.Q2, g.u, g.v, g.w);
    // shading normal (bsdf.py:46-51): normalize, (identity perturbation), normalize again, two-sided flip, bend
    g.s1 = fnormalize(g.gnrm, 1e-12f, g.ln);
    g.sh = fnormalize(g.s1, 1e-12f, g.ls);
    V3 cp = ld3(P.campos + (size_t)b * 3);
    g.vv = cp - g.gpos;
    g.view = fnormalize(g.vv, 1e-12f, g.lv);
    g.front = !P.two_sided || dot3(g.ggeo, g.view) > 0.f;
    g.shf = g.front ? g.sh : neg(g.sh);
    g.geof = g.front ? g.ggeo : neg(g.ggeo);
    g.tpre = dot3(g.view, g.shf) / 0.1f;
    g.t = fminf(fmaxf(g.tpre, 0.f), 1.f);
    V3 diff = g.shf - g.geof;  // torch.lerp: two algebraically equal branches
    g.out = g.t < 0.5f ? g.geof + diff * g.t : g.shf - diff * (1.f - g.t);
    const float* m = P.w2c + (size_t)b * 16;
    g.cam = V3{(g.out.x * __ldg(m + 0) + g.out.y * __ldg(m + 1)) + g.out.z * __ldg(m + 2),
               (g.out.x * __ldg(m + 4) + g.out.y * __ldg(m + 5)) + g.out.z * __ldg(m + 6),
               (g.out.x * __ldg(m + 8) + g.out.y * __ldg(m + 9)) + g.out.z * __ldg(m + 10)};
    g.cn = snormalize(g.cam, g.lc);
}

// vertex attributes -> 16-byte records (see GbParams::pn / q4); one thread per (image, vertex)
__global__ void __launch_bounds__(256) gb_pack_kernel(const float* __restrict__ v_pos, const float* __restrict__ v_nrm,
                                                      const float* __restrict__ prior, int B, int Bq, int64_t V, float4* __restrict__ pn,
                                                      float4* __restrict__ q4)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * V) return;
    V3 p = ld3(v_pos + i * 3), n = ld3(v_nrm + i * 3);
    pn[i * 2] = make_float4(p.x, p.y, p.z, n.x);       // record = (P.xyz, N.xyz, 0, 0): 32 bytes, 32-byte aligned
    pn[i * 2 + 1] = make_float4(n.y, n.z, 0.f, 0.f);
    if (i < (int64_t)Bq * V) {
        V3 q = ld3(prior + i * 3);
        q4[i] = make_float4(q.x, q.y, q.z, 0.f);
    }
}

__global__ void __launch_bounds__(256) gb_fwd_kernel(GbParams P, float* __restrict__ gb_pos, float* __restrict__ gb_geo,
                                                     float* __restrict__ gb_shn, float* __restrict__ gb_cam, float* __restrict__ gb_tex)
{
    int ip = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (ip >= P.H * P.W) return;
    int px = ip % P.W, py = ip / P.W;
    size_t ri = ((size_t)b * P.H * P.spp + (size_t)py * P.spp) * ((size_t)P.W * P.spp) + (size_t)px * P.spp;
    float4 r = ldg4(P.rast + ri * 4);
    size_t po = ((size_t)b * P.H * P.W + ip) * 3;
    int f = (int)r.w - 1;
    V3 z{0.f, 0.f, 0.f};
    if (f < 0 || f >= P.F) {
        if (gb_pos) st3(gb_pos + po, z);
        if (gb_geo) st3(gb_geo + po, z);
        if (gb_shn) st3(gb_shn + po, z);
        if (gb_cam) st3(gb_cam + po, z);
        if (gb_tex) st3(gb_tex + po, z);
        return;
    }
    GbPixel g;
    gb_forward(P, b, f, r.x, r.y, g);
    if (gb_pos) st3(gb_pos + po, g.gpos);
    if (gb_geo) st3(gb_geo + po, g.ggeo);
    if (gb_shn) st3(gb_shn + po, g.out);
    if (gb_cam) st3(gb_cam + po, g.cn);
    if (gb_tex) st3(gb_tex + po, g.gtex);
}

// Vertex-gradient accumulator: one 48-byte row per (image, vertex), three 16-byte vector reductions per touched
// vertex instead of twelve scalar ones (the backward is bound by the SMs' RED issue rate, not by bytes):
//   A0 = (d_v_pos.xyz, d_clip.w)   A1 = (d_v_nrm.xyz, d_clip.x)   A2 = (d_prior.xyz, d_clip.y)
__device__ __forceinline__ void red4(float* p, float x, float y, float z, float w)
{
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(x, y, z, w));   // red.global.add.v4.f32 (sm_90+)
}

// LIST: threads walk the compact covered-pixel list written by the rasterizer (dense warps; DMTet renders cover ~20 %
// of the image); otherwise one thread per pixel of the [B,H,W] grid (spp > 1 or no list).
// launch bounds measured at C1 (B2A_GB_MINB sweep, round 1): (128,4) 123 regs, no spills 38.3 us | (128,5) 96 regs 38.3-39.1 us |
// (128,6) 80 regs 44.8 us | (128,8) 64 regs 48.1 us - occupancy bought with spills loses
// MTX: the fused render-geometry path - packed vertex records are present (P.pn) and clip positions are recomputed from P.mtx
template <bool LIST, bool CAMGRAD, bool MTX>
__global__ void __launch_bounds__(128, 5) gb_bwd_kernel(GbParams P, const float* __restrict__ pos_clip, const int4* __restrict__ cov_list,
                                                     const int* __restrict__ cov_count, const float* __restrict__ d_gb_pos,
                                                     const float* __restrict__ d_gb_geo, const float* __restrict__ d_gb_shn,
                                                     const float* __restrict__ d_gb_cam, const float* __restrict__ d_gb_tex,
                                                     float* __restrict__ acc, float* __restrict__ d_w2c, float* __restrict__ d_campos,
                                                     float* __restrict__ zero_buf, int64_t zero_n)
{
    const int HW = P.H * P.W;
    const int64_t total = LIST ? (int64_t)*cov_count : (int64_t)P.B * HW;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t span = (int64_t)gridDim.x * blockDim.x;
    // zero-initialise the shared-prior gradient for the finalize pass that follows (instead of a memset in the stream)
    for (int64_t i = first; i < zero_n; i += span) zero_buf[i] = 0.f;
    for (int64_t it = first - (threadIdx.x & 31); it < total; it += span) {   // warp-uniform trip count
        const int64_t idx = it + (threadIdx.x & 31);
        bool active = idx < total;
        int b = 0, ip = 0, px = 0, py = 0, f = -1;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        GbPixel g;
        if (active) {
            int64_t flat = idx;
            if (LIST) {   // 16-byte entry: pixel + vertex ids, so the vertex gathers do not wait for rast -> tri
                const int4 e = __ldg(cov_list + idx);
                flat = e.x; g.i0 = e.y; g.i1 = e.z; g.i2 = e.w;
            }
            b = (int)(flat / HW); ip = (int)(flat - (int64_t)b * HW);
            px = ip % P.W; py = ip / P.W;
            size_t ri = ((size_t)b * P.H * P.spp + (size_t)py * P.spp) * ((size_t)P.W * P.spp) + (size_t)px * P.spp;
            r = ldg4(P.rast + ri * 4);
            f = (int)r.w - 1;
            active = LIST || (f >= 0 && f < P.F);
        }
        V3 zero{0.f, 0.f, 0.f};
        V3 d_cam = zero, d_vv = zero;
        g.out = zero;
        if (active) {
            size_t po = ((size_t)b * HW + ip) * 3;
            V3 g_pos = d_gb_pos ? ld3(d_gb_pos + po) : zero;
            V3 g_geo = d_gb_geo ? ld3(d_gb_geo + po) : zero;
            V3 g_shn = d_gb_shn ? ld3(d_gb_shn + po) : zero;
            V3 g_cn = d_gb_cam ? ld3(d_gb_cam + po) : zero;
            V3 g_tex = d_gb_tex ? ld3(d_gb_tex + po) : zero;
            // clip-space positions: fetched up front with the other vertex attributes (one latency level, not a second)
            float4 p0 = make_float4(0.f, 0.f, 0.f, 1.f), p1 = p0, p2 = p0;
            if (!MTX && LIST && pos_clip) {
                const float* pb = pos_clip + (size_t)b * P.V * 4;
                p0 = ldg4(pb + (size_t)g.i0 * 4); p1 = ldg4(pb + (size_t)g.i1 * 4); p2 = ldg4(pb + (size_t)g.i2 * 4);
            }
            gb_forward<MTX>(P, b, LIST ? -1 : f, r.x, r.y, g);
            // camera-space normal
            d_cam = snormalize_bwd(g.cam, g.cn, g.lc, g_cn);
            const float* m = P.w2c + (size_t)b * 16;
            V3 d_out = V3{g_shn.x + ((d_cam.x * __ldg(m + 0) + d_cam.y * __ldg(m + 4)) + d_cam.z * __ldg(m + 8)),
                          g_shn.y + ((d_cam.x * __ldg(m + 1) + d_cam.y * __ldg(m + 5)) + d_cam.z * __ldg(m + 9)),
                          g_shn.z + ((d_cam.x * __ldg(m + 2) + d_cam.y * __ldg(m + 6)) + d_cam.z * __ldg(m + 10))};
            // bend (lerp) and clamp
            V3 d_geof = d_out * (1.f - g.t), d_shf = d_out * g.t;
            float d_t = dot3(d_out, g.shf - g.geof);
            float d_dot = (g.tpre >= 0.f && g.tpre <= 1.f) ? d_t / 0.1f : 0.f;
            V3 d_view = g.shf * d_dot;
            d_shf = d_shf + g.view * d_dot;
            // two-sided flip
            V3 d_sh = g.front ? d_shf : neg(d_shf);
            V3 d_ggeo = g_geo + (g.front ? d_geof : neg(d_geof));
            // double normalisation of the smooth normal
            V3 d_s1 = fnormalize_bwd(g.s1, g.sh, g.ls, 1e-12f, d_sh);
            V3 d_gnrm = fnormalize_bwd(g.gnrm, g.s1, g.ln, 1e-12f, d_s1);
            // view vector
            d_vv = fnormalize_bwd(g.vv, g.view, g.lv, 1e-12f, d_view);
            V3 d_gpos = g_pos - d_vv;
            // face normal: ggeo = (u+v+w) fn ; fn = safe_normalize(e1 x e2)
            V3 d_fn = d_ggeo * ((g.u + g.v) + g.w);
            V3 d_n = snormalize_bwd(g.n, g.fn, g.nlen, d_fn);
            V3 e1 = g.P1 - g.P0, e2 = g.P2 - g.P0;
            V3 d_e1 = cross3(e2, d_n), d_e2 = cross3(d_n, e1);
            V3 dp0 = d_gpos * g.u - (d_e1 + d_e2), dp1 = d_gpos * g.v + d_e1, dp2 = d_gpos * g.w + d_e2;
            // barycentric gradients -> clip-space positions (rasterize backward, oracle/raster_ref.c orc_rasterize_bwd)
            float c0x = 0.f, c0y = 0.f, c0w = 0.f, c1x = 0.f, c1y = 0.f, c1w = 0.f, c2x = 0.f, c2y = 0.f, c2w = 0.f;
            float du = (dot3(d_gpos, g.P0 - g.P2) + dot3(d_gnrm, g.N0 - g.N2)) + dot3(g_tex, g.Q0 - g.Q2);
            float dv = (dot3(d_gpos, g.P1 - g.P2) + dot3(d_gnrm, g.N1 - g.N2)) + dot3(g_tex, g.Q1 - g.Q2);
            if ((MTX || pos_clip) && (du != 0.f || dv != 0.f)) {
                if (MTX) {          // clip = [P,1] . M^T (xfm_fwd_kernel): x, y, w rows; z is not needed
                    const float* M = P.mtx + (size_t)b * 16;
                    const float m0 = __ldg(M), m1 = __ldg(M + 1), m2 = __ldg(M + 2), m3 = __ldg(M + 3);
                    const float m4 = __ldg(M + 4), m5 = __ldg(M + 5), m6 = __ldg(M + 6), m7 = __ldg(M + 7);
                    const float mc = __ldg(M + 12), md = __ldg(M + 13), me = __ldg(M + 14), mf = __ldg(M + 15);
                    p0.x = ((m0 * g.P0.x + m1 * g.P0.y) + m2 * g.P0.z) + m3; p0.y = ((m4 * g.P0.x + m5 * g.P0.y) + m6 * g.P0.z) + m7;
                    p0.w = ((mc * g.P0.x + md * g.P0.y) + me * g.P0.z) + mf;
                    p1.x = ((m0 * g.P1.x + m1 * g.P1.y) + m2 * g.P1.z) + m3; p1.y = ((m4 * g.P1.x + m5 * g.P1.y) + m6 * g.P1.z) + m7;
                    p1.w = ((mc * g.P1.x + md * g.P1.y) + me * g.P1.z) + mf;
                    p2.x = ((m0 * g.P2.x + m1 * g.P2.y) + m2 * g.P2.z) + m3; p2.y = ((m4 * g.P2.x + m5 * g.P2.y) + m6 * g.P2.z) + m7;
                    p2.w = ((mc * g.P2.x + md * g.P2.y) + me * g.P2.z) + mf;
                } else if (!LIST) {
                    const float* pb = pos_clip + (size_t)b * P.V * 4;
                    p0 = ldg4(pb + (size_t)g.i0 * 4); p1 = ldg4(pb + (size_t)g.i1 * 4); p2 = ldg4(pb + (size_t)g.i2 * 4);
                }
                float fx, fy;
                pixel_ndc(px * P.spp, py * P.spp, P.H * P.spp, P.W * P.spp, fx, fy);
                float q0x = p0.x - fx * p0.w, q0y = p0.y - fy * p0.w;
                float q1x = p1.x - fx * p1.w, q1y = p1.y - fy * p1.w;
                float q2x = p2.x - fx * p2.w, q2y = p2.y - fy * p2.w;
                float a0 = q1x * q2y - q1y * q2x, a1 = q2x * q0y - q2y * q0x, a2 = q0x * q1y - q0y * q1x;
                float iw = 1.f / ((a0 + a1) + a2);
                float uu = a0 * iw, vv = a1 * iw;
                float gs = uu * du + vv * dv;
                float ga0 = (du - gs) * iw, ga1 = (dv - gs) * iw, ga2 = -gs * iw;
                c0x = ga2 * q1y - ga1 * q2y; c0y = ga1 * q2x - ga2 * q1x;
                c1x = ga0 * q2y - ga2 * q0y; c1y = ga2 * q0x - ga0 * q2x;
                c2x = ga1 * q0y - ga0 * q1y; c2y = ga0 * q1x - ga1 * q0x;
                c0w = -(fx * c0x + fy * c0y); c1w = -(fx * c1x + fy * c1y); c2w = -(fx * c2x + fy * c2y);
            }
            float* a = acc + (size_t)b * P.V * 12;
            float* a0p = a + (size_t)g.i0 * 12;
            float* a1p = a + (size_t)g.i1 * 12;
            float* a2p = a + (size_t)g.i2 * 12;
            red4(a0p, dp0.x, dp0.y, dp0.z, c0w); red4(a0p + 4, d_gnrm.x * g.u, d_gnrm.y * g.u, d_gnrm.z * g.u, c0x); red4(a0p + 8, g_tex.x * g.u, g_tex.y * g.u, g_tex.z * g.u, c0y);
            red4(a1p, dp1.x, dp1.y, dp1.z, c1w); red4(a1p + 4, d_gnrm.x * g.v, d_gnrm.y * g.v, d_gnrm.z * g.v, c1x); red4(a1p + 8, g_tex.x * g.v, g_tex.y * g.v, g_tex.z * g.v, c1y);
            red4(a2p, dp2.x, dp2.y, dp2.z, c2w); red4(a2p + 4, d_gnrm.x * g.w, d_gnrm.y * g.w, d_gnrm.z * g.w, c2x); red4(a2p + 8, g_tex.x * g.w, g_tex.y * g.w, g_tex.z * g.w, c2y);
        }
        // camera gradients: d_w2c[i][j] += d_cam[i] out[j] ; d_campos += d_vv  (warp-reduced when the warp is in one image)
        if (CAMGRAD && (d_w2c || d_campos)) {
            const unsigned am = __ballot_sync(0xffffffffu, active);
            if (am) {
                const int b0 = __shfl_sync(0xffffffffu, b, __ffs(am) - 1);
                const bool uniform = __all_sync(0xffffffffu, !active || b == b0);
                float vals[12] = {d_cam.x * g.out.x, d_cam.x * g.out.y, d_cam.x * g.out.z, d_cam.y * g.out.x, d_cam.y * g.out.y, d_cam.y * g.out.z,
                                  d_cam.z * g.out.x, d_cam.z * g.out.y, d_cam.z * g.out.z, d_vv.x, d_vv.y, d_vv.z};
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    float* dst = i < 9 ? (d_w2c ? d_w2c + (i / 3) * 4 + i % 3 : nullptr) : (d_campos ? d_campos + (i - 9) : nullptr);
                    const int stride = i < 9 ? 16 : 3;
                    if (!dst) continue;
                    if (uniform) {
                        float s = warp_sum(active ? vals[i] : 0.f);
                        if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(dst + (size_t)b0 * stride, s);
                    } else if (active && vals[i] != 0.f) {
                        atomicAdd(dst + (size_t)b * stride, vals[i]);
                    }
                }
            }
        }
    }
}

// accumulator rows -> gradient tensors (each nullable), one thread per (image, vertex): three 16-byte loads, the outputs,
// and (rezero) the row handed back zeroed.  With a shared prior (Bq == 1) d_prior sums over the batch with scalar
// reductions into a zero-initialised [V,3] (B-way contention per address, but only 3 V B of them).
__global__ void __launch_bounds__(256) gb_bwd_finalize_kernel(float* __restrict__ acc, int rezero, int B, int Bq, int64_t V, float* __restrict__ d_v_pos,
                                                              float* __restrict__ d_v_nrm, float* __restrict__ d_prior, float* __restrict__ d_clip)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int b = blockIdx.y;
    float4* a = reinterpret_cast<float4*>(acc + ((size_t)b * V + v) * 12);
    const float4 a0 = a[0], a1 = a[1], a2 = a[2];
    const size_t o = ((size_t)b * V + v) * 3;
    if (d_v_pos) { d_v_pos[o] = a0.x; d_v_pos[o + 1] = a0.y; d_v_pos[o + 2] = a0.z; }
    if (d_v_nrm) { d_v_nrm[o] = a1.x; d_v_nrm[o + 1] = a1.y; d_v_nrm[o + 2] = a1.z; }
    if (d_clip) reinterpret_cast<float4*>(d_clip)[(size_t)b * V + v] = make_float4(a1.w, a2.w, 0.f, a0.w);
    if (d_prior) {
        if (Bq == 1) {
            float* q = d_prior + (size_t)v * 3;
            if (a2.x != 0.f) atomicAdd(q, a2.x);
            if (a2.y != 0.f) atomicAdd(q + 1, a2.y);
            if (a2.z != 0.f) atomicAdd(q + 2, a2.z);
        } else {
            d_prior[o] = a2.x; d_prior[o + 1] = a2.y; d_prior[o + 2] = a2.z;
        }
    }
    if (rezero) a[0] = a[1] = a[2] = make_float4(0.f, 0.f, 0.f, 0.f);   // a caller that keeps the accumulator needs no memset next time
}

int gb_check(const float* rast, int spp, const int32_t* tri, const float* v_pos, const float* v_nrm, const float* prior_pos, int Bq,
             const float* w2c, const float* campos, int B, int64_t V, int64_t F, int H, int W)
{
    B2A_CHECK_ARG(rast && tri && v_pos && v_nrm && prior_pos && w2c && campos, "null pointer");
    B2A_CHECK_ARG(B > 0 && B <= 65535 && (Bq == 1 || Bq == B) && V > 0 && F >= 0 && H > 0 && W > 0 && spp >= 1 &&
                      (int64_t)H * W * spp * spp < (1ll << 31), "shape");
    B2A_CHECK_ARG(((uintptr_t)rast & 15) == 0, "rast must be 16-byte aligned");
    return 0;
}

}  // namespace

B2A_API int b2a_gbuffer_pack_bytes(int B, int Bq, int64_t V, size_t* bytes)
{
    B2A_CHECK_ARG(bytes && B > 0 && (Bq == 1 || Bq == B) && V >= 0, "shape");
    *bytes = b2a_align((size_t)B * V * 2 * sizeof(float4)) + b2a_align((size_t)Bq * V * sizeof(float4));
    return 0;
}

namespace {
// packed vertex records live in a caller-owned buffer written by the forward and re-used by the backward
void gb_packed(void* packed, int B, int Bq, int64_t V, GbParams* P)
{
    P->pn = (const float4*)packed;
    P->q4 = packed ? (const float4*)((char*)packed + b2a_align((size_t)B * V * 2 * sizeof(float4))) : nullptr;
}
}  // namespace

B2A_API int b2a_gbuffer_fwd(const float* rast, int spp, const int32_t* tri, const float* v_pos, const float* v_nrm,
                            const float* prior_pos, int Bq, const float* w2c, const float* campos, int two_sided, int B, int64_t V,
                            int64_t F, int H, int W, void* packed, size_t packed_bytes, float* gb_pos, float* gb_geo_nrm,
                            float* gb_shading_nrm, float* gb_cam_nrm, float* gb_tex_pos, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = gb_check(rast, spp, tri, v_pos, v_nrm, prior_pos, Bq, w2c, campos, B, V, F, H, W);
    if (rc) return rc;
    GbParams P{rast, tri, v_pos, v_nrm, prior_pos, w2c, campos, spp, Bq, two_sided, B, H, W, V, F, nullptr, nullptr, nullptr};
    if (packed) {
        size_t need;
        b2a_gbuffer_pack_bytes(B, Bq, V, &need);
        B2A_CHECK_ARG(packed_bytes >= need && ((uintptr_t)packed & 31) == 0, "packed buffer (32-byte aligned)");
        gb_packed(packed, B, Bq, V, &P);
        gb_pack_kernel<<<b2a_blocks((int64_t)B * V, 256), 256, 0, stream>>>(v_pos, v_nrm, prior_pos, B, Bq, V, (float4*)P.pn, (float4*)P.q4);
    }
    gb_fwd_kernel<<<dim3(b2a_blocks((int64_t)H * W, 256), B), 256, 0, stream>>>(P, gb_pos, gb_geo_nrm, gb_shading_nrm, gb_cam_nrm, gb_tex_pos);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_gbuffer_bwd_workspace_bytes(int B, int64_t V, size_t* bytes)
{
    B2A_CHECK_ARG(bytes && B > 0 && V >= 0, "shape");
    *bytes = b2a_align((size_t)B * V * 12 * sizeof(float));
    return 0;
}

namespace {

// Fused tail of the geometry backward (b2a_render_geometry_bwd): accumulator rows -> gradient tensors AND the adjoint of the
// clip transform in the same pass.  One thread per (image, vertex):
//   d_clip = (A1.w, A2.w, 0, A0.w) [g-buffer / rasterize part] + d_clip_up [antialias part, nullable]
//   d_v_pos = A0.xyz + M^T d_clip        d_v_nrm = A1.xyz        d_prior = A2.xyz (summed over the batch when the prior is shared)
//   d_mtx[r][c] += sum_v d_clip[r] * [v,1][c]   (nullable; block-reduced, 16 atomics per block)
// and the row is handed back zeroed.  Replaces gb_bwd_finalize_kernel + xfm_bwd_kernel (one grid-wide pass, no d_clip round trip).
__global__ void __launch_bounds__(256) gb_bwd_finalize_xfm_kernel(float* __restrict__ acc, int rezero, int B, int Bq, int64_t V,
                                                                  const float* __restrict__ mtx, const float* __restrict__ v_pos,
                                                                  const float* __restrict__ d_clip_up, float* __restrict__ d_v_pos,
                                                                  float* __restrict__ d_v_nrm, float* __restrict__ d_prior, float* __restrict__ d_mtx)
{
    __shared__ float m[16];
    __shared__ float macc[16];
    const int b = blockIdx.y;
    if (threadIdx.x < 16) { m[threadIdx.x] = mtx[(size_t)b * 16 + threadIdx.x]; macc[threadIdx.x] = 0.f; }
    __syncthreads();
    // launched with programmatic stream serialization right behind the scatter kernel: everything above overlaps that kernel's
    // tail; the accumulator is only read once the whole preceding grid has completed and flushed
    cudaGridDependencySynchronize();
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    float x = 0.f, y = 0.f, z = 0.f, one = 0.f;
    if (v < V) {
        float4* a = reinterpret_cast<float4*>(acc + ((size_t)b * V + v) * 12);
        const float4 a0 = a[0], a1 = a[1], a2 = a[2];
        g = make_float4(a1.w, a2.w, 0.f, a0.w);
        if (d_clip_up) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(d_clip_up) + (size_t)b * V + v);
            g.x += u.x; g.y += u.y; g.z += u.z; g.w += u.w;
        }
        const size_t o = ((size_t)b * V + v) * 3;
        if (d_v_pos) {
            d_v_pos[o] = a0.x + (((m[0] * g.x + m[4] * g.y) + m[8] * g.z) + m[12] * g.w);
            d_v_pos[o + 1] = a0.y + (((m[1] * g.x + m[5] * g.y) + m[9] * g.z) + m[13] * g.w);
            d_v_pos[o + 2] = a0.z + (((m[2] * g.x + m[6] * g.y) + m[10] * g.z) + m[14] * g.w);
        }
        if (d_v_nrm) { d_v_nrm[o] = a1.x; d_v_nrm[o + 1] = a1.y; d_v_nrm[o + 2] = a1.z; }
        if (d_prior) {
            if (Bq == 1) {
                float* q = d_prior + (size_t)v * 3;
                if (a2.x != 0.f) atomicAdd(q, a2.x);
                if (a2.y != 0.f) atomicAdd(q + 1, a2.y);
                if (a2.z != 0.f) atomicAdd(q + 2, a2.z);
            } else {
                d_prior[o] = a2.x; d_prior[o + 1] = a2.y; d_prior[o + 2] = a2.z;
            }
        }
        if (rezero) a[0] = a[1] = a[2] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d_mtx) { x = v_pos[o]; y = v_pos[o + 1]; z = v_pos[o + 2]; one = 1.f; }
    }
    if (d_mtx) {
        if (__ballot_sync(0xffffffffu, g.x != 0.f || g.y != 0.f || g.z != 0.f || g.w != 0.f)) {
            const float gr[4] = {g.x, g.y, g.z, g.w}, h[4] = {x, y, z, one};
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float s = warp_sum(gr[r] * h[c]);
                    if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(&macc[r * 4 + c], s);
                }
        }
        __syncthreads();
        if (threadIdx.x < 16 && macc[threadIdx.x] != 0.f) atomicAdd(&d_mtx[(size_t)b * 16 + threadIdx.x], macc[threadIdx.x]);
    }
}

// shared body of b2a_gbuffer_bwd (mtx == NULL: clip positions gathered, d_clip written, separate clip-transform adjoint) and
// b2a_render_geometry_bwd (mtx != NULL: clip positions recomputed, clip-transform adjoint fused into the finalize pass)
int gb_bwd_impl(const char* who, const float* rast, int spp, const float* pos_clip, const float* mtx, const int32_t* tri, const float* v_pos,
                const float* v_nrm, const float* prior_pos, int Bq, const float* w2c, const float* campos, int two_sided, int B, int64_t V,
                int64_t F, int H, int W, const void* packed, size_t packed_bytes, const int32_t* cov_list, const int32_t* cov_count,
                const float* d_gb_pos, const float* d_gb_geo_nrm, const float* d_gb_shading_nrm, const float* d_gb_cam_nrm,
                const float* d_gb_tex_pos, const float* d_clip_up, void* workspace, size_t workspace_bytes, int workspace_is_zero,
                float* d_v_pos, float* d_v_nrm, float* d_prior_pos, float* d_clip, float* d_mtx, float* d_w2c, float* d_campos,
                cudaStream_t stream)
{
    int rc = gb_check(rast, spp, tri, v_pos, v_nrm, prior_pos, Bq, w2c, campos, B, V, F, H, W);
    if (rc) return rc;
    B2A_CHECK_ARG(!d_clip || (pos_clip && ((uintptr_t)pos_clip & 15) == 0 && ((uintptr_t)d_clip & 15) == 0), "pos_clip / d_clip");
    B2A_CHECK_ARG(((uintptr_t)d_clip_up & 15) == 0, "d_clip_up must be 16-byte aligned");
    B2A_CHECK_ARG(workspace && ((uintptr_t)workspace & 15) == 0 && workspace_bytes >= (size_t)B * V * 12 * sizeof(float), "workspace");
    B2A_CHECK_ARG((cov_list == nullptr) == (cov_count == nullptr) && (!cov_list || spp == 1), "covered-pixel list");
    float* acc = (float*)workspace;
    if (!workspace_is_zero) B2A_CUDA_OK(cudaMemsetAsync(acc, 0, (size_t)B * V * 12 * sizeof(float), stream));
    GbParams P{rast, tri, v_pos, v_nrm, prior_pos, w2c, campos, spp, Bq, two_sided, B, H, W, V, F, nullptr, nullptr, mtx};
    if (packed) {
        size_t need;
        b2a_gbuffer_pack_bytes(B, Bq, V, &need);
        B2A_CHECK_ARG(packed_bytes >= need && ((uintptr_t)packed & 31) == 0, "packed buffer");
        gb_packed(const_cast<void*>(packed), B, Bq, V, &P);
    }
    const float* pcl = d_clip ? pos_clip : nullptr;     // legacy path: clip positions gathered when d_clip is wanted; with mtx they are recomputed
    const bool cam = d_w2c || d_campos;   // camera gradients cost registers (12 warp reductions): separate instantiation
    float* zbuf = (d_prior_pos && Bq == 1) ? d_prior_pos : nullptr;      // zeroed by the main kernel for finalize's atomics
    const int64_t zn = zbuf ? V * 3 : 0;
    const bool have_gb = d_gb_pos || d_gb_geo_nrm || d_gb_shading_nrm || d_gb_cam_nrm || d_gb_tex_pos;
    if (!have_gb) {
        if (zbuf) B2A_CUDA_OK(cudaMemsetAsync(zbuf, 0, (size_t)zn * sizeof(float), stream));
    } else if (cov_list) {
        // about one covered pixel per thread at typical coverage (~25 %); the grid-stride loop absorbs the rest
        unsigned lblocks = b2a_blocks(((int64_t)B * H * W + 3) / 4, 128);
        if (lblocks > 148u * 32u) lblocks = 148u * 32u;
#define GB_LAUNCH_LIST(CAM, MTX)                                                                                                          \
    gb_bwd_kernel<true, CAM, MTX><<<lblocks, 128, 0, stream>>>(P, pcl, (const int4*)cov_list, cov_count, d_gb_pos, d_gb_geo_nrm, d_gb_shading_nrm, \
                                                               d_gb_cam_nrm, d_gb_tex_pos, acc, d_w2c, d_campos, zbuf, zn)
        const bool fused = mtx && P.pn;
        if (cam) { if (fused) GB_LAUNCH_LIST(true, true); else GB_LAUNCH_LIST(true, false); }
        else { if (fused) GB_LAUNCH_LIST(false, true); else GB_LAUNCH_LIST(false, false); }
#undef GB_LAUNCH_LIST
    } else {
        unsigned blocks = b2a_blocks((int64_t)B * H * W, 128);
#define GB_LAUNCH_GRID(CAM, MTX)                                                                                                       \
    gb_bwd_kernel<false, CAM, MTX><<<blocks, 128, 0, stream>>>(P, pcl, nullptr, nullptr, d_gb_pos, d_gb_geo_nrm, d_gb_shading_nrm, d_gb_cam_nrm, \
                                                               d_gb_tex_pos, acc, d_w2c, d_campos, zbuf, zn)
        const bool fused = mtx && P.pn;
        if (cam) { if (fused) GB_LAUNCH_GRID(true, true); else GB_LAUNCH_GRID(true, false); }
        else { if (fused) GB_LAUNCH_GRID(false, true); else GB_LAUNCH_GRID(false, false); }
#undef GB_LAUNCH_GRID
    }
    if (mtx) {
        // programmatic dependent launch: the per-vertex pass is queued while the scatter kernel drains (its blocks start as SMs
        // free up and wait in cudaGridDependencySynchronize) - removes the launch gap between the two kernels of the call
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(b2a_blocks(V, 256), B);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = have_gb ? 1 : 0;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const int rz = workspace_is_zero, iB = B, iBq = Bq;
        const int64_t iV = V;
        B2A_CUDA_OK(cudaLaunchKernelEx(&cfg, gb_bwd_finalize_xfm_kernel, acc, rz, iB, iBq, iV, mtx, v_pos, d_clip_up, d_v_pos, d_v_nrm, d_prior_pos, d_mtx));
    } else if (d_v_pos || d_v_nrm || d_prior_pos || d_clip || workspace_is_zero) {
        gb_bwd_finalize_kernel<<<dim3(b2a_blocks(V, 256), B), 256, 0, stream>>>(acc, workspace_is_zero, B, Bq, V, d_v_pos, d_v_nrm, d_prior_pos, d_clip);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { b2a_set_error("%s: CUDA error %s", who, cudaGetErrorName(e)); return 1; }
    return 0;
}

}  // namespace

B2A_API int b2a_gbuffer_bwd(const float* rast, int spp, const float* pos_clip, const int32_t* tri, const float* v_pos,
                            const float* v_nrm, const float* prior_pos, int Bq, const float* w2c, const float* campos, int two_sided,
                            int B, int64_t V, int64_t F, int H, int W, const void* packed, size_t packed_bytes, const int32_t* cov_list,
                            const int32_t* cov_count, const float* d_gb_pos, const float* d_gb_geo_nrm, const float* d_gb_shading_nrm, const float* d_gb_cam_nrm,
                            const float* d_gb_tex_pos, void* workspace, size_t workspace_bytes, int workspace_is_zero, float* d_v_pos,
                            float* d_v_nrm, float* d_prior_pos, float* d_clip, float* d_w2c, float* d_campos, b2a_stream_t stream_)
{
    return gb_bwd_impl(__func__, rast, spp, pos_clip, nullptr, tri, v_pos, v_nrm, prior_pos, Bq, w2c, campos, two_sided, B, V, F, H, W, packed, packed_bytes,
                       cov_list, cov_count, d_gb_pos, d_gb_geo_nrm, d_gb_shading_nrm, d_gb_cam_nrm, d_gb_tex_pos, nullptr, workspace, workspace_bytes,
                       workspace_is_zero, d_v_pos, d_v_nrm, d_prior_pos, d_clip, nullptr, d_w2c, d_campos, (cudaStream_t)stream_);
}

// ------------------------------------------------------------------------------------------------------------------------
// Fused geometry half of render_mesh (reference model/render/render.py:270-296 head + :160-209 render_layer + :72-75):
// clip transform -> rasterize -> g-buffer -> antialias analysis, and its adjoint, as ONE C-ABI call per direction.
// ------------------------------------------------------------------------------------------------------------------------
B2A_API int b2a_render_geometry_fwd(const float* v_pos, const float* v_nrm, const float* prior_pos, int Bq, const float* mtx, const float* w2c,
                                    const float* campos, const int32_t* tri, const int32_t* opp, int two_sided, int B, int64_t V, int64_t F,
                                    int H, int W, int spp, void* raster_ws, size_t raster_ws_bytes, void* packed, size_t packed_bytes,
                                    float* clip, float* rast, int32_t* cov_list, int32_t* cov_count, float* gb_pos, float* gb_geo_nrm,
                                    float* gb_shading_nrm, float* gb_cam_nrm, float* gb_tex_pos, void* aa_ctx, size_t aa_ctx_bytes,
                                    b2a_stream_t stream)
{
    B2A_CHECK_ARG(v_pos && v_nrm && prior_pos && mtx && w2c && campos && tri && clip && rast, "null pointer");
    B2A_CHECK_ARG(spp >= 1 && H > 0 && W > 0, "shape");
    int rc = b2a_xfm_points_fwd(v_pos, mtx, B, B, V, clip, stream);
    if (rc) return rc;
    rc = b2a_rasterize_fwd(clip, tri, B, V, F, H * spp, W * spp, raster_ws, raster_ws_bytes, rast, cov_list, cov_count, stream);
    if (rc) return rc;
    rc = b2a_gbuffer_fwd(rast, spp, tri, v_pos, v_nrm, prior_pos, Bq, w2c, campos, two_sided, B, V, F, H, W, packed, packed_bytes, gb_pos, gb_geo_nrm,
                         gb_shading_nrm, gb_cam_nrm, gb_tex_pos, stream);
    if (rc) return rc;
    if (aa_ctx) {
        B2A_CHECK_ARG(opp, "antialias analysis needs the edge adjacency table");
        rc = b2a_antialias_prepare(rast, clip, tri, opp, B, V, F, H * spp, W * spp, aa_ctx, aa_ctx_bytes, stream);
    }
    return rc;
}

B2A_API int b2a_render_geometry_bwd(const float* rast, int spp, const float* mtx, const int32_t* tri, const float* v_pos, const float* v_nrm,
                                    const float* prior_pos, int Bq, const float* w2c, const float* campos, int two_sided, int B, int64_t V,
                                    int64_t F, int H, int W, const void* packed, size_t packed_bytes, const int32_t* cov_list,
                                    const int32_t* cov_count, const float* d_gb_pos, const float* d_gb_geo_nrm, const float* d_gb_shading_nrm,
                                    const float* d_gb_cam_nrm, const float* d_gb_tex_pos, const float* d_clip_up, void* workspace,
                                    size_t workspace_bytes, int workspace_is_zero, float* d_v_pos, float* d_v_nrm, float* d_prior_pos,
                                    float* d_mtx, float* d_w2c, float* d_campos, b2a_stream_t stream_)
{
    B2A_CHECK_ARG(mtx && packed, "mtx and the packed vertex records of the forward are required");
    return gb_bwd_impl(__func__, rast, spp, nullptr, mtx, tri, v_pos, v_nrm, prior_pos, Bq, w2c, campos, two_sided, B, V, F, H, W, packed, packed_bytes,
                       cov_list, cov_count, d_gb_pos, d_gb_geo_nrm, d_gb_shading_nrm, d_gb_cam_nrm, d_gb_tex_pos, d_clip_up, workspace, workspace_bytes,
                       workspace_is_zero, d_v_pos, d_v_nrm, d_prior_pos, nullptr, d_mtx, d_w2c, d_campos, (cudaStream_t)stream_);
}
