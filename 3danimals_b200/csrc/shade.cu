// shade.cu - directional-light diffuse shading of the g-buffer on sm_100a (FAST arithmetic file, contract 1e-4).
// Replaces the elementwise torch sequence of DirectionalLight.shade (reference model/render/light.py:186-193):
//     shading = ambient + diffuse * clamp(dot(light_dir, normal), min=0)       [B,H,W,1]
//     shaded  = shading * kd                                                   [B,H,W,3]
// with light_params [Bl,5] = (dir.xyz, ambient, diffuse) per image (Bl = B) or shared (Bl = 1).  The light-direction MLP
// that produces light_params stays PyTorch.  Five elementwise kernels forward and ~ten backward become one streaming
// pass per direction: 40 B/pixel forward (24 read + 16 written), 64 B/pixel backward - HBM/L2-bound, 4 pixels per
// thread through 16-byte accesses.  The clamp's sub-gradient follows torch.clamp (passes where dot >= 0).
#include "common.cuh"

namespace {

struct L5 { float x, y, z, amb, diff; };

__device__ __forceinline__ L5 load_light(const float* __restrict__ light, int Bl, int b)
{
    const float* l = light + (size_t)(Bl == 1 ? 0 : b) * 5;
    return L5{__ldg(l), __ldg(l + 1), __ldg(l + 2), __ldg(l + 3), __ldg(l + 4)};
}

// 4 pixels = 12 floats = 3 float4 (pixel-major xyz xyz xyz xyz)
__device__ __forceinline__ void unpack12(const float4 a, const float4 b, const float4 c, float (&v)[12])
{
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
}
__device__ __forceinline__ void store12(float* p, const float (&v)[12])
{
    float4* q = reinterpret_cast<float4*>(p);
    q[0] = make_float4(v[0], v[1], v[2], v[3]); q[1] = make_float4(v[4], v[5], v[6], v[7]); q[2] = make_float4(v[8], v[9], v[10], v[11]);
}

template <bool KDV>
__device__ __forceinline__ void load_kd4(const float* __restrict__ kd, int64_t kd_stride, size_t p, float (&k)[12])
{
    if (KDV) {
        const float4* k4 = reinterpret_cast<const float4*>(kd + p * 3);
        unpack12(__ldg(k4), __ldg(k4 + 1), __ldg(k4 + 2), k);
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float* kp = kd + (p + j) * kd_stride;
            k[3 * j] = __ldg(kp); k[3 * j + 1] = __ldg(kp + 1); k[3 * j + 2] = __ldg(kp + 2);
        }
    }
}

// grid (blocks over HW/4 quads, B).  VEC: HW % 4 == 0 and 16-byte aligned bases; else one pixel per thread.
// KDV: kd rows are dense (stride 3) and 16-byte aligned; otherwise kd is a channel slice of a wider NHWC tensor (row stride
// kd_stride floats, e.g. the first 3 of the texture field's 9 channels) read with scalar loads - no gather copy.
template <bool VEC, bool KDV>
__global__ void __launch_bounds__(256) shade_fwd_kernel(const float* __restrict__ kd, int64_t kd_stride, const float* __restrict__ nrm,
                                                        const float* __restrict__ light, int Bl, int64_t HW, float* __restrict__ shaded,
                                                        float* __restrict__ shading)
{
    const int b = blockIdx.y;
    const L5 L = load_light(light, Bl, b);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (VEC) {
        if (i * 4 >= HW) return;
        const size_t p = (size_t)b * HW + (size_t)i * 4;
        const float4* n4 = reinterpret_cast<const float4*>(nrm + p * 3);
        float k[12], n[12], o[12], s[4];
        load_kd4<KDV>(kd, kd_stride, p, k);
        unpack12(__ldg(n4), __ldg(n4 + 1), __ldg(n4 + 2), n);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float c = L.x * n[3 * j] + L.y * n[3 * j + 1] + L.z * n[3 * j + 2];
            s[j] = L.amb + L.diff * fmaxf(c, 0.f);
            o[3 * j] = s[j] * k[3 * j]; o[3 * j + 1] = s[j] * k[3 * j + 1]; o[3 * j + 2] = s[j] * k[3 * j + 2];
        }
        store12(shaded + p * 3, o);
        if (shading) reinterpret_cast<float4*>(shading + p)[0] = make_float4(s[0], s[1], s[2], s[3]);
    } else {
        if (i >= HW) return;
        const size_t p = (size_t)b * HW + (size_t)i;
        const float c = L.x * __ldg(nrm + p * 3) + L.y * __ldg(nrm + p * 3 + 1) + L.z * __ldg(nrm + p * 3 + 2);
        const float s = L.amb + L.diff * fmaxf(c, 0.f);
        const float* kp = kd + p * kd_stride;
        shaded[p * 3] = s * __ldg(kp); shaded[p * 3 + 1] = s * __ldg(kp + 1); shaded[p * 3 + 2] = s * __ldg(kp + 2);
        if (shading) shading[p] = s;
    }
}

__device__ __forceinline__ void shade_bwd_pixel(const L5& L, const float* k, const float* n, const float* g, float gs, float* dk, float* dn, float (&acc)[5])
{
    const float c = L.x * n[0] + L.y * n[1] + L.z * n[2];
    const float cl = fmaxf(c, 0.f);
    const float s = L.amb + L.diff * cl;
    const float ds = g[0] * k[0] + g[1] * k[1] + g[2] * k[2] + gs;     // d shading
    dk[0] = s * g[0]; dk[1] = s * g[1]; dk[2] = s * g[2];
    const float dc = c >= 0.f ? ds * L.diff : 0.f;                      // torch.clamp(min=0) passes the gradient where c >= 0
    dn[0] = dc * L.x; dn[1] = dc * L.y; dn[2] = dc * L.z;
    acc[0] += dc * n[0]; acc[1] += dc * n[1]; acc[2] += dc * n[2];
    acc[3] += ds; acc[4] += ds * cl;
}

template <bool VEC, bool KDV>
__global__ void __launch_bounds__(256) shade_bwd_kernel(const float* __restrict__ kd, int64_t kd_stride, const float* __restrict__ nrm,
                                                        const float* __restrict__ light, int Bl, int64_t HW, const float* __restrict__ d_shaded,
                                                        const float* __restrict__ d_shading, float* __restrict__ d_kd, float* __restrict__ d_nrm,
                                                        float* __restrict__ d_light)
{
    __shared__ float s_red[8][5];
    const int b = blockIdx.y;
    const L5 L = load_light(light, Bl, b);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (VEC) {
        if (i * 4 < HW) {
            const size_t p = (size_t)b * HW + (size_t)i * 4;
            const float4* n4 = reinterpret_cast<const float4*>(nrm + p * 3);
            const float4* g4 = reinterpret_cast<const float4*>(d_shaded + p * 3);
            float k[12], n[12], g[12], dk[12], dn[12];
            load_kd4<KDV>(kd, kd_stride, p, k);
            unpack12(__ldg(n4), __ldg(n4 + 1), __ldg(n4 + 2), n);
            unpack12(__ldg(g4), __ldg(g4 + 1), __ldg(g4 + 2), g);
            float4 gs4 = d_shading ? __ldg(reinterpret_cast<const float4*>(d_shading + p)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float gs[4] = {gs4.x, gs4.y, gs4.z, gs4.w};
#pragma unroll
            for (int j = 0; j < 4; j++) shade_bwd_pixel(L, k + 3 * j, n + 3 * j, g + 3 * j, gs[j], dk + 3 * j, dn + 3 * j, acc);
            if (d_kd) store12(d_kd + p * 3, dk);
            if (d_nrm) store12(d_nrm + p * 3, dn);
        }
    } else if (i < HW) {
        const size_t p = (size_t)b * HW + (size_t)i;
        float k[3], n[3], g[3], dk[3], dn[3];
#pragma unroll
        for (int c = 0; c < 3; c++) { k[c] = __ldg(kd + p * kd_stride + c); n[c] = __ldg(nrm + p * 3 + c); g[c] = __ldg(d_shaded + p * 3 + c); }
        shade_bwd_pixel(L, k, n, g, d_shading ? __ldg(d_shading + p) : 0.f, dk, dn, acc);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (d_kd) d_kd[p * 3 + c] = dk[c];
            if (d_nrm) d_nrm[p * 3 + c] = dn[c];
        }
    }
    if (d_light) {   // per-image light gradient: warp -> block -> one atomic per block and component
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int c = 0; c < 5; c++) {
            const float v = warp_sum(acc[c]);
            if (lane == 0) s_red[w][c] = v;
        }
        __syncthreads();
        if (threadIdx.x < 5) {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < 8; k++) v += s_red[k][threadIdx.x];
            if (v != 0.f) atomicAdd(d_light + (size_t)(Bl == 1 ? 0 : b) * 5 + threadIdx.x, v);
        }
    }
}

bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

B2A_API int b2a_shade_directional_fwd(const float* kd, int64_t kd_stride, const float* normal, const float* light, int Bl, int B, int64_t HW,
                                      float* shaded, float* shading, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(kd && normal && light && shaded, "null pointer");
    B2A_CHECK_ARG(B > 0 && B <= 65535 && HW > 0 && HW < (1ll << 31) && (Bl == 1 || Bl == B) && kd_stride >= 3, "shape");
    const bool vec = HW % 4 == 0 && al16(normal) && al16(shaded) && al16(shading);
    const bool kdv = kd_stride == 3 && al16(kd);
    const dim3 gv(b2a_blocks(HW / 4, 256), B);
    if (vec && kdv) shade_fwd_kernel<true, true><<<gv, 256, 0, stream>>>(kd, kd_stride, normal, light, Bl, HW, shaded, shading);
    else if (vec) shade_fwd_kernel<true, false><<<gv, 256, 0, stream>>>(kd, kd_stride, normal, light, Bl, HW, shaded, shading);
    else shade_fwd_kernel<false, false><<<dim3(b2a_blocks(HW, 256), B), 256, 0, stream>>>(kd, kd_stride, normal, light, Bl, HW, shaded, shading);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_shade_directional_bwd(const float* kd, int64_t kd_stride, const float* normal, const float* light, int Bl,
                                      const float* d_shaded, const float* d_shading, int B, int64_t HW, float* d_kd, float* d_normal,
                                      float* d_light, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(kd && normal && light && d_shaded, "null pointer");
    B2A_CHECK_ARG(B > 0 && B <= 65535 && HW > 0 && HW < (1ll << 31) && (Bl == 1 || Bl == B) && kd_stride >= 3, "shape");
    const bool vec = HW % 4 == 0 && al16(normal) && al16(d_shaded) && al16(d_shading) && al16(d_kd) && al16(d_normal);
    const bool kdv = kd_stride == 3 && al16(kd);
    const dim3 gv(b2a_blocks(HW / 4, 256), B);
    if (vec && kdv)
        shade_bwd_kernel<true, true><<<gv, 256, 0, stream>>>(kd, kd_stride, normal, light, Bl, HW, d_shaded, d_shading, d_kd, d_normal, d_light);
    else if (vec)
        shade_bwd_kernel<true, false><<<gv, 256, 0, stream>>>(kd, kd_stride, normal, light, Bl, HW, d_shaded, d_shading, d_kd, d_normal, d_light);
    else
        shade_bwd_kernel<false, false><<<dim3(b2a_blocks(HW, 256), B), 256, 0, stream>>>(kd, kd_stride, normal, light, Bl, HW, d_shaded, d_shading, d_kd,
                                                                                          d_normal, d_light);
    B2A_LAUNCH_OK();
    return 0;
}
