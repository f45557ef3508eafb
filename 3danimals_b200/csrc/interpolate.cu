// interpolate.cu - barycentric attribute interpolation on sm_100a.
// Replaces nvdiffrast.torch.interpolate (reference call sites model/render/render.py:24, :182-209):
//   out = u*A[i0] + v*A[i1] + (1-u-v)*A[i2], zeros on empty pixels; attr batch 1 broadcasts over images.
// Generic-C path used by the nvdiffrast-compatible shim; the training fast path is gbuffer.cu.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) interp_fwd_kernel(const float* __restrict__ attr, const float* __restrict__ rast, const int* __restrict__ tri,
                                                         int Ba, int64_t V, int64_t F, int HW, int C, float* __restrict__ out)
{
    int ip = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (ip >= HW) return;
    size_t pi = (size_t)b * HW + ip;
    float4 r = ldg4(rast + pi * 4);
    float* o = out + pi * C;
    int f = (int)r.w - 1;
    if (f < 0 || f >= F) {
        for (int c = 0; c < C; c++) o[c] = 0.f;
        return;
    }
    const float* ab = attr + (size_t)(Ba == 1 ? 0 : b) * V * C;
    const float* A0 = ab + (size_t)__ldg(tri + (size_t)f * 3) * C;
    const float* A1 = ab + (size_t)__ldg(tri + (size_t)f * 3 + 1) * C;
    const float* A2 = ab + (size_t)__ldg(tri + (size_t)f * 3 + 2) * C;
    float u = r.x, v = r.y, w = 1.f - u - v;
    for (int c = 0; c < C; c++) o[c] = (u * __ldg(A0 + c) + v * __ldg(A1 + c)) + w * __ldg(A2 + c);
}

__global__ void __launch_bounds__(256) interp_bwd_kernel(const float* __restrict__ attr, const float* __restrict__ rast, const int* __restrict__ tri,
                                                         const float* __restrict__ d_out, int Ba, int64_t V, int64_t F, int HW, int C,
                                                         float* __restrict__ d_attr, float* __restrict__ d_rast)
{
    int ip = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (ip >= HW) return;
    size_t pi = (size_t)b * HW + ip;
    float4 r = ldg4(rast + pi * 4);
    int f = (int)r.w - 1;
    float du = 0.f, dv = 0.f;
    if (f >= 0 && f < F) {
        size_t o0 = (size_t)__ldg(tri + (size_t)f * 3) * C, o1 = (size_t)__ldg(tri + (size_t)f * 3 + 1) * C, o2 = (size_t)__ldg(tri + (size_t)f * 3 + 2) * C;
        const float* ab = attr + (size_t)(Ba == 1 ? 0 : b) * V * C;
        float* gab = d_attr ? d_attr + (size_t)(Ba == 1 ? 0 : b) * V * C : nullptr;
        const float* g = d_out + pi * C;
        float u = r.x, v = r.y, w = 1.f - u - v;
        for (int c = 0; c < C; c++) {
            float gc = g[c];
            if (gc == 0.f) continue;
            if (gab) { atomicAdd(gab + o0 + c, u * gc); atomicAdd(gab + o1 + c, v * gc); atomicAdd(gab + o2 + c, w * gc); }
            float a2 = __ldg(ab + o2 + c);
            du += gc * (__ldg(ab + o0 + c) - a2);
            dv += gc * (__ldg(ab + o1 + c) - a2);
        }
    }
    if (d_rast) reinterpret_cast<float4*>(d_rast)[pi] = make_float4(du, dv, 0.f, 0.f);
}

}  // namespace

B2A_API int b2a_interpolate_fwd(const float* attr, const float* rast, const int32_t* tri, int B, int Ba, int64_t V, int64_t F, int H, int W,
                                int C, float* out, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(attr && rast && tri && out, "null pointer");
    B2A_CHECK_ARG(B > 0 && B <= 65535 && (Ba == 1 || Ba == B) && C > 0 && (int64_t)H * W < (1ll << 31), "shape");
    interp_fwd_kernel<<<dim3(b2a_blocks((int64_t)H * W, 256), B), 256, 0, stream>>>(attr, rast, tri, Ba, V, F, H * W, C, out);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_interpolate_bwd(const float* attr, const float* rast, const int32_t* tri, const float* d_out, int B, int Ba, int64_t V,
                                int64_t F, int H, int W, int C, float* d_attr, float* d_rast, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(attr && rast && tri && d_out, "null pointer");
    B2A_CHECK_ARG(B > 0 && B <= 65535 && (Ba == 1 || Ba == B) && C > 0 && (int64_t)H * W < (1ll << 31), "shape");
    interp_bwd_kernel<<<dim3(b2a_blocks((int64_t)H * W, 256), B), 256, 0, stream>>>(attr, rast, tri, d_out, Ba, V, F, H * W, C, d_attr, d_rast);
    B2A_LAUNCH_OK();
    return 0;
}
