// normals.cu - smooth vertex normals on sm_100a.
// Replaces mesh.auto_normals (reference model/render/mesh.py:276-304): un-normalised face normals
// cross(v1-v0, v2-v0) splatted to the three corners (area weighting), fallback (0,0,1) when |n|^2 <= 1e-20,
// then safe_normalize (render/util.py:28-32).  One launch covers the whole batch (grid.y = image).
#include "common.cuh"

namespace {

// nsum is the library's own buffer (kept for the backward): rows are padded to 16 bytes so that each corner takes ONE
// vector reduction instead of three scalar ones - the splat is bound by the SMs' RED issue rate.
__global__ void normals_splat_kernel(const float* __restrict__ v_pos, const int* __restrict__ tri, int64_t V, int64_t F,
                                     float* __restrict__ nsum)
{
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int b = blockIdx.y;
    const float* p = v_pos + (size_t)b * V * 3;
    float4* s = reinterpret_cast<float4*>(nsum) + (size_t)b * V;
    int i0 = __ldg(tri + f * 3), i1 = __ldg(tri + f * 3 + 1), i2 = __ldg(tri + f * 3 + 2);
    float ax = p[(size_t)i0 * 3], ay = p[(size_t)i0 * 3 + 1], az = p[(size_t)i0 * 3 + 2];
    float e1x = p[(size_t)i1 * 3] - ax, e1y = p[(size_t)i1 * 3 + 1] - ay, e1z = p[(size_t)i1 * 3 + 2] - az;
    float e2x = p[(size_t)i2 * 3] - ax, e2y = p[(size_t)i2 * 3 + 1] - ay, e2z = p[(size_t)i2 * 3 + 2] - az;
    float nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
    const float4 n4 = make_float4(nx, ny, nz, 0.f);
    atomicAdd(s + i0, n4); atomicAdd(s + i1, n4); atomicAdd(s + i2, n4);
}

__global__ void normals_normalize_kernel(const float* __restrict__ nsum, int64_t n, float* __restrict__ v_nrm)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s4 = reinterpret_cast<const float4*>(nsum)[i];
    float x = s4.x, y = s4.y, z = s4.z;
    float d = (x * x + y * y) + z * z;
    if (!(d > 1e-20f)) { x = 0.f; y = 0.f; z = 1.f; d = 1.f; }
    float inv = 1.f / sqrtf(fmaxf(d, 1e-20f));
    v_nrm[i * 3] = x * inv; v_nrm[i * 3 + 1] = y * inv; v_nrm[i * 3 + 2] = z * inv;
}

// d_nsum = (g - n (n.g)) / |nsum|   (zero where the constant fallback normal was used)
__global__ void normals_normalize_bwd_kernel(const float* __restrict__ nsum, const float* __restrict__ d_nrm, int64_t n,
                                             float* __restrict__ d_nsum)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s4 = reinterpret_cast<const float4*>(nsum)[i];
    float x = s4.x, y = s4.y, z = s4.z;
    float d = (x * x + y * y) + z * z;
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (d > 1e-20f) {
        float inv = 1.f / sqrtf(d);
        float nx = x * inv, ny = y * inv, nz = z * inv;
        float gx = d_nrm[i * 3], gy = d_nrm[i * 3 + 1], gz = d_nrm[i * 3 + 2];
        float dt = nx * gx + ny * gy + nz * gz;
        ox = (gx - nx * dt) * inv; oy = (gy - ny * dt) * inv; oz = (gz - nz * dt) * inv;
    }
    reinterpret_cast<float4*>(d_nsum)[i] = make_float4(ox, oy, oz, 0.f);   // padded like nsum: 16-byte gathers in the splat backward
}

__global__ void normals_splat_bwd_kernel(const float* __restrict__ v_pos, const int* __restrict__ tri, const float* __restrict__ d_nsum,
                                         int64_t V, int64_t F, float* __restrict__ d_v_pos)
{
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int b = blockIdx.y;
    const float* p = v_pos + (size_t)b * V * 3;
    const float4* gs = reinterpret_cast<const float4*>(d_nsum) + (size_t)b * V;
    float* gp = d_v_pos + (size_t)b * V * 3;
    const int j0 = __ldg(tri + f * 3), j1 = __ldg(tri + f * 3 + 1), j2 = __ldg(tri + f * 3 + 2);
    const float4 g0 = __ldg(gs + j0), g1 = __ldg(gs + j1), g2 = __ldg(gs + j2);
    size_t i0 = (size_t)j0 * 3, i1 = (size_t)j1 * 3, i2 = (size_t)j2 * 3;
    float gx = g0.x + g1.x + g2.x, gy = g0.y + g1.y + g2.y, gz = g0.z + g1.z + g2.z;
    if (gx == 0.f && gy == 0.f && gz == 0.f) return;
    float e1x = p[i1] - p[i0], e1y = p[i1 + 1] - p[i0 + 1], e1z = p[i1 + 2] - p[i0 + 2];
    float e2x = p[i2] - p[i0], e2y = p[i2 + 1] - p[i0 + 1], e2z = p[i2 + 2] - p[i0 + 2];
    // n = e1 x e2 :  d e1 = e2 x g,  d e2 = g x e1
    float a1x = e2y * gz - e2z * gy, a1y = e2z * gx - e2x * gz, a1z = e2x * gy - e2y * gx;
    float a2x = gy * e1z - gz * e1y, a2y = gz * e1x - gx * e1z, a2z = gx * e1y - gy * e1x;
    atomicAdd(gp + i1, a1x); atomicAdd(gp + i1 + 1, a1y); atomicAdd(gp + i1 + 2, a1z);
    atomicAdd(gp + i2, a2x); atomicAdd(gp + i2 + 1, a2y); atomicAdd(gp + i2 + 2, a2z);
    atomicAdd(gp + i0, -(a1x + a2x)); atomicAdd(gp + i0 + 1, -(a1y + a2y)); atomicAdd(gp + i0 + 2, -(a1z + a2z));
}

}  // namespace

B2A_API int b2a_vertex_normals_fwd(const float* v_pos, const int32_t* tri, int B, int64_t V, int64_t F, float* nsum, float* v_nrm,
                                   b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(v_pos && tri && nsum && v_nrm, "null pointer");
    B2A_CHECK_ARG(B > 0 && B <= 65535 && V > 0 && F >= 0, "shape");
    B2A_CHECK_ARG(((uintptr_t)nsum & 15) == 0, "nsum must be 16-byte aligned");
    B2A_CUDA_OK(cudaMemsetAsync(nsum, 0, (size_t)B * V * 4 * sizeof(float), stream));
    if (F > 0) normals_splat_kernel<<<dim3(b2a_blocks(F, 256), B), 256, 0, stream>>>(v_pos, tri, V, F, nsum);
    normals_normalize_kernel<<<b2a_blocks((int64_t)B * V, 256), 256, 0, stream>>>(nsum, (int64_t)B * V, v_nrm);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_vertex_normals_bwd(const float* v_pos, const int32_t* tri, const float* nsum, const float* d_v_nrm, int B, int64_t V,
                                   int64_t F, float* scratch, float* d_v_pos, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(v_pos && tri && nsum && d_v_nrm && scratch && d_v_pos, "null pointer");
    B2A_CHECK_ARG(B > 0 && B <= 65535 && V > 0 && F >= 0, "shape");
    B2A_CHECK_ARG(((uintptr_t)nsum & 15) == 0 && ((uintptr_t)scratch & 15) == 0, "nsum / scratch must be 16-byte aligned");
    normals_normalize_bwd_kernel<<<b2a_blocks((int64_t)B * V, 256), 256, 0, stream>>>(nsum, d_v_nrm, (int64_t)B * V, scratch);
    if (F > 0) normals_splat_bwd_kernel<<<dim3(b2a_blocks(F, 256), B), 256, 0, stream>>>(v_pos, tri, scratch, V, F, d_v_pos);
    B2A_LAUNCH_OK();
    return 0;
}
