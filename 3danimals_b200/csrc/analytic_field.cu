// analytic_field.cu - the benchmark's stand-in pixel shader (NOT part of the reference path).
// SURVEY.md §8d defines M1a as the hot path with the texture / DINO field MLPs "replaced by a fixed analytic colour
// function so both sides do identical work".  That function is y = act(x W) of the canonical position x [N,3]:
//     squash = 1 (texture): out [N,3C] = three copies of sigmoid(x W)      (kd | ks | normal, C = 3)
//     squash = 0 (dino)   : out [N,C]  = sin(x W)                          (C = 16)
// (the same function as oracle/pipeline_ref.py analytic_shader).  Written as one kernel per direction so the stand-in
// costs ~2 launches instead of a K=3 cuBLAS SIMT sgemm (~90 us) plus a dozen elementwise/autograd kernels per step:
// M1a measures the hot-path kernels, and whatever the stand-in costs is noise on that number.
// Compiled in the EXACT regime (accurate sinf / expf).
#include "common.cuh"

namespace {
constexpr int AF_MAXC = 16;

template <bool SQUASH>
__global__ void __launch_bounds__(256) af_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, int C, int64_t N, float* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float x0 = __ldg(x + i * 3), x1 = __ldg(x + i * 3 + 1), x2 = __ldg(x + i * 3 + 2);
    if (SQUASH) {
        float* o = out + i * 3 * C;
        for (int c = 0; c < C; c++) {
            const float z = x0 * __ldg(w + c) + x1 * __ldg(w + C + c) + x2 * __ldg(w + 2 * C + c);
            const float y = 1.f / (1.f + expf(-z));
            o[c] = y; o[C + c] = y; o[2 * C + c] = y;
        }
    } else {
        float* o = out + i * C;
        if ((C & 3) == 0) {
            for (int c = 0; c < C; c += 4) {
                float y[4];
#pragma unroll
                for (int j = 0; j < 4; j++) y[j] = sinf(x0 * __ldg(w + c + j) + x1 * __ldg(w + C + c + j) + x2 * __ldg(w + 2 * C + c + j));
                reinterpret_cast<float4*>(o + c)[0] = make_float4(y[0], y[1], y[2], y[3]);
            }
        } else {
            for (int c = 0; c < C; c++) o[c] = sinf(x0 * __ldg(w + c) + x1 * __ldg(w + C + c) + x2 * __ldg(w + 2 * C + c));
        }
    }
}

template <bool SQUASH>
__global__ void __launch_bounds__(256) af_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, int C, int64_t N,
                                                     const float* __restrict__ g, float* __restrict__ dx)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float x0 = __ldg(x + i * 3), x1 = __ldg(x + i * 3 + 1), x2 = __ldg(x + i * 3 + 2);
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
    const float* gp = g + i * (SQUASH ? 3 * C : C);
    const bool vec = !SQUASH && (C & 3) == 0;
    for (int c0 = 0; c0 < C; c0 += 4) {
        float gv[4] = {0.f, 0.f, 0.f, 0.f};
        if (vec) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(gp + c0));
            gv[0] = t.x; gv[1] = t.y; gv[2] = t.z; gv[3] = t.w;
        } else {
            for (int j = 0; j < 4 && c0 + j < C; j++)
                gv[j] = SQUASH ? (__ldg(gp + c0 + j) + __ldg(gp + C + c0 + j)) + __ldg(gp + 2 * C + c0 + j) : __ldg(gp + c0 + j);
        }
        for (int j = 0; j < 4 && c0 + j < C; j++) {
            const int c = c0 + j;
            const float w0 = __ldg(w + c), w1 = __ldg(w + C + c), w2 = __ldg(w + 2 * C + c);
            const float z = x0 * w0 + x1 * w1 + x2 * w2;
            float gz;
            if (SQUASH) {
                const float y = 1.f / (1.f + expf(-z));
                gz = gv[j] * (y * (1.f - y));
            } else {
                gz = gv[j] * cosf(z);
            }
            d0 += gz * w0; d1 += gz * w1; d2 += gz * w2;
        }
    }
    dx[i * 3] = d0; dx[i * 3 + 1] = d1; dx[i * 3 + 2] = d2;
}
}  // namespace

B2A_API int b2a_analytic_field_fwd(const float* x, const float* weight, int C, int squash, int64_t N, float* out, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(x && weight && out, "null pointer");
    B2A_CHECK_ARG(C > 0 && C <= AF_MAXC && N >= 0 && N < (1ll << 40), "shape");
    B2A_CHECK_ARG(squash || (C & 3) != 0 || ((uintptr_t)out & 15) == 0, "out must be 16-byte aligned");
    if (N == 0) return 0;
    if (squash) af_fwd_kernel<true><<<b2a_blocks(N, 256), 256, 0, stream>>>(x, weight, C, N, out);
    else af_fwd_kernel<false><<<b2a_blocks(N, 256), 256, 0, stream>>>(x, weight, C, N, out);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_analytic_field_bwd(const float* x, const float* weight, int C, int squash, int64_t N, const float* d_out, float* d_x,
                                   b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(x && weight && d_out && d_x, "null pointer");
    B2A_CHECK_ARG(C > 0 && C <= AF_MAXC && N >= 0 && N < (1ll << 40), "shape");
    B2A_CHECK_ARG(squash || (C & 3) != 0 || ((uintptr_t)d_out & 15) == 0, "d_out must be 16-byte aligned");
    if (N == 0) return 0;
    if (squash) af_bwd_kernel<true><<<b2a_blocks(N, 256), 256, 0, stream>>>(x, weight, C, N, d_out, d_x);
    else af_bwd_kernel<false><<<b2a_blocks(N, 256), 256, 0, stream>>>(x, weight, C, N, d_out, d_x);
    B2A_LAUNCH_OK();
    return 0;
}
