// articulation.cu - the articulation-angle constraints as one kernel per direction.
//
// Reference: InstancePredictorBase.apply_articulation_constraints (model/predictors/InstancePredictorBase.py:435-511) and Fauna's split
// form (InstancePredictorFauna.py:149-212): `angles *= output_multiplier`, optional root-bone mask, tanh, then a config-dependent
// sequence of per-(bone, axis) scalings written as mask algebra (`m * (a * f) + (1 - m) * a`, `m * a`), and `* max_arti_angle / 180 * pi`
// - ~40 small torch kernels per step in front of skinning (R5).  For masks in {0, 1} every one of those statements is, per element,
// ONE fp32 multiplication (or division) by a constant that depends on (bone, axis) only, so the whole method is
//     out = post_S(...post_1(tanh(pre_P(...pre_1(x)))))        with pre_i / post_j = "times (or divided by) table[stage][bone][axis]".
// The stage tables are built on the host from the config (3danimals_b200/predictors.py); applying them in the reference's order, each
// individually rounded (this file is compiled with -fmad=false), reproduces the reference's fp32 arithmetic.
#include "common.cuh"

namespace {

constexpr int MAX_STAGES = 24;

struct StageArgs {
    const float* pre;        // [n_pre, K, 3]
    const float* post;       // [n_post, K, 3]
    int n_pre, n_post, K;
    unsigned div_mask;       // bit j: post stage j divides instead of multiplying
};

__global__ void __launch_bounds__(256) arti_fwd_kernel(const float* __restrict__ x, StageArgs s, int64_t n, float* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int kc = (int)(i % (3 * s.K));
    float v = x[i];
    for (int j = 0; j < s.n_pre; j++) v = v * __ldg(s.pre + (size_t)j * 3 * s.K + kc);
    v = tanhf(v);
    for (int j = 0; j < s.n_post; j++) {
        const float f = __ldg(s.post + (size_t)j * 3 * s.K + kc);
        v = ((s.div_mask >> j) & 1u) ? v / f : v * f;
    }
    out[i] = v;
}

// autograd's order: the post stages backwards, tanh' on the forward's tanh value, the pre stages backwards.  torch's tanh_backward
// evaluates 1 - t*t with ONE rounding (a fused multiply-add, on CPU and GPU builds alike: the golden gradients pin it), hence the explicit fma
__global__ void __launch_bounds__(256) arti_bwd_kernel(const float* __restrict__ x, StageArgs s, int64_t n, const float* __restrict__ d_out,
                                                       float* __restrict__ d_x)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int kc = (int)(i % (3 * s.K));
    float v = x[i];
    for (int j = 0; j < s.n_pre; j++) v = v * __ldg(s.pre + (size_t)j * 3 * s.K + kc);
    const float t = tanhf(v);
    float g = d_out[i];
    for (int j = s.n_post - 1; j >= 0; j--) {
        const float f = __ldg(s.post + (size_t)j * 3 * s.K + kc);
        g = ((s.div_mask >> j) & 1u) ? g / f : g * f;
    }
    g = g * __fmaf_rn(-t, t, 1.f);
    for (int j = s.n_pre - 1; j >= 0; j--) g = g * __ldg(s.pre + (size_t)j * 3 * s.K + kc);
    d_x[i] = g;
}

int check_stages(const char* who, const float* pre, int n_pre, const float* post, int n_post, int K, unsigned div_mask)
{
    if (n_pre < 0 || n_post < 0 || n_pre > MAX_STAGES || n_post > MAX_STAGES || K <= 0 || (n_pre && !pre) || (n_post && !post) ||
        (n_post < 32 && (div_mask >> n_post) != 0u)) {
        b2a_set_error("%s: invalid argument: stage tables", who);
        return 2;
    }
    return 0;
}

}  // namespace

// x, out: [rows, K, 3] (rows = batch * frames); pre [n_pre, K, 3], post [n_post, K, 3] device tables; bit j of post_div_mask makes post
// stage j a division.  out may alias x (the reference scales its argument in place; nothing downstream reads it).
B2A_API int b2a_articulation_constraints_fwd(const float* x, const float* pre, int n_pre, const float* post, int n_post, int post_div_mask,
                                             int64_t rows, int K, float* out, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(x && out && rows >= 0, "null pointer / rows");
    if (int rc = check_stages(__func__, pre, n_pre, post, n_post, K, (unsigned)post_div_mask)) return rc;
    const int64_t n = rows * K * 3;
    if (n == 0) return 0;
    StageArgs s{pre, post, n_pre, n_post, K, (unsigned)post_div_mask};
    arti_fwd_kernel<<<b2a_blocks(n, 256), 256, 0, stream>>>(x, s, n, out);
    B2A_LAUNCH_OK();
    return 0;
}

// d_x [rows, K, 3] is WRITTEN (x is the forward's input).
B2A_API int b2a_articulation_constraints_bwd(const float* x, const float* pre, int n_pre, const float* post, int n_post, int post_div_mask,
                                             int64_t rows, int K, const float* d_out, float* d_x, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(x && d_out && d_x && rows >= 0, "null pointer / rows");
    if (int rc = check_stages(__func__, pre, n_pre, post, n_post, K, (unsigned)post_div_mask)) return rc;
    const int64_t n = rows * K * 3;
    if (n == 0) return 0;
    StageArgs s{pre, post, n_pre, n_post, K, (unsigned)post_div_mask};
    arti_bwd_kernel<<<b2a_blocks(n, 256), 256, 0, stream>>>(x, s, n, d_out, d_x);
    B2A_LAUNCH_OK();
    return 0;
}
