// estimate_bones.cu - bone placement heuristic on sm_100a, no host synchronisation.
// Replaces estimate_bones (reference model/geometry/skinning.py:49-248) for body_bones_mode in {z_minmax, z_minmax_y+},
// resample=False: the MagicPony / Ponymation configuration (bone_y_threshold=None: leg quadrants from the whole-batch x
// quantiles, :156-161) and the 3D-Fauna one (bone_y_threshold = q: quadrants centred on the medians of the vertices below
// the q-quantile of y, with margins from their 5 % / 95 % quantiles in x and z, :163-175 - seven quantiles, six of them
// over a data-dependent subset).
//
// The reference runs ~80 torch ops per call: two full sorts (xs.quantile(0.05/0.95) over the whole batch), boolean-mask
// gathers and a Python loop over (b,f) x 4 legs, each with host syncs.  Here:
//   eb_hist<0,1,2>  : exact order statistics by radix select over sortable float keys, 11+11+10 bits; every pass is one
//                     streaming read of the x coordinates with shared-memory privatised histograms.  The four targets
//                     (floor/ceil ranks of the two quantiles) are selected together.
//   eb_final        : one block per (b,f): quantile lerp (torch.quantile 'linear': rank = q*(n-1) in fp32, fused lerp),
//                     deterministic mean, masked arg-max/arg-min of z (body end points), masked arg-min of y per leg
//                     quadrant (first index wins, like torch.argmin on the masked subset), joints -> bones.
// Algorithmic bytes: 3 x 4 n (histogram passes over x) + 12 n (final) = 24 B per vertex.
#include "common.cuh"

namespace {

constexpr int EB_T = 4;            // order statistics selected together
constexpr int EB_BINS = 2048;
constexpr int EB_HIST_THREADS = 256;

constexpr int EB_SETS = 5;         // histogram sets: 0 = primary selection, 1..4 = the masked selections of the Fauna variant
constexpr size_t EB_SET_WORDS = (size_t)3 * EB_T * EB_BINS;

struct EbWorkspace {
    unsigned* hist;   // [EB_SETS][3][EB_T][EB_BINS]
};

// one selection: two quantiles (floor / ceil rank each -> EB_T = 4 order statistics) of one coordinate
struct EbSel {
    int comp;         // 0 x, 1 y, 2 z
    float q0, q1;
};
// the masked selections of the Fauna variant (skinning.py:167-170), sets 1..4
__device__ __forceinline__ EbSel eb_masked_sel(int g)
{
    return g == 0 ? EbSel{0, 0.05f, 0.95f} : (g == 1 ? EbSel{0, 0.5f, 0.5f} : (g == 2 ? EbSel{2, 0.05f, 0.95f} : EbSel{2, 0.5f, 0.5f}));
}

__device__ __forceinline__ unsigned sortable_key(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ranks of the four order statistics: floor/ceil of q*(n-1) for the two quantiles (fp32 arithmetic, as torch.quantile)
__device__ __forceinline__ void eb_ranks(int64_t n, float q0, float q1, unsigned rank[EB_T], float w[2])
{
    const float last = (float)(n > 0 ? n - 1 : 0);
    const float qs[2] = {q0, q1};
#pragma unroll
    for (int i = 0; i < 2; i++) {
        float r = qs[i] * last;
        float lo = floorf(r);
        rank[2 * i] = (unsigned)lo;
        rank[2 * i + 1] = (unsigned)ceilf(r);
        w[i] = r - lo;
    }
}

// torch lerp (ATen/native/Lerp.h), fused like the CUDA build
__device__ __forceinline__ float torch_lerp(float a, float b, float w)
{
    float d = b - a;
    return fabsf(w) < 0.5f ? fmaf(w, d, a) : fmaf(-d, 1.f - w, b);
}

// One warp: smallest bin whose inclusive cumulative count exceeds k; *krem = k - (count before that bin).  Each lane owns
// nbins/32 consecutive bins (register-resident), one shuffle scan locates the lane, that lane walks its bins.
template <int NB>
__device__ __forceinline__ void eb_select_bin_warp(const unsigned* __restrict__ hist, unsigned k, unsigned* bin, unsigned* krem)
{
    constexpr int PER = NB / 32;
    const int lane = threadIdx.x & 31;
    unsigned h[PER];
    unsigned local = 0;
#pragma unroll
    for (int i = 0; i < PER; i++) { h[i] = hist[lane * PER + i]; local += h[i]; }
    unsigned inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const unsigned before = inc - local;
    unsigned rb = 0, rk = 0;
    const bool mine = before <= k && k < inc;
    if (mine) {
        unsigned acc = before;
#pragma unroll
        for (int i = 0; i < PER; i++) {
            if (k >= acc && k < acc + h[i]) { rb = lane * PER + i; rk = k - acc; }
            acc += h[i];
        }
    }
    const unsigned who = __ballot_sync(0xffffffffu, mine);
    const int src = who ? __ffs(who) - 1 : 0;
    *bin = __shfl_sync(0xffffffffu, rb, src);
    *krem = __shfl_sync(0xffffffffu, rk, src);
}

// prefix (already selected high bits) and remaining rank of every target after `passes` completed passes.  Warp t of
// the block resolves target t (the four selections run concurrently); results are broadcast through shared memory.
// Needs blockDim.x >= 128.  smem: 2 * EB_T unsigned.
__device__ void eb_resolve(const unsigned* __restrict__ hset, int passes, int64_t n, float q0, float q1, int* smem, unsigned prefix[EB_T],
                           unsigned krem[EB_T])
{
    float w[2];
    eb_ranks(n, q0, q1, krem, w);
#pragma unroll
    for (int t = 0; t < EB_T; t++) prefix[t] = 0u;
    const int warp = threadIdx.x >> 5;
    if (warp < EB_T && passes > 0) {
        const int t = warp;
        unsigned pf = 0u, kr = krem[t];
        for (int p = 0; p < passes; p++) {
            unsigned bin, k2;
            // pass 0 has one shared histogram (slot 0); later passes one per target
            const unsigned* h = hset + ((size_t)p * EB_T + (p == 0 ? 0 : t)) * EB_BINS;
            if (p == 2) eb_select_bin_warp<1024>(h, kr, &bin, &k2);
            else eb_select_bin_warp<EB_BINS>(h, kr, &bin, &k2);
            pf = (pf << (p == 2 ? 10 : 11)) | bin;
            kr = k2;
        }
        if ((threadIdx.x & 31) == 0) { smem[t] = (int)pf; smem[EB_T + t] = (int)kr; }
    }
    __syncthreads();
    if (passes > 0) {
#pragma unroll
        for (int t = 0; t < EB_T; t++) { prefix[t] = (unsigned)smem[t]; krem[t] = (unsigned)smem[EB_T + t]; }
    }
    __syncthreads();
}

// number of keys a set's pass-0 histogram holds (the size of a masked subset); blockDim.x >= 32.  smem: 1 int.
__device__ int64_t eb_set_count(const unsigned* __restrict__ hset, int* smem)
{
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned s = 0;
        for (int i = threadIdx.x; i < EB_BINS; i += 32) s += hset[i];
        s = (unsigned)warp_sum_i((int)s);
        if (threadIdx.x == 0) smem[0] = (int)s;
    }
    __syncthreads();
    const int64_t c = (int64_t)(unsigned)smem[0];
    __syncthreads();
    return c;
}

// the two quantile values of a finished selection (three passes done): torch.quantile 'linear'
__device__ void eb_quantiles(const unsigned* __restrict__ hset, int64_t n, float q0, float q1, int* smem, float& v0, float& v1)
{
    unsigned prefix[EB_T], krem[EB_T];
    eb_resolve(hset, 3, n, q0, q1, smem, prefix, krem);
    float w[2];
    unsigned rk[EB_T];
    eb_ranks(n, q0, q1, rk, w);
    v0 = torch_lerp(key_to_float(prefix[0]), key_to_float(prefix[1]), w[0]);
    v1 = torch_lerp(key_to_float(prefix[2]), key_to_float(prefix[3]), w[1]);
}

// Histogram pass PASS of selection `sel` into set `hset`.  MASKED: only vertices with y < y_thr take part, where y_thr is
// the q_y quantile of y over everything (set 0, finished) and the subset size is the total of this set's pass-0 histogram;
// blockIdx.y picks one of the four masked selections.
template <int PASS, bool MASKED>
__global__ void __launch_bounds__(EB_HIST_THREADS) eb_hist_kernel(const float* __restrict__ verts, int64_t n, EbWorkspace ws, EbSel sel0, float q_y)
{
    __shared__ unsigned s_hist[(PASS == 0 ? 1 : EB_T) * EB_BINS];
    __shared__ int s_scan[34];
    const EbSel sel = MASKED ? eb_masked_sel((int)blockIdx.y) : sel0;
    unsigned* hset = ws.hist + (size_t)(MASKED ? 1 + blockIdx.y : 0) * EB_SET_WORDS;
    float y_thr = 0.f;
    int64_t n_sel = n;
    if (MASKED) {
        float dummy;
        eb_quantiles(ws.hist, n, q_y, q_y, s_scan, y_thr, dummy);      // set 0 holds the finished y selection
        if (PASS > 0) n_sel = eb_set_count(hset, s_scan);
    }
    constexpr int NH = PASS == 0 ? 1 : EB_T;
    constexpr int SHIFT = PASS == 0 ? 21 : (PASS == 1 ? 10 : 0);       // position of this pass's digit
    constexpr unsigned MASK = PASS == 2 ? 1023u : 2047u;
    constexpr int PSHIFT = PASS == 1 ? 21 : 10;                         // key >> PSHIFT = bits selected so far
    for (int i = threadIdx.x; i < NH * EB_BINS; i += blockDim.x) s_hist[i] = 0u;
    unsigned prefix[EB_T], krem[EB_T];
    eb_resolve(hset, PASS, n_sel, sel.q0, sel.q1, s_scan, prefix, krem);   // includes __syncthreads
    __syncthreads();
#pragma unroll 4
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (MASKED && !(__ldg(verts + i * 3 + 1) < y_thr)) continue;
        unsigned key = sortable_key(__ldg(verts + i * 3 + sel.comp));
        unsigned digit = (key >> SHIFT) & MASK;
        if (PASS == 0) {
            atomicAdd(&s_hist[digit], 1u);
        } else {
            unsigned hi = key >> PSHIFT;
#pragma unroll
            for (int t = 0; t < EB_T; t++)
                if (hi == prefix[t]) atomicAdd(&s_hist[t * EB_BINS + digit], 1u);
        }
    }
    __syncthreads();
    unsigned* g = hset + (size_t)PASS * EB_T * EB_BINS;
    for (int i = threadIdx.x; i < NH * EB_BINS; i += blockDim.x)
        if (s_hist[i]) atomicAdd(g + i, s_hist[i]);
}

struct ArgVal {
    float v;
    int i;
};
// lexicographic (value, index) minimum: the first index among equal values wins, like torch.argmin / argmax
__device__ __forceinline__ ArgVal arg_min2(ArgVal a, ArgVal b) { return (b.v < a.v || (b.v == a.v && b.i < a.i)) ? b : a; }

__device__ ArgVal block_argmin(ArgVal x, ArgVal* smem /* 32 */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgVal y{__shfl_xor_sync(0xffffffffu, x.v, o), __shfl_xor_sync(0xffffffffu, x.i, o)};
        x = arg_min2(x, y);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) smem[warp] = x;
    __syncthreads();
    if (warp == 0) {
        ArgVal y = lane < nwarp ? smem[lane] : ArgVal{3.4e38f, 0x7fffffff};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ArgVal z{__shfl_xor_sync(0xffffffffu, y.v, o), __shfl_xor_sync(0xffffffffu, y.i, o)};
            y = arg_min2(y, z);
        }
        if (lane == 0) smem[0] = y;
    }
    __syncthreads();
    return smem[0];
}

__device__ double block_sum_d(double x, double* smem /* 32 */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) smem[warp] = x;
    __syncthreads();
    if (warp == 0) {
        double y = lane < nwarp ? smem[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) y += __shfl_xor_sync(0xffffffffu, y, o);
        if (lane == 0) smem[0] = y;
    }
    __syncthreads();
    return smem[0];
}

// torch.linspace(0, 1, steps)[i] in fp32 (ATen RangeFactories: symmetric evaluation from both ends)
__device__ __forceinline__ float linspace01(int i, int steps)
{
    if (steps <= 1) return 0.f;
    float step = 1.f / (float)(steps - 1);
    return i < steps / 2 ? step * (float)i : 1.f - step * (float)(steps - 1 - i);
}

constexpr int EB_MAX_JOINTS = 65;

__global__ void __launch_bounds__(1024) eb_final_kernel(const float* __restrict__ verts, int N, int64_t V, int n_body, int n_leg, int mode,
                                                        int at0, int at1, int at2, int at3, EbWorkspace ws, float q_y, float* __restrict__ bones,
                                                        int* __restrict__ attach_out, float* __restrict__ stats_out)
{
    __shared__ int s_scan[34];
    __shared__ ArgVal s_arg[32];
    __shared__ double s_dbl[32];
    __shared__ float s_joints[EB_MAX_JOINTS * 3];
    __shared__ float s_pts[7 * 3];   // point_a, point_b, mid, 4 feet
    const int inst = blockIdx.x;
    const float* vp = verts + (size_t)inst * V * 3;
    const int64_t n = (int64_t)N * V;

    // leg quadrants: x_margin from the whole-batch quantiles (skinning.py:157), or (Fauna, :163-175) centre (x0, z0) and
    // margins (x_margin, z_margin) from the quantiles of the vertices below the y threshold
    float x_margin = 0.f, z_margin = 0.f, x0 = 0.f, z0 = 0.f;
    const bool fauna = q_y > 0.f;
    if (n_leg > 0 && !fauna) {
        float q05, q95;
        eb_quantiles(ws.hist, n, 0.05f, 0.95f, s_scan, q05, q95);
        x_margin = (q95 - q05) * 0.2f;
    } else if (n_leg > 0) {
        float lo, hi, med, dummy;
        const unsigned* h1 = ws.hist + 1 * EB_SET_WORDS;
        const int64_t m = eb_set_count(h1, s_scan);
        eb_quantiles(h1, m, 0.05f, 0.95f, s_scan, lo, hi);
        x_margin = (hi - lo) * 0.2f;
        eb_quantiles(ws.hist + 2 * EB_SET_WORDS, m, 0.5f, 0.5f, s_scan, med, dummy);
        x0 = med;
        eb_quantiles(ws.hist + 3 * EB_SET_WORDS, m, 0.05f, 0.95f, s_scan, lo, hi);
        z_margin = (hi - lo) * 0.2f;
        eb_quantiles(ws.hist + 4 * EB_SET_WORDS, m, 0.5f, 0.5f, s_scan, med, dummy);
        z0 = med;
    }

    // mean (deterministic: fixed strided partial sums in double, fixed tree)
    double sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll 4
    for (int64_t v = threadIdx.x; v < V; v += blockDim.x) {
        sx += (double)__ldg(vp + v * 3); sy += (double)__ldg(vp + v * 3 + 1); sz += (double)__ldg(vp + v * 3 + 2);
    }
    sx = block_sum_d(sx, s_dbl); sy = block_sum_d(sy, s_dbl); sz = block_sum_d(sz, s_dbl);
    const float mx = (float)(sx / (double)V), my = (float)(sy / (double)V), mz = (float)(sz / (double)V);
    (void)mx;

    // body end points: arg-max / arg-min z (among vertices with y > mean_y - 0.5 in mode 1; skinning.py:70-87) and the
    // four feet: lowest y per quadrant (skinning.py:155-161, :183-184)
    const float ythr = my - 0.5f;
    ArgVal amax{3.4e38f, 0x7fffffff}, amin{3.4e38f, 0x7fffffff};
    ArgVal foot[4] = {{3.4e38f, 0x7fffffff}, {3.4e38f, 0x7fffffff}, {3.4e38f, 0x7fffffff}, {3.4e38f, 0x7fffffff}};
#pragma unroll 4
    for (int64_t v = threadIdx.x; v < V; v += blockDim.x) {
        float x = __ldg(vp + v * 3), y = __ldg(vp + v * 3 + 1), z = __ldg(vp + v * 3 + 2);
        bool up = mode == 0 || y > ythr;
        float zmax = up ? z : -1e6f, zmin = up ? z : 1e6f;
        amax = arg_min2(amax, ArgVal{-zmax, (int)v});
        amin = arg_min2(amin, ArgVal{zmin, (int)v});
        if (n_leg > 0) {
            const float xr = x - x0, zr = z - z0;       // x0 = z0 = 0 and z_margin = 0 outside the Fauna variant
            bool q[4] = {xr > x_margin && zr > z_margin, xr > x_margin && z < z0, xr < -x_margin && z < z0, xr < -x_margin && zr > z_margin};
#pragma unroll
            for (int k = 0; k < 4; k++) foot[k] = arg_min2(foot[k], ArgVal{q[k] ? y : __int_as_float(0x7f800000), (int)v});
        }
    }
    amax = block_argmin(amax, s_arg);
    amin = block_argmin(amin, s_arg);
    if (n_leg > 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) foot[k] = block_argmin(foot[k], s_arg);
    }
    if (threadIdx.x == 0) {
        int ia = amax.i, ib = amin.i;
        s_pts[0] = 0.f; s_pts[1] = vp[(size_t)ia * 3 + 1]; s_pts[2] = vp[(size_t)ia * 3 + 2];
        s_pts[3] = 0.f; s_pts[4] = vp[(size_t)ib * 3 + 1]; s_pts[5] = vp[(size_t)ib * 3 + 2];
        s_pts[6] = 0.f; s_pts[7] = n_leg > 0 ? my + 0.5f : my; s_pts[8] = mz;
        for (int k = 0; k < 4; k++)
            for (int c = 0; c < 3; c++) s_pts[9 + k * 3 + c] = n_leg > 0 ? vp[(size_t)foot[k].i * 3 + c] : 0.f;
        if (stats_out) {
            float* so = stats_out + (size_t)inst * 8;
            so[0] = x_margin; so[1] = mx; so[2] = my; so[3] = mz;
            so[4] = __int_as_float(ia); so[5] = __int_as_float(ib);
            so[6] = x0; so[7] = z0;
        }
    }
    __syncthreads();

    // joints along point_a -> mid -> point_b (skinning.py:101-108)
    const int J2 = n_body / 2 + 1, n_joints = n_body + 1;
    for (int j = threadIdx.x; j < n_joints * 3; j += blockDim.x) {
        int jj = j / 3, c = j % 3;
        float val;
        if (jj < J2 - 1) {
            float bl = linspace01(jj, J2);
            val = s_pts[c] * (1.f - bl) + s_pts[6 + c] * bl;
        } else {
            float bl = linspace01(jj - (J2 - 1), J2);
            val = s_pts[3 + c] * bl + s_pts[6 + c] * (1.f - bl);
        }
        s_joints[j] = val;
    }
    __syncthreads();
    const int K = n_body + 4 * n_leg;
    float* ob = bones + (size_t)inst * K * 6;
    const int half = n_body / 2;
    // body bones (skinning.py:118-131): k < half: (joint k+1, joint k); then i = n_body-1 .. half: (joint i, joint i+1)
    for (int j = threadIdx.x; j < n_body * 6; j += blockDim.x) {
        int k = j / 6, e = (j / 3) % 2, c = j % 3;
        int ja, jb;
        if (k < half) { ja = k + 1; jb = k; } else { int i = n_body - 1 - (k - half); ja = i; jb = i + 1; }
        ob[j] = s_joints[(e == 0 ? ja : jb) * 3 + c];
    }
    if (n_leg > 0) {
        // attachment joint: given, or (auto) the body bone whose end joint is closest in z to the foot (skinning.py:190-192)
        __shared__ int s_attach[4];
        if (threadIdx.x < 4) {
            int at = threadIdx.x == 0 ? at0 : (threadIdx.x == 1 ? at1 : (threadIdx.x == 2 ? at2 : at3));
            if (at < 0) {
                float best = 3.4e38f;
                float fz = s_pts[9 + threadIdx.x * 3 + 2];
                for (int k = 0; k < n_body; k++) {
                    int jb = k < half ? k : (n_body - 1 - (k - half)) + 1;
                    float dist = fabsf(s_joints[jb * 3 + 2] - fz);
                    if (dist < best) { best = dist; at = k; }
                }
            }
            s_attach[threadIdx.x] = at;
            if (attach_out && inst == 0) attach_out[threadIdx.x] = at;
        }
        __syncthreads();
        // leg bones (skinning.py:195-198, build_kinematic_chain :25-35): bone i = (joint i+1, joint i), joint j = lerp(foot, body joint)
        for (int j = threadIdx.x; j < 4 * n_leg * 6; j += blockDim.x) {
            int leg = j / (n_leg * 6), r = j % (n_leg * 6);
            int i = r / 6, e = (r / 3) % 2, c = r % 3;
            int at = s_attach[leg];
            int jb = at < half ? at : (n_body - 1 - (at - half)) + 1;   // end joint of the body bone
            float bj = s_joints[jb * 3 + c], ft = s_pts[9 + leg * 3 + c];
            float bl = linspace01(e == 0 ? i + 1 : i, n_leg + 1);
            ob[n_body * 6 + j] = ft * (1.f - bl) + bj * bl;
        }
    }
}

size_t eb_layout(void* base, EbWorkspace* ws)
{
    if (ws) ws->hist = (unsigned*)base;
    return b2a_align((size_t)EB_SETS * EB_SET_WORDS * sizeof(unsigned));
}

}  // namespace

B2A_API int b2a_estimate_bones_workspace_bytes(size_t* bytes)
{
    B2A_CHECK_ARG(bytes, "null pointer");
    *bytes = eb_layout(nullptr, nullptr);
    return 0;
}

B2A_API int b2a_estimate_bones(const float* verts, int N, int64_t V, int n_body_bones, int n_leg_bones, int mode, float bone_y_threshold,
                               int attach0, int attach1, int attach2, int attach3, void* workspace, size_t workspace_bytes, float* bones,
                               int32_t* attach_out, float* stats_out, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(verts && bones && workspace, "null pointer");
    B2A_CHECK_ARG(N > 0 && V > 0 && (int64_t)N * V < (1ll << 31), "shape");
    B2A_CHECK_ARG(n_body_bones >= 2 && n_body_bones % 2 == 0 && n_body_bones + 1 <= EB_MAX_JOINTS && n_leg_bones >= 0 && n_leg_bones <= 16,
                  "bone counts");
    B2A_CHECK_ARG(mode == 0 || mode == 1, "mode");
    B2A_CHECK_ARG(bone_y_threshold >= 0.f && bone_y_threshold <= 1.f, "bone_y_threshold must be a quantile in [0,1] (0 = off)");
    B2A_CHECK_ARG(attach0 < n_body_bones && attach1 < n_body_bones && attach2 < n_body_bones && attach3 < n_body_bones, "attach index");
    EbWorkspace ws;
    B2A_CHECK_ARG(eb_layout(workspace, &ws) <= workspace_bytes, "workspace too small");
    const int64_t n = (int64_t)N * V;
    const float q_y = n_leg_bones > 0 ? bone_y_threshold : 0.f;
    if (n_leg_bones > 0) {
        const bool fauna = q_y > 0.f;
        B2A_CUDA_OK(cudaMemsetAsync(ws.hist, 0, (size_t)(fauna ? EB_SETS : 1) * EB_SET_WORDS * sizeof(unsigned), stream));
        unsigned blocks = b2a_blocks(n, EB_HIST_THREADS * 2);
        if (blocks > 148u * 4u) blocks = 148u * 4u;
        // primary selection: x at 5 % / 95 % (MagicPony), or y at the threshold quantile (Fauna)
        const EbSel s0 = fauna ? EbSel{1, q_y, q_y} : EbSel{0, 0.05f, 0.95f};
        eb_hist_kernel<0, false><<<blocks, EB_HIST_THREADS, 0, stream>>>(verts, n, ws, s0, q_y);
        eb_hist_kernel<1, false><<<blocks, EB_HIST_THREADS, 0, stream>>>(verts, n, ws, s0, q_y);
        eb_hist_kernel<2, false><<<blocks, EB_HIST_THREADS, 0, stream>>>(verts, n, ws, s0, q_y);
        if (fauna) {   // four masked selections side by side (grid.y): x 5/95 %, x median, z 5/95 %, z median of {y < y_thr}
            unsigned mb = blocks > 148u ? 148u : blocks;
            eb_hist_kernel<0, true><<<dim3(mb, 4), EB_HIST_THREADS, 0, stream>>>(verts, n, ws, s0, q_y);
            eb_hist_kernel<1, true><<<dim3(mb, 4), EB_HIST_THREADS, 0, stream>>>(verts, n, ws, s0, q_y);
            eb_hist_kernel<2, true><<<dim3(mb, 4), EB_HIST_THREADS, 0, stream>>>(verts, n, ws, s0, q_y);
        }
    }
    eb_final_kernel<<<N, 1024, 0, stream>>>(verts, N, V, n_body_bones, n_leg_bones, mode, attach0, attach1, attach2, attach3, ws, q_y, bones,
                                            attach_out, stats_out);
    B2A_LAUNCH_OK();
    return 0;
}
