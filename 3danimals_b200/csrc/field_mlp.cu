// field_mlp.cu - the field MLPs of the pixel shader (reference model/networks/MLPs.py:34-101 CoordMLP, evaluated by
// material.sample / dino_net.sample, model/render/render.py:54,61) on the 5th-generation tensor cores of sm_100a.
//
// This is the one dense contraction on the hot path: per covered pixel a harmonic embedding (63 / 51 wide) followed by
// 8 (texture) or 5 (DINO) layers of 256 x 256, ~1.6 MFLOP per pixel forward.  The reference runs it as fp32 SIMT GEMMs; the horse
// configs' contract is fp32 (1e-4), which single-pass bf16 / tf32 operands cannot hold once pre-activations grow
// (profiles/mlp_precision_study_r1.txt).  Here every fp32 operand is split on the fly into two bf16 terms (x = hi + lo) and
// each product is three tcgen05 MMAs with fp32 accumulation in tensor memory:   a.b ~ a_hi.b_hi + a_hi.b_lo + a_lo.b_hi
// (the dropped lo.lo term is 2^-18 relative) - fp32-grade results at the bf16 rate / 3.  `passes = 1` keeps only hi.hi: the
// arithmetic of the bird config's fp16/bf16 autocast (train_magicpony_bird.yaml:52).
//
// Kernels
//   mlp_pack_weights   W [N,K] fp32 (optionally transposed) -> bf16 hi | lo images of the canonical K-major no-swizzle shared-
//                      memory layout, one 32-wide K chunk after the other: the GEMM CTAs fetch them with ONE bulk async copy
//                      (cp.async.bulk, TMA engine) per chunk and mbarrier transaction counts.
//   mlp_rows_gemm      OUT[rows, N] = epilogue(A[rows, K] . W^T): A is fp32 in global memory (activations / gradients), staged
//                      by the CTA's threads (optional ReLU on load, hi/lo split, 16-byte shared stores in the canonical layout),
//                      128 rows per CTA, all N (<= 256) columns in one 128 x N accumulator of tensor memory; the MMAs of chunk i
//                      overlap the staging of chunk i+1 (two stages, tcgen05.commit -> mbarrier); epilogue: tcgen05.ld 32 columns
//                      per thread, + bias (vector or per-image row), ReLU-derivative mask, or sigmoid, 16-byte global stores.
//   mlp_wgrad          dW[M, N] += P^T Q over a range of rows (both operands MN-major: row-major activations / gradients with
//                      the reduction index slow), persistent over its K range, accumulator 128 x N in tensor memory, one
//                      vector reduction per element at the end.
//   mlp_embed fwd/bwd  harmonic embedding (HarmonicEmbedding.py: [x, sin(x f), cos(x f)], x mirrored to |x| first when the field
//                      is symmetric) and its adjoint.
//   mlp_colsum_segments  per-image column sums of a gradient (adjoint of the per-image bias of the first hidden layer).
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int KC = 32;            // K chunk staged per pipeline step (two stages of 48 KB at N = 256: two CTAs per SM)
constexpr int TILE_M = 128;       // rows per CTA = MMA M
constexpr int GEMM_THREADS = 256;                   // staging / epilogue threads (warps 0-7)
constexpr int GEMM_LAUNCH = GEMM_THREADS + 32;      // + the issuer warp (warp 8): bulk copies, proxy fence, tcgen05.mma, commit
constexpr int WG_WARPS = 16;                        // staging warps of the weight-gradient kernel (one CTA per SM); warp 16 issues
constexpr int WG_LAUNCH_THREADS = WG_WARPS * 32 + 32;

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// eight consecutive floats of a 32-byte aligned address in ONE 256-bit load (LDG.E.ENL2.256): a whole sector per lane and instruction.
// With two 128-bit loads every instruction touched 32 half sectors and the LSU data pipe ran at 90 % of its wavefront rate (ncu,
// profiles/ncu_r2_mlp_wgrad_detail.txt) - the pipe, not HBM, bounded the staging.
__device__ __forceinline__ void ldg8(const float* p, float (&v)[8])
{
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}

// bulk async copy global -> shared (TMA engine, no tensor map), completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// shared-memory matrix descriptor, no swizzle (layout type 0), version 1 (sm_100): start address, leading-dimension and
// stride-dimension byte offsets in 16-byte units (cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
           (1ull << 46);
}
// instruction descriptor, kind::f16: D fp32, A/B bf16, M x N, operand majors (0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, int a_mn_major, int b_mn_major)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 columns of fp32 accumulators: thread t of the warp receives row (lane base + t), columns [col, col + 32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v)
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
        "%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// x = hi + lo with both terms bf16 (round to nearest even); packs two consecutive elements.  The PACKED conversion
// (cvt.rn.bf16x2.f32 -> F2FP.BF16.F32.PACK_AB, ALU pipe) - the scalar __float2bfloat16_rn compiles to F2F.BF16.F32 on the
// conversion pipe (16 lanes per clock and SM), which at 2 conversions per staged element was ~30 % of a wgrad chunk.
// The bf16 -> f32 widening is a shift / mask.
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem)
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));      // first source -> upper half
    return r;
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo)
{
    hi = pack_bf16x2(a, b);
    const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
    lo = pack_bf16x2(a - ah, b - bh);
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight packing.  Canonical K-major no-swizzle image of a [Npad rows, KC k] chunk: byte(n, k) = (k/8) * Npad*16 + n*16 + (k%8)*2
// (core matrices of 8 rows x 16 bytes; consecutive row groups 128 bytes apart = SBO, the two K halves of an MMA Npad*16 bytes
// apart = LBO).  packed = for every chunk c: [hi image | lo image], each Npad*64*2 bytes.
// W is [N, K] row-major (transpose = 0: B[n][k] = W[n*ldw + k]) or its transpose (transpose = 1: B[n][k] = W[k*ldw + n]).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mlp_pack_weights_kernel(const float* __restrict__ W, int ldw, int N, int K, int transpose, int Npad, int Kpad,
                                                               uint16_t* __restrict__ packed)
{
    const int64_t total = (int64_t)Npad * Kpad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % Kpad), n = (int)(i / Kpad);
        float w = 0.f;
        if (n < N && k < K) w = transpose ? __ldg(W + (size_t)k * ldw + n) : __ldg(W + (size_t)n * ldw + k);
        const __nv_bfloat16 h = __float2bfloat16_rn(w);
        const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
        const int c = k / KC, kk = k % KC;
        const size_t img = (size_t)Npad * KC;                                 // elements per hi (or lo) image
        const size_t off = (size_t)c * 2 * img + (size_t)(kk / 8) * Npad * 8 + (size_t)n * 8 + (kk % 8);
        packed[off] = __bfloat16_as_ushort(h);
        packed[off + img] = __bfloat16_as_ushort(l);
    }
}

// All the weight images of a network in ONE launch (a step packs 26: every layer, both orientations - each a 3 us kernel plus a
// launch gap on a GPU-bound stream).  Jobs travel in the kernel's parameter space.
constexpr int PACK_MAX_JOBS = 32;
struct PackJob {
    const float* W;
    uint16_t* out;
    int ldw, N, K, transpose, Npad, Kpad;
    long long first;                      // running offset of this job's Npad * Kpad elements
};
struct PackJobs {
    PackJob j[PACK_MAX_JOBS];
    int n;
    long long total;
};
__global__ void __launch_bounds__(256) mlp_pack_many_kernel(const PackJobs J)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < J.total; i += (long long)gridDim.x * blockDim.x) {
        int q = 0;
        while (q + 1 < J.n && i >= J.j[q + 1].first) q++;
        const PackJob& b = J.j[q];
        const long long e = i - b.first;
        const int k = (int)(e % b.Kpad), n = (int)(e / b.Kpad);
        float w = 0.f;
        if (n < b.N && k < b.K) w = b.transpose ? __ldg(b.W + (size_t)k * b.ldw + n) : __ldg(b.W + (size_t)n * b.ldw + k);
        const __nv_bfloat16 h = __float2bfloat16_rn(w);
        const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
        const int c = k / KC, kk = k % KC;
        const size_t img = (size_t)b.Npad * KC;
        const size_t off = (size_t)c * 2 * img + (size_t)(kk / 8) * b.Npad * 8 + (size_t)n * 8 + (kk % 8);
        b.out[off] = __bfloat16_as_ushort(h);
        b.out[off + img] = __bfloat16_as_ushort(l);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// clock64 trace of ONE CTA (profiling builds only: B2A_NVCC_DEFINES=-DB2A_MLP_TRACE python 3danimals_b200/build.py --force; read with
// scripts/dbg_gemm_trace.py / dbg_wgrad_trace.py through b2a_debug_mlp_trace).  Compiles to nothing otherwise.
// ---------------------------------------------------------------------------------------------------------------------
#ifdef B2A_MLP_TRACE
__device__ unsigned long long g_mlp_trace[128];
#define MLP_STAMP(on, i) do { if (on) g_mlp_trace[(i)] = clock64(); } while (0)
#else
#define MLP_STAMP(on, i) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------------------------------
// Rows GEMM
// ---------------------------------------------------------------------------------------------------------------------
enum { EPI_BIAS = 0, EPI_MASK = 1, EPI_SIGMOID = 2 };

struct GemmParams {
    const float* A;            // [rows, lda] fp32
    int64_t lda;
    int64_t rows;
    int K;                     // valid K (<= Kpad)
    int Kpad;                  // multiple of KC
    int N;                     // valid output columns
    int Npad;                  // MMA N: multiple of 16, <= 256
    const uint16_t* packed;    // mlp_pack_weights image
    int relu_on_load;          // A := max(A, 0) while staging (the stored tensor is the pre-activation)
    int passes;                // 3: hi.hi + hi.lo + lo.hi   1: hi.hi only
    float* out;                // [rows, ldo]
    int64_t ldo;
    const float* bias;         // EPI_BIAS / EPI_SIGMOID: [N] (bias_rows == NULL) or [n_img, N] gathered through bias_rows; nullable
    const int* bias_rows;      // [rows] image index of each row (nullable)
    const float* mask_src;     // EPI_MASK: [rows, ldm] pre-activation whose sign gates the output (when mask_bits is NULL)
    int64_t ldm;
    const uint32_t* mask_bits; // EPI_MASK: [rows, Npad/32] sign bits (bit j of word g = pre-activation[row, 32 g + j] > 0): 1/32 of the bytes
    uint32_t* bits_out;        // EPI_BIAS: (nullable) writes those sign bits of the OUTPUT for the backward pass
};

// The K-major A image: k group kg (8 k values = 16 bytes per row) starts at kg * A_LBO.  A_LBO = 128 rows * 16 B + 32: with the dense
// 2048 the four k groups a warp writes in one 128-bit store instruction (lane = (row, kg)) fall on the same banks - a 4-way conflict on
// every staging store; 32 bytes of padding per group spread them (the descriptor's leading-dimension byte offset is free to choose).
constexpr uint32_t A_LBO = TILE_M * 16 + 32;
constexpr uint32_t A_IMG = (KC / 8) * A_LBO;
constexpr int A_ITEMS = TILE_M * (KC / 8) / GEMM_THREADS;   // (row, 8-wide k group) items per thread and chunk

// global -> registers: the fp32 values of this thread's items of chunk c (issued one chunk ahead of their use)
__device__ __forceinline__ void a_load(const GemmParams& P, int64_t row0, int c, int tid, float (&v)[A_ITEMS][8])
{
#pragma unroll
    for (int it = 0; it < A_ITEMS; it++) {
        const int item = it * GEMM_THREADS + tid;
        const int kg = item % (KC / 8), r = item / (KC / 8);
        const int64_t row = row0 + r;
        const int k0 = c * KC + kg * 8;
        if (row < P.rows && k0 + 8 <= P.K) {
            const float* src = P.A + row * P.lda + k0;
            if ((reinterpret_cast<uintptr_t>(src) & 31) == 0) ldg8(src, v[it]);
            else {
                const float4 x = __ldg(reinterpret_cast<const float4*>(src));
                const float4 y = __ldg(reinterpret_cast<const float4*>(src) + 1);
                v[it][0] = x.x; v[it][1] = x.y; v[it][2] = x.z; v[it][3] = x.w; v[it][4] = y.x; v[it][5] = y.y; v[it][6] = y.z; v[it][7] = y.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) v[it][j] = (row < P.rows && k0 + j < P.K) ? __ldg(P.A + row * P.lda + k0 + j) : 0.f;
        }
    }
}
// registers -> hi / lo images of the stage (canonical K-major layout: byte(r, k) = (k/8) * A_LBO + r*16 + (k%8)*2)
__device__ __forceinline__ void a_store(const GemmParams& P, uint8_t* st, uint32_t a_bytes, int tid, float (&v)[A_ITEMS][8])
{
#pragma unroll
    for (int it = 0; it < A_ITEMS; it++) {
        const int item = it * GEMM_THREADS + tid;
        const int kg = item % (KC / 8), r = item / (KC / 8);
        if (P.relu_on_load) {
#pragma unroll
            for (int j = 0; j < 8; j++) v[it][j] = fmaxf(v[it][j], 0.f);
        }
        uint4 hi, lo;
        split2(v[it][0], v[it][1], hi.x, lo.x); split2(v[it][2], v[it][3], hi.y, lo.y);
        split2(v[it][4], v[it][5], hi.z, lo.z); split2(v[it][6], v[it][7], hi.w, lo.w);
        const uint32_t off = (uint32_t)kg * A_LBO + (uint32_t)r * 16;
        *reinterpret_cast<uint4*>(st + off) = hi;
        *reinterpret_cast<uint4*>(st + a_bytes + off) = lo;
    }
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_LAUNCH, 2) mlp_rows_gemm_kernel(GemmParams P)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    // stage s: [A_hi A_IMG | A_lo A_IMG | B_hi Npad*64 B | B_lo Npad*64 B]
    const uint32_t a_bytes = A_IMG, b_bytes = (uint32_t)P.Npad * KC * 2;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    const uint32_t w_bytes = P.passes == 3 ? 2 * b_bytes : b_bytes;          // one product per element: the hi image only
    __shared__ __align__(8) uint64_t full_b[2], mma_done[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool tr = blockIdx.x == 700 && tid == 0, itr = blockIdx.x == 700 && tid == GEMM_THREADS;      // (trace builds) staging thread 0, issuer
    (void)tr; (void)itr;
    MLP_STAMP(tr, 0);
    const int64_t row0 = (int64_t)blockIdx.x * TILE_M;
    const uint32_t tmem_cols = P.Npad <= 32 ? 32u : (P.Npad <= 64 ? 64u : (P.Npad <= 128 ? 128u : 256u));
    const int nchunks = P.Kpad / KC;

    // Warps 0-7 stage the activations; warp 8 is the ISSUER (weights' bulk copies, the generic->async proxy fence, the MMAs).
    // Why a separate warp: fence.proxy.async is a CTA memory barrier, and a memory barrier waits for the executing thread's outstanding
    // global loads - with the fence in the staging threads every chunk drained the prefetch and paid a full HBM latency.  The staging
    // threads order their shared-memory stores with the CTA barrier alone; the issuer executes the proxy fence after that barrier (it has
    // no loads in flight), then issues.  Two chunks of activations stay in flight per staging thread (register double buffer, buffer =
    // chunk parity = stage).
    const bool issuer = warp == GEMM_THREADS / 32;
    float va[2][A_ITEMS][8];
    if (!issuer) {
        a_load(P, row0, 0, tid, va[0]);    // chunks 0 and 1 are on their way while the barriers / tensor memory are set up
        if (nchunks > 1) a_load(P, row0, 1, tid, va[1]);
    }
    if (tid == 0) {
        mbar_init(&full_b[0], 1); mbar_init(&full_b[1], 1);
        mbar_init(&mma_done[0], 1); mbar_init(&mma_done[1], 1);
        fence_barrier_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t idesc = instr_desc(TILE_M, P.Npad, 0, 0);
    MLP_STAMP(tr, 1);

    // weights of chunk c: one bulk copy (hi | lo are adjacent) into stage c & 1, issued ONE CHUNK AHEAD of its use (chunk 0 here,
    // chunk c+1 at the top of iteration c) - issued in the iteration that consumes it, the issuer sat out the L2 -> shared latency of
    // 32 KB in front of every chunk's MMAs
    auto weights = [&](int c) {
        const int s = c & 1;
        mbar_expect_tx(&full_b[s], w_bytes);
        bulk_g2s(smem + (size_t)s * stage_bytes + 2 * a_bytes, P.packed + (size_t)c * P.Npad * KC * 2, w_bytes, &full_b[s]);
    };
    if (issuer && lane == 0) weights(0);
    auto chunk = [&](int c, float (&v)[A_ITEMS][8]) {
        const int s = c & 1;
        uint8_t* st = smem + (size_t)s * stage_bytes;
        if (c >= 2) mbar_wait(&mma_done[s], ((c >> 1) - 1) & 1);        // the MMAs that read this stage (chunk c-2) have completed
        MLP_STAMP(tr && c < 8, 2 + 3 * c);
        if (!issuer) {
            a_store(P, st, a_bytes, tid, v);                             // activations of chunk c: registers -> hi / lo images
            if (c + 2 < nchunks) a_load(P, row0, c + 2, tid, v);         // chunk c+2 into the buffer just drained
        } else if (lane == 0 && c + 1 < nchunks) {
            // the NEXT chunk's weights, requested while the staging warps work on this one: the other stage's last readers were the
            // MMAs of chunk c-1 (the clock64 trace showed the 32 KB copy landing 0.3-0.7 us after the barrier when it was requested
            // only after this chunk's MMAs - the L2 is the busy unit of this kernel)
            if (c >= 1) mbar_wait(&mma_done[s ^ 1], ((c - 1) >> 1) & 1);
            weights(c + 1);
        }
        MLP_STAMP(tr && c < 8, 3 + 3 * c);
        __syncthreads();
        MLP_STAMP(tr && c < 8, 4 + 3 * c);
        if (issuer && lane == 0) {
            fence_proxy_async();        // the staged images (generic-proxy stores, ordered by the barrier) -> visible to the tensor core
            mbar_wait(&full_b[s], (c >> 1) & 1);
            MLP_STAMP(itr && c < 8, 64 + 2 * c);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(st), a_lo = a_hi + a_bytes, b_hi = a_hi + 2 * a_bytes, b_lo = b_hi + b_bytes;
            const uint32_t a_lbo = A_LBO, b_lbo = (uint32_t)P.Npad * 16;
#pragma unroll
            for (int ks = 0; ks < KC / 16; ks++) {
                const uint64_t dah = smem_desc(a_hi + ks * 2 * a_lbo, a_lbo, 128), dal = smem_desc(a_lo + ks * 2 * a_lbo, a_lbo, 128);
                const uint64_t dbh = smem_desc(b_hi + ks * 2 * b_lbo, b_lbo, 128), dbl = smem_desc(b_lo + ks * 2 * b_lbo, b_lbo, 128);
                umma(tmem_d, dah, dbh, idesc, (c | ks) != 0);
                if (P.passes == 3) {
                    umma(tmem_d, dah, dbl, idesc, 1);
                    umma(tmem_d, dal, dbh, idesc, 1);
                }
            }
            umma_commit(&mma_done[s]);      // arrives when every MMA issued so far has completed (implies fence::before_thread_sync)
            MLP_STAMP(itr && c < 8, 65 + 2 * c);
        }
    };
    for (int c = 0; c < nchunks; c += 2) {
        chunk(c, va[0]);
        if (c + 1 < nchunks) chunk(c + 1, va[1]);
    }
    // the ReLU-derivative words of this thread's row (EPI_MASK): requested before the wait for the tensor core, consumed after it
    const int quad = warp & 3;
    const int64_t wrow0 = row0 + quad * 32;
    const int ngroups = (P.Npad + 31) / 32;
    uint32_t mw[4] = {0u, 0u, 0u, 0u};
    if (EPI == EPI_MASK && P.mask_bits && !issuer && wrow0 + lane < P.rows) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int g = (warp >> 2) + 2 * i;
            if (g < ngroups) mw[i] = __ldg(P.mask_bits + (wrow0 + lane) * ngroups + g);
        }
    }
    // all MMAs done: the last commit covers every earlier one
    MLP_STAMP(tr, 40);
    mbar_wait(&mma_done[(nchunks - 1) & 1], ((nchunks - 1) >> 1) & 1);
    __syncwarp();                   // tcgen05.ld is .sync.aligned: the K-tail path of the staging and the spin loop can leave lanes diverged
    tc_fence_after();
    MLP_STAMP(tr, 41);

    // Epilogue.  Warp w reads the 32 accumulator lanes of its quadrant (w % 4): thread t holds 32 consecutive columns of row
    // 32*(w%4) + t.  The 32 x 32 block goes through a shared-memory tile (the pipeline stages are free now; rows 36 words apart:
    // 16-byte aligned and conflict-free for the quarter-warp phases of 128-bit accesses) so that global memory sees whole 128-byte
    // row segments: 8 lanes x float4 cover one row segment, a warp instruction writes four rows.  Warps 0-3 take the even column
    // groups, 4-7 the odd.
    float* tile = reinterpret_cast<float*>(smem) + warp * (32 * 36);
    const bool vec_out = (P.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(P.out) & 15) == 0;
    for (int g = issuer ? ngroups : (warp >> 2), gi = 0; g < ngroups; g += 2, gi++) {
        float v[32];
        __syncwarp();
        tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * 32), v);
        const int c0 = g * 32;
        if (EPI == EPI_BIAS || EPI == EPI_SIGMOID) {
            const int64_t row = wrow0 + lane;
            if (P.bias && row < P.rows) {
                const float* b = P.bias + (P.bias_rows ? (size_t)__ldg(P.bias_rows + row) * P.N : 0);
#pragma unroll
                for (int j = 0; j < 32; j++)
                    if (c0 + j < P.N) v[j] += __ldg(b + c0 + j);
            }
            if (EPI == EPI_SIGMOID) {
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = 1.f / (1.f + expf(-v[j]));
            }
            if (EPI == EPI_BIAS && P.bits_out && row < P.rows) {
                uint32_t w = 0;
#pragma unroll
                for (int j = 0; j < 32; j++) w |= (v[j] > 0.f ? 1u : 0u) << j;
                P.bits_out[row * ngroups + g] = w;
            }
        }
        if (EPI == EPI_MASK && P.mask_bits) {
            const uint32_t w = gi == 0 ? mw[0] : (gi == 1 ? mw[1] : (gi == 2 ? mw[2] : mw[3]));
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = ((w >> j) & 1u) ? v[j] : 0.f;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; q++)
            *reinterpret_cast<float4*>(tile + lane * 36 + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
        const int rsub = lane >> 3, col = c0 + 4 * (lane & 7);
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int r = it * 4 + rsub;
            const int64_t row = wrow0 + r;
            if (row >= P.rows || col >= P.N) continue;
            float4 x = *reinterpret_cast<const float4*>(tile + r * 36 + 4 * (lane & 7));
            if (EPI == EPI_MASK && !P.mask_bits) {
                const float* ms = P.mask_src + row * P.ldm + col;
                if (!(__ldg(ms) > 0.f)) x.x = 0.f;
                if (col + 1 < P.N && !(__ldg(ms + 1) > 0.f)) x.y = 0.f;
                if (col + 2 < P.N && !(__ldg(ms + 2) > 0.f)) x.z = 0.f;
                if (col + 3 < P.N && !(__ldg(ms + 3) > 0.f)) x.w = 0.f;
            }
            float* o = P.out + row * P.ldo + col;
            if (vec_out && col + 4 <= P.N) *reinterpret_cast<float4*>(o) = x;
            else {
                o[0] = x.x;
                if (col + 1 < P.N) o[1] = x.y;
                if (col + 2 < P.N) o[2] = x.z;
                if (col + 3 < P.N) o[3] = x.w;
            }
        }
    }
    MLP_STAMP(tr, 42);
    tc_fence_before();
    __syncthreads();
    MLP_STAMP(tr, 43);
    if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
}

int gemm_smem_bytes(int Npad)
{
    const int stages = 2 * (2 * (int)A_IMG + 2 * Npad * KC * 2), tiles = (GEMM_THREADS / 32) * 32 * 36 * 4;      // the epilogue reuses the stages
    return stages > tiles ? stages : tiles;
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient: C[M, N] += sum over rows r of P[r, m] * Q[r, n]   (dW = dz^T . a)
// Both operands are row-major with the reduction index (the row) slow, i.e. MN-major for the tensor core.  Canonical MN-major
// no-swizzle image of a [KC rows, MN columns] chunk: byte(mn, k) = (k/8) * MN*16 + (mn/8) * 128 + (k%8) * 16 + (mn%8) * 2
// (core matrix = 8 k-rows of 16 bytes = 8 consecutive mn elements; mn groups 128 bytes apart = SBO, k groups MN*16 bytes apart = LBO).
// A warp stages (8 k-rows) x (32 mn columns) per item: lane = kq + 8*mg reads 32 bytes of row kq (128 contiguous bytes per row
// over the four mg lanes) and stores 16 bytes per image; the warp's stores cover 512 contiguous bytes (conflict-free).
// One CTA = one 128-wide M tile x all N columns over a contiguous range of rows; accumulator in tensor memory for the whole range.
// ---------------------------------------------------------------------------------------------------------------------
struct WgradParams {
    const float* Pm;  int64_t ldp;  int relu_p;      // [rows, >= M]
    const float* Qm;  int64_t ldq;  int relu_q;      // [rows, >= N]
    int64_t rows;
    int M, N, Npad;                                   // M: valid columns of P (tiles of 128), N: valid columns of Q
    int passes;
    float* out;  int64_t ldo;  int transpose_out;     // out[m*ldo + n] += C[m][n]   (transpose_out: out[n*ldo + m])
    int64_t rows_per_cta;                             // multiple of KC
};

template <int ITEMS>
__device__ __forceinline__ void mn_load(const float* __restrict__ X, int64_t ldx, int64_t rows, int ncols, int64_t r0, int col0, int warp, int lane,
                                        float (&v)[ITEMS][8])
{
    const int kq = lane & 7, mg = lane >> 3;
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        const int witem = it * WG_WARPS + warp;                   // (k group, quad of mn groups)
        const int kg = witem & (KC / 8 - 1), mq = witem / (KC / 8);
        const int64_t row = r0 + kg * 8 + kq;
        const int col = col0 + mq * 32 + mg * 8;
        if (row < rows && col + 8 <= ncols) {
            const float* src = X + row * ldx + col;
            if ((reinterpret_cast<uintptr_t>(src) & 31) == 0) ldg8(src, v[it]);
            else {
                const float4 x = __ldg(reinterpret_cast<const float4*>(src));
                const float4 y = __ldg(reinterpret_cast<const float4*>(src) + 1);
                v[it][0] = x.x; v[it][1] = x.y; v[it][2] = x.z; v[it][3] = x.w; v[it][4] = y.x; v[it][5] = y.y; v[it][6] = y.z; v[it][7] = y.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) v[it][j] = (row < rows && col + j < ncols) ? __ldg(X + row * ldx + col + j) : 0.f;
        }
    }
}
template <int ITEMS>
__device__ __forceinline__ void mn_store(uint8_t* hi_img, uint8_t* lo_img, int MN, int relu, int warp, int lane, float (&v)[ITEMS][8])
{
    const int kq = lane & 7, mg = lane >> 3;
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        const int witem = it * WG_WARPS + warp;
        const int kg = witem & (KC / 8 - 1), mq = witem / (KC / 8);
        if (mq * 32 >= MN) continue;                               // narrow operands: fewer warp items than warps
        if (relu) {
#pragma unroll
            for (int j = 0; j < 8; j++) v[it][j] = fmaxf(v[it][j], 0.f);
        }
        uint4 hi, lo;
        split2(v[it][0], v[it][1], hi.x, lo.x); split2(v[it][2], v[it][3], hi.y, lo.y);
        split2(v[it][4], v[it][5], hi.z, lo.z); split2(v[it][6], v[it][7], hi.w, lo.w);
        const uint32_t off = (uint32_t)kg * (MN * 16) + (uint32_t)(mq * 4 + mg) * 128 + (uint32_t)kq * 16;
        *reinterpret_cast<uint4*>(hi_img + off) = hi;
        *reinterpret_cast<uint4*>(lo_img + off) = lo;
    }
}

// PI / QI = P / Q items per thread = (KC/8) * (width/32) / 16 warps: 2 for width 256, 1 for 128 and below (then some warps idle).
// Sixteen staging warps: with eight, a chunk's conversion took 1.8 us of a 2.0 us chunk period on two warps per scheduler (clock64 trace) -
// the kernel was bound by the issue rate of its own staging code, at 41 % of the HBM rate.
// One CTA covers every 128-row M tile of the result (P is staged Mpad wide, one tensor-memory accumulator per M tile), so each operand
// is read from HBM exactly once.
template <int PI, int QI>
__global__ void __launch_bounds__(WG_LAUNCH_THREADS, 1) mlp_wgrad_kernel(WgradParams W)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int Mpad = PI * 128;                                             // 128 or 256
    const int mtiles = Mpad / TILE_M;
    const uint32_t p_bytes = (uint32_t)KC * Mpad * 2, q_bytes = (uint32_t)KC * W.Npad * 2;
    const uint32_t stage_bytes = 2 * p_bytes + 2 * q_bytes;
    __shared__ __align__(8) uint64_t mma_done[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t r_begin = (int64_t)blockIdx.x * W.rows_per_cta;
    int64_t r_end = r_begin + W.rows_per_cta;
    if (r_end > W.rows) r_end = W.rows;
    if (r_begin >= r_end) return;
    const int nchunks = (int)((r_end - r_begin + KC - 1) / KC);
    const uint32_t acc_cols = W.Npad < 32 ? 32u : (uint32_t)W.Npad;       // columns of one accumulator (a power of two >= 32)
    const uint32_t tmem_cols = acc_cols * mtiles;

    // two chunks of both operands in flight per thread (register double buffer, buffer = chunk parity = stage), as in the rows GEMM
    const bool issuer = warp == WG_WARPS;                // the last warp: proxy fence + MMAs (see mlp_rows_gemm_kernel)
    const bool wtr = blockIdx.x == 70 && tid == 0;       // (trace builds)
    (void)wtr;
    MLP_STAMP(wtr, 96);
    float vp[2][PI][8], vq[2][QI][8];
    if (!issuer) {
        mn_load<PI>(W.Pm, W.ldp, r_end, W.M, r_begin, 0, warp, lane, vp[0]);
        mn_load<QI>(W.Qm, W.ldq, r_end, W.N, r_begin, 0, warp, lane, vq[0]);
    }
    if (!issuer && nchunks > 1) {
        mn_load<PI>(W.Pm, W.ldp, r_end, W.M, r_begin + KC, 0, warp, lane, vp[1]);
        mn_load<QI>(W.Qm, W.ldq, r_end, W.N, r_begin + KC, 0, warp, lane, vq[1]);
    }
    if (tid == 0) {
        mbar_init(&mma_done[0], 1); mbar_init(&mma_done[1], 1);
        fence_barrier_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t idesc = instr_desc(TILE_M, W.Npad, 1, 1);

    auto chunk = [&](int c, float (&p)[PI][8], float (&q)[QI][8]) {
        const int s = c & 1;
        uint8_t* st = smem + (size_t)s * stage_bytes;
        if (c >= 2) mbar_wait(&mma_done[s], ((c >> 1) - 1) & 1);
        if (!issuer) {
            mn_store<PI>(st, st + p_bytes, Mpad, W.relu_p, warp, lane, p);
            mn_store<QI>(st + 2 * p_bytes, st + 2 * p_bytes + q_bytes, W.Npad, W.relu_q, warp, lane, q);
            if (c + 2 < nchunks) {
                const int64_t r0 = r_begin + (int64_t)(c + 2) * KC;
                mn_load<PI>(W.Pm, W.ldp, r_end, W.M, r0, 0, warp, lane, p);
                mn_load<QI>(W.Qm, W.ldq, r_end, W.N, r0, 0, warp, lane, q);
            }
        }
        __syncthreads();
        if (issuer && lane == 0) {
            fence_proxy_async();
            tc_fence_after();
            const uint32_t p_hi = smem_u32(st), p_lo = p_hi + p_bytes, q_hi = p_hi + 2 * p_bytes, q_lo = q_hi + q_bytes;
            const uint32_t p_lbo = (uint32_t)Mpad * 16, q_lbo = (uint32_t)W.Npad * 16;
#pragma unroll
            for (int ks = 0; ks < KC / 16; ks++) {
                const uint64_t dqh = smem_desc(q_hi + ks * 2 * q_lbo, q_lbo, 128), dql = smem_desc(q_lo + ks * 2 * q_lbo, q_lbo, 128);
                for (int mt = 0; mt < mtiles; mt++) {                      // M tile mt = mn groups [16 mt, 16 mt + 16) of the P image
                    const uint32_t po = ks * 2 * p_lbo + mt * (TILE_M / 8) * 128;
                    const uint64_t dph = smem_desc(p_hi + po, p_lbo, 128), dpl = smem_desc(p_lo + po, p_lbo, 128);
                    const uint32_t td = tmem_d + mt * acc_cols;
                    umma(td, dph, dqh, idesc, (c | ks) != 0);
                    if (W.passes == 3) {
                        umma(td, dph, dql, idesc, 1);
                        umma(td, dpl, dqh, idesc, 1);
                    }
                }
            }
            umma_commit(&mma_done[s]);
        }
    };
    for (int c = 0; c < nchunks; c += 2) {
        chunk(c, vp[0], vq[0]);
        if (c + 1 < nchunks) chunk(c + 1, vp[1], vq[1]);
    }
    MLP_STAMP(wtr, 97);
    mbar_wait(&mma_done[(nchunks - 1) & 1], ((nchunks - 1) >> 1) & 1);
    __syncwarp();
    tc_fence_after();
    MLP_STAMP(wtr, 98);
    // epilogue: thread t of quadrant q holds row m0 + 32q + t, 32 consecutive n.  The 148 partial results are reduced into the shared
    // result with 128-bit reductions; the 32 x 32 block goes through a shared-memory tile first (the stages are free now) so that a warp
    // instruction covers four whole 128-byte lines instead of 32 lines of 16 bytes each - the L2's reduction rate bounds this phase
    // (2.4 M float4 reductions per launch) - and the CTAs walk the column groups in different orders so that they do not all hit the
    // same lines at the same time.
    const int quad = warp & 3;
    const int ngroups = (W.Npad + 31) / 32;
    const int ngi = ngroups * mtiles, step = WG_WARPS / 4;
    const int iters = (ngi + step - 1) / step;
    float* tile = reinterpret_cast<float*>(smem) + warp * (32 * 36);
    for (int k = 0; k < (issuer ? 0 : iters); k++) {
        const int gi = (warp >> 2) + ((k + (int)blockIdx.x) % iters) * step;
        if (gi >= ngi) continue;
        const int mt = gi / ngroups, g = gi - mt * ngroups;
        float v[32];
        __syncwarp();
        tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * acc_cols + g * 32), v);
        const int n0 = g * 32;
        if (!W.transpose_out && n0 + 32 <= W.N && (W.ldo & 3) == 0) {
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; q++)
                *reinterpret_cast<float4*>(tile + lane * 36 + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int r = it * 4 + (lane >> 3);
                const int m = mt * TILE_M + quad * 32 + r;
                if (m < W.M)
                    atomicAdd(reinterpret_cast<float4*>(W.out + (size_t)m * W.ldo + n0 + 4 * (lane & 7)),
                              *reinterpret_cast<const float4*>(tile + r * 36 + 4 * (lane & 7)));
            }
        } else {
            const int m = mt * TILE_M + quad * 32 + lane;
            if (m < W.M) {
#pragma unroll
                for (int j = 0; j < 32; j++)
                    if (n0 + j < W.N) atomicAdd(W.transpose_out ? W.out + (size_t)(n0 + j) * W.ldo + m : W.out + (size_t)m * W.ldo + n0 + j, v[j]);
            }
        }
    }
    MLP_STAMP(wtr, 99);
    tc_fence_before();
    __syncthreads();
    MLP_STAMP(wtr, 100);
    if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------
// Harmonic embedding (model/networks/HarmonicEmbedding.py; CoordMLP.forward :75-84): E = [x', sin(x' f_0..f_{n-1}), cos(...)],
// x' = (|x|, y, z) when the field is symmetric; freq f_i = scalar * 2^i; layout of the reference: for the sin block the index is
// d * n + i (d = coordinate), then the cos block.  E rows are padded to ldE floats (zeros).  Accurate sinf / cosf: arguments reach
// ~1e3 rad at the top octave.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mlp_embed_fwd_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int n_harm, float scalar, int symmetrize,
                                                            int concat_pts, float* __restrict__ E, int64_t ldE)
{
    // 32 lanes per row: lane j < 3 * n_harm owns (coordinate d, octave i) = (j / n_harm, j % n_harm) -> one sin and one cos; the
    // remaining lanes write the raw coordinates and the zero padding.  A warp writes one row: contiguous segments.
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    float p[3] = {__ldg(x + r * ldx), __ldg(x + r * ldx + 1), __ldg(x + r * ldx + 2)};
    if (symmetrize) p[0] = fabsf(p[0]);
    float* e = E + r * ldE;
    const int o = concat_pts ? 3 : 0, nh3 = 3 * n_harm;
    for (int j = lane; j < nh3; j += 32) {
        const int d = j / n_harm, i = j - d * n_harm;
        const float a = p[d] * (scalar * (float)(1u << i));
        float sn, cs;
        sincosf(a, &sn, &cs);          // one argument reduction for both (same values as sinf / cosf)
        e[o + j] = sn;
        e[o + nh3 + j] = cs;
    }
    if (concat_pts && lane < 3) e[lane] = p[lane];
    for (int k = o + 2 * nh3 + lane; k < ldE; k += 32) e[k] = 0.f;
}

// d_x[d] = dE[x'_d] + sum_i f_i (dE[sin] cos(a) - dE[cos] sin(a)); symmetric fields: d_x[0] *= sign(x[0]) (torch.abs: 0 at 0).
// One warp per row, like the forward: lane j owns (coordinate, octave) j, the row of dE is read as contiguous segments, the three sums
// are warp reductions (a thread per row read its 63 floats with a row-sized stride: 45 us for 185 k rows).
__global__ void __launch_bounds__(256) mlp_embed_bwd_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int n_harm, float scalar, int symmetrize,
                                                            int concat_pts, const float* __restrict__ dE, int64_t ldE, float* __restrict__ d_x, int64_t lddx)
{
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float x0 = __ldg(x + r * ldx);
    float p[3] = {x0, __ldg(x + r * ldx + 1), __ldg(x + r * ldx + 2)};
    if (symmetrize) p[0] = fabsf(p[0]);
    const float* g = dE + r * ldE;
    const int o = concat_pts ? 3 : 0, nh3 = 3 * n_harm;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int j = lane; j < nh3; j += 32) {
        const int d = j / n_harm, i = j - d * n_harm;
        const float f = scalar * (float)(1u << i);
        float sn, cs;
        sincosf(p[d] * f, &sn, &cs);
        const float t = f * (__ldg(g + o + j) * cs - __ldg(g + o + nh3 + j) * sn);
        acc[0] += d == 0 ? t : 0.f; acc[1] += d == 1 ? t : 0.f; acc[2] += d == 2 ? t : 0.f;
    }
    acc[0] = warp_sum(acc[0]); acc[1] = warp_sum(acc[1]); acc[2] = warp_sum(acc[2]);
    if (lane < 3) {
        float v = (lane == 0 ? acc[0] : (lane == 1 ? acc[1] : acc[2])) + (concat_pts ? __ldg(g + lane) : 0.f);
        if (lane == 0 && symmetrize) v = x0 > 0.f ? v : (x0 < 0.f ? -v : 0.f);
        d_x[r * lddx + lane] = v;
    }
}

// out[s, n] += sum of G[r, n] over the rows r of segment s = [seg_start[s], seg_start[s+1]) - the adjoint of a per-image bias when
// the rows are grouped by image.  grid (segments, ceil(N/32), row splits); 8 warps stride the block's share of the rows, lanes take
// 32 columns; one atomic per (block, column) into the zero-initialised result.
__global__ void __launch_bounds__(256) mlp_colsum_segments_kernel(const float* __restrict__ G, int64_t ldg, const int64_t* __restrict__ seg_start, int N,
                                                                  float* __restrict__ out)
{
    __shared__ float part[8][32];
    const int sgm = blockIdx.x, n = blockIdx.y * 32 + (threadIdx.x & 31), warp = threadIdx.x >> 5;
    const int64_t a0 = seg_start[sgm], b0 = seg_start[sgm + 1];
    const int64_t share = (b0 - a0 + gridDim.z - 1) / gridDim.z;
    const int64_t a = a0 + share * blockIdx.z, b = a + share < b0 ? a + share : b0;
    float acc = 0.f;
    if (n < N)
        for (int64_t r = a + warp; r < b; r += 8) acc += __ldg(G + r * ldg + n);
    part[warp][threadIdx.x & 31] = acc;
    __syncthreads();
    if (warp == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) t += part[w][threadIdx.x];
        if (t != 0.f) atomicAdd(out + (size_t)sgm * N + n, t);
    }
}

// N = 256, 32-byte aligned rows: a warp reads one whole row per instruction (lane = 8 consecutive columns, one 256-bit load), four
// rows in flight per warp; the block's 8 warps stride its share of the segment's rows and are reduced through shared memory.
// (The generic kernel above reads 128 bytes per row and block - 1.8 TB/s; this form streams whole rows.)
__global__ void __launch_bounds__(256) mlp_colsum256_kernel(const float* __restrict__ G, int64_t ldg, const int64_t* __restrict__ seg_start,
                                                            float* __restrict__ out)
{
    __shared__ float part[8][256];
    const int sgm = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t a0 = seg_start[sgm], b0 = seg_start[sgm + 1];
    const int64_t share = (b0 - a0 + gridDim.y - 1) / gridDim.y;
    const int64_t a = a0 + share * blockIdx.y, b = a + share < b0 ? a + share : b0;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int64_t r = a + warp;
    for (; r + 24 < b; r += 32) {
        float v0[8], v1[8], v2[8], v3[8];
        ldg8(G + r * ldg + lane * 8, v0);
        ldg8(G + (r + 8) * ldg + lane * 8, v1);
        ldg8(G + (r + 16) * ldg + lane * 8, v2);
        ldg8(G + (r + 24) * ldg + lane * 8, v3);
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] += (v0[j] + v1[j]) + (v2[j] + v3[j]);
    }
    for (; r < b; r += 8) {
        float v0[8];
        ldg8(G + r * ldg + lane * 8, v0);
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] += v0[j];
    }
#pragma unroll
    for (int j = 0; j < 8; j++) part[warp][lane * 8 + j] = acc[j];
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) t += part[w][threadIdx.x];
    if (t != 0.f) atomicAdd(out + (size_t)sgm * 256 + threadIdx.x, t);
}

// out[r, :] = src[idx[r], :] for the covered rows of a dense [n, C] image (C floats per row): one thread per (row, 16-byte piece) when
// rows are 16-byte multiples, else per element.  (torch's index_select took 98 us for 185 k rows of 64 bytes.)
__global__ void __launch_bounds__(256) rows_gather_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, int64_t N, int C,
                                                          float* __restrict__ out)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((C & 3) == 0) {
        const int q = C >> 2;
        if (t >= N * q) return;
        const int64_t r = t / q;
        const int k = (int)(t - r * q);
        reinterpret_cast<float4*>(out)[t] = __ldg(reinterpret_cast<const float4*>(src + __ldg(idx + r) * C) + k);
    } else {
        if (t >= N * C) return;
        const int64_t r = t / C;
        out[t] = __ldg(src + __ldg(idx + r) * C + (t - r * C));
    }
}

// dst[idx[r], :] = src[r, :] (dst rows not named by idx are left as they are: the caller zero-fills)
__global__ void __launch_bounds__(256) rows_scatter_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, int64_t N, int C,
                                                           float* __restrict__ dst)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((C & 3) == 0) {
        const int q = C >> 2;
        if (t >= N * q) return;
        const int64_t r = t / q;
        const int k = (int)(t - r * q);
        reinterpret_cast<float4*>(dst + __ldg(idx + r) * C)[k] = __ldg(reinterpret_cast<const float4*>(src) + t);
    } else {
        if (t >= N * C) return;
        const int64_t r = t / C;
        dst[__ldg(idx + r) * C + (t - r * C)] = __ldg(src + t);
    }
}

}  // namespace

#ifdef B2A_MLP_TRACE
B2A_API int b2a_debug_mlp_trace(unsigned long long* out128)
{
    return cudaMemcpyFromSymbol(out128, g_mlp_trace, sizeof(unsigned long long) * 128) == cudaSuccess ? 0 : 1;
}
#endif

// Covered-row gather / scatter between the dense shaded image [n, C] and the compact rows [N, C] the field networks are evaluated on
// (3danimals_b200/render/render.py _sample_field; reference: the dense material.sample / dino_net.sample at render.py:54,61).
// idx [N] int64 row numbers (unique).  scatter: dst is zero-filled here (cudaMemsetAsync) when zero_fill != 0.
B2A_API int b2a_rows_gather(const float* src, const int64_t* idx, int64_t N, int C, float* out, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(N >= 0 && C > 0 && ((src && idx && out) || N == 0), "arguments");
    B2A_CHECK_ARG((C & 3) != 0 || ((((uintptr_t)src | (uintptr_t)out) & 15) == 0), "16-byte alignment");
    if (N == 0) return 0;
    const int64_t n = (C & 3) == 0 ? N * (C >> 2) : N * C;
    rows_gather_kernel<<<b2a_blocks(n, 256), 256, 0, stream>>>(src, idx, N, C, out);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_rows_scatter(const float* src, const int64_t* idx, int64_t N, int C, float* dst, int64_t dst_rows, int zero_fill, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(dst && N >= 0 && C > 0 && dst_rows >= N && ((src && idx) || N == 0), "arguments");
    B2A_CHECK_ARG((C & 3) != 0 || ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0), "16-byte alignment");
    if (zero_fill) B2A_CUDA_OK(cudaMemsetAsync(dst, 0, (size_t)dst_rows * C * sizeof(float), stream));
    if (N == 0) return 0;
    const int64_t n = (C & 3) == 0 ? N * (C >> 2) : N * C;
    rows_scatter_kernel<<<b2a_blocks(n, 256), 256, 0, stream>>>(src, idx, N, C, dst);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_mlp_packed_bytes(int N, int K, size_t* bytes)
{
    B2A_CHECK_ARG(bytes && N > 0 && N <= 256 && K > 0, "shape");
    const int Npad = (N + 15) / 16 * 16, Kpad = (K + KC - 1) / KC * KC;
    *bytes = (size_t)Npad * Kpad * 2 * sizeof(uint16_t);
    return 0;
}

B2A_API int b2a_mlp_pack_weights(const float* W, int64_t ldw, int N, int K, int transpose, void* packed, size_t packed_bytes, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t need;
    int rc = b2a_mlp_packed_bytes(N, K, &need);
    if (rc) return rc;
    B2A_CHECK_ARG(W && packed && packed_bytes >= need && ((uintptr_t)packed & 15) == 0, "packed buffer");
    const int Npad = (N + 15) / 16 * 16, Kpad = (K + KC - 1) / KC * KC;
    mlp_pack_weights_kernel<<<b2a_blocks((int64_t)Npad * Kpad, 256), 256, 0, stream>>>(W, (int)ldw, N, K, transpose, Npad, Kpad, (uint16_t*)packed);
    B2A_LAUNCH_OK();
    return 0;
}

// jobs: HOST array of n x 6 int64 {W (device pointer), ldw, N, K, transpose, packed (device pointer, b2a_mlp_packed_bytes(N, K) bytes,
// 16-byte aligned)}: b2a_mlp_pack_weights for each, in one launch.  n <= 32.
B2A_API int b2a_mlp_pack_weights_many(const int64_t* jobs, int n, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(jobs && n > 0 && n <= PACK_MAX_JOBS, "1..32 jobs");
    PackJobs J;
    long long total = 0;
    for (int i = 0; i < n; i++) {
        const int64_t* q = jobs + 6 * i;
        PackJob& b = J.j[i];
        b.W = (const float*)(uintptr_t)q[0]; b.ldw = (int)q[1]; b.N = (int)q[2]; b.K = (int)q[3]; b.transpose = (int)q[4];
        b.out = (uint16_t*)(uintptr_t)q[5];
        B2A_CHECK_ARG(b.W && b.out && b.N > 0 && b.N <= 256 && b.K > 0 && ((uintptr_t)b.out & 15) == 0, "job");
        b.Npad = (b.N + 15) / 16 * 16; b.Kpad = (b.K + KC - 1) / KC * KC;
        b.first = total;
        total += (long long)b.Npad * b.Kpad;
    }
    J.n = n; J.total = total;
    unsigned blocks = b2a_blocks(total, 256);
    if (blocks > 148u * 8u) blocks = 148u * 8u;
    mlp_pack_many_kernel<<<blocks, 256, 0, stream>>>(J);
    B2A_LAUNCH_OK();
    return 0;
}

// out[rows, N] = epilogue(op(A)[rows, K] . Wpacked^T).  epilogue 0: + bias (vector [N], or row bias_rows[r] of a [n_img, N] table;
// nullable); 1: zero where mask_src[r, n] <= 0 (the ReLU derivative of the stored pre-activation); 2: sigmoid(. + bias).
// relu_on_load: op(A) = max(A, 0).  passes: 3 (fp32-grade split products) or 1 (single bf16 product).
B2A_API int b2a_mlp_rows_gemm(const float* A, int64_t lda, int64_t rows, int K, const void* packed, int N, int relu_on_load, int passes, int epilogue,
                              const float* bias, const int32_t* bias_rows, const float* mask_src, int64_t ldm, const uint32_t* mask_bits,
                              uint32_t* bits_out, float* out, int64_t ldo, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(A && packed && out, "null pointer");
    B2A_CHECK_ARG(rows >= 0 && K > 0 && N > 0 && N <= 256 && lda >= K && ldo >= N && (passes == 1 || passes == 3), "shape");
    B2A_CHECK_ARG(epilogue >= 0 && epilogue <= 2 && (epilogue != EPI_MASK || mask_bits || (mask_src && ldm >= N)), "epilogue");
    B2A_CHECK_ARG(((uintptr_t)A & 15) == 0 && (lda & 3) == 0 && ((uintptr_t)packed & 15) == 0 && ((uintptr_t)out & 15) == 0, "alignment (16 bytes, lda % 4 == 0)");
    if (rows == 0) return 0;
    GemmParams P;
    P.A = A; P.lda = lda; P.rows = rows; P.K = K; P.Kpad = (K + KC - 1) / KC * KC; P.N = N; P.Npad = (N + 15) / 16 * 16;
    P.packed = (const uint16_t*)packed; P.relu_on_load = relu_on_load; P.passes = passes; P.out = out; P.ldo = ldo; P.bias = bias;
    P.bias_rows = bias_rows; P.mask_src = mask_src; P.ldm = ldm; P.mask_bits = mask_bits; P.bits_out = bits_out;
    const int smem = gemm_smem_bytes(P.Npad);
    const unsigned grid = b2a_blocks(rows, TILE_M);
#define MLP_LAUNCH(E)                                                                                                     \
    do {                                                                                                                  \
        B2A_CUDA_OK(cudaFuncSetAttribute(mlp_rows_gemm_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));    \
        mlp_rows_gemm_kernel<E><<<grid, GEMM_LAUNCH, smem, stream>>>(P);                                                   \
    } while (0)
    if (epilogue == EPI_BIAS) MLP_LAUNCH(EPI_BIAS);
    else if (epilogue == EPI_MASK) MLP_LAUNCH(EPI_MASK);
    else MLP_LAUNCH(EPI_SIGMOID);
#undef MLP_LAUNCH
    B2A_LAUNCH_OK();
    return 0;
}

// C[M, N] += P^T Q over all rows: out[m*ldo + n] (or out[n*ldo + m] with transpose_out) is ACCUMULATED with reductions - the caller
// zero-initialises it.  relu_p / relu_q: the operand is max(., 0) of the stored tensor.  M, N <= 256.
B2A_API int b2a_mlp_wgrad(const float* P, int64_t ldp, int relu_p, const float* Q, int64_t ldq, int relu_q, int64_t rows, int M, int N, int passes,
                          float* out, int64_t ldo, int transpose_out, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(P && Q && out, "null pointer");
    B2A_CHECK_ARG(rows >= 0 && M > 0 && M <= 256 && N > 0 && N <= 256 && ldp >= M && ldq >= N && (passes == 1 || passes == 3), "shape");
    B2A_CHECK_ARG(((uintptr_t)P & 15) == 0 && ((uintptr_t)Q & 15) == 0 && (ldp & 3) == 0 && (ldq & 3) == 0 && ((uintptr_t)out & 15) == 0, "alignment");
    if (rows == 0) return 0;
    WgradParams W;
    W.Pm = P; W.ldp = ldp; W.relu_p = relu_p; W.Qm = Q; W.ldq = ldq; W.relu_q = relu_q; W.rows = rows; W.M = M; W.N = N;
    W.Npad = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));      // whole warp items of 32 columns
    W.passes = passes; W.out = out; W.ldo = ldo; W.transpose_out = transpose_out;
    const int mtiles = (M + TILE_M - 1) / TILE_M;
    int64_t nsplit = 148;                                                  // one persistent CTA per SM
    int64_t per = (rows + nsplit - 1) / nsplit;
    per = (per + KC - 1) / KC * KC;
    if (per < 8 * KC) per = 8 * KC;
    W.rows_per_cta = per;
    nsplit = (rows + per - 1) / per;
    int smem = 2 * (2 * KC * mtiles * TILE_M * 2 + 2 * KC * W.Npad * 2);
    if (smem < WG_WARPS * 32 * 36 * 4) smem = WG_WARPS * 32 * 36 * 4;      // the epilogue's transposition tiles reuse the stages
    const dim3 grid((unsigned)nsplit);
#define WG_LAUNCH(PI, QI)                                                                                                \
    do {                                                                                                                 \
        B2A_CUDA_OK(cudaFuncSetAttribute(mlp_wgrad_kernel<PI, QI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        mlp_wgrad_kernel<PI, QI><<<grid, WG_LAUNCH_THREADS, smem, stream>>>(W);                                               \
    } while (0)
    if (mtiles == 2) {
        if (W.Npad == 256) WG_LAUNCH(2, 2);
        else WG_LAUNCH(2, 1);
    } else {
        if (W.Npad == 256) WG_LAUNCH(1, 2);
        else WG_LAUNCH(1, 1);
    }
#undef WG_LAUNCH
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_mlp_embed_fwd(const float* x, int64_t ldx, int64_t rows, int n_harmonic, float scalar, int symmetrize, int concat_pts, float* E,
                              int64_t ldE, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(x && E && rows >= 0 && n_harmonic >= 0 && ldx >= 3 && ldE >= 6 * n_harmonic + (concat_pts ? 3 : 0), "shape");
    B2A_CHECK_ARG(n_harmonic <= 24, "at most 24 octaves");
    if (rows) mlp_embed_fwd_kernel<<<b2a_blocks(rows * 32, 256), 256, 0, stream>>>(x, ldx, rows, n_harmonic, scalar, symmetrize, concat_pts, E, ldE);
    B2A_LAUNCH_OK();
    return 0;
}

B2A_API int b2a_mlp_embed_bwd(const float* x, int64_t ldx, int64_t rows, int n_harmonic, float scalar, int symmetrize, int concat_pts, const float* dE,
                              int64_t ldE, float* d_x, int64_t lddx, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(x && dE && d_x && rows >= 0 && n_harmonic >= 0 && ldx >= 3 && lddx >= 3 && ldE >= 6 * n_harmonic + (concat_pts ? 3 : 0), "shape");
    if (rows) mlp_embed_bwd_kernel<<<b2a_blocks(rows * 32, 256), 256, 0, stream>>>(x, ldx, rows, n_harmonic, scalar, symmetrize, concat_pts, dE, ldE, d_x, lddx);
    B2A_LAUNCH_OK();
    return 0;
}

// out[s, n] (zeroed here, then accumulated) = sum over rows seg_start[s] <= r < seg_start[s+1] of G[r, n]; seg_start: device int64 [n_seg + 1]
B2A_API int b2a_mlp_colsum_segments(const float* G, int64_t ldg, const int64_t* seg_start, int n_seg, int N, float* out, b2a_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2A_CHECK_ARG(G && seg_start && out && n_seg > 0 && N > 0 && ldg >= N, "shape");
    B2A_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)n_seg * N * sizeof(float), stream));
    if (N == 256 && (ldg & 7) == 0 && ((uintptr_t)G & 31) == 0) {
        int splits = (8 * 148 + n_seg - 1) / n_seg;
        if (splits > 1024) splits = 1024;
        mlp_colsum256_kernel<<<dim3(n_seg, splits), 256, 0, stream>>>(G, ldg, seg_start, out);
        B2A_LAUNCH_OK();
        return 0;
    }
    int splits = (4 * 148) / (n_seg * ((N + 31) / 32));
    if (splits < 1) splits = 1;
    if (splits > 64) splits = 64;
    mlp_colsum_segments_kernel<<<dim3(n_seg, (N + 31) / 32, splits), 256, 0, stream>>>(G, ldg, seg_start, N, out);
    B2A_LAUNCH_OK();
    return 0;
}
