// Decoder-MLP layers on the 5th-generation tensor cores (tcgen05, kind::tf32, accumulators in TMEM).
//
// Forward: one persistent, warp-specialised CTA per SM computes  Y[M,128] (+)= act(bn(X))[M,K] * W[128,K]^T + bias  for its share
// of 64-pixel tiles, as the TRANSPOSED product  Y^T[out, px] = W[out, K] * act(bn(X))[px, K]^T  (17 warps):
//   warps 0-7   converters: the raw (pre-BatchNorm) input rows travel global -> shared with 16-byte cp.async, each piece straight to
//                           its place in the 128-byte-swizzled K-major UMMA image, up to three tiles ahead (four 32 KB stages); the thread
//                           that copied a piece converts it IN PLACE (BatchNorm + Softplus in the log2 domain, TF32 rounding), then
//                           fence.proxy.async, mbarrier
//   warp  8     MMA issuer: one elected lane issues K/8 tcgen05.mma (M = 128 output channels, N = 64 pixels, K = 8) per tile into one of
//                           four 64-column TMEM accumulators, tcgen05.commit -> mbarrier (which also frees the input stage)
//   warps 9-16  epilogue  : two warps per TMEM lane quarter, 32 pixels each; a TMEM lane is an OUTPUT CHANNEL, so bias and the layer's
//                           BatchNorm sum / sum-of-squares are thread-local, and Y leaves straight from the registers: a warp
//                           instruction writes 32 channels of one pixel = one whole 128-byte line (no staging pass, no named barrier)
// The layer moves 1 KiB per pixel against 32.8 kFLOP: HBM is the bound.  Round 1 computed Y (not Y^T) in 128-pixel tiles with two
// 64 KB stages, register loads and a shared-memory staging pass for coalesced row stores: 0.57 of the HBM peak; this form 0.69
// (0.863 -> 0.714 ms per step at config 3; lookahead 1 / 2 / 3 tiles: 0.93 / 0.73 / 0.71 ms).
//
// TF32 is what the reference computes these 1x1 convolutions in on any Ampere+ GPU (cuDNN allow_tf32 default,
// SURVEY.md §8 a-4); the strict-FP32 CUDA-core path (gemm.cuh) remains selectable and is the GPU-side reference.
#include "gemm.cuh"
#include "tc_common.cuh"
#include <type_traits>

// timing-ablation switches for tools/variants.py (never set in the product build): 1 no global loads in the producers, 2 no activation
// math, 4 no MMA issue, 8 no epilogue math, 16 no epilogue global traffic
#ifndef GA_ABLATE
#define GA_ABLATE 0
#endif
// GA_TC_TIMING=1: every warp of CTA 0 prints the cycles it spent inside mbarrier waits (tools/variants.py; never set in the product build)
#ifndef GA_TC_TIMING
#define GA_TC_TIMING 0
#endif
#if GA_TC_TIMING
#define TWAIT(slot, call) do { const long long t0_ = clock64(); call; tw_[slot] += clock64() - t0_; } while (0)
#else
#define TWAIT(slot, call) call
#endif
#if GA_TC_TIMING
#define TMARK(var) const long long var = clock64()
#define TADD(slot, var) tw_[slot] += clock64() - var
#else
#define TMARK(var)
#define TADD(slot, var)
#endif

namespace ga {
namespace {

using namespace tc;

constexpr int kBM = 128, kBN = 128;
constexpr int kChunkBytes = kBM * 128;          // one 32-channel chunk of a 128-row tile: 16 KB
constexpr int kMaxChunks = 4;                    // K <= 128
constexpr int kStageBytes = kMaxChunks * kChunkBytes;   // 64 KB
constexpr uint32_t kTmemCols = 256;

struct TcFwdParams {
    const float *X; int ldx; int K;
    const float *a, *b;            // folded BatchNorm of the producer layer (nullptr: raw input, no activation)
    const float *W; int ldw;       // [128][K] row-major
    const float *bias;             // nullable
    float *Y; int ldy; int accumulate;
    double *sum, *sumsq;           // nullable
    int M;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---------------------------------------------------------------------------------------------------------------------
// Forward layer (see the file header).  Operand images: A = W [4 chunks][128 out rows][128 B], B = the X tile [4 chunks][64 px rows]
// [128 B], both K-major with the 128-byte swizzle (16-byte units XOR row % 8); a chunk is 32 input channels.
constexpr int kTPx = 64, kTStages = 4;
constexpr int kTChunk = kTPx * 128;                 // 8 KB: [64 px rows][32 channels]
constexpr int kTStageBytes = kMaxChunks * kTChunk;   // 32 KB
constexpr int kTThreads = 17 * 32;
constexpr int kTMmaWarp = 8, kTEpiWarp0 = 9;
#ifndef GA_FWDT_LOOK
#define GA_FWDT_LOOK 3
#endif
constexpr int kTLook = GA_FWDT_LOOK;                 // tiles of copies in flight per converter thread (1..3)

struct alignas(1024) TcFwdTSmem {
    unsigned char w[kStageBytes];                    // W: [4 chunks][128 out rows][128 B], K-major, 128-byte swizzle
    unsigned char x[kTStages][kTStageBytes];         // X tile: [4 chunks][64 px rows][128 B]
    float sa[kBN], sb[kBN], sbias[kBN];
    double red[2][kBN];
    uint64_t full[kTStages], mma_done[kTStages], tmem_empty[kTStages];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kTThreads, 1)
tc_fwd_kernel(const TcFwdParams p)
{
    extern __shared__ unsigned char smem_raw[];
    TcFwdTSmem &sm = *reinterpret_cast<TcFwdTSmem *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int num_tiles = (p.M + kTPx - 1) / kTPx;
    const int gstep = gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < kTStages; ++s) {
            mbar_init(&sm.full[s], 8 * kArrivalsPerWarp); mbar_init(&sm.mma_done[s], 1); mbar_init(&sm.tmem_empty[s], 8 * kArrivalsPerWarp);
        }
        fence_barrier_init();
    }
    if (warp == kTMmaWarp) tmem_alloc(&sm.tmem_base, kTmemCols);
    pdl_wait();
    for (int i = tid; i < kBN * (p.K / 4); i += kTThreads) {          // weights -> shared, K-major, 128-byte swizzle
        const int n = i / (p.K / 4), q = i % (p.K / 4);
        const float4 v = *reinterpret_cast<const float4 *>(p.W + (size_t)n * p.ldw + q * 4);
        *reinterpret_cast<float4 *>(sm.w + (q >> 3) * kChunkBytes + sw128_offset(n, q & 7)) = to_tf32(v);
    }
    for (int i = tid; i < kBN; i += kTThreads) {
        sm.sbias[i] = p.bias ? p.bias[i] : 0.f;
        if (i < p.K) { sm.sa[i] = p.a ? p.a[i] : 1.f; sm.sb[i] = p.a ? p.b[i] : 0.f; }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = sm.tmem_base;

    if (warp < kTMmaWarp) {
        // ================================ converters ================================
        // warp w owns channel chunk c = w & 3 and pixel rows r0 + 8 i (i = 0..7): a warp instruction covers 4 rows x 128 B.
        const int rl = lane >> 3, u = lane & 7;
        const int c = warp & 3, r0 = (warp >> 2) * 4 + rl;
        const int k = c * 32 + u * 4;
        const bool kin = k < p.K, chunk_used = c * 32 < p.K;
        float4 av = make_float4(1.f, 1.f, 1.f, 1.f), bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kin) { av = *reinterpret_cast<const float4 *>(&sm.sa[k]); bv = *reinterpret_cast<const float4 *>(&sm.sb[k]); }
        const bool act = p.a != nullptr && !(GA_ABLATE & 2);
        {   // log2 domain: z * log2(e) comes straight out of the FMA
            const float L2E = 1.44269504089f;
            av.x *= L2E; av.y *= L2E; av.z *= L2E; av.w *= L2E; bv.x *= L2E; bv.y *= L2E; bv.z *= L2E; bv.w *= L2E;
        }
        const int M = p.M, ldx = p.ldx;
        const uint32_t soff = (uint32_t)c * kTChunk + sw128_offset(r0, u);      // + i * 1024
        const float *x0 = p.X + (size_t)r0 * ldx + k;
        auto issue = [&](int t, int st) {
            if (!chunk_used || (GA_ABLATE & 1)) return;
            const uint32_t dst = smem_u32(sm.x[st]) + soff;
            const float *src = x0 + (size_t)t * kTPx * ldx;
            const bool whole = (t + 1) * kTPx <= M;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool ok = kin && (whole || t * kTPx + r0 + 8 * i < M);      // rows >= M and channels >= K are zero-filled, never read
                cp_async16(dst + i * 1024, ok ? src + (size_t)(8 * i) * ldx : p.X, ok);
            }
        };
        // the stage of tile t + kTStages is free once the tensor core has read tile t (mma_done): the first refill blocks, the others only probe
        int tl = blockIdx.x, sl = 0, nl = 0, ahead = 0;
        auto refill = [&]() {
            while (ahead < kTLook && tl < num_tiles) {
                if (ahead == 0) warp_wait(&sm.mma_done[sl], (nl & 1) ^ 1, lane);
                else if (!warp_test(&sm.mma_done[sl], (nl & 1) ^ 1, lane)) break;
                issue(tl, sl);
                cp_async_commit();                         // one group per tile
                ++ahead; tl += gstep; if (++sl == kTStages) { sl = 0; ++nl; }
            }
        };
        refill();
        int s = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gstep) {
            if (ahead >= 3) cp_async_wait<2>(); else if (ahead == 2) cp_async_wait<1>(); else cp_async_wait<0>();
            if (chunk_used) {
                unsigned char *xs = sm.x[s] + soff;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 x = (GA_ABLATE & 1) ? make_float4(0.5f, 0.25f, 0.125f, 1.f) : *reinterpret_cast<const float4 *>(xs + i * 1024);
                    if (act) {
                        x.x = softplus_log2(fmaf(x.x, av.x, bv.x)); x.y = softplus_log2(fmaf(x.y, av.y, bv.y));
                        x.z = softplus_log2(fmaf(x.z, av.z, bv.z)); x.w = softplus_log2(fmaf(x.w, av.w, bv.w));
                    }
                    *reinterpret_cast<float4 *>(xs + i * 1024) = to_tf32(x);       // rows >= M hold act(bn(0)): their columns are never stored
                }
            }
            fence_proxy_async_smem();
            warp_arrive(&sm.full[s], lane);
            if (++s == kTStages) s = 0;
            --ahead;
            refill();
        }
    } else if (warp == kTMmaWarp) {
        // ================================ MMA issuer ================================
        constexpr uint32_t idesc = make_idesc_tf32(128, kTPx, false, false);
        const uint32_t w_lo = desc_lo(smem_u32(sm.w), 16);
        const int nk = (GA_ABLATE & 4) ? 0 : p.K / 8;               // 16 (K = 128) or 9 (K = 72)
        int s = 0, n = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gstep) {
            warp_wait(&sm.full[s], n & 1, lane);
            warp_wait(&sm.tmem_empty[s], (n & 1) ^ 1, lane);
            tc_fence_after_sync();
            if (lane == 0) {
                constexpr uint32_t hi = desc_hi(1024);
                const uint32_t x_lo = desc_lo(smem_u32(sm.x[s]), 16);
                const uint32_t d = tmem_base + (uint32_t)s * kTPx;
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (k < nk)
                        mma_tf32_lohi(d, w_lo + (uint32_t)(k >> 2) * (kChunkBytes >> 4) + (uint32_t)(k & 3) * 2u,
                                      x_lo + (uint32_t)(k >> 2) * (kTChunk >> 4) + (uint32_t)(k & 3) * 2u, hi, idesc, k > 0);
                mma_commit(&sm.mma_done[s]);
            }
            __syncwarp();
            if (++s == kTStages) { s = 0; ++n; }
        }
    } else {
        // ================================ epilogue: one output channel per thread ================================
        const int ew = warp - kTEpiWarp0, q = warp & 3, half = ew >> 2;
        const int c = q * 32 + lane;                 // TMEM lane == output channel
        const float bias = sm.sbias[c];
        const int ldy = p.ldy;
        const bool accumulate = p.accumulate && !(GA_ABLATE & 16);
        const size_t ystep = (size_t)gstep * kTPx * ldy;
        float *yp = p.Y + ((size_t)blockIdx.x * kTPx + half * 32) * ldy + c;        // this thread's column, first of its 32 rows
        double dsum = 0.0, dsq = 0.0;
        int s = 0, n = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gstep, yp += ystep) {
            const int rows = p.M - (tile * kTPx + half * 32);                       // valid rows of this warp's slice (may be <= 0 or >= 32)
            float ev[16];
            auto ldev = [&](int ph) {                // existing Y (the fan-in's other half): fetched before the accumulator is needed
#pragma unroll
                for (int j = 0; j < 16; ++j) ev[j] = (ph * 16 + j < rows) ? yp[(ph * 16 + j) * ldy] : 0.f;
            };
            if (accumulate) ldev(0);
            warp_wait(&sm.mma_done[s], n & 1, lane);
            tc_fence_after_sync();
            float cs = 0.f, cq = 0.f;
#pragma unroll 1
            for (int ph = 0; ph < 2; ++ph) {
                float v[16];
                tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)s * kTPx + half * 32 + ph * 16, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += bias;
                if (accumulate) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += ev[j];
                    if (ph == 0) ldev(1);
                }
                if (rows >= (ph + 1) * 16) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (!(GA_ABLATE & 16)) yp[(ph * 16 + j) * ldy] = v[j];
                        if (!(GA_ABLATE & 8)) { cs += v[j]; cq = fmaf(v[j], v[j], cq); }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (ph * 16 + j < rows) {
                            if (!(GA_ABLATE & 16)) yp[(ph * 16 + j) * ldy] = v[j];
                            if (!(GA_ABLATE & 8)) { cs += v[j]; cq = fmaf(v[j], v[j], cq); }
                        }
                }
            }
            tc_fence_before_sync();
            warp_arrive(&sm.tmem_empty[s], lane);          // accumulator drained
            dsum += (double)cs; dsq += (double)cq;
            if (++s == kTStages) { s = 0; ++n; }
        }
        if (p.sum) {                                 // the two pixel halves of a channel meet in shared memory: 2 atomics per channel per CTA
            if (half == 1) { sm.red[0][c] = dsum; sm.red[1][c] = dsq; }
            named_bar_sync(1, 8 * 32);
            if (half == 0) { atomicAdd(&p.sum[c], dsum + sm.red[0][c]); atomicAdd(&p.sumsq[c], dsq + sm.red[1][c]); }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kTMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

// Host-side launcher shared by the decoder orchestration and the unit-test entry point.
int launch_tc_fwd(const float *X, int ldx, int K, const float *a, const float *b, const float *W, int ldw, const float *bias, float *Y,
                  int ldy, int accumulate, double *sum, double *sumsq, int M, cudaStream_t st)
{
    GA_REQUIRE(K % 8 == 0 && K >= 8 && K <= 128, "tcgen05 layer: K=%d must be a multiple of 8 in [8,128]", K);
    GA_REQUIRE(ldx % 4 == 0 && ldw % 4 == 0 && ldy % 4 == 0, "tcgen05 layer: leading dimensions must be multiples of 4");
    TcFwdParams p{X, ldx, K, a, b, W, ldw, bias, Y, ldy, accumulate, sum, sumsq, M};
    static PerDeviceOnce attr_t_set;
    if (attr_t_set.first()) {
        GA_CHECK_CUDA(cudaFuncSetAttribute(tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcFwdTSmem) + 1024));
    }
    const int tiles = cdiv(M, kTPx);
    const int grid = tiles < kNumSMs ? tiles : kNumSMs;
    {
        ProfScope _ps("mlp_tc_fwd", st);
        launch_k(tc_fwd_kernel, grid, kTThreads, sizeof(TcFwdTSmem) + 1024, st, p);
    }
    GA_CHECK_LAUNCH("tc_fwd_kernel");
    return GA_OK;
}

}  // namespace ga

// Unit-test / building-block entry: Y[M,128] (+)= act(bn(X))[M,K] W[128,K]^T + bias on the tensor cores.
extern "C" int ga_tc_linear_forward(int32_t M, int32_t K, const float *X, int32_t ldx, const float *bn_a, const float *bn_b, const float *W,
                                    int32_t ldw, const float *bias, float *Y, int32_t ldy, int32_t accumulate, double *sum, double *sumsq,
                                    void *stream)
{
    GA_REQUIRE(M > 0 && X && W && Y, "bad arguments");
    return ga::launch_tc_fwd(X, ldx, K, bn_a, bn_b, W, ldw, bias, Y, ldy, accumulate, sum, sumsq, M, static_cast<cudaStream_t>(stream));
}

// =====================================================================================================================
// Backward of one hidden layer on the tensor cores: data gradient AND weight gradient from ONE pass over the tile.
//
//   G  = dY_l      [px, out] = ga (dZ_l - m1 - xhat_l m2)          BatchNorm backward applied while loading (dZ_l, Y_l)
//   X  = x_l       [px, in ] = softplus(a Y_{l-1} + b)             recomputed while loading Y_{l-1}
//   dW_l [out,in] += G^T X      K = pixel: A = G and B = X are consumed MN-major straight from their pixel-major images
//                               (128-byte swizzle with a 32-byte base, UMMA layout type 1); the accumulator lives in TMEM for
//                               the whole kernel, one flush (16-byte vector reductions) at the end
//   dX^T [in, px]  = W_l^T G^T  A = W_l^T [in rows][out] (transposed once per CTA), B = G [px rows][out] K-major; computed
//                               TRANSPOSED so that a TMEM lane is an input channel: the epilogue thread's BatchNorm
//                               scalars are constants and sum(dZ), sum(dZ xhat) need no cross-thread reduction
//   dZ_{l-1} = dX * sigmoid(z_{l-1})   (sigmoid(z_{l-1}) and xhat_{l-1} recovered from x = softplus(z) in the X tile: no re-read)
// G is stored in two swizzles of the same [px][channel] image (K-major operands reject layout type 1), X once.  The raw operands
// travel global -> shared with 16-byte cp.async (no register behind a load in flight) and are converted IN PLACE; dZ_{l-1} leaves
// straight from the epilogue registers (one whole 128-byte line per warp instruction).  32-pixel tiles, three stages, 25 warps:
// 8 converters for G, 8 for X, 1 MMA issuer, 8 epilogue (two per TMEM lane quarter, 16 pixels each).
// Round-2 log (tools/variants.py on the device, profiles/r2_tc_bwd_ncu.md): register loads one tile ahead -> cp.async in place
// 1.97 -> 1.63 ms per step; shared-memory staging + named barrier + row stores -> direct stores 1.63 -> 1.55 ms.  What did NOT pay:
// bulk L2 prefetch from the MMA warp (3.1 ms: 128 small TMA requests per tile), dedicated copy warps / copies issued by the
// epilogue warps with mbarrier completion (1.8 / 1.7 ms: one more parked wait per tile on the ring), spinning instead of
// suspended mbarrier waits (1.8 ms), "arrive before prefetch" (2.0 ms).  With every load, store, MMA and activation ablated the
// skeleton (waits, converter LDS/STS, tcgen05.ld) still takes 58 us of the 129 us per launch: the kernel is bound by the per-tile
// hand-offs of its 32-pixel tiles, which the 227 KB of shared memory do not allow to grow (64 KB W^T + 3 x 48 KB stages).
// =====================================================================================================================
namespace ga {
namespace {

#ifndef GA_BWD_PX
#define GA_BWD_PX 32
#endif
constexpr int kPx = GA_BWD_PX;            // pixels per tile: 32, or 64 (two stages) when W^T lives in tensor memory
constexpr int kBlk = kPx / 8, kPh = kPx / 16;
static_assert(kPx == 32 || kPx == 64, "tc_bwd tile: 32 pixels (default) or 64; 48 was measured (no gain) and is not supported");
#ifndef GA_BWD_WT_TMEM
#define GA_BWD_WT_TMEM 1
#endif
constexpr int kBStages = GA_BWD_PX == 64 ? 2 : GA_BWD_WT_TMEM ? 4 : 3;   // W^T in tensor memory frees 64 KB of shared memory: a fourth 48 KB stage
constexpr int kGkChunk = kPx * 128;       // 4 KB: [32 px rows][32 channels]
constexpr int kGkTile = 4 * kGkChunk;     // 16 KB
constexpr int kMnTile = 4 * kGkChunk;     // 16 KB: MN-major image [4 x 32-channel chunks][32 px][128 B] (32-byte-base swizzle)
#ifndef GA_BWD_EWARPS
#define GA_BWD_EWARPS 8
#endif
#ifndef GA_BWD_G_MMADONE
#define GA_BWD_G_MMADONE 1
#endif
#ifndef GA_BWD_LOOK
#define GA_BWD_LOOK (GA_BWD_PX == 64 ? 1 : GA_BWD_WT_TMEM ? 3 : 2)
#endif
constexpr int kLook = GA_BWD_LOOK;               // tiles of raw operand copies in flight per producer thread (1 or 2; < kBStages)
constexpr int kBwdEWarps = GA_BWD_EWARPS;  // epilogue warps: 4 (each 2 x 16 pixels) or 8 (two per TMEM lane quarter, 16 pixels each)
constexpr int kBwdThreads = (17 + kBwdEWarps) * 32;      // 8 G-producer warps, 8 X-producer warps, 1 MMA warp, epilogue warps
constexpr int kBwdMmaWarp = 16, kBwdEpiWarp0 = 17;
constexpr uint32_t kBwdTmemCols = GA_BWD_WT_TMEM ? 512 : 256;    // [0,128): dW accumulator; 128 + 32 s: dX^T accumulator of stage s; [256,384): W^T
constexpr uint32_t kWtCol = 256;

struct TcBwdParams {
    const float *dZ, *Y; int ldg;                      // layer l: [M][ldg], 128 output channels from the pointer
    const float *ga, *m1, *m2, *mu, *rstd;             // BatchNorm backward of layer l (ga == nullptr: G = dZ)
    const float *Yprev; int ldp;                       // layer l-1 pre-BN output, 128 channels
    const float *pa, *pb, *pmu, *prstd;                // BatchNorm (folded) of layer l-1
    const float *W; int ldw;                           // [128 out][128 in]
    float *dW; int lddw;                               // += (atomics)
    float *dZprev; int ldo;
    int mode;                                          // 0 final, 1 store raw dX, 2 accumulate raw dX, 3 final on (existing + dX)
    double *s1, *s2;
    int M;
    int kin;                                           // valid input channels (128, or 72 for the [features|uv] input)
    int x_raw;                                         // X = Yprev as is (the network input: no BatchNorm / Softplus; modes 1, 2 only)
};

struct alignas(1024) TcBwdSmem {
#if !GA_BWD_WT_TMEM
    unsigned char wt[4 * 128 * 128];                   // 64 KB: W^T, rows = input channel, K = output channel
#endif
    unsigned char gk[kBStages][kGkTile];               // G   rows = pixel,          K = output channel   (dgrad B operand)
    unsigned char gm[kBStages][kMnTile];               // G   MN-major: M = output channel, K = pixel     (wgrad A operand)
    unsigned char xm[kBStages][kMnTile];               // X   MN-major: N = input channel,  K = pixel     (wgrad B operand)
    float gA[128], gB[128], gC[128];                   // dY = gA dZ + gB Y + gC  (BatchNorm backward folded to two FMAs)
    float pa[128], pb[128], pbeta[128], pinvg[128];    // layer l-1: z log2e = y pa + pb ; xhat = (z - pbeta) * pinvg
    uint64_t full[kBStages], empty[kBStages], mma_done[kBStages], tmem_empty[kBStages];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kBwdThreads, 1)
tc_bwd_kernel(const TcBwdParams p)
{
    extern __shared__ unsigned char smem_raw[];
    TcBwdSmem &sm = *reinterpret_cast<TcBwdSmem *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int num_tiles = (p.M + kPx - 1) / kPx;
#if GA_TC_TIMING
    long long tw_[4] = {0, 0, 0, 0};
    const long long tstart_ = clock64();
#endif

    if (tid == 0) {
        for (int s = 0; s < kBStages; ++s) {
            mbar_init(&sm.full[s], 16 * kArrivalsPerWarp); mbar_init(&sm.empty[s], kBwdEWarps * kArrivalsPerWarp);        // arrivals are per WARP (warp_arrive)
            mbar_init(&sm.mma_done[s], 1); mbar_init(&sm.tmem_empty[s], kBwdEWarps * kArrivalsPerWarp);
        }
        fence_barrier_init();
    }
    if (warp == kBwdMmaWarp) tmem_alloc(&sm.tmem_base, kBwdTmemCols);
    pdl_wait();            // barrier init and TMEM allocation above overlap the previous kernel's tail
#if !GA_BWD_WT_TMEM
    // W^T -> shared (K-major over the OUTPUT channel): element (in, out) <- W[out][in]
    // thread = (input channel, group of 4 output channels): four coalesced row reads of W, ONE 16-byte store of the transposed quad
    // (the 8 lanes of a store phase have 8 different `in & 7`, hence 8 different 16-byte units: conflict-free)
    for (int i = tid; i < 128 * 32; i += kBwdThreads) {
        const int in = i & 127, o = (i >> 7) * 4;
        if (in < p.kin) {
            const float *wp = p.W + (size_t)o * p.ldw + in;
            const float4 v = to_tf32(make_float4(wp[0], wp[p.ldw], wp[2 * p.ldw], wp[3 * p.ldw]));
            *reinterpret_cast<float4 *>(sm.wt + (o >> 5) * (128 * 128) + in * 128 + ((((o & 31) >> 2) ^ (in & 7)) << 4)) = v;
        }
    }
#endif
    for (int i = tid; i < 128; i += kBwdThreads) {
        if (p.ga) {   // ga (dZ - m1 - (Y - mu) rstd m2)
            const float ga = p.ga[i], k = p.rstd[i] * p.m2[i];
            sm.gA[i] = ga; sm.gB[i] = -ga * k; sm.gC[i] = ga * (k * p.mu[i] - p.m1[i]);
        } else { sm.gA[i] = 1.f; sm.gB[i] = 0.f; sm.gC[i] = 0.f; }
        if (p.x_raw) { sm.pa[i] = 1.f; sm.pb[i] = 0.f; sm.pbeta[i] = 0.f; sm.pinvg[i] = 0.f; }
        else {
            // folded BatchNorm of layer l-1 in the log2 domain; xhat = (y - mu) rstd = (z - beta) / gamma with
            // gamma = a / rstd, beta = b + mu a  (a, b: folded scale / shift)
            const float a = p.pa[i], b = p.pb[i];
            sm.pa[i] = a * 1.44269504089f; sm.pb[i] = b * 1.44269504089f;
            sm.pbeta[i] = b + p.pmu[i] * a; sm.pinvg[i] = p.prstd[i] / a;
        }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = sm.tmem_base;
#if GA_BWD_WT_TMEM
    // W^T -> TENSOR MEMORY: the dgrad's A operand (rows = input channel = TMEM lane, one column per output channel) is read by the
    // tensor core from there, so it costs neither shared memory (64 KB) nor shared-memory bandwidth (it was re-read for every tile).
    // Four warps, one per lane quarter: lanes read W[out][in] coalesced over `in`.
    if (warp >= kBwdEpiWarp0 && warp < kBwdEpiWarp0 + 4) {
        const int in = (warp & 3) * 32 + lane;
        const float *wp = p.W + in;
#pragma unroll 1
        for (int cc = 0; cc < 8; ++cc) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = in < p.kin ? to_tf32(wp[(size_t)(cc * 16 + j) * p.ldw]) : 0.f;
            tmem_st_32x16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kWtCol + cc * 16, v);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
#endif
#if GA_TC_TIMING
    const long long tpro_ = clock64() - tstart_;
    long long tloop_ = 0;
#endif

    // ================================ converters: warps 0-7 build G (two swizzles), warps 8-15 build X =================
    // Both wgrad operands are consumed MN-major (pixel = K index), which for 32-bit elements means the 128-byte swizzle with a
    // 32-byte base (layout type 1; the physical layout was decoded with tools/exp_umma_probe.cu):
    //     byte(px, c) = (c / 32) * LBO + (px / 4) * SBO + (px % 4) * 128 + ((((c % 32) / 8) ^ (px % 4)) * 32) + (c % 8) * 4
    // so a pixel's 32 channels stay one permuted 128-byte row, exactly like the K-major image the dgrad needs (16-byte units
    // XOR px % 8): the two images differ only by a permutation of the 16-byte pieces inside each 128-byte row.  A warp instruction
    // covers 4 pixels x 128 B (full-line global requests) and every quarter-warp touches one whole row (conflict-free).
    // Warp w owns the 32-channel chunk j = w & 3 (its BatchNorm coefficients are thread constants) and pixel rows p0 + 8 e.
    const int M = p.M;
    if (warp < kBwdMmaWarp) {
        const int c16 = lane & 7, pxl = lane >> 3;
        const int pw = warp & 7;
        const int j = pw & 3, p0 = (pw >> 2) * 4 + pxl;                      // block e is pixel p0 + 8 e:  px & 7 == p0, px & 3 == pxl
        const int ch = j * 32 + c16 * 4;                                     // this thread's 4 channels (all tiles, all blocks)
        const uint32_t k_off = (uint32_t)j * kGkChunk + (uint32_t)p0 * 128u + (uint32_t)((c16 ^ p0) << 4);                       // + e * 1024
        const uint32_t mn_off = (uint32_t)j * kGkChunk + (uint32_t)p0 * 128u + (uint32_t)((((c16 >> 1) ^ pxl) << 5) | ((c16 & 1) << 4));
        int s = 0, n = 0;
        const int gstep = gridDim.x;
        int tile = blockIdx.x;
        if (warp < 8) {
            const float4 cA = *reinterpret_cast<const float4 *>(&sm.gA[ch]), cB = *reinterpret_cast<const float4 *>(&sm.gB[ch]),
                         cC = *reinterpret_cast<const float4 *>(&sm.gC[ch]);
            const int ldg = p.ldg;
            const bool has_y = p.ga != nullptr;
            const float *dz0 = p.dZ + (size_t)p0 * ldg + ch, *y0 = (has_y ? p.Y : p.dZ) + (size_t)p0 * ldg + ch;   // block e adds 8 e rows
            const size_t tile_stride = (size_t)kPx * ldg;
            // The raw dZ / Y pieces travel global -> shared with 16-byte cp.async, each straight to the address its converted value will
            // occupy (dZ: the K-major image, Y: the MN-major image), kLook tiles ahead and without a register; once a tile has landed, the
            // thread that copied a piece reads it back, folds the BatchNorm backward and overwrites both images IN PLACE -- every piece is
            // private to one thread, so the conversion needs no synchronisation beyond cp.async.wait_group.
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const int row8 = 8 * ldg;
            const ptrdiff_t ydelta = y0 - dz0;
            // running source pointer of the NEXT tile to copy, stage addresses as base + stage * size: a whole tile (every tile but possibly
            // the last) is 8 plain copies with no predicate or select -- this bookkeeping used to be 15 % of the kernel's instructions
            const float *pz = dz0 + (size_t)tile * tile_stride;
            const size_t pstep = (size_t)gstep * tile_stride;
            const uint32_t dk0 = smem_u32(sm.gk[0]) + k_off, dm0 = smem_u32(sm.gm[0]) + mn_off;
            auto issue = [&](int t, int st) {
                if (!(GA_ABLATE & 1)) {
                    const uint32_t dk = dk0 + (uint32_t)st * kGkTile, dm = dm0 + (uint32_t)st * kMnTile;
                    if ((t + 1) * kPx <= M) {
#pragma unroll
                        for (int e = 0; e < kBlk; ++e) cp_async16_full(dk + e * 1024, pz + e * row8);
                        if (has_y) {
#pragma unroll
                            for (int e = 0; e < kBlk; ++e) cp_async16_full(dm + e * 1024, pz + ydelta + e * row8);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < kBlk; ++e) {
                            const bool ok = t * kPx + e * 8 + p0 < M;          // rows >= M are zero-filled
                            cp_async16(dk + e * 1024, ok ? pz + e * row8 : dz0, ok);
                            if (has_y) cp_async16(dm + e * 1024, ok ? pz + ydelta + e * row8 : dz0, ok);
                        }
                    }
                }
                pz += pstep;
            };
            // Copies run up to two tiles ahead.  The stage of the tile after next is the one the epilogue frees LAST, so it is only probed:
            // if it is still busy the copy is issued one iteration later (blocking then would serialise conversion behind the epilogue).
            int tl = tile, sl = 0, nl = 0, ahead = 0;      // next tile to copy, its stage / round; tiles copied but not yet converted
            auto refill = [&]() {
                while (ahead < kLook && tl < num_tiles) {
                    uint64_t *freed = GA_BWD_G_MMADONE ? &sm.mma_done[sl] : &sm.empty[sl];   // the epilogue never touches the G images
                    if (ahead == 0) TWAIT(0, warp_wait(freed, (nl & 1) ^ 1, lane));
                    else if (!warp_test(freed, (nl & 1) ^ 1, lane)) break;
                    issue(tl, sl);
                    cp_async_commit();                     // one group per tile
                    ++ahead; tl += gstep; if (++sl == kBStages) { sl = 0; ++nl; }
                }
            };
            refill();
            for (; tile < num_tiles; tile += gstep) {
                if (ahead >= 3) cp_async_wait<2>(); else if (ahead == 2) cp_async_wait<1>(); else cp_async_wait<0>();      // this tile's pieces have landed (younger groups stay in flight)
                unsigned char *gk = sm.gk[s] + k_off, *gm = sm.gm[s] + mn_off;
                const bool ragged = (tile + 1) * kPx > M;
#pragma unroll
                for (int e = 0; e < kBlk; ++e) {
                    float4 za, ya;
                    if (GA_ABLATE & 1) { za = ya = make_float4(0.5f, 0.25f, 0.125f, 1.f); }
                    else {
                        za = *reinterpret_cast<const float4 *>(gk + e * 1024);
                        ya = has_y ? *reinterpret_cast<const float4 *>(gm + e * 1024) : za;
                    }
                    float4 o;
                    o.x = fmaf(cA.x, za.x, fmaf(cB.x, ya.x, cC.x)); o.y = fmaf(cA.y, za.y, fmaf(cB.y, ya.y, cC.y));
                    o.z = fmaf(cA.z, za.z, fmaf(cB.z, ya.z, cC.z)); o.w = fmaf(cA.w, za.w, fmaf(cB.w, ya.w, cC.w));
                    if (ragged && tile * kPx + e * 8 + p0 >= M) o = zero4;       // rows >= M carry zeros (gC alone would not be zero)
                    o = to_tf32(o);
                    *reinterpret_cast<float4 *>(gk + e * 1024) = o;
                    *reinterpret_cast<float4 *>(gm + e * 1024) = o;
                }
                fence_proxy_async_smem();
                warp_arrive(&sm.full[s], lane);
                if (++s == kBStages) { s = 0; ++n; }
                --ahead;
                refill();
            }
        } else {
            const int ldp = p.ldp;
            const bool raw = p.x_raw != 0 || (GA_ABLATE & 2);
            const bool x_read_by_epilogue = (p.mode == 0 || p.mode == 3) && !(GA_ABLATE & 8);      // final modes recover sigmoid / xhat from the X tile
            const float4 av = *reinterpret_cast<const float4 *>(&sm.pa[ch]), bv = *reinterpret_cast<const float4 *>(&sm.pb[ch]);
            const bool okc = ch < p.kin;
            const float *x0 = p.Yprev + (size_t)p0 * ldp + ch;
            const size_t tile_stride = (size_t)kPx * ldp;
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const int row8 = 8 * ldp;
            const float *px = x0 + (size_t)tile * tile_stride;
            const size_t pstep = (size_t)gstep * tile_stride;
            const uint32_t dm0 = smem_u32(sm.xm[0]) + mn_off;
            auto issue = [&](int t, int st) {
                if (!(GA_ABLATE & 1)) {
                    const uint32_t dm = dm0 + (uint32_t)st * kMnTile;
                    if (okc && (t + 1) * kPx <= M) {
#pragma unroll
                        for (int e = 0; e < kBlk; ++e) cp_async16_full(dm + e * 1024, px + e * row8);
                    } else {
#pragma unroll
                        for (int e = 0; e < kBlk; ++e) {
                            const bool ok = okc && t * kPx + e * 8 + p0 < M;     // channels past kin (the 72-wide input layer) are never read
                            cp_async16(dm + e * 1024, ok ? px + e * row8 : p.Yprev, ok);
                        }
                    }
                }
                px += pstep;
            };
            int tl = tile, sl = 0, nl = 0, ahead = 0;
            auto refill = [&]() {
                while (ahead < kLook && tl < num_tiles) {
                    uint64_t *freed = (GA_BWD_G_MMADONE && !x_read_by_epilogue) ? &sm.mma_done[sl] : &sm.empty[sl];
                    if (ahead == 0) TWAIT(0, warp_wait(freed, (nl & 1) ^ 1, lane));
                    else if (!warp_test(freed, (nl & 1) ^ 1, lane)) break;
                    issue(tl, sl);
                    cp_async_commit();
                    ++ahead; tl += gstep; if (++sl == kBStages) { sl = 0; ++nl; }
                }
            };
            refill();
            for (; tile < num_tiles; tile += gstep) {
                if (ahead >= 3) cp_async_wait<2>(); else if (ahead == 2) cp_async_wait<1>(); else cp_async_wait<0>();
                unsigned char *xm = sm.xm[s] + mn_off;
                const bool ragged = (tile + 1) * kPx > M;
#pragma unroll
                for (int e = 0; e < kBlk; ++e) {
                    float4 o = (GA_ABLATE & 1) ? make_float4(0.5f, 0.25f, 0.125f, 1.f) : *reinterpret_cast<const float4 *>(xm + e * 1024);
                    if (!raw) {
                        o.x = softplus_log2(fmaf(o.x, av.x, bv.x)); o.y = softplus_log2(fmaf(o.y, av.y, bv.y));
                        o.z = softplus_log2(fmaf(o.z, av.z, bv.z)); o.w = softplus_log2(fmaf(o.w, av.w, bv.w));
                    }
                    if (!okc || (ragged && tile * kPx + e * 8 + p0 >= M)) o = zero4;      // softplus(bn(0)) is not 0: rows >= M must be
                    *reinterpret_cast<float4 *>(xm + e * 1024) = to_tf32(o);
                }
                fence_proxy_async_smem();
                warp_arrive(&sm.full[s], lane);
                if (++s == kBStages) { s = 0; ++n; }
                --ahead;
                refill();
            }
        }
    } else if (warp == kBwdMmaWarp) {
        // ================================ MMA issuer ================================
        constexpr uint32_t idesc_dgrad = make_idesc_tf32(128, kPx, false, false);
        constexpr uint32_t idesc_wgrad = make_idesc_tf32(128, 128, true, true);      // both operands MN-major (K = pixel)
#if !GA_BWD_WT_TMEM
        const uint32_t wt_addr = smem_u32(sm.wt);
#endif
        int it = 0, s = 0, n = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            TWAIT(0, warp_wait(&sm.full[s], n & 1, lane));
            TWAIT(1, warp_wait(&sm.tmem_empty[s], (n & 1) ^ 1, lane));
            tc_fence_after_sync();
            TMARK(tm_);
            if (lane == 0 && (GA_ABLATE & 4)) mma_commit(&sm.mma_done[s]);
            if (lane == 0 && !(GA_ABLATE & 4)) {
                constexpr uint32_t hi = desc_hi(1024);
                // MN-major images: 32-channel chunks kGkChunk apart (LBO), 4-pixel swizzle atoms 512 B apart (SBO), layout type 1
                constexpr uint32_t hi_mn = (512u >> 4) | (1u << 14) | (1u << 29);
#if !GA_BWD_WT_TMEM
                const uint32_t wt_lo = desc_lo(wt_addr, 16);
#endif
                const uint32_t gk_lo = desc_lo(smem_u32(sm.gk[s]), 16), gm_lo = desc_lo(smem_u32(sm.gm[s]), kGkChunk),
                               xm_lo = desc_lo(smem_u32(sm.xm[s]), kGkChunk);
                const uint32_t d_dx = tmem_base + 128 + (uint32_t)s * kPx;
                // dX^T[in, px] = sum_out W^T[in, out] G[px, out]: 16 steps of 8 output channels
#pragma unroll
                for (int k = 0; k < 16; ++k)
#if GA_BWD_WT_TMEM
                    mma_tf32_ts(d_dx, tmem_base + kWtCol + (uint32_t)k * 8u, gk_lo + (uint32_t)(k >> 2) * (kGkChunk >> 4) + (uint32_t)(k & 3) * 2u, hi, idesc_dgrad, k > 0);
#else
                    mma_tf32_lohi(d_dx, wt_lo + (uint32_t)(k >> 2) * ((128 * 128) >> 4) + (uint32_t)(k & 3) * 2u,
                                  gk_lo + (uint32_t)(k >> 2) * (kGkChunk >> 4) + (uint32_t)(k & 3) * 2u, hi, idesc_dgrad, k > 0);
#endif
                // dW[out, in] += sum_px G^T[out, px] X^T[in, px]: 4 steps of 8 pixels
#pragma unroll
                for (int j = 0; j < kPx / 8; ++j)
                    mma_tf32_lohi(tmem_base, gm_lo + (uint32_t)j * (1024u >> 4), xm_lo + (uint32_t)j * (1024u >> 4), hi_mn, idesc_wgrad, (it > 0) || (j > 0));
                mma_commit(&sm.mma_done[s]);
            }
            __syncwarp();
            TADD(2, tm_);
            if (++s == kBStages) { s = 0; ++n; }
        }
    } else {
        // ================================ epilogue: one input channel per thread ================================
        // 4 warps (TMEM lane quarter = warp & 3).  A thread's BatchNorm scalars are constants; sigmoid(z_{l-1}) and
        // xhat_{l-1} are recovered from x = softplus(z) in the X^T tile (no global re-read); 2 x 16 pixels per tile.
        const int ew = warp - kBwdEpiWarp0;          // 0..kBwdEWarps-1
        constexpr int kPhStep = kBwdEWarps / 4;
        const int q = warp & 3;
        const int c = q * 32 + lane;                 // TMEM lane == input channel
        // xhat = (z - beta) / gamma with z = x + ln2 * log2(sigmoid):  xhat = ln2/gamma * ls + (x / gamma - beta / gamma)
        const float cinvg = sm.pinvg[c], cbg = sm.pbeta[c] * cinvg, cl2g = 0.69314718056f * cinvg;
        const bool final_mode = (p.mode == 0 || p.mode == 3) && !(GA_ABLATE & 8);
        const bool add_existing = p.mode >= 2 && c < p.kin && !(GA_ABLATE & 16);     // uniform per warp when kin is a multiple of 32
        uint32_t xoff[4];                            // this channel's word inside the row of a pixel with px % 4 == i (MN-major image of X)
#pragma unroll
        for (int i = 0; i < 4; ++i) xoff[i] = (uint32_t)(c >> 5) * kGkChunk + (uint32_t)((((lane >> 3) ^ i) << 5) + (lane & 7) * 4);
        const int ldo = p.ldo;
        double d1 = 0.0, d2 = 0.0;
        int s = 0, n = 0;
        // running pointers (advanced by one grid round of tiles): the channel column this thread accumulates into, and the rows this warp stores
        const size_t ostep = (size_t)gridDim.x * kPx * ldo;
        const float *evp = p.dZprev + (size_t)blockIdx.x * kPx * ldo + c;
        const bool store_col = c < p.kin && !(GA_ABLATE & 16);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, evp += ostep) {
            const int m0 = tile * kPx;
            const bool full = m0 + kPx <= p.M;            // only the last tile can be ragged: one uniform branch instead of a test per row
            // raw dX already accumulated by an earlier layer of this fan-in: fetched BEFORE waiting for the tensor core, and the second
            // half while the first is being processed, so the latency hides behind the MMA and the sigmoid math (rows >= M are never read)
            float ev[16];
            auto ldev = [&](int ph) {
                const float *ep = evp + ph * 16 * ldo;
                if (full) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) ev[j] = ep[j * ldo];
                } else {
                    const int rows = p.M - (m0 + ph * 16);
#pragma unroll
                    for (int j = 0; j < 16; ++j) ev[j] = (j < rows) ? ep[j * ldo] : 0.f;
                }
            };
            if (add_existing) ldev(ew >> 2);
            TWAIT(0, warp_wait(&sm.mma_done[s], n & 1, lane));
            tc_fence_after_sync();
            TMARK(te_);
            float t1 = 0.f, t2 = 0.f;
#pragma unroll 1
            for (int ph = ew >> 2; ph < kPh; ph += kPhStep) {
                float v[16];
                tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + 128 + (uint32_t)s * kPx + ph * 16, v);
                if (add_existing) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += ev[j];
                    if (ph + kPhStep < kPh) ldev(ph + kPhStep);
                }
                if (final_mode) {
                    const unsigned char *xr = sm.xm[s] + ph * (16 * 128);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        // the producers stored x + half a TF32 ulp (bit pattern + 0x1000, nothing truncated): undo it -> exact fp32 x
                        const float x = __uint_as_float(*reinterpret_cast<const uint32_t *>(xr + j * 128 + xoff[j & 3]) - 0x1000u);
                        float em, ls;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(em) : "f"(-1.44269504089f * x));
                        const float sg = 1.f - em;                                  // sigmoid(z) = 1 - exp(-softplus(z))
                        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(ls) : "f"(fmaxf(sg, 1e-30f)));
                        const float dx = v[j] * sg;                                 // 0 for pixels >= M (G and X are zero there)
                        v[j] = dx;
                        t1 += dx;
                        t2 = fmaf(dx, fmaf(ls, cl2g, fmaf(x, cinvg, -cbg)), t2);
                    }
                }
                // straight from the registers: a warp instruction writes 32 consecutive channels of one pixel = one whole 128-byte line
                if (store_col) {
                    float *oc = const_cast<float *>(evp) + ph * 16 * ldo;
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) oc[j * ldo] = v[j];
                    } else {
                        const int rows = p.M - (m0 + ph * 16);
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (j < rows) oc[j * ldo] = v[j];
                    }
                }
            }
            tc_fence_before_sync();
            warp_arrive(&sm.tmem_empty[s], lane);
            d1 += (double)t1; d2 += (double)t2;
            TADD(2, te_);
            TMARK(ts_);
            warp_arrive(&sm.empty[s], lane);               // the X tile has been read: nothing of this stage is needed any more
            TADD(3, ts_);
            if (++s == kBStages) { s = 0; ++n; }
        }
#if GA_TC_TIMING
        tloop_ = clock64() - tstart_;
#endif
        if (final_mode && p.s1) { atomicAdd(&p.s1[c], d1); atomicAdd(&p.s2[c], d2); }
        // flush the weight-gradient accumulator: lane == output channel; every MMA was covered by the last mma_done wait
        tc_fence_after_sync();
#pragma unroll 1
        for (int cc = (ew >> 2) * (8 / kPhStep); cc < ((ew >> 2) + 1) * (8 / kPhStep); ++cc) {      // with 8 warps each pair splits the columns
            float v[16];
            tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + cc * 16, v);
#pragma unroll
            for (int g = 0; g < 4; ++g)          // kin and lddw are multiples of 4: 16-byte vector reductions
                if (cc * 16 + g * 4 < p.kin) red_add_v4(p.dW + (size_t)c * p.lddw + cc * 16 + g * 4, v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
    }
#if GA_TC_TIMING
    if (blockIdx.x == 0 && lane == 0) printf("TCT bwd pro %lld loopend %lld mode=%d kin=%d warp %d total %lld wait0 %lld wait1 %lld ph2 %lld ph3 %lld tiles %d\n", tpro_, tloop_, p.mode, p.kin, warp, clock64() - tstart_, tw_[0], tw_[1], tw_[2], tw_[3], (num_tiles + (int)gridDim.x - 1) / (int)gridDim.x);
#endif
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kBwdMmaWarp) tmem_dealloc(tmem_base, kBwdTmemCols);
}

}  // namespace

int launch_tc_bwd(const float *dZ, const float *Y, int ldg, const float *ga, const float *m1, const float *m2, const float *mu, const float *rstd,
                  const float *Yprev, int ldp, const float *pa, const float *pb, const float *pmu, const float *prstd, const float *W, int ldw,
                  float *dW, int lddw, float *dZprev, int ldo, int mode, double *s1, double *s2, int M, int kin, int x_raw, cudaStream_t st)
{
    GA_REQUIRE(kin % 4 == 0 && kin >= 4 && kin <= 128 && (!x_raw || mode == 1 || mode == 2), "tcgen05 backward: bad kin / mode");
    GA_REQUIRE(ldg % 4 == 0 && ldp % 4 == 0 && ldw % 4 == 0 && ldo % 4 == 0, "tcgen05 backward: leading dimensions must be multiples of 4");
    GA_REQUIRE(lddw % 4 == 0 && (reinterpret_cast<uintptr_t>(dW) & 15) == 0, "tcgen05 backward: dW must be 16-byte aligned with lddw a multiple of 4");
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        GA_CHECK_CUDA(cudaFuncSetAttribute(tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcBwdSmem) + 1024));
    }
    TcBwdParams p{dZ, Y, ldg, ga, m1, m2, mu, rstd, Yprev, ldp, pa, pb, pmu, prstd, W, ldw, dW, lddw, dZprev, ldo, mode, s1, s2, M, kin, x_raw};
    const int tiles = cdiv(M, kPx);
    const int grid = tiles < kNumSMs ? tiles : kNumSMs;
    {
        ProfScope _ps("mlp_tc_bwd", st);
        launch_k(tc_bwd_kernel, grid, kBwdThreads, sizeof(TcBwdSmem) + 1024, st, p);
    }
    GA_CHECK_LAUNCH("tc_bwd_kernel");
    return GA_OK;
}

}  // namespace ga

// Unit-test / building-block entry for the fused backward layer (see tc_bwd_kernel).
extern "C" int ga_tc_linear_backward(int32_t M, const float *dZ, const float *Y, int32_t ldg, const float *bwd_coef /*[5][128] ga,m1,m2,mu,rstd or NULL*/,
                                     const float *Yprev, int32_t ldp, const float *prev_coef /*[4][128] a,b,mu,rstd*/, const float *W, int32_t ldw,
                                     float *dW, int32_t lddw, float *dZprev, int32_t ldo, int32_t mode, double *s1, double *s2, void *stream)
{
    GA_REQUIRE(M > 0 && dZ && Yprev && prev_coef && W && dW && dZprev, "bad arguments");
    const float *bc = bwd_coef;
    return ga::launch_tc_bwd(dZ, Y, ldg, bc, bc ? bc + 128 : nullptr, bc ? bc + 256 : nullptr, bc ? bc + 384 : nullptr, bc ? bc + 512 : nullptr, Yprev,
                             ldp, prev_coef, prev_coef + 128, prev_coef + 256, prev_coef + 384, W, ldw, dW, lddw, dZprev, ldo, mode, s1, s2, M,
                             128, 0, static_cast<cudaStream_t>(stream));
}
