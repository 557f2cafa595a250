// Decoder-MLP layers on the 5th-generation tensor cores (tcgen05, kind::tf32, accumulators in TMEM).
//
// One persistent, warp-specialised CTA per SM computes  Y[M,128] (+)= act(bn(X))[M,K] * W[128,K]^T + bias  for its share
// of 128-pixel tiles:
//   warps 0-3  producers : coalesced LDG of the raw (pre-BatchNorm) input tile, BatchNorm + Softplus applied in
//                          registers, STS into the 128-byte-swizzled K-major UMMA layout, fence.proxy.async, mbarrier
//   warp  4    MMA issuer: one elected lane issues K/8 tcgen05.mma (M=128, N=128, K=8) per tile into one of two TMEM
//                          accumulators, tcgen05.commit -> mbarrier
//   warps 5-8  epilogue  : tcgen05.ld the accumulator (lane = pixel row), + bias, stage the tile in the just-consumed
//                          input buffer (XOR-swizzled, conflict-free both ways), coalesced global stores, per-channel
//                          sum / sum-of-squares for the layer's BatchNorm (fp32 per tile, double across tiles)
// Two input stages + two accumulators overlap load/transform, MMA and store.  The layer is HBM-bound (64 KB in + 64 KB
// out per 2.1 MFLOP... DESIGN.md §4), so the point of the tensor core here is to take the math off the critical path.
//
// TF32 is what the reference computes these 1x1 convolutions in on any Ampere+ GPU (cuDNN allow_tf32 default,
// SURVEY.md §8 a-4); the strict-FP32 CUDA-core path (gemm.cuh) remains selectable and is the GPU-side reference.
#include "gemm.cuh"
#include "tc_common.cuh"

namespace ga {
namespace {

using namespace tc;

constexpr int kBM = 128, kBN = 128;
constexpr int kChunkBytes = kBM * 128;          // one 32-channel chunk of a 128-row tile: 16 KB
constexpr int kMaxChunks = 4;                    // K <= 128
constexpr int kStageBytes = kMaxChunks * kChunkBytes;   // 64 KB
constexpr int kTcThreads = 9 * 32;
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

struct alignas(1024) TcFwdSmem {
    unsigned char w[kStageBytes];
    unsigned char a[2][kStageBytes];
    float sa[kBN], sb[kBN], sbias[kBN];
    uint64_t full[2], empty[2], mma_done[2], tmem_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__global__ void __launch_bounds__(kTcThreads, 1)
tc_fwd_kernel(const TcFwdParams p)
{
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B operands need 1024-byte aligned tiles: align the dynamic window by hand (1 KB of slack is requested)
    TcFwdSmem &sm = *reinterpret_cast<TcFwdSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunks = (p.K + 31) / 32;
    const int num_tiles = (p.M + kBM - 1) / kBM;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&sm.full[s], 128); mbar_init(&sm.empty[s], 128);
            mbar_init(&sm.mma_done[s], 1); mbar_init(&sm.tmem_empty[s], 128);
        }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(&sm.tmem_base, kTmemCols);
    // weights -> shared, K-major, 128-byte swizzle
    for (int i = tid; i < kBN * (p.K / 4); i += kTcThreads) {
        const int n = i / (p.K / 4), q = i % (p.K / 4);
        const float4 v = *reinterpret_cast<const float4 *>(p.W + (size_t)n * p.ldw + q * 4);
        *reinterpret_cast<float4 *>(sm.w + (q >> 3) * kChunkBytes + sw128_offset(n, q & 7)) = v;
    }
    for (int i = tid; i < kBN; i += kTcThreads) {
        sm.sbias[i] = p.bias ? p.bias[i] : 0.f;
        if (i < p.K) { sm.sa[i] = p.a ? p.a[i] : 1.f; sm.sb[i] = p.a ? p.b[i] : 0.f; }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = sm.tmem_base;

    if (warp < 4) {
        // ================================ producers ================================
        const int rl = lane >> 3, u = lane & 7;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int s = it & 1, n = it >> 1;
            mbar_wait(&sm.empty[s], (n & 1) ^ 1);
            unsigned char *dst = sm.a[s];
            const int m0 = tile * kBM;
            for (int c = 0; c < chunks; ++c) {
                const int k = c * 32 + u * 4;
                const bool kin = k < p.K;
                float4 av = make_float4(1.f, 1.f, 1.f, 1.f), bv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kin) { av = *reinterpret_cast<const float4 *>(&sm.sa[k]); bv = *reinterpret_cast<const float4 *>(&sm.sb[k]); }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {       // 4 loads in flight per thread, 8 per chunk
                        const int r = (half * 4 + j) * 16 + warp * 4 + rl;
                        const int m = m0 + r;
                        v[j] = (kin && m < p.M) ? *reinterpret_cast<const float4 *>(p.X + (size_t)m * p.ldx + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = (half * 4 + j) * 16 + warp * 4 + rl;
                        float4 x = v[j];
                        if (p.a) {
                            x.x = softplus_f(fmaf(x.x, av.x, bv.x)); x.y = softplus_f(fmaf(x.y, av.y, bv.y));
                            x.z = softplus_f(fmaf(x.z, av.z, bv.z)); x.w = softplus_f(fmaf(x.w, av.w, bv.w));
                        }
                        *reinterpret_cast<float4 *>(dst + c * kChunkBytes + sw128_offset(r, u)) = x;
                    }
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(&sm.full[s]);
        }
    } else if (warp == 4) {
        // ================================ MMA issuer ================================
        constexpr uint32_t idesc = make_idesc_tf32(kBM, kBN, false, false);
        const uint32_t w_addr = smem_u32(sm.w);
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int s = it & 1, n = it >> 1;
            mbar_wait(&sm.full[s], n & 1);
            mbar_wait(&sm.tmem_empty[s], (n & 1) ^ 1);
            tc_fence_after_sync();
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(sm.a[s]);
                const int nk = p.K / 8;
                for (int k = 0; k < nk; ++k) {
                    const uint32_t off = (uint32_t)(k >> 2) * kChunkBytes + (uint32_t)(k & 3) * 32u;
                    mma_tf32(tmem_base + (uint32_t)s * kBN, make_smem_desc(a_addr + off, 16, 1024), make_smem_desc(w_addr + off, 16, 1024), idesc,
                             k > 0);
                }
                mma_commit(&sm.mma_done[s]);
            }
            __syncwarp();
        }
    } else {
        // ================================ epilogue ================================
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;               // tile row == TMEM lane
        const int et = (warp - 5) * 32 + lane;       // 0..127: channel owned for the statistics
        double dsum = 0.0, dsq = 0.0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int s = it & 1, n = it >> 1;
            const int m0 = tile * kBM;
            mbar_wait(&sm.mma_done[s], n & 1);
            tc_fence_after_sync();
            unsigned char *stg = sm.a[s];            // the MMA has finished reading this stage: reuse it as staging
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                float v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)s * kBN + cc * 32, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int un = cc * 8 + j;
                    float4 o = make_float4(v[4 * j] + sm.sbias[un * 4], v[4 * j + 1] + sm.sbias[un * 4 + 1], v[4 * j + 2] + sm.sbias[un * 4 + 2],
                                           v[4 * j + 3] + sm.sbias[un * 4 + 3]);
                    *reinterpret_cast<float4 *>(stg + row * 512 + ((un ^ (row & 31)) << 4)) = o;
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&sm.tmem_empty[s]);          // accumulator drained
            named_bar_sync(1, 128);
            // coalesced row stores (each warp: 32 rows, one 512-byte row per instruction)
            const int wr = warp - 5;
            for (int r = wr * 32; r < wr * 32 + 32; ++r) {
                const int m = m0 + r;
                if (m >= p.M) break;
                float4 *sp = reinterpret_cast<float4 *>(stg + r * 512 + ((lane ^ (r & 31)) << 4));
                float4 o = *sp;
                float4 *gp = reinterpret_cast<float4 *>(p.Y + (size_t)m * p.ldy + lane * 4);
                if (p.accumulate) {
                    const float4 e = *gp;
                    o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
                    *sp = o;
                }
                *gp = o;
            }
            if (p.sum) {
                if (p.accumulate) named_bar_sync(1, 128);
                float cs = 0.f, cq = 0.f;
                const int rows = min(kBM, p.M - m0);
                for (int r = 0; r < rows; ++r) {
                    const float x = *reinterpret_cast<const float *>(stg + r * 512 + (((et >> 2) ^ (r & 31)) << 4) + (et & 3) * 4);
                    cs += x; cq = fmaf(x, x, cq);
                }
                dsum += (double)cs; dsq += (double)cq;
            }
            mbar_arrive(&sm.empty[s]);               // staging consumed: the producers may refill this stage
        }
        if (p.sum) { atomicAdd(&p.sum[et], dsum); atomicAdd(&p.sumsq[et], dsq); }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

// Host-side launcher shared by the decoder orchestration and the unit-test entry point.
int launch_tc_fwd(const float *X, int ldx, int K, const float *a, const float *b, const float *W, int ldw, const float *bias, float *Y,
                  int ldy, int accumulate, double *sum, double *sumsq, int M, cudaStream_t st)
{
    GA_REQUIRE(K % 8 == 0 && K >= 8 && K <= 128, "tcgen05 layer: K=%d must be a multiple of 8 in [8,128]", K);
    GA_REQUIRE(ldx % 4 == 0 && ldw % 4 == 0 && ldy % 4 == 0, "tcgen05 layer: leading dimensions must be multiples of 4");
    static bool attr_set = false;
    if (!attr_set) {
        GA_CHECK_CUDA(cudaFuncSetAttribute(tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcFwdSmem) + 1024));
        attr_set = true;
    }
    TcFwdParams p{X, ldx, K, a, b, W, ldw, bias, Y, ldy, accumulate, sum, sumsq, M};
    const int tiles = cdiv(M, kBM);
    const int grid = tiles < kNumSMs ? tiles : kNumSMs;
    {
        ProfScope _ps("mlp_tc_fwd", st);
        tc_fwd_kernel<<<grid, kTcThreads, sizeof(TcFwdSmem) + 1024, st>>>(p);
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
