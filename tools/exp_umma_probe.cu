// Experiment (not part of the product): decode the physical shared-memory layout tcgen05.mma kind::tf32 reads for an MN-major operand.
// One word of the operand region is set to 1.0 at a time; the other operand (K-major, known layout) holds k+1, so D[m][0] reveals (m, k).
//   usage: exp_umma_probe <operand 0=A 1=B> <lbo> <sbo> <ltype> [start_off_bytes]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"
using namespace ga::tc;

__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo, int ltype)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)ltype << 61;
    return d;
}

constexpr int kWords = 8192;       // 32 KB probe region

__global__ void __launch_bounds__(128) probe_kernel(int operand, uint32_t lbo, uint32_t sbo, int ltype, uint32_t start_off, int *map)
{
    extern __shared__ unsigned char raw[];
    unsigned char *base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    float *probe = reinterpret_cast<float *>(base);              // 32 KB
    unsigned char *other = base + kWords * 4;                     // 16 KB: K-major SW128 [128 rows][32 k], value k + 1 (k < 8)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 128);
    for (int i = tid; i < kWords; i += 128) probe[i] = 0.f;
    for (int idx = tid; idx < 128 * 32; idx += 128) {
        const int r = idx / 32, k = idx % 32;
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)k * 4u;
        *reinterpret_cast<float *>(other + (off ^ (((off >> 7) & 7u) << 4))) = (float)(k + 1);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = make_idesc_tf32(128, 128, operand == 0, operand == 1);      // operand 2 / 3: K-major A / B probed with the given layout type
    const uint64_t dprobe = mkdesc(smem_u32(probe) + start_off, lbo, sbo, ltype), dother = mkdesc(smem_u32(other), 16, 1024, 2);
    for (int w = 0; w < kWords; ++w) {
        if (tid == 0) { if (w) probe[w - 1] = 0.f; probe[w] = 1.f; }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
        if (tid == 0) {
            if (operand == 0 || operand == 2) mma_tf32(tmem_base, dprobe, dother, idesc, false); else mma_tf32(tmem_base, dother, dprobe, idesc, false);
            mma_commit(&bar);
        }
        mbar_wait(&bar, w & 1);
        tc_fence_after_sync();
        float v[16];
        tmem_ld_32x16(tmem_base + ((uint32_t)(warp * 32) << 16), v);
        // operand A probe: row m = this thread, D[m][n] = k+1 for every n.  operand B probe: column n0 holds A(m,k)=k+1 for every m -> thread 0 scans
        if (operand == 0 || operand == 2) { if (v[0] != 0.f) map[w] = (warp * 32 + lane) * 16 + (int)v[0] - 1; }
        else if (warp == 0) {      // B probe: column n0 of every row holds k0 + 1; the whole warp loads (aligned instruction), lane 0 (row 0) scans
            for (int cc = 0; cc < 8; ++cc) {
                float u[16];
                tmem_ld_32x16(tmem_base + cc * 16, u);
                if (lane == 0) for (int j = 0; j < 16; ++j) if (u[j] != 0.f) map[w] = (cc * 16 + j) * 16 + (int)u[j] - 1;
            }
        }
        tc_fence_before_sync();
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

int main(int argc, char **argv)
{
    const int operand = atoi(argv[1]); const uint32_t lbo = atoi(argv[2]), sbo = atoi(argv[3]); const int ltype = atoi(argv[4]);
    const uint32_t start = argc > 5 ? atoi(argv[5]) : 0;
    int *map; cudaMalloc(&map, kWords * 4); cudaMemset(map, 0xff, kWords * 4);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
    probe_kernel<<<1, 128, 49 * 1024 + 1024>>>(operand, lbo, sbo, ltype, start, map);
    cudaError_t e = cudaDeviceSynchronize();
    printf("# operand %d lbo %u sbo %u ltype %d start %u : %s\n", operand, lbo, sbo, ltype, start, cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<int> h(kWords); cudaMemcpy(h.data(), map, kWords * 4, cudaMemcpyDeviceToHost);
    // inverse: for (k, m) -> byte offset
    std::vector<int> inv(8 * 128, -1); int hits = 0;
    for (int w = 0; w < kWords; ++w) if (h[w] >= 0) { const int m = h[w] / 16, k = h[w] % 16; if (k < 8 && m < 128) { inv[k * 128 + m] = w * 4; ++hits; } }
    printf("# hits %d\n", hits);
    for (int k = 0; k < 8; ++k) { printf("k%d:", k); for (int m = 0; m < 128; ++m) printf(" %x", inv[k * 128 + m]); printf("\n"); }
    return 0;
}
