// Experiment (not part of the product): which physical shared-memory layouts / descriptor settings does tcgen05.mma kind::tf32 accept
// for MN-major operands, and does a K-major operand work with the 32-byte-base 128 B swizzle?  Prints one line per combination.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I gaussianavatar_b200/csrc -o tools/bin/exp_umma_layouts tools/exp_umma_layouts.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"
using namespace ga::tc;

struct Cfg { int mode, phys, swz, swap, ltype; };

__device__ __forceinline__ uint32_t swizzle(uint32_t off, int swz)
{
    if (swz == 1) return off ^ (((off >> 7) & 7u) << 4);      // Swizzle<3,4,3>
    if (swz == 2) return off ^ (((off >> 7) & 3u) << 5);      // Swizzle<2,5,2>
    if (swz == 3) return off ^ (((off >> 7) & 7u) << 5) ;     // Swizzle<3,5,2> (probe)
    return off;
}
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
__host__ __device__ inline float aval(int k, int m) { return (float)(((k * 7 + m * 3) % 5) - 2); }
__host__ __device__ inline float bval(int n, int k) { return (float)(((n * 5 + k * 3) % 7) - 3); }

// D[m][n] = sum_k A(k,m) * B(n,k), M = N = 128, K = 32
__global__ void __launch_bounds__(128) exp_kernel(Cfg c, float *out)
{
    extern __shared__ unsigned char raw[];
    unsigned char *base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    unsigned char *sa = base, *sb = base + 16384;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 128);
    // strides of the MN-major arrangement: j = 32-element chunk along M/N (4 of them), kg = group of 8 K rows (4 of them)
    const uint32_t LBOp = c.phys == 0 ? 4096u : 1024u, SBOp = c.phys == 0 ? 1024u : 4096u;
    for (int idx = tid; idx < 32 * 128; idx += 128) {
        const int k = idx / 128, x = idx % 128;          // x = m (A) or n (B)
        // MN-major image of the "activation" operand: logical [k][x]
        const uint32_t off = (uint32_t)(x >> 5) * LBOp + (uint32_t)(k >> 3) * SBOp + (uint32_t)(k & 7) * 128u + (uint32_t)(x & 31) * 4u;
        // K-major image of the same / the other operand: rows = x, 128 B = 32 k
        const uint32_t offk = (uint32_t)x * 128u + (uint32_t)k * 4u;
        if (c.mode == 0) {          // A MN-major (probe), B K-major (known good)
            *reinterpret_cast<float *>(sa + swizzle(off, c.swz)) = aval(k, x);
            *reinterpret_cast<float *>(sb + swizzle(offk, 1)) = bval(x, k);
        } else if (c.mode == 1) {   // A K-major (known good), B MN-major (probe)
            *reinterpret_cast<float *>(sa + swizzle(offk, 1)) = aval(k, x);
            *reinterpret_cast<float *>(sb + swizzle(off, c.swz)) = bval(x, k);
        } else {                    // both K-major; B written with swizzle c.swz and described with layout type c.ltype
            *reinterpret_cast<float *>(sa + swizzle(offk, 1)) = aval(k, x);
            *reinterpret_cast<float *>(sb + swizzle(offk, c.swz)) = bval(x, k);
        }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(128, 128, c.mode == 0, c.mode == 1);
        for (int kg = 0; kg < 4; ++kg) {
            const uint32_t lbo_f = c.swap ? SBOp : LBOp, sbo_f = c.swap ? LBOp : SBOp;
            const uint64_t kmaj_a = mkdesc(smem_u32(sa) + kg * 32, 16, 1024, 2), kmaj_b = mkdesc(smem_u32(sb) + kg * 32, 16, 1024, 2);
            uint64_t da, db;
            if (c.mode == 0) { da = mkdesc(smem_u32(sa) + kg * SBOp, lbo_f, sbo_f, c.ltype); db = kmaj_b; }
            else if (c.mode == 1) { da = kmaj_a; db = mkdesc(smem_u32(sb) + kg * SBOp, lbo_f, sbo_f, c.ltype); }
            else { da = kmaj_a; db = mkdesc(smem_u32(sb) + kg * 32, 16, 1024, c.ltype); }
            mma_tf32(tmem_base, da, db, idesc, kg > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    for (int cc = 0; cc < 4; ++cc) {
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + cc * 32, v);
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 128 + cc * 32 + j] = v[j];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

int main(int argc, char **argv)
{
    const int start = argc > 1 ? atoi(argv[1]) : 0;
    float *out; cudaMalloc(&out, 128 * 128 * 4);
    std::vector<float> h(128 * 128), ref(128 * 128);
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) { float s = 0; for (int k = 0; k < 32; ++k) s += aval(k, m) * bval(n, k); ref[m * 128 + n] = s; }
    cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024);
    std::vector<Cfg> cfgs;
    for (int mode = 0; mode < 2; ++mode) for (int phys = 0; phys < 2; ++phys) for (int swz = 0; swz < 4; ++swz) for (int swap = 0; swap < 2; ++swap)
        for (int lt : {1, 2, 0}) cfgs.push_back({mode, phys, swz, swap, lt});
    for (int swz = 0; swz < 4; ++swz) for (int lt : {1, 2, 0, 4, 6}) cfgs.push_back({2, 0, swz, 0, lt});
    for (int ci = start; ci < (int)cfgs.size(); ++ci) {
        const Cfg &c = cfgs[ci];
        cudaMemset(out, 0xff, 128 * 128 * 4);
        exp_kernel<<<1, 128, 33 * 1024>>>(c, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d phys %d swz %d swap %d ltype %d : CUDA ERROR %s\nRESTART %d\n", c.mode, c.phys, c.swz, c.swap, c.ltype, cudaGetErrorString(e), ci + 1); return 2; }
        cudaMemcpy(h.data(), out, 128 * 128 * 4, cudaMemcpyDeviceToHost);
        int bad = 0; for (int i = 0; i < 128 * 128; ++i) bad += (h[i] != ref[i]);
        printf("mode %d phys %d swz %d swap %d ltype %d : %s (%d mismatches)\n", c.mode, c.phys, c.swz, c.swap, c.ltype, bad ? "no" : "MATCH", bad);
    }
    printf("RESTART -1\n");
    return 0;
}
