// Differentiable 3D-Gaussian tile rasterizer for sm_100a — forward (K1..K6) and backward (K7, fused K8/K9).
//
// Replaces what the reference reaches through `GaussianRasterizer(raster_settings)(...)`
// (/root/reference gaussian_renderer/__init__.py:36-48; [UPSTREAM] diff-gaussian-rasterization, not vendored).
// Algorithm and constants: SURVEY.md §8 "a-8 forward spec" / "a-9 backward spec".  Written from that specification,
// not from upstream source.
//
// Arithmetic contract: every value that feeds an INTEGER decision (depth key bits, radius, tile rect)
// is computed with explicit round-to-nearest intrinsics in exactly the operation order of oracle/raster_oracle_impl.inc,
// so keys / radii / ranges are bit-exact against the CPU oracle regardless of nvcc's FMA contraction.
#include <cub/cub.cuh>

#include "common.cuh"

namespace ga {
namespace {

constexpr int kTile = 16;
constexpr int kBlock = kTile * kTile;

struct GeomViews {
    float *depth;
    float2 *xy;
    float4 *conic_o;
    float *cov3d;
    uint32_t *tiles;
    uint32_t *offsets;
    ushort4 *rect;
    void *scan_temp;
    size_t scan_temp_bytes;
    size_t total;
};

size_t scan_temp_bytes_for(int P)
{
    size_t b = 0;
    cub::DeviceScan::InclusiveSum(nullptr, b, (uint32_t *)nullptr, (uint32_t *)nullptr, P);
    return b;
}

GeomViews carve_geom(void *buf, int P)
{
    Carver c(buf);
    GeomViews g;
    const size_t n = P > 0 ? (size_t)P : 1;
    g.depth = c.take<float>(n);
    g.xy = c.take<float2>(n);
    g.conic_o = c.take<float4>(n);
    g.cov3d = c.take<float>(6 * n);
    g.tiles = c.take<uint32_t>(n);
    g.offsets = c.take<uint32_t>(n);
    g.rect = c.take<ushort4>(n);
    g.scan_temp_bytes = scan_temp_bytes_for(P > 0 ? P : 1);
    g.scan_temp = c.take<char>(g.scan_temp_bytes);
    g.total = c.used();
    return g;
}

struct ImgViews {
    float *final_T;
    uint32_t *n_contrib;
    uint2 *ranges;
    size_t total;
};

ImgViews carve_img(void *buf, int H, int W)
{
    Carver c(buf);
    ImgViews v;
    const size_t hw = (size_t)H * W, T = (size_t)cdiv(W, kTile) * cdiv(H, kTile);
    v.final_T = c.take<float>(hw ? hw : 1);
    v.n_contrib = c.take<uint32_t>(hw ? hw : 1);
    v.ranges = c.take<uint2>(T ? T : 1);
    v.total = c.used();
    return v;
}

struct BinViews {
    uint64_t *keys_unsorted, *keys;
    uint32_t *vals_unsorted, *vals;
    void *sort_temp;
    size_t sort_temp_bytes;
    size_t total;
};

int sort_end_bit(int H, int W)
{
    const int T = cdiv(W, kTile) * cdiv(H, kTile);
    int bits = 0;
    while ((1 << bits) < T) ++bits;  // bits to represent tile ids 0..T-1
    return 32 + (bits > 0 ? bits : 1);
}

BinViews carve_bin(void *buf, int64_t R, int H, int W)
{
    Carver c(buf);
    BinViews b;
    const size_t n = R > 0 ? (size_t)R : 1;
    b.keys_unsorted = c.take<uint64_t>(n);
    b.keys = c.take<uint64_t>(n);
    b.vals_unsorted = c.take<uint32_t>(n);
    b.vals = c.take<uint32_t>(n);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (uint64_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr,
                                    (uint32_t *)nullptr, (int)n, 0, sort_end_bit(H, W));
    b.sort_temp_bytes = tb;
    b.sort_temp = c.take<char>(tb);
    b.total = c.used();
    return b;
}

// ---- contraction-proof arithmetic (mirrors oracle/raster_oracle_impl.inc op for op) ----
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float div_(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ float xf_row(const float *__restrict__ M, int r, float x, float y, float z)
{
    float t = mul_(M[r], x);
    t = fma_(M[4 + r], y, t);
    t = fma_(M[8 + r], z, t);
    return add_(t, M[12 + r]);
}

struct Cov3 {
    float v[6];
};

__device__ __forceinline__ void quat_to_rot(float qr, float qx, float qy, float qz, float R[3][3])
{
    R[0][0] = sub_(1.f, mul_(2.f, add_(mul_(qy, qy), mul_(qz, qz))));
    R[0][1] = mul_(2.f, sub_(mul_(qx, qy), mul_(qr, qz)));
    R[0][2] = mul_(2.f, add_(mul_(qx, qz), mul_(qr, qy)));
    R[1][0] = mul_(2.f, add_(mul_(qx, qy), mul_(qr, qz)));
    R[1][1] = sub_(1.f, mul_(2.f, add_(mul_(qx, qx), mul_(qz, qz))));
    R[1][2] = mul_(2.f, sub_(mul_(qy, qz), mul_(qr, qx)));
    R[2][0] = mul_(2.f, sub_(mul_(qx, qz), mul_(qr, qy)));
    R[2][1] = mul_(2.f, add_(mul_(qy, qz), mul_(qr, qx)));
    R[2][2] = sub_(1.f, mul_(2.f, add_(mul_(qx, qx), mul_(qy, qy))));
}

// K1: one thread per Gaussian.
__global__ void __launch_bounds__(256)
preprocess_fwd_kernel(int P, int H, int W, int gx, int gy, float tanfovx, float tanfovy, float mod,
                      const float *__restrict__ means3D, const float *__restrict__ scales,
                      const float *__restrict__ rots, const float *__restrict__ opac,
                      const float *__restrict__ view_g, const float *__restrict__ proj_g, float *__restrict__ depth,
                      float2 *__restrict__ xy, float4 *__restrict__ conic_o, float *__restrict__ cov3d,
                      uint32_t *__restrict__ tiles, ushort4 *__restrict__ rect, int32_t *__restrict__ radii)
{
    __shared__ float view[16], proj[16];
    if (threadIdx.x < 16) view[threadIdx.x] = view_g[threadIdx.x];
    else if (threadIdx.x < 32) proj[threadIdx.x - 16] = proj_g[threadIdx.x - 16];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;

    radii[i] = 0;
    tiles[i] = 0;
    depth[i] = 0.f;
    xy[i] = make_float2(0.f, 0.f);
    conic_o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    rect[i] = make_ushort4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < 6; ++k) cov3d[6 * (size_t)i + k] = 0.f;

    const float px = means3D[3 * (size_t)i], py = means3D[3 * (size_t)i + 1], pz = means3D[3 * (size_t)i + 2];
    const float tvx = xf_row(view, 0, px, py, pz), tvy = xf_row(view, 1, px, py, pz), tvz = xf_row(view, 2, px, py, pz);
    if (tvz <= 0.2f) return;

    const float hx = xf_row(proj, 0, px, py, pz), hy = xf_row(proj, 1, px, py, pz), hw = xf_row(proj, 3, px, py, pz);
    const float p_w = div_(1.f, add_(hw, 0.0000001f));
    const float ndc_x = mul_(hx, p_w), ndc_y = mul_(hy, p_w);

    const float sv[3] = {mul_(mod, scales[3 * (size_t)i]), mul_(mod, scales[3 * (size_t)i + 1]), mul_(mod, scales[3 * (size_t)i + 2])};
    float Rm[3][3];
    quat_to_rot(rots[4 * (size_t)i], rots[4 * (size_t)i + 1], rots[4 * (size_t)i + 2], rots[4 * (size_t)i + 3], Rm);
    float Mm[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k) Mm[a][k] = mul_(sv[a], Rm[k][a]);
    float Sg[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int k = j; k < 3; ++k) {
            float t = mul_(Mm[0][j], Mm[0][k]);
            t = fma_(Mm[1][j], Mm[1][k], t);
            t = fma_(Mm[2][j], Mm[2][k], t);
            Sg[j][k] = t;
        }
    const float S00 = Sg[0][0], S01 = Sg[0][1], S02 = Sg[0][2], S11 = Sg[1][1], S12 = Sg[1][2], S22 = Sg[2][2];

    const float focal_x = div_((float)W, mul_(2.f, tanfovx)), focal_y = div_((float)H, mul_(2.f, tanfovy));
    const float limx = mul_(1.3f, tanfovx), limy = mul_(1.3f, tanfovy);
    const float txtz = div_(tvx, tvz), tytz = div_(tvy, tvz);
    const float tx = mul_(fminf(limx, fmaxf(-limx, txtz)), tvz);
    const float ty = mul_(fminf(limy, fmaxf(-limy, tytz)), tvz);
    const float tz2 = mul_(tvz, tvz);
    const float J00 = div_(focal_x, tvz), J02 = div_(-mul_(focal_x, tx), tz2);
    const float J11 = div_(focal_y, tvz), J12 = div_(-mul_(focal_y, ty), tz2);
    float m0[3], m1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        m0[k] = fma_(view[k * 4 + 2], J02, mul_(view[k * 4 + 0], J00));
        m1[k] = fma_(view[k * 4 + 2], J12, mul_(view[k * 4 + 1], J11));
    }
    const float v00 = fma_(S02, m0[2], fma_(S01, m0[1], mul_(S00, m0[0])));
    const float v01 = fma_(S12, m0[2], fma_(S11, m0[1], mul_(S01, m0[0])));
    const float v02 = fma_(S22, m0[2], fma_(S12, m0[1], mul_(S02, m0[0])));
    const float v10 = fma_(S02, m1[2], fma_(S01, m1[1], mul_(S00, m1[0])));
    const float v11 = fma_(S12, m1[2], fma_(S11, m1[1], mul_(S01, m1[0])));
    const float v12 = fma_(S22, m1[2], fma_(S12, m1[1], mul_(S02, m1[0])));
    const float ca = add_(fma_(m0[2], v02, fma_(m0[1], v01, mul_(m0[0], v00))), 0.3f);
    const float cb = fma_(m0[2], v12, fma_(m0[1], v11, mul_(m0[0], v10)));
    const float cc = add_(fma_(m1[2], v12, fma_(m1[1], v11, mul_(m1[0], v10))), 0.3f);

    const float det = sub_(mul_(ca, cc), mul_(cb, cb));
    if (det == 0.f) return;
    const float det_inv = div_(1.f, det);
    const float conx = mul_(cc, det_inv), cony = mul_(-cb, det_inv), conz = mul_(ca, det_inv);
    const float mid = mul_(0.5f, add_(ca, cc));
    const float sq = __fsqrt_rn(fmaxf(0.1f, sub_(mul_(mid, mid), det)));
    const float lambda1 = add_(mid, sq), lambda2 = sub_(mid, sq);
    const int radius = (int)ceilf(mul_(3.f, __fsqrt_rn(fmaxf(lambda1, lambda2))));

    const float pixx = (float)__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn((double)ndc_x, 1.0), (double)W), -1.0), 0.5);
    const float pixy = (float)__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn((double)ndc_y, 1.0), (double)H), -1.0), 0.5);

    const float rf = (float)radius;
    int rminx = __float2int_rz(div_(sub_(pixx, rf), 16.f)); rminx = min(gx, max(0, rminx));
    int rminy = __float2int_rz(div_(sub_(pixy, rf), 16.f)); rminy = min(gy, max(0, rminy));
    int rmaxx = __float2int_rz(div_(add_(add_(pixx, rf), 15.f), 16.f)); rmaxx = min(gx, max(0, rmaxx));
    int rmaxy = __float2int_rz(div_(add_(add_(pixy, rf), 15.f), 16.f)); rmaxy = min(gy, max(0, rmaxy));
    const int area = (rmaxx - rminx) * (rmaxy - rminy);
    if (area == 0) return;

    float *c3 = cov3d + 6 * (size_t)i;
    c3[0] = S00; c3[1] = S01; c3[2] = S02; c3[3] = S11; c3[4] = S12; c3[5] = S22;
    depth[i] = tvz;
    radii[i] = radius;
    xy[i] = make_float2(pixx, pixy);
    conic_o[i] = make_float4(conx, cony, conz, opac[i]);
    tiles[i] = (uint32_t)area;
    rect[i] = make_ushort4((unsigned short)rminx, (unsigned short)rminy, (unsigned short)rmaxx, (unsigned short)rmaxy);
}

// K3: one thread per Gaussian writes its (tile<<32 | depth bits, index) instances.
__global__ void __launch_bounds__(256)
duplicate_with_keys_kernel(int P, int gx, const float *__restrict__ depth, const uint32_t *__restrict__ offsets,
                           const uint32_t *__restrict__ tiles, const ushort4 *__restrict__ rect,
                           uint64_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    if (tiles[i] == 0) return;
    uint32_t off = (i == 0) ? 0u : offsets[i - 1];
    const uint32_t dbits = __float_as_uint(depth[i]);
    const ushort4 r = rect[i];
    for (int y = r.y; y < r.w; ++y)
        for (int x = r.x; x < r.z; ++x) {
            const uint64_t key = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
            keys[off] = key;
            vals[off] = (uint32_t)i;
            ++off;
        }
}

// K5: [start,end) of every tile in the sorted list.
__global__ void __launch_bounds__(256)
tile_ranges_kernel(int64_t R, const uint64_t *__restrict__ keys, uint2 *__restrict__ ranges)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= R) return;
    const uint32_t t = (uint32_t)(keys[k] >> 32);
    if (k == 0) ranges[t].x = 0;
    else {
        const uint32_t pt = (uint32_t)(keys[k - 1] >> 32);
        if (pt != t) {
            ranges[pt].y = (uint32_t)k;
            ranges[t].x = (uint32_t)k;
        }
    }
    if (k == R - 1) ranges[t].y = (uint32_t)R;
}

// Squared distance from a Gaussian's centre beyond which it is certain that the compositing loop skips it
// (alpha = min(0.99, o exp(power)) < 1/255):  power <= -0.5 lmin |d|^2 with lmin the smaller eigenvalue of the conic.
// Conservative by construction (1 % slack on alpha, lmin pushed down by its round-off, 0.1 % on the radius), so culling a
// (Gaussian, sub-tile) pair never changes a result: every pair that could pass the exact per-pixel tests is still
// evaluated with the exact arithmetic.  Returns -1 (always culled) for opacity below the threshold, +inf (never culled) for
// an indefinite or NaN conic.
__device__ __forceinline__ float cull_radius2(const float4 co)
{
    if (co.w <= 0.f) return -1.f;
    const float thr = logf(255.f * co.w) + 0.01f;
    if (thr <= 0.f) return -1.f;
    const float mid = 0.5f * (co.x + co.z), hd = 0.5f * (co.x - co.z);
    const float lmin = mid - sqrtf(hd * hd + co.y * co.y) - 1e-5f * fabsf(mid);
    if (!(lmin > 0.f)) return INFINITY;
    return 2.002f * thr / lmin;
}

// Squared distance from point c to the pixel rectangle [x0,x1] x [y0,y1] (0 inside).
__device__ __forceinline__ float rect_dist2(const float2 c, float x0, float x1, float y0, float y1)
{
    const float ex = fmaxf(fmaxf(x0 - c.x, c.x - x1), 0.f), ey = fmaxf(fmaxf(y0 - c.y, c.y - y1), 0.f);
    return ex * ex + ey * ey;
}

// K6: one CTA per 16x16 tile, one thread per pixel; the tile's depth-ordered list is staged through shared memory 256
// entries at a time.  Each warp owns an 8x4-pixel sub-tile: 32 list entries are tested against the sub-tile in parallel
// (one per lane, cull_radius2) and only the survivors are walked by the per-pixel loop, in list order.
__global__ void __launch_bounds__(kBlock)
render_fwd_kernel(int H, int W, const uint2 *__restrict__ ranges, const uint32_t *__restrict__ point_list,
                  const float2 *__restrict__ xy, const float4 *__restrict__ conic_o, const float *__restrict__ colors,
                  const float *__restrict__ bg, float *__restrict__ final_T, uint32_t *__restrict__ n_contrib,
                  float *__restrict__ out)
{
    // two staging buffers: the next 256 entries are fetched into registers while the current ones are composited and stored
    // before the round's single barrier, so the gather latency hides behind the compositing
    __shared__ float2 s_xy[2][kBlock];
    __shared__ float4 s_co[2][kBlock];
    __shared__ float s_rgb[2][3][kBlock];
    __shared__ float s_rc2[2][kBlock];

    const int tid = threadIdx.y * kTile + threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int sx = blockIdx.x * kTile + (warp & 1) * 8, sy = blockIdx.y * kTile + (warp >> 1) * 4;
    const int px = sx + (lane & 7), py = sy + (lane >> 3);
    const float x0f = (float)sx, x1f = (float)(sx + 7), y0f = (float)sy, y1f = (float)(sy + 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
    int todo = (int)(range.y - range.x);
    const int rounds = (todo + kBlock - 1) / kBlock;

    bool done = !inside;
    float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t last = 0;

    // entry of this thread in the batch being fetched
    float2 f_xy = make_float2(0.f, 0.f); float4 f_co = make_float4(0.f, 0.f, 0.f, 0.f); float f_r = 0.f, f_g = 0.f, f_b = 0.f;
    auto fetch = [&](int r) {
        const uint32_t i = range.x + (uint32_t)r * kBlock + tid;
        if (i < range.y) {
            const uint32_t g = point_list[i];
            f_xy = xy[g]; f_co = conic_o[g];
            f_r = colors[3 * (size_t)g]; f_g = colors[3 * (size_t)g + 1]; f_b = colors[3 * (size_t)g + 2];
        }
    };
    auto stage = [&](int b) {
        s_xy[b][tid] = f_xy; s_co[b][tid] = f_co;
        s_rgb[b][0][tid] = f_r; s_rgb[b][1][tid] = f_g; s_rgb[b][2][tid] = f_b;
        s_rc2[b][tid] = cull_radius2(f_co);
    };
    if (rounds > 0) { fetch(0); stage(0); }
    __syncthreads();

    for (int r = 0; r < rounds; ++r, todo -= kBlock) {
        const int b = r & 1;
        const bool more = r + 1 < rounds;
        if (more) fetch(r + 1);
        const int n = min(kBlock, todo);
        if (!__all_sync(0xffffffffu, done)) {
            for (int j0 = 0; j0 < n; j0 += 32) {
                const int jt = j0 + lane;
                bool hit = false;
                if (jt < n) hit = !(rect_dist2(s_xy[b][jt], x0f, x1f, y0f, y1f) > s_rc2[b][jt]);
                uint32_t mask = __ballot_sync(0xffffffffu, hit);
                while (mask) {
                    const int j = j0 + __ffs(mask) - 1;
                    mask &= mask - 1;
                    if (done) continue;
                    const float2 c = s_xy[b][j];
                    const float dx = c.x - pxf, dy = c.y - pyf;
                    const float4 co = s_co[b][j];
                    const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
                    if (power > 0.f) continue;
                    const float alpha = fminf(0.99f, co.w * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = T * (1.f - alpha);
                    if (test_T < 0.0001f) { done = true; continue; }
                    const float w = alpha * T;
                    C0 += s_rgb[b][0][j] * w; C1 += s_rgb[b][1][j] * w; C2 += s_rgb[b][2][j] * w;
                    T = test_T;
                    last = (uint32_t)(r * kBlock + j + 1);
                }
                if (__all_sync(0xffffffffu, done)) break;
            }
        }
        if (more) stage(b ^ 1);          // buffer b^1 was last read in round r-1: every warp has passed that round's barrier
        if (__syncthreads_count(done) == kBlock) break;
    }
    if (inside) {
        const size_t pid = (size_t)py * W + px, HW = (size_t)H * W;
        final_T[pid] = T;
        n_contrib[pid] = last;
        out[pid] = C0 + T * bg[0];
        out[HW + pid] = C1 + T * bg[1];
        out[2 * HW + pid] = C2 + T * bg[2];
    }
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Reduce 8 per-lane values across the warp with 9 shuffles (recursive halving: after the xor-16 / 8 / 4 steps a lane
// holds ONE of the eight sums over 8 lanes, the xor-2 / 1 steps finish it).  Returns the sum of value
// `idx = 4*bit4(lane) + 2*bit3(lane) + bit2(lane)` over all 32 lanes (replicated over the lane's 4-lane group).
__device__ __forceinline__ float warp_reduce8(float (&v)[8], int lane)
{
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = h4 ? v[i] : v[i + 4];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
        v[i] = (h4 ? v[i + 4] : v[i]) + recv;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = h3 ? v[i] : v[i + 2];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
        v[i] = (h3 ? v[i + 2] : v[i]) + recv;
    }
    {
        const float send = h2 ? v[0] : v[1];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
        v[0] = (h2 ? v[1] : v[0]) + recv;
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    return v[0];
}

// K7: per pixel, back to front.  Per-Gaussian gradients are reduced warp-wide with shuffles (9 per Gaussian for the 8
// components GaussianAvatar consumes), then across the CTA's eight warps in shared memory, so each (tile, Gaussian) pair
// issues at most one global atomic per component.  kOpacity adds dL/dopacity (API completeness; the avatar's opacity is
// a constant without gradient, model/avatar_model.py:80).
#ifndef GA_RENDER_BWD_MINB
#define GA_RENDER_BWD_MINB 1
#endif
template <bool kOpacity>
__global__ void __launch_bounds__(kBlock, GA_RENDER_BWD_MINB)
render_bwd_kernel(int H, int W, const uint2 *__restrict__ ranges, const uint32_t *__restrict__ point_list,
                  const float *__restrict__ bg, const float2 *__restrict__ xy, const float4 *__restrict__ conic_o,
                  const float *__restrict__ colors, const float *__restrict__ final_T,
                  const uint32_t *__restrict__ n_contrib, const float *__restrict__ dL_dout,
                  float2 *__restrict__ d_mean2D, float4 *__restrict__ d_conic_op, float *__restrict__ d_colors)
{
    // staging and accumulators are double-buffered: round r composites buffer r & 1 while the entries of round r + 1 sit in
    // registers (stored before the round's single barrier) and the previous round's accumulators are flushed after it
    __shared__ uint32_t s_id[2][kBlock];
    __shared__ float2 s_xy[2][kBlock];
    __shared__ float4 s_co[2][kBlock];
    __shared__ float s_rgb[2][3][kBlock];
    __shared__ float s_acc[2][9][kBlock];
    __shared__ float s_rc2[2][kBlock];
    __shared__ uint32_t s_max[kBlock / 32];

    const int tid = threadIdx.y * kTile + threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int sx = blockIdx.x * kTile + (warp & 1) * 8, sy = blockIdx.y * kTile + (warp >> 1) * 4;      // 8x4 sub-tile per warp
    const int px = sx + (lane & 7), py = sy + (lane >> 3);
    const float x0f = (float)sx, x1f = (float)(sx + 7), y0f = (float)sy, y1f = (float)(sy + 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
    if (range.y <= range.x) return;  // uniform for the CTA

    const size_t pid = (size_t)py * W + px, HW = (size_t)H * W;
    const float T_final = inside ? final_T[pid] : 0.f;
    const uint32_t last_contributor = inside ? n_contrib[pid] : 0u;
    float dpix0 = 0.f, dpix1 = 0.f, dpix2 = 0.f;
    if (inside) { dpix0 = dL_dout[pid]; dpix1 = dL_dout[HW + pid]; dpix2 = dL_dout[2 * HW + pid]; }
    const float bg_dot = bg[0] * dpix0 + bg[1] * dpix1 + bg[2] * dpix2;

    // nothing past the furthest blended entry of any pixel of this tile contributes
    uint32_t m = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_max[warp] = m;
#pragma unroll
    for (int k = 0; k < 9; ++k) { s_acc[0][k][tid] = 0.f; s_acc[1][k][tid] = 0.f; }
    __syncthreads();
    uint32_t todo = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) todo = max(todo, s_max[w]);
    if (todo == 0) return;

    float T = T_final;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;
    const float ddelx_dx = 0.5f * (float)W, ddely_dy = 0.5f * (float)H;
    const int rounds = (int)((todo + kBlock - 1) / kBlock);
    const int vidx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);   // component this lane ends up owning

    uint32_t f_id = 0; float2 f_xy = make_float2(0.f, 0.f); float4 f_co = make_float4(0.f, 0.f, 0.f, 0.f); float f_r = 0.f, f_g = 0.f, f_b = 0.f;
    auto fetch = [&](int r) {
        const uint32_t progress = (uint32_t)r * kBlock + tid;
        if (progress < todo) {
            f_id = point_list[range.x + todo - 1 - progress];
            f_xy = xy[f_id]; f_co = conic_o[f_id];
            f_r = colors[3 * (size_t)f_id]; f_g = colors[3 * (size_t)f_id + 1]; f_b = colors[3 * (size_t)f_id + 2];
        }
    };
    auto stage = [&](int b) {
        s_id[b][tid] = f_id; s_xy[b][tid] = f_xy; s_co[b][tid] = f_co;
        s_rgb[b][0][tid] = f_r; s_rgb[b][1][tid] = f_g; s_rgb[b][2][tid] = f_b;
        s_rc2[b][tid] = cull_radius2(f_co);
    };
    fetch(0); stage(0);
    __syncthreads();

    for (int r = 0; r < rounds; ++r) {
        const int b = r & 1;
        const bool more = r + 1 < rounds;
        const bool have = (uint32_t)r * kBlock + tid < todo;
        if (more) fetch(r + 1);
        const int n = (int)min((uint32_t)kBlock, todo - (uint32_t)r * kBlock);
        for (int j0 = 0; j0 < n; j0 += 32) {
          const int jt = j0 + lane;
          bool hit = false;
          if (jt < n) hit = !(rect_dist2(s_xy[b][jt], x0f, x1f, y0f, y1f) > s_rc2[b][jt]);
          uint32_t mask = __ballot_sync(0xffffffffu, hit);
          while (mask) {
            const int j = j0 + __ffs(mask) - 1;
            mask &= mask - 1;
            const uint32_t index = todo - 1 - ((uint32_t)r * kBlock + j);  // 0-based position in the tile list
            bool active = index < last_contributor;
            float2 c; float4 co; float dx = 0.f, dy = 0.f, G = 0.f, alpha = 0.f;
            if (active) {
                c = s_xy[b][j]; co = s_co[b][j];
                dx = c.x - pxf; dy = c.y - pyf;
                const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
                if (power > 0.f) active = false;
                else {
                    G = expf(power);
                    alpha = fminf(0.99f, co.w * G);
                    if (alpha < 1.0f / 255.0f) active = false;
                }
            }
            if (!__any_sync(0xffffffffu, active)) continue;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = 0.f;
            float vo = 0.f;
            if (active) {
                T = T / (1.f - alpha);
                const float dchannel_dcolor = alpha * T;
                const float col0 = s_rgb[b][0][j], col1 = s_rgb[b][1][j], col2 = s_rgb[b][2][j];
                acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0; lc0 = col0;
                acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1; lc1 = col1;
                acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2; lc2 = col2;
                float dL_dalpha = (col0 - acc0) * dpix0 + (col1 - acc1) * dpix1 + (col2 - acc2) * dpix2;
                v[0] = dchannel_dcolor * dpix0; v[1] = dchannel_dcolor * dpix1; v[2] = dchannel_dcolor * dpix2;
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                const float dL_dG = co.w * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                const float dG_ddelx = -gdx * co.x - gdy * co.y;
                const float dG_ddely = -gdy * co.z - gdx * co.y;
                v[3] = dL_dG * dG_ddelx * ddelx_dx;
                v[4] = dL_dG * dG_ddely * ddely_dy;
                v[5] = -0.5f * gdx * dx * dL_dG;
                v[6] = -0.5f * gdx * dy * dL_dG;
                v[7] = -0.5f * gdy * dy * dL_dG;
                vo = G * dL_dalpha;
            }
            const float red = warp_reduce8(v, lane);
            if ((lane & 3) == 0) atomicAdd(&s_acc[b][vidx][j], red);
            if (kOpacity) {
                vo = warp_sum(vo);
                if (lane == 0) atomicAdd(&s_acc[b][8][j], vo);
            }
          }
        }
        if (more) stage(b ^ 1);      // buffer b^1 was last read in round r-1, whose barrier every warp has passed
        __syncthreads();
        if (have) {
            const uint32_t g = s_id[b][tid];
            float a[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) { a[k] = s_acc[b][k][tid]; s_acc[b][k][tid] = 0.f; }      // reused in round r+2, after the next barrier
            if (a[0] != 0.f) atomicAdd(&d_colors[3 * (size_t)g], a[0]);
            if (a[1] != 0.f) atomicAdd(&d_colors[3 * (size_t)g + 1], a[1]);
            if (a[2] != 0.f) atomicAdd(&d_colors[3 * (size_t)g + 2], a[2]);
            // 8- and 16-byte vector reductions (sm_90+): one L2 atomic per float2 / float4 instead of one per component
            if (a[3] != 0.f || a[4] != 0.f)
                asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(&d_mean2D[g]), "f"(a[3]), "f"(a[4]) : "memory");
            if (a[5] != 0.f || a[6] != 0.f || a[7] != 0.f || (kOpacity && a[8] != 0.f))
                red_add_v4(&d_conic_op[g].x, a[5], a[6], a[7], kOpacity ? a[8] : 0.f);
        }
    }
}

// Fused K8 (conic -> cov2D -> cov3D / mean) + K9 (mean2D -> mean3D, cov3D -> scale / rotation): one thread per Gaussian.
__global__ void __launch_bounds__(256)
preprocess_bwd_kernel(int P, int H, int W, float tanfovx, float tanfovy, float mod, const float *__restrict__ means3D,
                      const float *__restrict__ scales, const float *__restrict__ rots,
                      const float *__restrict__ view_g, const float *__restrict__ proj_g,
                      const int32_t *__restrict__ radii, const float *__restrict__ cov3d,
                      const float2 *__restrict__ d_mean2D, const float4 *__restrict__ d_conic_op,
                      float *__restrict__ d_means3D, float *__restrict__ d_scales, float *__restrict__ d_rots,
                      float *__restrict__ d_opac, float *__restrict__ d_means2D_out)
{
    __shared__ float view[16], proj[16];
    if (threadIdx.x < 16) view[threadIdx.x] = view_g[threadIdx.x];
    else if (threadIdx.x < 32) proj[threadIdx.x - 16] = proj_g[threadIdx.x - 16];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const size_t i3 = 3 * (size_t)i;
    const float4 gco = d_conic_op[i];
    const float2 g2 = d_mean2D[i];
    if (d_opac) d_opac[i] = gco.w;
    if (d_means2D_out) { d_means2D_out[i3] = g2.x; d_means2D_out[i3 + 1] = g2.y; d_means2D_out[i3 + 2] = 0.f; }
    if (radii[i] <= 0) {
        d_means3D[i3] = d_means3D[i3 + 1] = d_means3D[i3 + 2] = 0.f;
        d_scales[i3] = d_scales[i3 + 1] = d_scales[i3 + 2] = 0.f;
        if (d_rots) { d_rots[4 * (size_t)i] = d_rots[4 * (size_t)i + 1] = d_rots[4 * (size_t)i + 2] = d_rots[4 * (size_t)i + 3] = 0.f; }
        return;
    }
    const float px = means3D[i3], py = means3D[i3 + 1], pz = means3D[i3 + 2];
    const float tvx = xf_row(view, 0, px, py, pz), tvy = xf_row(view, 1, px, py, pz), tvz = xf_row(view, 2, px, py, pz);
    const float fx = (float)W / (2.f * tanfovx), fy = (float)H / (2.f * tanfovy);
    const float limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
    const float txtz = tvx / tvz, tytz = tvy / tvz;
    const float tx = fminf(limx, fmaxf(-limx, txtz)) * tvz, ty = fminf(limy, fmaxf(-limy, tytz)) * tvz;
    const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    const float tz = 1.f / tvz, tz2 = tz * tz, tz3 = tz2 * tz;
    const float J00 = fx * tz, J02 = -fx * tx * tz2, J11 = fy * tz, J12 = -fy * ty * tz2;
    float m0[3], m1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        m0[k] = view[k * 4 + 0] * J00 + view[k * 4 + 2] * J02;
        m1[k] = view[k * 4 + 1] * J11 + view[k * 4 + 2] * J12;
    }
    const float *c3 = cov3d + 6 * (size_t)i;
    const float Sg[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    float v0[3], v1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        v0[k] = Sg[k][0] * m0[0] + Sg[k][1] * m0[1] + Sg[k][2] * m0[2];
        v1[k] = Sg[k][0] * m1[0] + Sg[k][1] * m1[1] + Sg[k][2] * m1[2];
    }
    const float a = m0[0] * v0[0] + m0[1] * v0[1] + m0[2] * v0[2] + 0.3f;
    const float b = m0[0] * v1[0] + m0[1] * v1[1] + m0[2] * v1[2];
    const float c = m1[0] * v1[0] + m1[1] * v1[1] + m1[2] * v1[2] + 0.3f;
    const float gA = gco.x, gBh = gco.y, gC = gco.z;  // gBh: half-convention off-diagonal (see a-9 spec)
    const float denom = a * c - b * b;
    const float denom2inv = 1.f / (denom * denom + 0.0000001f);
    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
    if (denom2inv != 0.f) {
        dL_da = denom2inv * (-c * c * gA + 2.f * b * c * gBh + (denom - a * c) * gC);
        dL_dc = denom2inv * (-a * a * gC + 2.f * a * b * gBh + (denom - a * c) * gA);
        dL_db = denom2inv * 2.f * (b * c * gA - (denom + 2.f * b * b) * gBh + a * b * gC);
    }
    float dc[6];
    dc[0] = m0[0] * m0[0] * dL_da + m0[0] * m1[0] * dL_db + m1[0] * m1[0] * dL_dc;
    dc[3] = m0[1] * m0[1] * dL_da + m0[1] * m1[1] * dL_db + m1[1] * m1[1] * dL_dc;
    dc[5] = m0[2] * m0[2] * dL_da + m0[2] * m1[2] * dL_db + m1[2] * m1[2] * dL_dc;
    dc[1] = 2.f * m0[0] * m0[1] * dL_da + (m0[0] * m1[1] + m0[1] * m1[0]) * dL_db + 2.f * m1[0] * m1[1] * dL_dc;
    dc[2] = 2.f * m0[0] * m0[2] * dL_da + (m0[0] * m1[2] + m0[2] * m1[0]) * dL_db + 2.f * m1[0] * m1[2] * dL_dc;
    dc[4] = 2.f * m0[2] * m0[1] * dL_da + (m0[1] * m1[2] + m0[2] * m1[1]) * dL_db + 2.f * m1[1] * m1[2] * dL_dc;
    float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float dm0 = 2.f * v0[k] * dL_da + v1[k] * dL_db;
        const float dm1 = 2.f * v1[k] * dL_dc + v0[k] * dL_db;
        dJ00 += view[k * 4 + 0] * dm0; dJ02 += view[k * 4 + 2] * dm0;
        dJ11 += view[k * 4 + 1] * dm1; dJ12 += view[k * 4 + 2] * dm1;
    }
    const float dL_dtx = x_grad_mul * -fx * tz2 * dJ02;
    const float dL_dty = y_grad_mul * -fy * tz2 * dJ12;
    const float dL_dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.f * fx * tx) * tz3 * dJ02 + (2.f * fy * ty) * tz3 * dJ12;
    float dmean[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) dmean[k] = view[k * 4 + 0] * dL_dtx + view[k * 4 + 1] * dL_dty + view[k * 4 + 2] * dL_dtz;

    const float hx = xf_row(proj, 0, px, py, pz), hy = xf_row(proj, 1, px, py, pz), hw = xf_row(proj, 3, px, py, pz);
    const float m_w = 1.f / (hw + 0.0000001f);
    const float mul1 = hx * m_w * m_w, mul2 = hy * m_w * m_w;
#pragma unroll
    for (int k = 0; k < 3; ++k)
        dmean[k] += (proj[4 * k + 0] * m_w - proj[4 * k + 3] * mul1) * g2.x + (proj[4 * k + 1] * m_w - proj[4 * k + 3] * mul2) * g2.y;
    d_means3D[i3] = dmean[0]; d_means3D[i3 + 1] = dmean[1]; d_means3D[i3 + 2] = dmean[2];

    const float sv[3] = {mod * scales[i3], mod * scales[i3 + 1], mod * scales[i3 + 2]};
    const float qr = rots[4 * (size_t)i], qx = rots[4 * (size_t)i + 1], qy = rots[4 * (size_t)i + 2], qz = rots[4 * (size_t)i + 3];
    float Rm[3][3];
    quat_to_rot(qr, qx, qy, qz, Rm);
    const float dS[3][3] = {{dc[0], 0.5f * dc[1], 0.5f * dc[2]}, {0.5f * dc[1], dc[3], 0.5f * dc[4]}, {0.5f * dc[2], 0.5f * dc[4], dc[5]}};
    float dM[3][3];
#pragma unroll
    for (int a2 = 0; a2 < 3; ++a2)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) t += (sv[a2] * Rm[j][a2]) * dS[j][k];
            dM[a2][k] = 2.f * t;
        }
#pragma unroll
    for (int a2 = 0; a2 < 3; ++a2) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) t += Rm[k][a2] * dM[a2][k];
        d_scales[i3 + a2] = mod * t;
    }
    if (d_rots) {
        float dR[3][3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int a2 = 0; a2 < 3; ++a2) dR[k][a2] = sv[a2] * dM[a2][k];
        d_rots[4 * (size_t)i] = 2.f * (-qz * dR[0][1] + qy * dR[0][2] + qz * dR[1][0] - qx * dR[1][2] - qy * dR[2][0] + qx * dR[2][1]);
        d_rots[4 * (size_t)i + 1] = 2.f * (qy * dR[0][1] + qz * dR[0][2] + qy * dR[1][0] - 2.f * qx * dR[1][1] - qr * dR[1][2] + qz * dR[2][0] + qr * dR[2][1] - 2.f * qx * dR[2][2]);
        d_rots[4 * (size_t)i + 2] = 2.f * (-2.f * qy * dR[0][0] + qx * dR[0][1] + qr * dR[0][2] + qx * dR[1][0] + qz * dR[1][2] - qr * dR[2][0] + qz * dR[2][1] - 2.f * qy * dR[2][2]);
        d_rots[4 * (size_t)i + 3] = 2.f * (-2.f * qz * dR[0][0] - qr * dR[0][1] + qx * dR[0][2] + qr * dR[1][0] - 2.f * qz * dR[1][1] + qy * dR[1][2] + qx * dR[2][0] + qy * dR[2][1]);
    }
}

int check_settings(const GaRasterSettings *s)
{
    GA_REQUIRE(s != nullptr, "settings is NULL");
    GA_REQUIRE(s->P >= 0 && s->H > 0 && s->W > 0, "bad raster dims P=%d H=%d W=%d", s->P, s->H, s->W);
    GA_REQUIRE(cdiv(s->W, kTile) <= 65535 && cdiv(s->H, kTile) <= 65535, "image too large");
    return GA_OK;
}

}  // namespace
}  // namespace ga

using namespace ga;

extern "C" size_t ga_raster_geom_bytes(int32_t P) { return carve_geom(nullptr, P).total; }
extern "C" size_t ga_raster_img_bytes(int32_t H, int32_t W) { return carve_img(nullptr, H, W).total; }
extern "C" size_t ga_raster_binning_bytes(int64_t R, int32_t H, int32_t W) { return carve_bin(nullptr, R, H, W).total; }
extern "C" size_t ga_raster_bwd_scratch_bytes(int32_t P)
{
    const size_t n = P > 0 ? (size_t)P : 1;
    return align_up(n * sizeof(float2)) + align_up(n * sizeof(float4));
}

extern "C" int ga_raster_forward_preprocess(const GaRasterSettings *s, const float *means3D, const float *scales,
                                            const float *rotations, const float *opacities, const float *viewmatrix,
                                            const float *projmatrix, void *geom, int32_t *radii,
                                            int64_t *num_rendered_host, void *stream_)
{
    if (int rc = check_settings(s)) return rc;
    GA_REQUIRE(num_rendered_host, "num_rendered_host is NULL");
    *num_rendered_host = 0;
    if (s->P == 0) return GA_OK;
    GA_REQUIRE(means3D && scales && rotations && opacities && viewmatrix && projmatrix && geom && radii, "NULL pointer argument");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    GeomViews g = carve_geom(geom, s->P);
    const int gx = cdiv(s->W, kTile), gy = cdiv(s->H, kTile);
    {
        ProfScope _ps("preprocess_fwd_kernel", stream);
        preprocess_fwd_kernel<<<cdiv(s->P, 256), 256, 0, stream>>>(s->P, s->H, s->W, gx, gy, s->tanfovx, s->tanfovy,
                                                                  s->scale_modifier, means3D, scales, rotations, opacities,
                                                                  viewmatrix, projmatrix, g.depth, g.xy, g.conic_o, g.cov3d,
                                                                  g.tiles, g.rect, radii);
    }
    GA_CHECK_LAUNCH("preprocess_fwd_kernel");
    size_t tb = g.scan_temp_bytes;
    {
        ProfScope _ps("cub_inclusive_sum", stream);
        GA_CHECK_CUDA(cub::DeviceScan::InclusiveSum(g.scan_temp, tb, g.tiles, g.offsets, s->P, stream));
    }
    uint32_t total = 0;
    GA_CHECK_CUDA(cudaMemcpyAsync(&total, g.offsets + (s->P - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    GA_CHECK_CUDA(cudaStreamSynchronize(stream));
    *num_rendered_host = (int64_t)total;
    return GA_OK;
}

extern "C" int ga_raster_forward_render(const GaRasterSettings *s, const float *colors, const float *bg, void *geom,
                                        void *binning, size_t binning_bytes, int64_t R, void *img, float *out_color,
                                        void *stream_)
{
    if (int rc = check_settings(s)) return rc;
    GA_REQUIRE(bg && img && out_color, "NULL pointer argument");
    GA_REQUIRE(R >= 0 && R < (int64_t)0x7fffffff, "num_rendered out of range: %lld", (long long)R);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int gx = cdiv(s->W, kTile), gy = cdiv(s->H, kTile);
    ImgViews iv = carve_img(img, s->H, s->W);
    GA_CHECK_CUDA(cudaMemsetAsync(iv.ranges, 0, sizeof(uint2) * (size_t)gx * gy, stream));
    BinViews b{};
    if (R > 0) {
        GA_REQUIRE(colors && geom && binning, "NULL pointer argument");
        b = carve_bin(binning, R, s->H, s->W);
        if (b.total > binning_bytes) {
            set_error("binning buffer too small: need %zu bytes, have %zu", b.total, binning_bytes);
            return GA_ERR_CAPACITY;
        }
        GeomViews g = carve_geom(geom, s->P);
        {
            ProfScope _ps("duplicate_with_keys_kernel", stream);
            duplicate_with_keys_kernel<<<cdiv(s->P, 256), 256, 0, stream>>>(s->P, gx, g.depth, g.offsets, g.tiles, g.rect,
                                                                           b.keys_unsorted, b.vals_unsorted);
        }
        GA_CHECK_LAUNCH("duplicate_with_keys_kernel");
        size_t tb = b.sort_temp_bytes;
        {
            ProfScope _ps("cub_radix_sort_pairs", stream);
            GA_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(b.sort_temp, tb, b.keys_unsorted, b.keys, b.vals_unsorted, b.vals,
                                                          (int)R, 0, sort_end_bit(s->H, s->W), stream));
        }
        {
            ProfScope _ps("tile_ranges_kernel", stream);
            tile_ranges_kernel<<<cdiv(R, 256), 256, 0, stream>>>(R, b.keys, iv.ranges);
        }
        GA_CHECK_LAUNCH("tile_ranges_kernel");
    }
    GeomViews g = carve_geom(geom, s->P);
    {
        ProfScope _ps("render_fwd_kernel", stream);
        render_fwd_kernel<<<dim3(gx, gy), dim3(kTile, kTile), 0, stream>>>(s->H, s->W, iv.ranges, b.vals, g.xy, g.conic_o,
                                                                          colors, bg, iv.final_T, iv.n_contrib, out_color);
    }
    GA_CHECK_LAUNCH("render_fwd_kernel");
    return GA_OK;
}

extern "C" int ga_raster_backward(const GaRasterSettings *s, const float *means3D, const float *colors,
                                  const float *scales, const float *rotations, const float *bg, const float *viewmatrix,
                                  const float *projmatrix, const int32_t *radii, const void *geom, const void *binning,
                                  const void *img, int64_t R, const float *dL_dout, void *scratch, float *d_means3D,
                                  float *d_colors, float *d_scales, float *d_rotations, float *d_opacities,
                                  float *d_means2D, void *stream_)
{
    if (int rc = check_settings(s)) return rc;
    if (s->P == 0) return GA_OK;
    GA_REQUIRE(means3D && colors && scales && rotations && bg && viewmatrix && projmatrix && radii && geom && img &&
                   dL_dout && scratch && d_means3D && d_colors && d_scales,
               "NULL pointer argument");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int gx = cdiv(s->W, kTile), gy = cdiv(s->H, kTile);
    GeomViews g = carve_geom(const_cast<void *>(geom), s->P);
    ImgViews iv = carve_img(const_cast<void *>(img), s->H, s->W);
    Carver sc(scratch);
    float2 *d_mean2D = sc.take<float2>(s->P);
    float4 *d_conic_op = sc.take<float4>(s->P);
    GA_CHECK_CUDA(cudaMemsetAsync(scratch, 0, ga_raster_bwd_scratch_bytes(s->P), stream));
    GA_CHECK_CUDA(cudaMemsetAsync(d_colors, 0, sizeof(float) * 3 * (size_t)s->P, stream));
    if (R > 0) {
        GA_REQUIRE(binning, "NULL binning buffer");
        BinViews b = carve_bin(const_cast<void *>(binning), R, s->H, s->W);
        {
            ProfScope _ps("render_bwd_kernel", stream);
            if (d_opacities)
                render_bwd_kernel<true><<<dim3(gx, gy), dim3(kTile, kTile), 0, stream>>>(s->H, s->W, iv.ranges, b.vals, bg, g.xy,
                                                                              g.conic_o, colors, iv.final_T, iv.n_contrib,
                                                                              dL_dout, d_mean2D, d_conic_op, d_colors);
            else
                render_bwd_kernel<false><<<dim3(gx, gy), dim3(kTile, kTile), 0, stream>>>(s->H, s->W, iv.ranges, b.vals, bg, g.xy,
                                                                              g.conic_o, colors, iv.final_T, iv.n_contrib,
                                                                              dL_dout, d_mean2D, d_conic_op, d_colors);
        }
        GA_CHECK_LAUNCH("render_bwd_kernel");
    }
    {
        ProfScope _ps("preprocess_bwd_kernel", stream);
        preprocess_bwd_kernel<<<cdiv(s->P, 256), 256, 0, stream>>>(s->P, s->H, s->W, s->tanfovx, s->tanfovy, s->scale_modifier,
                                                                  means3D, scales, rotations, viewmatrix, projmatrix, radii,
                                                                  g.cov3d, d_mean2D, d_conic_op, d_means3D, d_scales,
                                                                  d_rotations, d_opacities, d_means2D);
    }
    GA_CHECK_LAUNCH("preprocess_bwd_kernel");
    return GA_OK;
}

extern "C" int ga_raster_views(const GaRasterSettings *s, const void *geom, const void *binning, const void *img,
                               int64_t R, GaRasterViews *out)
{
    if (int rc = check_settings(s)) return rc;
    GA_REQUIRE(out, "out is NULL");
    memset(out, 0, sizeof(*out));
    if (geom) {
        GeomViews g = carve_geom(const_cast<void *>(geom), s->P);
        out->depth = g.depth; out->xy = reinterpret_cast<const float *>(g.xy);
        out->conic_opacity = reinterpret_cast<const float *>(g.conic_o); out->cov3d = g.cov3d;
        out->tiles_touched = g.tiles; out->offsets = g.offsets; out->rect = reinterpret_cast<const uint16_t *>(g.rect);
    }
    if (binning && R > 0) {
        BinViews b = carve_bin(const_cast<void *>(binning), R, s->H, s->W);
        out->keys_unsorted = b.keys_unsorted; out->keys_sorted = b.keys;
        out->vals_unsorted = b.vals_unsorted; out->vals_sorted = b.vals;
    }
    if (img) {
        ImgViews iv = carve_img(const_cast<void *>(img), s->H, s->W);
        out->final_T = iv.final_T; out->n_contrib = iv.n_contrib; out->ranges = reinterpret_cast<const uint32_t *>(iv.ranges);
    }
    return GA_OK;
}
