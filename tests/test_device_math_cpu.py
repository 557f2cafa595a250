"""CPU checks of device-side math lifted from the product sources and compiled for the host (no GPU needed).

1. The sub-tile culling of the compositing kernels must be CONSERVATIVE -- a (Gaussian, sub-tile) pair may only be skipped when
every pixel of the sub-tile would have been skipped by the exact per-pixel test (alpha < 1/255 or power > 0) anyway; that is what makes
the culled kernels bit-identical to the un-culled algorithm.  The test lifts the source text of `cull_radius2` / `rect_dist2` out of
csrc/raster.cu, compiles it for the host and brute-forces random conics, opacities and rectangles (including extreme anisotropy)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r"""
#include <cmath>
#include <cstdio>
#include <cstdint>
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
#define __device__
#define __forceinline__ inline
%(functions)s
static uint64_t s = 0x9E3779B97F4A7C15ull;
static inline double rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; }
int main()
{
    long culled = 0, violations = 0, evaluated = 0;
    for (long it = 0; it < 400000; ++it) {
        // random 2D covariance: rotation angle, two axis lengths over six orders of magnitude of anisotropy (+0.3 dilation as in K1)
        const double th = rnd() * 3.14159265358979, l1 = std::exp(rnd() * 9.0 - 2.0), l2 = l1 * std::exp(-rnd() * 12.0);
        const double c = std::cos(th), sn = std::sin(th);
        const float a = (float)(c * c * l1 + sn * sn * l2 + 0.3), b = (float)(c * sn * (l1 - l2)), cc = (float)(sn * sn * l1 + c * c * l2 + 0.3);
        const float det = a * cc - b * b;
        if (!(det > 0.f)) continue;
        float4 co; co.x = cc / det; co.y = -b / det; co.z = a / det;
        co.w = (it %% 7 == 0) ? 1.0f : (float)std::exp(-rnd() * 8.0);            // opacity: 1 (the avatar's constant) or down to 3e-4
        const float rc2 = cull_radius2(co);
        const float reach = 4.f * std::sqrt((float)l1 + 0.3f) + 12.f;
        for (int k = 0; k < 8; ++k) {
            const int sx = (int)(rnd() * 1000), sy = (int)(rnd() * 1000);
            float2 ctr; ctr.x = (float)(sx + 3.5 + (rnd() * 2 - 1) * reach); ctr.y = (float)(sy + 1.5 + (rnd() * 2 - 1) * reach);
            const float d2 = rect_dist2(ctr, (float)sx, (float)(sx + 7), (float)sy, (float)(sy + 3));
            ++evaluated;
            if (!(d2 > rc2)) continue;                                            // the kernels evaluate this pair exactly
            ++culled;
            for (int py = sy; py < sy + 4; ++py)
                for (int px = sx; px < sx + 8; ++px) {
                    const float dx = ctr.x - (float)px, dy = ctr.y - (float)py;
                    const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
                    if (power > 0.f) continue;
                    const float alpha = std::fmin(0.99f, co.w * std::exp(power));
                    if (!(alpha < 1.0f / 255.0f)) ++violations;
                }
        }
    }
    std::printf("%%ld %%ld %%ld\n", evaluated, culled, violations);
    return 0;
}
"""


def _extract(src, name):
    m = re.search(r"__device__ __forceinline__ float " + name + r"\(.*?\n}\n", src, re.S)
    assert m, f"{name} not found in raster.cu"
    return m.group(0)


def test_subtile_culling_is_conservative(tmp_path):
    src = open(os.path.join(ROOT, "gaussianavatar_b200", "csrc", "raster.cu")).read()
    code = HARNESS % {"functions": _extract(src, "cull_radius2") + "\n" + _extract(src, "rect_dist2")}
    cpp, exe = tmp_path / "cull.cpp", tmp_path / "cull"
    cpp.write_text(code)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    for flags in (["-O2", "-ffp-contract=off"], ["-O2", "-ffp-contract=fast", "-march=native"]):
        r = subprocess.run([cxx, "-std=c++17", *flags, "-o", str(exe), str(cpp)], capture_output=True, text=True)
        if r.returncode != 0 and "-march=native" in flags:
            continue
        assert r.returncode == 0, r.stderr
        out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300).stdout.split()
        evaluated, culled, violations = (int(v) for v in out)
        assert evaluated > 1_000_000 and culled > 100_000, (evaluated, culled)   # the test really exercises the cull branch
        assert violations == 0, f"{violations} pixels of culled pairs would have been blended"


LAYOUT_HARNESS = r"""
#include <cstdio>
#include <cstdint>
#include <set>
#define __device__
#define __forceinline__ inline
%(functions)s
// decoded on the device (profiles/r1_umma_mn_major_probe.md): byte offset of element (k, m) of an MN-major kind::tf32 operand
static uint32_t probe_byte(int k, int m, uint32_t lbo, uint32_t sbo)
{ return (uint32_t)(m / 32) * lbo + (uint32_t)(k / 4) * sbo + (uint32_t)(k %% 4) * 128u + (uint32_t)((((m %% 32) / 8) ^ (k %% 4)) * 32) + (uint32_t)(m %% 8) * 4u; }
int main()
{
    int bad = 0;
    std::set<uint32_t> seen;
    for (int r = 0; r < 32; ++r)                 // r = K index (pixel) inside a 32-row chunk, v = 16-byte unit = channels 4v..4v+3
        for (int v = 0; v < 8; ++v) {
            const uint32_t off = mn_unit(r, v);
            bad += off != probe_byte(r, 4 * v, 4096, 512);          // SBO = 512: 4-row atoms are contiguous, so row r sits at r * 128
            bad += (off %% 16) != 0 || off >= 4096;
            seen.insert(off);
        }
    bad += seen.size() != 256;                   // a bijection onto the chunk's 256 sixteen-byte slots
    seen.clear();
    for (int r = 0; r < 128; ++r)                // K-major 128-byte swizzle: 16-byte unit u of row r
        for (int u = 0; u < 8; ++u) {
            const uint32_t off = sw128_offset(r, u);
            bad += off != (uint32_t)r * 128u + (uint32_t)((u ^ (r & 7)) << 4);
            seen.insert(off);
        }
    bad += seen.size() != 1024;
    std::printf("%%d\n", bad);
    return 0;
}
"""


def test_umma_layout_helpers_match_the_decoded_hardware_layout(tmp_path):
    """`mn_unit` (conv_tc.cu) and `sw128_offset` (tc_common.cuh), lifted from the product sources and compiled for the host: the
    MN-major helper must reproduce the layout decoded on the device word by word, and both must be bijections onto their tiles."""
    conv = open(os.path.join(ROOT, "gaussianavatar_b200", "csrc", "conv_tc.cu")).read()
    tcc = open(os.path.join(ROOT, "gaussianavatar_b200", "csrc", "tc_common.cuh")).read()
    f1 = re.search(r"__device__ __forceinline__ uint32_t mn_unit\(.*?\}\n", conv, re.S)
    f2 = re.search(r"__device__ __forceinline__ uint32_t sw128_offset\(.*?\}\n", tcc, re.S)
    assert f1 and f2
    cpp, exe = tmp_path / "layout.cpp", tmp_path / "layout"
    cpp.write_text(LAYOUT_HARNESS % {"functions": f1.group(0) + "\n" + f2.group(0)})
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-std=c++17", "-O1", "-o", str(exe), str(cpp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)], capture_output=True, text=True, timeout=60).stdout.strip() == "0"
