// warp_u8c3.cuh -- the per-pixel arithmetic of the uint8 x 3 perspective warp, shared by the staged
// kernel (warp_fast.cu, windows in shared memory) and the direct-gather kernel (warp_generic.cu,
// windows straight from global memory).  Semantics: cv2.warpPerspective 4.13 on BGR frames
// (reference call sites vis_homo.py:89,91), SURVEY.md Appendix A.
#pragma once
#include "bevk_common.cuh"

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    return __byte_perm(a, b, sel);
}
__device__ __forceinline__ void st_stream(uint32_t *p, uint32_t v)
{
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// 2-tap window along one axis: first index (clamped into the image) and the weight each of the
// two window positions receives.  Taps outside [0, n) contribute nothing (border value 0).
__device__ __forceinline__ void window(int s, int frac, int n, int &first, int &w0, int &w1)
{
    first = min(max(s, 0), n - 2);
    const int t0 = 32 - frac, t1 = frac;  // weights of taps s and s+1
    w0 = (first == s ? t0 : 0) + (first == s + 1 ? t1 : 0);
    w1 = (first + 1 == s ? t0 : 0) + (first + 1 == s + 1 ? t1 : 0);
}

// Frame-invariant description of one dst pixel.
struct Pix {
    uint32_t addr;   // byte offset (4-aligned) inside a stage of the first window word, row 0
    uint32_t sh;     // 8 * (window start & 3): funnel-shift amount that byte-aligns the window
    uint32_t w03;    // bilinear: column weights as bytes 0 and 3 (dp4a with [B0 B1 B2 B3] -> ch. 0)
                     // nearest : 0x00ffffff when the tap is inside the image, else 0
    uint32_t w16;    // column weights as 16-bit halves (dp2a lo/hi with [B1 B4 B2 B5] -> ch. 1, 2)
    uint32_t b0, b1; // row weights * 64
};

// One pixel out of its staged window: cv2's fixed-point bilinear, result [c0, c1, c2, 0].
// f0 / g0 = window bytes 0..3 of row 0 / 1, f1 / g1 = window bytes 4.. (already byte-aligned).
__device__ __forceinline__ uint32_t lerp_aligned(const Pix &q, uint32_t f0, uint32_t f1, uint32_t g0,
                                                 uint32_t g1)
{
    // F = [B0 B1 B2 B3], G = [B1 B4 B2 B5]
    const uint32_t fg = prmt(f0, f1, 0x5241u), gg = prmt(g0, g1, 0x5241u);
    // horizontal pass: h[row][channel] = a0 * tap0 + a1 * tap1
    const uint32_t h00 = __dp4a(f0, q.w03, 0u);
    const uint32_t h01 = __dp2a_lo(q.w16, fg, 0u);
    const uint32_t h02 = __dp2a_hi(q.w16, fg, 0u);
    const uint32_t h10 = __dp4a(g0, q.w03, 0u);
    const uint32_t h11 = __dp2a_lo(q.w16, gg, 0u);
    const uint32_t h12 = __dp2a_hi(q.w16, gg, 0u);
    // vertical pass, scaled by 64: byte 2 of t is (sum w*p + 2^14) >> 15
    const uint32_t t0 = q.b1 * h10 + (q.b0 * h00 + 32768u);
    const uint32_t t1 = q.b1 * h11 + (q.b0 * h01 + 32768u);
    const uint32_t t2 = q.b1 * h12 + (q.b0 * h02 + 32768u);
    return prmt(prmt(t0, t1, 0x4462u), t2, 0x7610u);
}

