// warp_formats.cuh -- further pixel-format policies of the staged warp kernel (warp_fast.cu):
// uint8 x 1 (grey frames, masks), uint8 x 4 (BGRA) and float32 x 3.  Same interface as PxU8C3
// (warp_u8c3.cuh) and PxF16C3 (warp_f16c3.cuh); semantics cv2.warpPerspective 4.13 (reference call
// sites vis_homo.py:89,91, bev/tool/compo.py:38,46,47), SURVEY.md Appendix A.
//
// As in the other policies the 2x2 window is clamped into the image and positions that no tap of
// cv2 falls on carry weight 0, so the frame loop has no border branches.
#pragma once
#include "bevk_common.cuh"
#include "warp_u8c3.cuh"  // prmt, st_stream, tap_weight, window

__device__ __forceinline__ void st_stream_u8(uint8_t *p, uint32_t v)
{
    asm volatile("st.global.cs.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- uint8 x 1 ------------------------------------------------------------------------------------
// A window row is 2 bytes starting on any byte: one aligned word, two when it starts on byte 3.
// 32 pixels are 32 bytes: every lane stores its byte (one full 32-byte sector per warp store).
struct PxU8C1 {
    static constexpr int kBpp = 1;
    static constexpr int kSegBytes = 32;
    static constexpr int kDtype = BEVK_U8;
    static constexpr bool kPairs = false;
    static constexpr int kLinearThreads = 256, kNearestThreads = 256;
    static constexpr int kLinearCtas = 3;
    static constexpr int kWinWords = 4;
    using Reg = Pix;       // addr, sh, w0 / w1 = dp2a tap weights of window rows 0 / 1 (nearest: w0 = mask)
    using Out = uint32_t;  // the byte
    struct Store {};

    template <bool LINEAR>
    static __device__ __forceinline__ Reg make(bool act, uint32_t A, int wc0, int wc1, int wr0, int wr1)
    {
        Reg q;
        q.addr = A & ~3u;
        q.sh = 8 * (A & 3);
        if (LINEAR) {
            q.w0 = tap_weight(wc0, wr0) | (tap_weight(wc1, wr0) << 16);
            q.w1 = tap_weight(wc0, wr1) | (tap_weight(wc1, wr1) << 16);
        } else {
            q.w0 = act ? 0xffu : 0u;
            q.w1 = 0;
        }
        return q;
    }
    template <bool LINEAR> static __device__ __forceinline__ Reg make(bool act, uint32_t A, uint32_t wpk)
    {
        return make<LINEAR>(act, A, (int)(wpk & 0xffu), (int)((wpk >> 8) & 0xffu), (int)((wpk >> 16) & 0xffu),
                            (int)(wpk >> 24));
    }
    template <bool LINEAR> static constexpr int last_word_offset() { return LINEAR ? 4 : 0; }

    template <bool LINEAR, typename LD>
    static __device__ __forceinline__ void load(const Reg &q, uint32_t ra, uint32_t rb, uint32_t last_a,
                                                uint32_t last_b, uint32_t (&w)[kWinWords], LD ld)
    {
        w[0] = ld(ra);
        if (LINEAR) {
            const bool second = q.sh == 24;  // the window's right tap sits in the next word
            w[1] = second ? ld(last_a) : 0u;
            w[2] = ld(rb);
            w[3] = second ? ld(last_b) : 0u;
        }
    }
    template <bool LINEAR>
    static __device__ __forceinline__ Out math(const Reg &q, const uint32_t (&w)[kWinWords])
    {
        if (LINEAR) {
            const uint32_t f = __funnelshift_r(w[0], w[1], q.sh), g = __funnelshift_r(w[2], w[3], q.sh);
            // byte 2 of the sum is (sum w*p + 2^14) >> 15 in cv2's scale (see lerp_xy)
            return (__dp2a_lo(q.w1, g, __dp2a_lo(q.w0, f, 32768u)) >> 16) & 0xffu;
        }
        return (w[0] >> q.sh) & q.w0;
    }
    static __device__ __forceinline__ Store store_setup(int) { return Store(); }
    static __device__ __forceinline__ uint32_t lane_offset(int lane) { return (uint32_t)lane; }
    static __device__ __forceinline__ bool lane_stores(int lane, int valid_px) { return lane < valid_px; }
    static __device__ __forceinline__ void store(uint8_t *d, Out v, bool ok, const Store &, int)
    {
        if (ok) st_stream_u8(d, v);
    }
    static __device__ __forceinline__ void store_zero(uint8_t *d, bool ok, int)
    {
        if (ok) st_stream_u8(d, 0u);
    }
};

// ---- uint8 x 4 ------------------------------------------------------------------------------------
// A pixel is one aligned word; a window row is two adjacent words; 32 pixels are 128 bytes and
// every lane stores its own word.
struct PxU8C4 {
    static constexpr int kBpp = 4;
    static constexpr int kSegBytes = 128;
    static constexpr int kDtype = BEVK_U8;
    static constexpr bool kPairs = false;
    static constexpr int kLinearThreads = 256, kNearestThreads = 256;
    static constexpr int kLinearCtas = 3;
    static constexpr int kWinWords = 4;
    using Reg = Pix;       // sh unused (0)
    using Out = uint32_t;  // [c0 c1 c2 c3]
    struct Store {};

    template <bool LINEAR>
    static __device__ __forceinline__ Reg make(bool act, uint32_t A, int wc0, int wc1, int wr0, int wr1)
    {
        Reg q;
        q.addr = A;
        q.sh = 0;
        if (LINEAR) {
            q.w0 = tap_weight(wc0, wr0) | (tap_weight(wc1, wr0) << 16);
            q.w1 = tap_weight(wc0, wr1) | (tap_weight(wc1, wr1) << 16);
        } else {
            q.w0 = act ? 0xffffffffu : 0u;
            q.w1 = 0;
        }
        return q;
    }
    template <bool LINEAR> static __device__ __forceinline__ Reg make(bool act, uint32_t A, uint32_t wpk)
    {
        return make<LINEAR>(act, A, (int)(wpk & 0xffu), (int)((wpk >> 8) & 0xffu), (int)((wpk >> 16) & 0xffu),
                            (int)(wpk >> 24));
    }
    template <bool LINEAR> static constexpr int last_word_offset() { return LINEAR ? 4 : 0; }

    template <bool LINEAR, typename LD>
    static __device__ __forceinline__ void load(const Reg &, uint32_t ra, uint32_t rb, uint32_t last_a,
                                                uint32_t last_b, uint32_t (&w)[kWinWords], LD ld)
    {
        w[0] = ld(ra);
        if (LINEAR) {
            w[1] = ld(last_a);
            w[2] = ld(rb);
            w[3] = ld(last_b);
        }
    }
    template <bool LINEAR>
    static __device__ __forceinline__ Out math(const Reg &q, const uint32_t (&w)[kWinWords])
    {
        if (LINEAR) {
            // per row: xa = [B0 B4 B1 B5] (channels 0, 1: left / right tap), ya = [B2 B6 B3 B7]
            const uint32_t xa = prmt(w[0], w[1], 0x5140u), ya = prmt(w[0], w[1], 0x7362u);
            const uint32_t xb = prmt(w[2], w[3], 0x5140u), yb = prmt(w[2], w[3], 0x7362u);
            const uint32_t t0 = __dp2a_lo(q.w1, xb, __dp2a_lo(q.w0, xa, 32768u));
            const uint32_t t1 = __dp2a_hi(q.w1, xb, __dp2a_hi(q.w0, xa, 32768u));
            const uint32_t t2 = __dp2a_lo(q.w1, yb, __dp2a_lo(q.w0, ya, 32768u));
            const uint32_t t3 = __dp2a_hi(q.w1, yb, __dp2a_hi(q.w0, ya, 32768u));
            // byte 2 of every t is the channel value
            return prmt(prmt(t0, t1, 0x0062u), prmt(t2, t3, 0x0062u), 0x5410u);
        }
        return w[0] & q.w0;
    }
    static __device__ __forceinline__ Store store_setup(int) { return Store(); }
    static __device__ __forceinline__ uint32_t lane_offset(int lane) { return 4u * (uint32_t)lane; }
    static __device__ __forceinline__ bool lane_stores(int lane, int valid_px) { return lane < valid_px; }
    static __device__ __forceinline__ void store(uint8_t *d, Out v, bool ok, const Store &, int)
    {
        if (ok) st_stream(reinterpret_cast<uint32_t *>(d), v);
    }
    static __device__ __forceinline__ void store_zero(uint8_t *d, bool ok, int)
    {
        if (ok) st_stream(reinterpret_cast<uint32_t *>(d), 0u);
    }
};

// ---- float32 x 3 ----------------------------------------------------------------------------------
// cv2's float path: tx = ax / 32, ty = ay / 32 in fp32, w00 = (1-ty)(1-tx),
// w01 = (1-ty) tx, w10 = ty (1-tx), w11 = ty tx, dst = ((p00 w00 + p01 w01) + p10 w10) + p11 w11 with
// every operation rounded (never fused).  A pixel is 3 aligned words, a window row 6.  Positions of
// the clamped window that no tap falls on carry weight 0 (value * 0 = 0 for every finite value;
// frames with Inf / NaN next to the image border take the direct-gather kernel's exact border
// code if that matters); a pixel with no tap inside the image is written as +0.
struct PixF32 {
    uint32_t addr;
    uint32_t sh;  // all ones if any tap is inside the image, else 0 (ANDed onto the result)
    float w00, w01, w10, w11;
};
struct PxF32C3 {
    static constexpr int kBpp = 12;
    static constexpr int kSegBytes = 384;
    static constexpr int kDtype = BEVK_F32;
    static constexpr bool kPairs = false;
    static constexpr int kLinearThreads = 256, kNearestThreads = 256;
    static constexpr int kLinearCtas = 3;
    static constexpr int kWinWords = 12;
    using Reg = PixF32;
    struct Out {
        uint32_t x, y, z;
    };
    struct Store {};

    template <bool LINEAR>
    static __device__ __forceinline__ Reg make(bool act, uint32_t A, int wc0, int wc1, int wr0, int wr1)
    {
        Reg q;
        q.addr = A;
        q.sh = act ? 0xffffffffu : 0u;
        if (LINEAR) {
            // wc / wr are 32 - frac, frac or 0: dividing by 32 is exact, so c0 = 1 - tx, c1 = tx ...
            const float c0 = __fmul_rn((float)wc0, 1.0f / 32.0f), c1 = __fmul_rn((float)wc1, 1.0f / 32.0f);
            const float r0 = __fmul_rn((float)wr0, 1.0f / 32.0f), r1 = __fmul_rn((float)wr1, 1.0f / 32.0f);
            q.w00 = __fmul_rn(r0, c0);
            q.w01 = __fmul_rn(r0, c1);
            q.w10 = __fmul_rn(r1, c0);
            q.w11 = __fmul_rn(r1, c1);
        } else {
            q.w00 = q.w01 = q.w10 = q.w11 = 0.0f;
        }
        return q;
    }
    template <bool LINEAR> static __device__ __forceinline__ Reg make(bool act, uint32_t A, uint32_t wpk)
    {
        return make<LINEAR>(act, A, (int)(wpk & 0xffu), (int)((wpk >> 8) & 0xffu), (int)((wpk >> 16) & 0xffu),
                            (int)(wpk >> 24));
    }
    template <bool LINEAR> static constexpr int last_word_offset() { return LINEAR ? 20 : 8; }

    template <bool LINEAR, typename LD>
    static __device__ __forceinline__ void load(const Reg &, uint32_t ra, uint32_t rb, uint32_t last_a,
                                                uint32_t last_b, uint32_t (&w)[kWinWords], LD ld)
    {
        if (LINEAR) {
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                w[i] = ld(ra + 4 * i);
                w[6 + i] = ld(rb + 4 * i);
            }
            w[5] = ld(last_a);
            w[11] = ld(last_b);
        } else {
            w[0] = ld(ra);
            w[1] = ld(ra + 4);
            w[2] = ld(last_a);
        }
    }
    static __device__ __forceinline__ uint32_t blend(const Reg &q, uint32_t p00, uint32_t p01, uint32_t p10,
                                                     uint32_t p11)
    {
        float r = __fadd_rn(__fmul_rn(__uint_as_float(p00), q.w00), __fmul_rn(__uint_as_float(p01), q.w01));
        r = __fadd_rn(r, __fmul_rn(__uint_as_float(p10), q.w10));
        r = __fadd_rn(r, __fmul_rn(__uint_as_float(p11), q.w11));
        return __float_as_uint(r) & q.sh;
    }
    template <bool LINEAR>
    static __device__ __forceinline__ Out math(const Reg &q, const uint32_t (&w)[kWinWords])
    {
        Out o;
        if (LINEAR) {
            o.x = blend(q, w[0], w[3], w[6], w[9]);
            o.y = blend(q, w[1], w[4], w[7], w[10]);
            o.z = blend(q, w[2], w[5], w[8], w[11]);
        } else {
            o.x = w[0] & q.sh;
            o.y = w[1] & q.sh;
            o.z = w[2] & q.sh;
        }
        return o;
    }
    static __device__ __forceinline__ Store store_setup(int) { return Store(); }
    static __device__ __forceinline__ uint32_t lane_offset(int lane) { return 12u * (uint32_t)lane; }
    static __device__ __forceinline__ bool lane_stores(int lane, int valid_px) { return lane < valid_px; }
    static __device__ __forceinline__ void store(uint8_t *d, Out v, bool ok, const Store &, int)
    {
        if (ok) {
            uint32_t *p = reinterpret_cast<uint32_t *>(d);
            st_stream(p, v.x);
            st_stream(p + 1, v.y);
            st_stream(p + 2, v.z);
        }
    }
    static __device__ __forceinline__ void store_zero(uint8_t *d, bool ok, int)
    {
        if (ok) {
            uint32_t *p = reinterpret_cast<uint32_t *>(d);
            st_stream(p, 0u);
            st_stream(p + 1, 0u);
            st_stream(p + 2, 0u);
        }
    }
};
