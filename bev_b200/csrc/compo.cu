// compo.cu -- alpha compositing of BEV frames (sm_100a): the blend of
// bev/tool/compo.py:5-24 (composite_reg_img), the caller of the three warps at compo.py:38,46,47.
//
//     out = uint8(clip(round(fg * (mask / 255) + bg * (1 - mask / 255)), max 255))
//
// The reference evaluates this in numpy float64, element by element, rounding half to even.  The
// kernel reproduces it bit for bit: mask / 255 and 1 - mask / 255 have only 256 possible values
// each (a shared-memory table built with IEEE divisions), the two products and the sum are
// unfused double operations in numpy's order, and rint() rounds half to even.  One pass over the
// bytes: 3 reads + 1 write per output byte, 16 bytes per thread and access, HBM-bound.
// bw_mode (compo.py:13-14): the foreground goes through cv2's BGR -> GRAY -> BGR, i.e.
// gray = (3735 B + 19235 G + 9798 R + 16384) >> 15 (cv2 4.13) replicated into the three channels.
#include "bevk_common.cuh"

namespace {

__device__ __forceinline__ uint32_t blend4(uint32_t bg, uint32_t fg, uint32_t mk, const double *s_m,
                                            const double *s_1m)
{
    uint32_t out = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const uint32_t m = (mk >> (8 * b)) & 255u;
        const double f = (double)((fg >> (8 * b)) & 255u), g = (double)((bg >> (8 * b)) & 255u);
        // fg * fg_mask + bg * (1 - fg_mask), every operation rounded on its own (numpy order)
        const double v = __dadd_rn(__dmul_rn(f, s_m[m]), __dmul_rn(g, s_1m[m]));
        double r = rint(v);  // np.round: half to even
        r = r > 255.0 ? 255.0 : r;
        out |= (uint32_t)(int)r << (8 * b);
    }
    return out;
}

// cv2.cvtColor(BGR2GRAY) + GRAY2BGR on 4 consecutive BGR pixels held in 3 words
__device__ __forceinline__ void gray3(uint32_t &w0, uint32_t &w1, uint32_t &w2)
{
    uint8_t px[12];
    px[0] = w0; px[1] = w0 >> 8; px[2] = w0 >> 16; px[3] = w0 >> 24;
    px[4] = w1; px[5] = w1 >> 8; px[6] = w1 >> 16; px[7] = w1 >> 24;
    px[8] = w2; px[9] = w2 >> 8; px[10] = w2 >> 16; px[11] = w2 >> 24;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t y = (3735u * px[3 * i] + 19235u * px[3 * i + 1] + 9798u * px[3 * i + 2] + 16384u) >> 15;
        px[3 * i] = px[3 * i + 1] = px[3 * i + 2] = (uint8_t)y;
    }
    w0 = px[0] | (px[1] << 8) | (px[2] << 16) | ((uint32_t)px[3] << 24);
    w1 = px[4] | (px[5] << 8) | (px[6] << 16) | ((uint32_t)px[7] << 24);
    w2 = px[8] | (px[9] << 8) | (px[10] << 16) | ((uint32_t)px[11] << 24);
}

// n12 = number of 12-byte groups (4 BGR pixels), tail_px = pixels after them (< 4); the buffers
// are 4-byte aligned
template <bool BW>
__global__ void __launch_bounds__(256) composite_kernel(const uint32_t *__restrict__ bg,
                                                        const uint32_t *__restrict__ fg,
                                                        const uint32_t *__restrict__ mk,
                                                        uint32_t *__restrict__ out, long long n12,
                                                        int tail_px)
{
    __shared__ double s_m[256], s_1m[256];
    {
        const double m = __ddiv_rn((double)threadIdx.x, 255.0);  // fg_mask.astype(float) / 255
        s_m[threadIdx.x] = m;
        s_1m[threadIdx.x] = __dsub_rn(1.0, m);
    }
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n12;
         i += (long long)gridDim.x * blockDim.x) {
        uint32_t f0 = __ldg(fg + 3 * i), f1 = __ldg(fg + 3 * i + 1), f2 = __ldg(fg + 3 * i + 2);
        if (BW) gray3(f0, f1, f2);
        const uint32_t o0 = blend4(__ldg(bg + 3 * i), f0, __ldg(mk + 3 * i), s_m, s_1m);
        const uint32_t o1 = blend4(__ldg(bg + 3 * i + 1), f1, __ldg(mk + 3 * i + 1), s_m, s_1m);
        const uint32_t o2 = blend4(__ldg(bg + 3 * i + 2), f2, __ldg(mk + 3 * i + 2), s_m, s_1m);
        out[3 * i] = o0;
        out[3 * i + 1] = o1;
        out[3 * i + 2] = o2;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < tail_px) {  // the last 1..3 pixels, byte by byte
        const long long o = 12 * n12 + 3 * threadIdx.x;
        const uint8_t *b8 = (const uint8_t *)bg + o, *f8 = (const uint8_t *)fg + o, *m8 = (const uint8_t *)mk + o;
        uint32_t f = f8[0] | (f8[1] << 8) | (f8[2] << 16);
        if (BW) {
            const uint32_t y = (3735u * f8[0] + 19235u * f8[1] + 9798u * f8[2] + 16384u) >> 15;
            f = y * 0x010101u;
        }
        const uint32_t r = blend4(b8[0] | (b8[1] << 8) | (b8[2] << 16), f, m8[0] | (m8[1] << 8) | (m8[2] << 16),
                                  s_m, s_1m);
        uint8_t *o8 = (uint8_t *)out + o;
        o8[0] = (uint8_t)r;
        o8[1] = (uint8_t)(r >> 8);
        o8[2] = (uint8_t)(r >> 16);
    }
}

}  // namespace

extern "C" int bevk_composite_u8c3(const void *bg, const void *fg, const void *fg_mask, void *out,
                                   int64_t n_pixels, int bw_mode, void *stream)
{
    int rc = bevk_require_device();
    if (rc) return rc;
    if (n_pixels < 0) BEVK_FAIL(BEVK_E_ARG, "bevk_composite_u8c3: n_pixels must be >= 0");
    if (n_pixels == 0) return BEVK_OK;
    if (!bg || !fg || !fg_mask || !out) BEVK_FAIL(BEVK_E_ARG, "bevk_composite_u8c3: null buffer");
    if (((uintptr_t)bg | (uintptr_t)fg | (uintptr_t)fg_mask | (uintptr_t)out) % 4 != 0)
        BEVK_FAIL(BEVK_E_ARG, "bevk_composite_u8c3: buffers must be 4-byte aligned");
    const long long n12 = n_pixels / 4;
    const int tail_px = (int)(n_pixels % 4);
    long long want = (n12 + 255) / 256;
    const long long cap = (long long)bevk_sm_count() * 8;
    want = want < 1 ? 1 : want;
    const int grid = (int)(want < cap ? want : cap);
    cudaStream_t st = (cudaStream_t)stream;
    if (bw_mode)
        composite_kernel<true><<<grid, 256, 0, st>>>((const uint32_t *)bg, (const uint32_t *)fg,
                                                     (const uint32_t *)fg_mask, (uint32_t *)out, n12, tail_px);
    else
        composite_kernel<false><<<grid, 256, 0, st>>>((const uint32_t *)bg, (const uint32_t *)fg,
                                                      (const uint32_t *)fg_mask, (uint32_t *)out, n12, tail_px);
    BEVK_CUDA(cudaGetLastError());
    return BEVK_OK;
}
