// proj.cu -- batched 3x3 projection kernels for points and rotated boxes (sm_100a).
//
// Replaces the ~25 eager ATen launches per call of bev/rbox_torch.py (reference) and the numpy
// passes of bev/rbox.py with ONE kernel per API function: every box / point is read once and
// written once.  These kernels are HBM-bound (52 B per box); no tensor cores -- the work is a
// per-element 2x2 rotate + 3x3 matvec + divide.
//
// Precision: tensors are float32 (or float64) in and out, but the geometry in between runs in
// float64 registers (own range-reduced sincos, FMA matvec, one IEEE reciprocal per corner) and is
// rounded once on output.  That is what keeps results within 1e-5 relative of the reference's
// float64 numpy path near the horizon, where the reference's own float32 torch path is 2.3e-4 off
// (SURVEY.md 0.5 / 8c).
//
// Layout: rows of 5 floats (20 B) are not 16 B aligned, but a block of 256 rows is one contiguous,
// 16-byte-aligned run: rows_bulk_kernel moves it with ONE cp.async.bulk into a shared-memory ring
// (mbarrier completion) and writes the 256 result rows back with one bulk store; threads only
// touch shared memory.  Unaligned buffers, float64 tensors and the tail of fewer than 256 rows go
// through rows_kernel, which stages the rows with coalesced 4 B accesses and an odd row pitch in
// shared memory (W | 1 words) so that the per-row reads are bank-conflict free.
#include "bevk_common.cuh"

namespace {

constexpr int kRows = 256;  // rows per block iteration == threads per block

struct Mat3 {
    double h[9];
};

// ---- float64 helpers ----------------------------------------------------------------------------
// sin/cos with |abs err| < 2e-11 for |x| < ~1e5: Cody-Waite reduction by pi/2 + Taylor kernels on
// [-pi/4, pi/4].  ~22 DP ops instead of the ~120 of the fully accurate library sincos, which
// would make the kernels FP64-bound.
__device__ __forceinline__ void sincos_fast64(double x, double &s, double &c)
{
    const double k = rint(x * 0.63661977236758134308);  // 2/pi
    double t = fma(k, -1.57079632679489655800e+00, x);  // pi/2 hi
    t = fma(k, -6.12323399573676603587e-17, t);         // pi/2 lo
    const double t2 = t * t;
    double ps = fma(t2, -2.50521083854417187751e-08, 2.75573192239858906526e-06);  // -1/11!, 1/9!
    ps = fma(ps, t2, -1.98412698412698412698e-04);                                // -1/7!
    ps = fma(ps, t2, 8.33333333333333333333e-03);                                 //  1/5!
    ps = fma(ps, t2, -1.66666666666666666667e-01);                                // -1/3!
    ps = fma(ps * t2, t, t);
    double pc = fma(t2, 2.08767569878680989792e-09, -2.75573192239858906526e-07);  // 1/12!, -1/10!
    pc = fma(pc, t2, 2.48015873015873015873e-05);                                 //  1/8!
    pc = fma(pc, t2, -1.38888888888888888889e-03);                                // -1/6!
    pc = fma(pc, t2, 4.16666666666666666667e-02);                                 //  1/4!
    pc = fma(pc, t2, -0.5);
    pc = fma(pc, t2, 1.0);
    const int q = (int)k & 3;
    const double ss = (q & 1) ? pc : ps;
    const double cc = (q & 1) ? ps : pc;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}

template <typename T> struct Tr;
template <> struct Tr<float> {
    static __device__ __forceinline__ void sincos(double r, double &s, double &c) { sincos_fast64(r, s, c); }
    static __device__ __forceinline__ double atan2(double y, double x) { return (double)atan2f((float)y, (float)x); }
    static __device__ __forceinline__ double sqrt(double v) { return (double)sqrtf((float)v); }
};
template <> struct Tr<double> {
    static __device__ __forceinline__ void sincos(double r, double &s, double &c) { ::sincos(r, &s, &c); }
    static __device__ __forceinline__ double atan2(double y, double x) { return ::atan2(y, x); }
    static __device__ __forceinline__ double sqrt(double v) { return ::sqrt(v); }
};

// [x, y, 1] through H with divide: one reciprocal, two multiplies
__device__ __forceinline__ void project(const Mat3 &H, double x, double y, double &u, double &v)
{
    const double X = fma(H.h[0], x, fma(H.h[1], y, H.h[2]));
    const double Y = fma(H.h[3], x, fma(H.h[4], y, H.h[5]));
    const double W = fma(H.h[6], x, fma(H.h[7], y, H.h[8]));
    // 1/W: hardware seed (2^-23) + one Newton step -> ~2^-45 relative, far inside the 1e-5
    // contract, at a third of the cost of the IEEE division (W = 0 still gives inf / nan)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(W));
    r = fma(r, fma(-W, r, 1.0), r);
    u = X * r;
    v = Y * r;
}

// ---- per-row functors: in[] / out[] are float64 registers ------------------------------------
struct FnCorners {  // xywhr -> 4 corners (+ optional homography)        rbox_torch.py:52-99
    static constexpr int IN = 5, OUT = 8;
    Mat3 H;
    int mode, use_h;
    template <typename T> __device__ __forceinline__ void operator()(const double *in, double *out) const
    {
        double s, c;
        Tr<T>::sincos(in[4], s, c);
        const double hw = in[2] * 0.5, hh = in[3] * 0.5;
        // template (tl, bl, br, tr) and rotation sign layout per mode (rbox_torch.py:45-48,65-82)
        double tx[4], ty[4], r01, r10;
        if (mode == BEVK_MODE_BEV) {
            tx[0] = -hw; ty[0] = -hh; tx[1] = -hw; ty[1] = hh;
            tx[2] = hw;  ty[2] = hh;  tx[3] = hw;  ty[3] = -hh;
            r01 = s; r10 = -s;
        } else {
            tx[0] = -hh; ty[0] = -hw; tx[1] = hh;  ty[1] = -hw;
            tx[2] = hh;  ty[2] = hw;  tx[3] = -hh; ty[3] = hw;
            r01 = -s; r10 = s;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double px = fma(c, tx[i], fma(r01, ty[i], in[0]));
            const double py = fma(r10, tx[i], fma(c, ty[i], in[1]));
            if (use_h)
                project(H, px, py, out[2 * i], out[2 * i + 1]);
            else {
                out[2 * i] = px;
                out[2 * i + 1] = py;
            }
        }
    }
};

struct FnFromCorners {  // 4 corners (+ optional homography first) -> xywhr     rbox.py:50-63
    static constexpr int IN = 8, OUT = 5;
    Mat3 H;
    int mode, use_h;
    template <typename T> __device__ __forceinline__ void operator()(const double *in, double *out) const
    {
        double tlx = in[0], tly = in[1], blx = in[2], bly = in[3], trx = in[6], try_ = in[7];
        if (use_h) {  // bottom-right (in[4], in[5]) is not used by the reference formula
            project(H, in[0], in[1], tlx, tly);
            project(H, in[2], in[3], blx, bly);
            project(H, in[6], in[7], trx, try_);
        }
        const double wx = trx - tlx, wy = try_ - tly, hx = blx - tlx, hy = bly - tly;
        out[0] = 0.5 * (blx + trx);
        out[1] = 0.5 * (bly + try_);
        out[2] = Tr<T>::sqrt(fma(wx, wx, wy * wy));
        out[3] = Tr<T>::sqrt(fma(hx, hx, hy * hy));
        const double vx = tlx - blx, vy = tly - bly;  // heading = tl - bl
        out[4] = (mode == BEVK_MODE_BEV) ? Tr<T>::atan2(vx, vy) : Tr<T>::atan2(vy, vx);
    }
};

struct FnSimilarity {  // rbox_world_bev                                   rbox_torch.py:123-168
    static constexpr int IN = 5, OUT = 5;
    Mat3 H;        // already divided by H[8] on the host
    double scale;  // sqrt(H00^2 + H01^2)
    int src_mode;
    template <typename T> __device__ __forceinline__ void operator()(const double *in, double *out) const
    {
        double s, c;
        Tr<T>::sincos(in[4], s, c);
        // yaw2v(src): bev -> (sin, cos), world -> (cos, sin)
        const double vx = (src_mode == BEVK_MODE_BEV) ? s : c;
        const double vy = (src_mode == BEVK_MODE_BEV) ? c : s;
        const double tx = fma(H.h[0], vx, H.h[1] * vy), ty = fma(H.h[3], vx, H.h[4] * vy);
        // v2yaw(target): target is the other system
        out[4] = (src_mode == BEVK_MODE_BEV) ? Tr<T>::atan2(ty, tx) : Tr<T>::atan2(tx, ty);
        out[0] = fma(H.h[0], in[0], fma(H.h[1], in[1], H.h[2]));  // affine: no divide (:155-156)
        out[1] = fma(H.h[3], in[0], fma(H.h[4], in[1], H.h[5]));
        out[2] = in[2] * scale;
        out[3] = in[3] * scale;
    }
};

struct FnRboxTT {  // rboxtt_world_bev: similarity of xywhr + the height tail (du, dv)   rbox.py:258-288
    static constexpr int IN = 7, OUT = 7;
    FnSimilarity sim;
    template <typename T> __device__ __forceinline__ void operator()(const double *in, double *out) const
    {
        sim.template operator()<T>(in, out);
        const Mat3 &H = sim.H;  // affine (asserted on the host): tail = H [x+du, y+dv, 1] - H [x, y, 1]
        const double ex = in[0] + in[5], ey = in[1] + in[6];
        out[5] = fma(H.h[0], ex, fma(H.h[1], ey, H.h[2])) - out[0];
        out[6] = fma(H.h[3], ex, fma(H.h[4], ey, H.h[5])) - out[1];
    }
};

struct FnZt2tt {  // rbox_zt2tt_world: (x, y, w, h, r, z, t) -> (x', y', w, h, r, du, dv)   rbox.py:228-256
    static constexpr int IN = 7, OUT = 7;
    double K[9], R[9], t[3];
    Mat3 Hwc;  // inv(K [r1 r2 t]): image -> world plane z = 0
    __device__ __forceinline__ void ground(double x, double y, double z, double &gx, double &gy) const
    {
        const double cx = fma(R[0], x, fma(R[1], y, fma(R[2], z, t[0])));
        const double cy = fma(R[3], x, fma(R[4], y, fma(R[5], z, t[1])));
        const double cz = fma(R[6], x, fma(R[7], y, fma(R[8], z, t[2])));
        const double u = fma(K[0], cx, fma(K[1], cy, K[2] * cz));
        const double v = fma(K[3], cx, fma(K[4], cy, K[5] * cz));
        const double d = fma(K[6], cx, fma(K[7], cy, K[8] * cz));
        const double dc = d < 1e-2 ? 1e-2 : d;  // np.clip(uvd[2], a_min=1e-2)
        const double a = u / dc, b = v / dc, c = d / dc;
        const double X = fma(Hwc.h[0], a, fma(Hwc.h[1], b, Hwc.h[2] * c));
        const double Y = fma(Hwc.h[3], a, fma(Hwc.h[4], b, Hwc.h[5] * c));
        const double W = fma(Hwc.h[6], a, fma(Hwc.h[7], b, Hwc.h[8] * c));
        gx = X / W;
        gy = Y / W;
    }
    template <typename T> __device__ __forceinline__ void operator()(const double *in, double *out) const
    {
        double lx, ly, hx, hy;
        ground(in[0], in[1], in[5], lx, ly);
        ground(in[0], in[1], in[5] + in[6], hx, hy);
        out[0] = lx;
        out[1] = ly;
        out[2] = in[2];
        out[3] = in[3];
        out[4] = in[4];
        out[5] = hx - lx;
        out[6] = hy - ly;
    }
};

struct FnPts2 {  // pts_world_bev, (N,2)                                           rbox.py:136-151
    static constexpr int IN = 2, OUT = 2;
    Mat3 H;
    template <typename T> __device__ __forceinline__ void operator()(const double *in, double *out) const
    {
        project(H, in[0], in[1], out[0], out[1]);
    }
};
struct FnPts3 {  // pts_world_bev, homogeneous (N,3): third column comes back as w/w
    static constexpr int IN = 3, OUT = 3;
    Mat3 H;
    template <typename T> __device__ __forceinline__ void operator()(const double *in, double *out) const
    {
        const double X = fma(H.h[0], in[0], fma(H.h[1], in[1], H.h[2] * in[2]));
        const double Y = fma(H.h[3], in[0], fma(H.h[4], in[1], H.h[5] * in[2]));
        const double W = fma(H.h[6], in[0], fma(H.h[7], in[1], H.h[8] * in[2]));
        out[0] = X / W;
        out[1] = Y / W;
        out[2] = W / W;
    }
};
struct FnXyvec {  // xywhr2xyvec                                              rbox_torch.py:101-112
    static constexpr int IN = 5, OUT = 4;
    int mode;
    template <typename T> __device__ __forceinline__ void operator()(const double *in, double *out) const
    {
        double s, c;
        Tr<T>::sincos(in[4], s, c);
        const double dx = (mode == BEVK_MODE_BEV) ? s : c, dy = (mode == BEVK_MODE_BEV) ? c : s;
        out[0] = in[0];
        out[1] = in[1];
        out[2] = fma(dx, in[3], in[0]);
        out[3] = fma(dy, in[3], in[1]);
    }
};
struct FnXy8vec {  // xy82xyvec                                               rbox_torch.py:114-121
    static constexpr int IN = 8, OUT = 4;
    template <typename T> __device__ __forceinline__ void operator()(const double *in, double *out) const
    {
        const double cx = 0.5 * (in[0] + in[4]), cy = 0.5 * (in[1] + in[5]);
        out[0] = cx;
        out[1] = cy;
        out[2] = cx + (in[2] - in[0]);
        out[3] = cy + (in[3] - in[1]);
    }
};
struct FnV2yaw {  // rbox_torch.py:24-31
    static constexpr int IN = 2, OUT = 1;
    int mode;
    template <typename T> __device__ __forceinline__ void operator()(const double *in, double *out) const
    {
        out[0] = (mode == BEVK_MODE_BEV) ? Tr<T>::atan2(in[0], in[1]) : Tr<T>::atan2(in[1], in[0]);
    }
};
struct FnYaw2v {  // rbox_torch.py:33-40
    static constexpr int IN = 1, OUT = 2;
    int mode;
    template <typename T> __device__ __forceinline__ void operator()(const double *in, double *out) const
    {
        double s, c;
        Tr<T>::sincos(in[0], s, c);
        out[0] = (mode == BEVK_MODE_BEV) ? s : c;
        out[1] = (mode == BEVK_MODE_BEV) ? c : s;
    }
};
struct FnYaw2mat {  // rbox_torch.py:42-50
    static constexpr int IN = 1, OUT = 4;
    int mode;
    template <typename T> __device__ __forceinline__ void operator()(const double *in, double *out) const
    {
        double s, c;
        Tr<T>::sincos(in[0], s, c);
        out[0] = c;
        out[3] = c;
        out[1] = (mode == BEVK_MODE_BEV) ? s : -s;
        out[2] = (mode == BEVK_MODE_BEV) ? -s : s;
    }
};

struct FnAngle {  // angle_world_bev: yaw2v(src) -> H[:2,:2] . v -> v2yaw(target)      rbox.py:162-171
    static constexpr int IN = 1, OUT = 1;
    double h00, h01, h10, h11;
    int src_mode;
    template <typename T> __device__ __forceinline__ void operator()(const double *in, double *out) const
    {
        double s, c;
        Tr<T>::sincos(in[0], s, c);
        const double vx = (src_mode == BEVK_MODE_BEV) ? s : c;
        const double vy = (src_mode == BEVK_MODE_BEV) ? c : s;
        const double tx = fma(h00, vx, h01 * vy), ty = fma(h10, vx, h11 * vy);
        out[0] = (src_mode == BEVK_MODE_BEV) ? Tr<T>::atan2(ty, tx) : Tr<T>::atan2(tx, ty);
    }
};
struct FnDist {  // dist_world_bev: lengths times the similarity's scale                rbox.py:153-160
    static constexpr int IN = 1, OUT = 1;
    double scale;
    template <typename T> __device__ __forceinline__ void operator()(const double *in, double *out) const
    {
        out[0] = in[0] * scale;
    }
};

// ---- the row kernel -----------------------------------------------------------------------------
template <typename T, typename F>
__global__ void __launch_bounds__(kRows) rows_kernel(const T *__restrict__ in, T *__restrict__ out,
                                                     long long n, const __grid_constant__ F f)
{
    constexpr int IN = F::IN, OUT = F::OUT;
    constexpr int PI = IN | 1, PO = OUT | 1;  // odd pitch: conflict-free row access
    __shared__ T s_in[kRows * PI];
    __shared__ T s_out[kRows * PO];
    const long long n_blocks = (n + kRows - 1) / kRows;
    for (long long b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const long long row0 = b * kRows;
        const int rows = (int)min((long long)kRows, n - row0);
        const T *gin = in + row0 * IN;
#pragma unroll
        for (int k = 0; k < IN; ++k) {
            const int i = k * kRows + threadIdx.x;
            if (i < rows * IN) s_in[(i / IN) * PI + (i % IN)] = __ldg(gin + i);
        }
        __syncthreads();
        if ((int)threadIdx.x < rows) {
            double a[IN], r[OUT];
#pragma unroll
            for (int k = 0; k < IN; ++k) a[k] = (double)s_in[threadIdx.x * PI + k];
            f.template operator()<T>(a, r);
#pragma unroll
            for (int k = 0; k < OUT; ++k) s_out[threadIdx.x * PO + k] = (T)r[k];
        }
        __syncthreads();
        T *gout = out + row0 * OUT;
#pragma unroll
        for (int k = 0; k < OUT; ++k) {
            const int i = k * kRows + threadIdx.x;
            if (i < rows * OUT) gout[i] = s_out[(i / OUT) * PO + (i % OUT)];
        }
        // the next iteration's s_in writes are ordered after this iteration's reads by the
        // barrier above; s_out reads are ordered before the next writes by the next first barrier
    }
}

// ---- the pipelined row kernel: whole blocks of 256 rows move with the TMA unit ------------------
// A block of 256 rows is one contiguous run in global memory (256 * IN * sizeof(T) bytes, always a
// multiple of 16), so it is fetched with ONE cp.async.bulk into a 3-deep shared-memory ring
// (completion on an mbarrier) and written back with ONE bulk store out of a 2-deep ring.  The
// threads only touch shared memory: thread t reads row t (stride IN words -- conflict-free for
// the odd widths, 64 / 128-bit accesses for the even ones), computes in float64 registers and
// writes row t of the output block.  Persistent CTAs, blocks strided by gridDim.x.
namespace bulk {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void store(void *dst, uint32_t src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// at most one bulk store may still be reading shared memory
__device__ __forceinline__ void store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make the threads' shared-memory writes visible to the async proxy (the bulk store)
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// row t of a block: W words at stride W, widest access the row alignment allows
template <typename T, int W> __device__ __forceinline__ void read_row(const T *base, int t, T (&v)[W])
{
    const T *r = base + t * W;
    if constexpr (sizeof(T) == 4 && W % 4 == 0) {
#pragma unroll
        for (int k = 0; k < W; k += 4) {
            const float4 q = *reinterpret_cast<const float4 *>(r + k);
            v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w;
        }
    } else if constexpr (W % 2 == 0 && sizeof(T) == 4) {
#pragma unroll
        for (int k = 0; k < W; k += 2) {
            const float2 q = *reinterpret_cast<const float2 *>(r + k);
            v[k] = q.x; v[k + 1] = q.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < W; ++k) v[k] = r[k];
    }
}
template <typename T, int W> __device__ __forceinline__ void write_row(T *base, int t, const T (&v)[W])
{
    T *r = base + t * W;
    if constexpr (sizeof(T) == 4 && W % 4 == 0) {
#pragma unroll
        for (int k = 0; k < W; k += 4)
            *reinterpret_cast<float4 *>(r + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
    } else if constexpr (W % 2 == 0 && sizeof(T) == 4) {
#pragma unroll
        for (int k = 0; k < W; k += 2) *reinterpret_cast<float2 *>(r + k) = make_float2(v[k], v[k + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < W; ++k) r[k] = v[k];
    }
}
}  // namespace bulk

constexpr int kInStages = 3, kOutStages = 2;

template <typename T, typename F>
__global__ void __launch_bounds__(kRows) rows_bulk_kernel(const T *__restrict__ in, T *__restrict__ out,
                                                          long long n_blocks, const __grid_constant__ F f)
{
    constexpr int IN = F::IN, OUT = F::OUT;
    constexpr uint32_t in_bytes = kRows * IN * sizeof(T), out_bytes = kRows * OUT * sizeof(T);
    __shared__ __align__(128) T s_in[kInStages][kRows * IN];
    __shared__ __align__(128) T s_out[kOutStages][kRows * OUT];
    __shared__ __align__(8) uint64_t s_bar[kInStages];
    const int tid = threadIdx.x;
    const uint32_t bar0 = bulk::smem_u32(s_bar);
    if (tid == 0) {
        for (int i = 0; i < kInStages; ++i) bulk::mbar_init(bar0 + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long first = blockIdx.x, step = gridDim.x;
    if (tid == 0)
        for (int j = 0; j < kInStages - 1; ++j) {
            const long long b = first + j * step;
            if (b < n_blocks) {
                bulk::mbar_expect_tx(bar0 + 8 * j, in_bytes);
                bulk::load(bulk::smem_u32(s_in[j]), in + b * (kRows * IN), in_bytes, bar0 + 8 * j);
            }
        }
    int slot = 0, oslot = 0;
    uint32_t phase = 0;
    for (long long b = first; b < n_blocks; b += step) {
        if (tid == 0) {
            // the slot consumed in the previous iteration is free: refill it two blocks ahead
            const long long nb = b + (kInStages - 1) * step;
            const int ns = slot == 0 ? kInStages - 1 : slot - 1;
            if (nb < n_blocks) {
                bulk::mbar_expect_tx(bar0 + 8 * ns, in_bytes);
                bulk::load(bulk::smem_u32(s_in[ns]), in + nb * (kRows * IN), in_bytes, bar0 + 8 * ns);
            }
            bulk::store_wait_read1();  // the store issued two iterations ago has left s_out[oslot]
        }
        bulk::mbar_wait(bar0 + 8 * slot, phase);
        T vin[IN];
        bulk::read_row<T, IN>(s_in[slot], tid, vin);
        __syncthreads();  // s_in[slot] is consumed; s_out[oslot] is free (thread 0 waited above)
        double a[IN], r[OUT];
#pragma unroll
        for (int k = 0; k < IN; ++k) a[k] = (double)vin[k];
        f.template operator()<T>(a, r);
        T vout[OUT];
#pragma unroll
        for (int k = 0; k < OUT; ++k) vout[k] = (T)r[k];
        bulk::write_row<T, OUT>(s_out[oslot], tid, vout);
        bulk::fence_async();
        __syncthreads();
        if (tid == 0) bulk::store(out + b * (kRows * OUT), bulk::smem_u32(s_out[oslot]), out_bytes);
        oslot ^= 1;
        if (++slot == kInStages) {
            slot = 0;
            phase ^= 1;
        }
    }
    if (tid == 0) bulk::store_wait_all();
}

template <typename F>
int launch_rows(const void *in, void *out, int64_t n, int dtype, const F &f, cudaStream_t stream,
                const char *name)
{
    int rc = bevk_require_device();
    if (rc) return rc;
    if (n < 0) BEVK_FAIL(BEVK_E_ARG, "%s: n must be >= 0", name);
    if (n == 0) return BEVK_OK;  // empty input: no launch
    if (!in || !out) BEVK_FAIL(BEVK_E_ARG, "%s: null buffer", name);
    if (dtype != BEVK_F32 && dtype != BEVK_F64)
        BEVK_FAIL(BEVK_E_ARG, "%s: dtype must be BEVK_F32 or BEVK_F64, got %d", name, dtype);
    const size_t es = dtype == BEVK_F32 ? 4 : 8;
    // whole 256-row blocks go through the bulk-copy pipeline when the buffers are 16-byte aligned
    // (float32 only: the float64 rings would not fit 4 CTAs per SM); the rest -- the tail rows, or
    // everything for unaligned buffers -- goes through the element-wise staging kernel
    long long bulk_blocks = 0;
    if (dtype == BEVK_F32 && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0) bulk_blocks = n / kRows;
    if (bulk_blocks > 0) {
        // 5 persistent CTAs per SM measured best (3: 74 %, 4: 78 %, 5: 81 %, 6: 80 %, 7: 77 % of HBM peak)
        const long long max_grid = (long long)bevk_sm_count() * 5;
        const int grid = (int)(bulk_blocks < max_grid ? bulk_blocks : max_grid);
        rows_bulk_kernel<float, F><<<grid, kRows, 0, stream>>>((const float *)in, (float *)out, bulk_blocks, f);
        BEVK_CUDA(cudaGetLastError());
    }
    const long long done = bulk_blocks * kRows;
    if (done < n) {
        const long long rest = n - done;
        const char *in2 = (const char *)in + (size_t)done * F::IN * es;
        char *out2 = (char *)out + (size_t)done * F::OUT * es;
        const long long n_blocks = (rest + kRows - 1) / kRows;
        // 8 resident blocks of 256 threads per SM; a whole number of waves when the batch is large
        const long long max_grid = (long long)bevk_sm_count() * 8;
        const int grid = (int)(n_blocks < max_grid ? n_blocks : max_grid);
        if (dtype == BEVK_F32)
            rows_kernel<float, F><<<grid, kRows, 0, stream>>>((const float *)in2, (float *)out2, rest, f);
        else
            rows_kernel<double, F><<<grid, kRows, 0, stream>>>((const double *)in2, (double *)out2, rest, f);
        BEVK_CUDA(cudaGetLastError());
    }
    return BEVK_OK;
}

int check_mode(int mode, const char *name)
{
    if (mode != BEVK_MODE_BEV && mode != BEVK_MODE_WORLD)
        BEVK_FAIL(BEVK_E_ARG, "%s: mode must be BEVK_MODE_BEV or BEVK_MODE_WORLD, got %d", name, mode);
    return BEVK_OK;
}

void set_h(Mat3 &m, const double *H)
{
    for (int i = 0; i < 9; ++i) m.h[i] = H ? H[i] : (i % 4 == 0 ? 1.0 : 0.0);
}

}  // namespace

extern "C" {

int bevk_pts_project(const void *pts, void *out, int64_t n, int dim, int dtype, const double H[9],
                     void *stream)
{
    if (!H) BEVK_FAIL(BEVK_E_ARG, "bevk_pts_project: H is null");
    if (dim == 2) {
        FnPts2 f;
        set_h(f.H, H);
        return launch_rows(pts, out, n, dtype, f, (cudaStream_t)stream, "bevk_pts_project");
    }
    if (dim == 3) {
        FnPts3 f;
        set_h(f.H, H);
        return launch_rows(pts, out, n, dtype, f, (cudaStream_t)stream, "bevk_pts_project");
    }
    BEVK_FAIL(BEVK_E_ARG, "bevk_pts_project: dim must be 2 or 3, got %d", dim);
}

int bevk_xywhr2xyxy(const void *xywhr, void *xy8, int64_t n, int mode, int dtype, const double *H,
                    void *stream)
{
    if (int rc = check_mode(mode, "bevk_xywhr2xyxy")) return rc;
    FnCorners f;
    set_h(f.H, H);
    f.mode = mode;
    f.use_h = H != nullptr;
    return launch_rows(xywhr, xy8, n, dtype, f, (cudaStream_t)stream, "bevk_xywhr2xyxy");
}

int bevk_xy82xywhr(const void *xy8, void *xywhr, int64_t n, int mode, int dtype, const double *H,
                   void *stream)
{
    if (int rc = check_mode(mode, "bevk_xy82xywhr")) return rc;
    FnFromCorners f;
    set_h(f.H, H);
    f.mode = mode;
    f.use_h = H != nullptr;
    return launch_rows(xy8, xywhr, n, dtype, f, (cudaStream_t)stream, "bevk_xy82xywhr");
}

int bevk_rbox_world_bev(const void *xywhr_in, void *xywhr_out, int64_t n, int src_mode, int dtype,
                        const double H[9], void *stream);

// shared by bevk_rbox_world_bev / bevk_rboxtt_world_bev: normalise H, check the reference's asserts
static int make_similarity(const double H[9], int src_mode, const char *name, FnSimilarity &f)
{
    if (int rc = check_mode(src_mode, name)) return rc;
    if (!H) BEVK_FAIL(BEVK_E_ARG, "%s: H is null", name);
    for (int i = 0; i < 9; ++i) f.H.h[i] = H[i] / H[8];  // rbox_torch.py:139
    if (!(fabs(f.H.h[6]) + fabs(f.H.h[7]) < 1e-5))
        BEVK_FAIL(BEVK_E_AFFINE, "%s: H is not affine (|H20|+|H21| = %g)", name,
                  fabs(f.H.h[6]) + fabs(f.H.h[7]));
    const double s0 = sqrt(f.H.h[0] * f.H.h[0] + f.H.h[1] * f.H.h[1]);
    const double s1 = sqrt(f.H.h[3] * f.H.h[3] + f.H.h[4] * f.H.h[4]);
    if (!(fabs(s0 - s1) < 1e-5))
        BEVK_FAIL(BEVK_E_AFFINE, "%s: H is not a similarity (scales %g vs %g)", name, s0, s1);
    f.scale = s0;
    f.src_mode = src_mode;
    return BEVK_OK;
}

int bevk_rbox_world_bev(const void *xywhr_in, void *xywhr_out, int64_t n, int src_mode, int dtype,
                        const double H[9], void *stream)
{
    FnSimilarity f;
    if (int rc = make_similarity(H, src_mode, "bevk_rbox_world_bev", f)) return rc;
    return launch_rows(xywhr_in, xywhr_out, n, dtype, f, (cudaStream_t)stream, "bevk_rbox_world_bev");
}

int bevk_rboxtt_world_bev(const void *in, void *out, int64_t n, int src_mode, int dtype, const double H[9],
                          void *stream)
{
    FnRboxTT f;
    if (int rc = make_similarity(H, src_mode, "bevk_rboxtt_world_bev", f.sim)) return rc;
    return launch_rows(in, out, n, dtype, f, (cudaStream_t)stream, "bevk_rboxtt_world_bev");
}

int bevk_rbox_zt2tt_world(const void *in, void *out, int64_t n, int dtype, const double K[9],
                          const double Rt[12], void *stream)
{
    if (!K || !Rt) BEVK_FAIL(BEVK_E_ARG, "bevk_rbox_zt2tt_world: K / Rt is null");
    FnZt2tt f;
    double Hcw[9];  // homo_from_KRt: K [r1 r2 t]
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            f.K[3 * i + j] = K[3 * i + j];
            f.R[3 * i + j] = Rt[4 * i + j];
        }
        f.t[i] = Rt[4 * i + 3];
    }
    for (int i = 0; i < 3; ++i) {
        const int cols[3] = {0, 1, 3};
        for (int j = 0; j < 3; ++j)
            Hcw[3 * i + j] = K[3 * i] * Rt[cols[j]] + K[3 * i + 1] * Rt[4 + cols[j]] + K[3 * i + 2] * Rt[8 + cols[j]];
    }
    if (!bevk_invert3x3(Hcw, f.Hwc.h)) BEVK_FAIL(BEVK_E_ARG, "bevk_rbox_zt2tt_world: K [r1 r2 t] is singular");
    return launch_rows(in, out, n, dtype, f, (cudaStream_t)stream, "bevk_rbox_zt2tt_world");
}

int bevk_angle_world_bev(const void *yaw_in, void *yaw_out, int64_t n, int src_mode, int dtype,
                         const double H[9], void *stream)
{
    if (int rc = check_mode(src_mode, "bevk_angle_world_bev")) return rc;
    if (!H) BEVK_FAIL(BEVK_E_ARG, "bevk_angle_world_bev: H is null");
    FnAngle f;
    f.h00 = H[0];
    f.h01 = H[1];
    f.h10 = H[3];
    f.h11 = H[4];
    f.src_mode = src_mode;
    return launch_rows(yaw_in, yaw_out, n, dtype, f, (cudaStream_t)stream, "bevk_angle_world_bev");
}

int bevk_dist_world_bev(const void *dist_in, void *dist_out, int64_t n, int dtype, const double H[9],
                        void *stream)
{
    if (!H) BEVK_FAIL(BEVK_E_ARG, "bevk_dist_world_bev: H is null");
    // the numpy twin compares COLUMN norms (rbox.py:154-156)
    const double s0 = sqrt(H[0] * H[0] + H[3] * H[3]), s1 = sqrt(H[1] * H[1] + H[4] * H[4]);
    if (!(fabs(s0 - s1) < 1e-5))
        BEVK_FAIL(BEVK_E_AFFINE, "bevk_dist_world_bev: H is not a similarity (scales %g vs %g)", s0, s1);
    FnDist f;
    f.scale = s0;
    return launch_rows(dist_in, dist_out, n, dtype, f, (cudaStream_t)stream, "bevk_dist_world_bev");
}

int bevk_xywhr2xyvec(const void *xywhr, void *xyvec, int64_t n, int mode, int dtype, void *stream)
{
    if (int rc = check_mode(mode, "bevk_xywhr2xyvec")) return rc;
    FnXyvec f;
    f.mode = mode;
    return launch_rows(xywhr, xyvec, n, dtype, f, (cudaStream_t)stream, "bevk_xywhr2xyvec");
}

int bevk_xy82xyvec(const void *xy8, void *xyvec, int64_t n, int dtype, void *stream)
{
    FnXy8vec f;
    return launch_rows(xy8, xyvec, n, dtype, f, (cudaStream_t)stream, "bevk_xy82xyvec");
}

int bevk_v2yaw(const void *v, void *yaw, int64_t n, int mode, int dtype, void *stream)
{
    if (int rc = check_mode(mode, "bevk_v2yaw")) return rc;
    FnV2yaw f;
    f.mode = mode;
    return launch_rows(v, yaw, n, dtype, f, (cudaStream_t)stream, "bevk_v2yaw");
}

int bevk_yaw2v(const void *yaw, void *v, int64_t n, int mode, int dtype, void *stream)
{
    if (int rc = check_mode(mode, "bevk_yaw2v")) return rc;
    FnYaw2v f;
    f.mode = mode;
    return launch_rows(yaw, v, n, dtype, f, (cudaStream_t)stream, "bevk_yaw2v");
}

int bevk_yaw2mat(const void *yaw, void *mat, int64_t n, int mode, int dtype, void *stream)
{
    if (int rc = check_mode(mode, "bevk_yaw2mat")) return rc;
    FnYaw2mat f;
    f.mode = mode;
    return launch_rows(yaw, mat, n, dtype, f, (cudaStream_t)stream, "bevk_yaw2mat");
}

static int host_roundtrip(const float *in, float *out, int64_t n, int in_w, int out_w, int which,
                          int mode, const double *H)
{
    int rc = bevk_require_device();
    if (rc) return rc;
    if (n == 0) return BEVK_OK;
    if (!in || !out) BEVK_FAIL(BEVK_E_ARG, "host projection: null buffer");
    float *d_in = nullptr, *d_out = nullptr;
    cudaStream_t st;
    BEVK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    BEVK_CUDA(cudaMallocAsync(&d_in, (size_t)n * in_w * sizeof(float), st));
    BEVK_CUDA(cudaMallocAsync(&d_out, (size_t)n * out_w * sizeof(float), st));
    BEVK_CUDA(cudaMemcpyAsync(d_in, in, (size_t)n * in_w * sizeof(float), cudaMemcpyHostToDevice, st));
    rc = which == 0 ? bevk_xywhr2xyxy(d_in, d_out, n, mode, BEVK_F32, H, st)
                    : bevk_xy82xywhr(d_in, d_out, n, mode, BEVK_F32, H, st);
    if (rc == BEVK_OK) {
        cudaError_t e = cudaMemcpyAsync(out, d_out, (size_t)n * out_w * sizeof(float),
                                        cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            bevk_set_error("host projection copy-back failed: %s", cudaGetErrorString(e));
            rc = BEVK_E_CUDA;
        }
    }
    cudaFreeAsync(d_in, st);
    cudaFreeAsync(d_out, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

int bevk_xywhr2xyxy_host(const float *xywhr, float *xy8, int64_t n, int mode, const double *H)
{
    return host_roundtrip(xywhr, xy8, n, 5, 8, 0, mode, H);
}

int bevk_xy82xywhr_host(const float *xy8, float *xywhr, int64_t n, int mode, const double *H)
{
    return host_roundtrip(xy8, xywhr, n, 8, 5, 1, mode, H);
}

}  // extern "C"
