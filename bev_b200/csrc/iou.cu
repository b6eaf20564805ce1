// iou.cu -- rotated-box IoU matrix (sm_100a): the O(N * M) step of the reference tracker's
// association, iou_batch_rbox (bev/tracker/rbox_tracker.py:87-92, used at :383-405), which calls
// the third-party d3d.box.box2d_iou(boxes1, boxes2, method="rbox") on [x, y, w, h, r + pi/2].
//
// Box convention (d3d's published one, also cv2.RotatedRect's): centre (x, y), side w along
// (cos r, sin r), side h along (-sin r, cos r); IoU = |A n B| / (|A| + |B| - |A n B|).
//
// One thread per (i, j) pair.  |A n B| is evaluated without building the clipped polygon: B's
// corners are taken into A's frame, where A is the axis-aligned rectangle [-hu, hu] x [-hv, hv],
// and the closed curve obtained by clamping B's boundary onto that rectangle encloses exactly
// A n B (the clamp is the nearest-point projection onto a convex set: it maps B onto A n B plus
// zero-area pieces of A's boundary).  Each edge is cut at the <= 4 parameters where it crosses the
// lines u = +-hu, v = +-hv (a 5-exchange sorting network, branch free), between two cuts the
// clamped curve is a straight segment, and the shoelace sum over <= 5 segments per edge is exact.
// The result is a continuous function of the inputs -- coincident edges (a track and the
// detection it came from) need no special case -- and lives entirely in registers.
//
// Pairs whose centres are further apart than the two circumscribed radii are answered 0 before any
// of that.  Per 32 x 32 tile of the matrix the 64 boxes are decoded once (sincos, half extents) into shared
// memory.  float32 or float64 boxes in, the same type out; arithmetic in float64.
#include "bevk_common.cuh"

namespace {

struct Rect {
    double cx, cy, ux, uy, hu, hv;
    double rad;  // circumscribed radius, slightly enlarged: pairs further apart cannot overlap
};

template <typename T>
__device__ __forceinline__ Rect load_rect(const T *row, double yaw_offset)
{
    Rect r;
    r.cx = (double)row[0];
    r.cy = (double)row[1];
    r.hu = 0.5 * fabs((double)row[2]);
    r.hv = 0.5 * fabs((double)row[3]);
    sincos((double)row[4] + yaw_offset, &r.uy, &r.ux);
    r.rad = sqrt(r.hu * r.hu + r.hv * r.hv) * (1.0 + 1e-9);
    return r;
}

__device__ __forceinline__ double clampd(double v, double h) { return fmin(fmax(v, -h), h); }

__device__ __forceinline__ void cswap(double &a, double &b)
{
    const double lo = fmin(a, b), hi = fmax(a, b);
    a = lo;
    b = hi;
}

// parameter in [0, 1] at which p + t d reaches the level b (0 when the edge is parallel to it)
__device__ __forceinline__ double crossing(double b, double p, double inv_d)
{
    const double t = (b - p) * inv_d;
    return fmin(fmax(t, 0.0), 1.0);  // fmax(NaN, 0) = 0
}

// shoelace contribution of the clamped image of the edge p -> p + d (local coordinates of A)
__device__ __forceinline__ double clamped_edge(double pu, double pv, double du, double dv, double hu, double hv)
{
    const double iu = du != 0.0 ? 1.0 / du : 0.0, iv = dv != 0.0 ? 1.0 / dv : 0.0;
    double t0 = crossing(-hu, pu, iu), t1 = crossing(hu, pu, iu);
    double t2 = crossing(-hv, pv, iv), t3 = crossing(hv, pv, iv);
    cswap(t0, t1);
    cswap(t2, t3);
    cswap(t0, t2);
    cswap(t1, t3);
    cswap(t1, t2);
    double acc = 0.0;
    double ax = clampd(pu, hu), ay = clampd(pv, hv);
    const double ts[5] = {t0, t1, t2, t3, 1.0};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double bx = clampd(fma(ts[k], du, pu), hu), by = clampd(fma(ts[k], dv, pv), hv);
        acc += ax * by - ay * bx;
        ax = bx;
        ay = by;
    }
    return acc;
}

// can the two rectangles overlap at all?  (centres closer than the circumscribed radii)
__device__ __forceinline__ bool may_overlap(const Rect &a, const Rect &b)
{
    const double rx = b.cx - a.cx, ry = b.cy - a.cy, reach = a.rad + b.rad;
    return rx * rx + ry * ry <= reach * reach;
}

__device__ __forceinline__ double pair_iou(const Rect &a, const Rect &b)
{
    // B's centre and half-axis vectors in A's frame
    const double rx = b.cx - a.cx, ry = b.cy - a.cy;
    const double cu = rx * a.ux + ry * a.uy, cv = -rx * a.uy + ry * a.ux;
    const double c = b.ux * a.ux + b.uy * a.uy, s = -b.ux * a.uy + b.uy * a.ux;  // cos / sin of r_b - r_a
    const double wu = c * b.hu, wv = s * b.hu;    // half side along B's u axis
    const double zu = -s * b.hv, zv = c * b.hv;   // half side along B's v axis
    // CCW corners: c - w - z, c + w - z, c + w + z, c - w + z
    const double p0u = cu - wu - zu, p0v = cv - wv - zv;
    const double p1u = cu + wu - zu, p1v = cv + wv - zv;
    const double p2u = cu + wu + zu, p2v = cv + wv + zv;
    const double p3u = cu - wu + zu, p3v = cv - wv + zv;
    double twice = clamped_edge(p0u, p0v, 2.0 * wu, 2.0 * wv, a.hu, a.hv);
    twice += clamped_edge(p1u, p1v, 2.0 * zu, 2.0 * zv, a.hu, a.hv);
    twice += clamped_edge(p2u, p2v, -2.0 * wu, -2.0 * wv, a.hu, a.hv);
    twice += clamped_edge(p3u, p3v, -2.0 * zu, -2.0 * zv, a.hu, a.hv);
    const double inter = fmax(0.5 * twice, 0.0);
    const double uni = 4.0 * (a.hu * a.hv + b.hu * b.hv) - inter;
    return uni > 0.0 ? inter / uni : 0.0;
}

// A block owns a 32 x 32 tile of the matrix.  Phase 1: every thread tests its 4 pairs with the
// cheap reach test, writes 0 for the ones that cannot overlap and queues the others in shared
// memory.  Phase 2: the queued pairs -- a few per tile in a tracking scene -- are evaluated by
// consecutive threads, so the expensive path runs with full warps instead of a lane here and there.
template <typename T>
__global__ void __launch_bounds__(256) iou_matrix_kernel(const T *__restrict__ b1, long long n, int stride1,
                                                         const T *__restrict__ b2, long long m, int stride2,
                                                         T *__restrict__ out, double yaw_offset)
{
    __shared__ Rect s_a[32], s_b[32];
    __shared__ unsigned short s_queue[1024];
    __shared__ int s_count;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const long long i0 = (long long)blockIdx.y * 32, j0 = (long long)blockIdx.x * 32;
    if (tid < 32) {
        if (j0 + tid < m) s_b[tid] = load_rect(b2 + (j0 + tid) * stride2, yaw_offset);
    } else if (tid < 64) {
        if (i0 + tid - 32 < n) s_a[tid - 32] = load_rect(b1 + (i0 + tid - 32) * stride1, yaw_offset);
    } else if (tid == 64) {
        s_count = 0;
    }
    __syncthreads();
    const long long j = j0 + threadIdx.x;
    if (j < m) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int li = threadIdx.y + 8 * k;
            if (i0 + li >= n) continue;
            if (may_overlap(s_a[li], s_b[threadIdx.x]))
                s_queue[atomicAdd(&s_count, 1)] = (unsigned short)(li * 32 + threadIdx.x);
            else
                out[(i0 + li) * m + j] = (T)0;
        }
    }
    __syncthreads();
    const int count = s_count;
    for (int q = tid; q < count; q += 256) {
        const int li = s_queue[q] >> 5, lj = s_queue[q] & 31;
        out[(i0 + li) * m + j0 + lj] = (T)pair_iou(s_a[li], s_b[lj]);
    }
}

}  // namespace

extern "C" int bevk_rbox_iou_matrix(const void *boxes1, int64_t n, int stride1, const void *boxes2, int64_t m,
                                    int stride2, void *out, int dtype, double yaw_offset, void *stream)
{
    if (n < 0 || m < 0) BEVK_FAIL(BEVK_E_ARG, "rbox_iou_matrix: box counts must be >= 0");
    if (stride1 < 5 || stride2 < 5)
        BEVK_FAIL(BEVK_E_ARG, "rbox_iou_matrix: rows need at least 5 columns [x, y, w, h, r] (strides %d, %d)",
                  stride1, stride2);
    if (dtype != BEVK_F32 && dtype != BEVK_F64)
        BEVK_FAIL(BEVK_E_ARG, "rbox_iou_matrix: dtype must be float32 or float64 (code %d)", dtype);
    int rc = bevk_require_device();
    if (rc) return rc;
    if (n == 0 || m == 0) return BEVK_OK;
    if (!boxes1 || !boxes2 || !out) BEVK_FAIL(BEVK_E_ARG, "rbox_iou_matrix: null buffer");
    const long long gx = (m + 31) / 32, gy = (n + 31) / 32;
    if (gy > 65535) BEVK_FAIL(BEVK_E_ARG, "rbox_iou_matrix: more than 2097120 rows in one call");
    const dim3 grid((unsigned)gx, (unsigned)gy, 1), block(32, 8, 1);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == BEVK_F32)
        iou_matrix_kernel<float><<<grid, block, 0, st>>>((const float *)boxes1, n, stride1, (const float *)boxes2,
                                                         m, stride2, (float *)out, yaw_offset);
    else
        iou_matrix_kernel<double><<<grid, block, 0, st>>>((const double *)boxes1, n, stride1,
                                                          (const double *)boxes2, m, stride2, (double *)out,
                                                          yaw_offset);
    BEVK_CUDA(cudaGetLastError());
    return BEVK_OK;
}
