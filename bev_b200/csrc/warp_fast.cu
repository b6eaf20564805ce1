// warp_fast.cu -- staged perspective warp for 3-channel frames (the BASELINE hot path:
// cv2.warpPerspective on BGR video frames, reference vis_homo.py:85-91), sm_100a.
//
// The kernel is written against a pixel-format policy PX (warp_u8c3.cuh: uint8 x 3,
// warp_f16c3.cuh: float16 x 3) that owns the per-pixel registers, the window loads, the
// interpolation and the store packing.
//
// Work item = (chunk of frames, homography group, dst tile).  A persistent CTA of 8 warps pulls
// items from a shared counter; a tile is 1024 dst pixels, 4 per thread, as 128x8, 64x16 or 32x32
// (picked per launch on the host, pick_tile_shape):
//
//   1. set-up, once per item: the exact FP64 coordinate pipeline of cv2 (bevk_map_pixel_xb) gives
//      each pixel its 2x2 source window, and frame-invariant registers are derived from it -- the
//      shared-memory offset of the window, a funnel-shift amount that byte-aligns it and the tap
//      weights laid out as the policy's operands.  Out-of-image taps get weight 0 and a clamped
//      address, so the frame loop has no border branches.  A block reduction yields the tile's
//      source bounding box.
//   2. frame loop: the bounding box of the next frames is fetched by the TMA unit as 2-D tensor
//      boxes (cp.async.bulk.tensor.2d, at most 3 requests per frame and tile, completion on an
//      mbarrier) into a 2-, 4- or 8-deep shared-memory ring of up to 4 frames per stage while the
//      warps interpolate the current frames out of shared memory.  The producer role rotates over
//      the warps (one elected lane); full[] / empty[] mbarriers are the only synchronisation in
//      the loop.  The box shape is picked per tile from a menu of 224 tensor maps over the source
//      batch viewed as a [frames*rows][row_bytes/4] uint32 (or /8 uint64) matrix: widths
//      64..2048 B, heights 1..32 rows.  A first version issued one cp.async.bulk per source row:
//      the TMA unit retired only one such ~300-byte request per ~70 cycles per SM, which capped
//      the kernel at 33 % of the HBM roofline (profiles/r01_fast_v1_*).
//   3. stores: the policy packs the lanes' pixels into words (one shuffle + PRMT for uint8) and
//      writes fully coalesced row segments with streaming stores.
//
// Tiles whose box does not stage (too wide / tall for the menu, larger than half the ring): a few
// of them fall back, inside the kernel, to aligned 32-bit global loads with the same arithmetic;
// when the host predicts 2-35 % of the tiles the launch is SPLIT -- the kernel only marks them in
// BevkWarpParams::hard and the direct-gather kernel (warp_generic.cu) follows on the stream over
// the marked tiles; above that the whole launch goes to the direct-gather kernel.
//
// HBM traffic per frame is the touched source footprint (bounding boxes overlap by a row /
// column and are re-served by L2: tiles are walked column-major so that neighbours run
// concurrently and near-/far-field tiles mix) plus the output, i.e. about 1.1x the algorithmic
// bytes of SURVEY.md 8d.  The binding resource is the shared-memory data pipe.
#include "bevk_common.cuh"
#include "warp_u8c3.cuh"
#include "warp_f16c3.cuh"

#include <cuda.h>  // CUtensorMap + enums only; the encoder is fetched through the runtime
#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int kThreads = 256, kWarps = kThreads / 32;
// Tile shapes.  A CTA always owns 1024 dst pixels, 4 per thread; SEGS = 32-pixel segments per tile
// row: 4 -> 128x8 (warp w owns tile row w), 2 -> 64x16 (rows 2w, 2w+1), 1 -> 32x32 (rows 4w..4w+3).
// Narrower tiles bound the width of the source box where the map minifies horizontally.
__host__ __device__ constexpr int tile_w(int segs) { return 32 * segs; }
__host__ __device__ constexpr int tile_h(int segs) { return kWarps * (4 / segs); }
// Tensor-map menu: one map per box shape.
//   narrow boxes: 12 widths (64..512 B in steps of 64, 640..1024 B in steps of 128) x 16 heights
//                 (1..8, 10..16 in steps of 2, 20..32 in steps of 4), uint32 elements;
//   wide boxes  :  4 widths (1280..2048 B in steps of 256) x 8 heights (4..32 in steps of 4),
//                 uint64 elements (a box row holds at most 256 elements).
// A tile's bounding box is fetched with one box whenever it has at most 32 rows (taller ones take
// up to kMaxBoxes).
constexpr int kNarrowW = 12, kNarrowH = 16, kWideW = 4, kWideH = 8;
constexpr int kMapCount = kNarrowW * kNarrowH + kWideW * kWideH;
constexpr int kMaxBoxWidth = 2048, kMaxBoxHeight = 32;
constexpr int kMaxBoxes = 3;
constexpr int kMaxStageFrames = 4;              // frames that share one ring stage / mbarrier phase
constexpr int kBarBytes = 256;                  // 28 mbarriers: ring depths 2, 4 and 8
constexpr int kPrefetchAhead = 2;               // depth-2 rings: L2 prefetch runs this many stages ahead
constexpr int kTailSlack = 64;                  // window words may run a few bytes past a stage

__host__ __device__ constexpr int map_width(int wi)
{
    return wi < 8 ? 64 * (wi + 1) : (wi < 12 ? 512 + 128 * (wi - 7) : 1024 + 256 * (wi - 11));
}
__host__ __device__ constexpr bool map_is_wide(int wi) { return wi >= kNarrowW; }
// heights of the narrow menu; the wide menu has 4 * (hi + 1)
__host__ __device__ constexpr int map_height(int hi)
{
    return hi < 8 ? hi + 1 : (hi < 12 ? 10 + 2 * (hi - 8) : 20 + 4 * (hi - 12));
}
// smallest menu entry that covers `bytes` / `rows`
__host__ __device__ __forceinline__ int map_width_index(int bytes)
{
    return bytes <= 512 ? (bytes + 63) / 64 - 1
                        : (bytes <= 1024 ? 7 + (bytes - 512 + 127) / 128 : 11 + (bytes - 1024 + 255) / 256);
}
__host__ __device__ __forceinline__ int map_height_index(int rows)
{
    return rows <= 8 ? rows - 1 : (rows <= 16 ? 8 + (rows - 9) / 2 : 12 + (rows - 17) / 4);
}
// box height the menu offers for `rows` (<= kMaxBoxHeight) at width index wi, and its map index
__host__ __device__ __forceinline__ int box_height(int wi, int rows)
{
    return map_is_wide(wi) ? (rows + 3) / 4 * 4 : map_height(map_height_index(rows));
}
__host__ __device__ __forceinline__ int box_map_index(int wi, int rows)
{
    return map_is_wide(wi) ? kNarrowW * kNarrowH + (wi - kNarrowW) * kWideH + (rows + 3) / 4 - 1
                           : wi * kNarrowH + map_height_index(rows);
}

struct WarpFastMaps {
    CUtensorMap m[kMapCount];
};

// Frame chunks of one launch.  Chunk k of a group with `count` frames covers frames
// [count * cum[k] >> 16, count * cum[k + 1] >> 16): the chunks shrink towards the end of the item
// list, so that the CTAs, which pull items from a shared counter, finish almost together.
constexpr int kMaxChunks = 12;
struct ChunkPlan {
    int n_chunks;
    uint32_t cum[kMaxChunks + 1];  // cum[0] = 0, cum[n_chunks] = 65536
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// 2-D tensor box global -> shared, completion counted in bytes on an mbarrier (the TMA path)
__device__ __forceinline__ void tma_box_g2s(uint32_t dst, const CUtensorMap *map, int x, int y,
                                            uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y)
        : "memory");
}
// the same box, global -> L2 only (no shared memory, no completion)
__device__ __forceinline__ void tma_box_prefetch(const CUtensorMap *map, int x, int y)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// Stops the compiler from re-deriving a loop-invariant value inside the frame loop.
__device__ __forceinline__ void keep(uint32_t &v) { asm volatile("" : "+r"(v)); }

// What the elected producer thread needs to fetch one frame's bounding box.
struct BoxPlan {
    int n_boxes;
    int x;                  // first column, in elements of the map (TMA wants 16-byte aligned box rows)
    int y0;                 // first source row of the box
    uint32_t bytes;         // sum of the box sizes (the mbarrier's transaction count per frame)
    int frame0, frame_step; // source frame of the item's frame i = frame0 + i * frame_step
    int src_h;
    int n_frames, fps;      // frames of the item / per ring stage
    uint32_t ring, full0;   // shared-memory addresses: ring, first barrier of the set in use
    uint32_t stride, frame_bytes, pitch;  // bytes per stage / staged frame / staged row
    int map_idx[kMaxBoxes];
    int row[kMaxBoxes];     // row offset of each box inside the stage
};

// A decoded work item (thread 0 decodes the next one while the CTA works on the current one).
struct ItemDesc {
    int item;                 // >= total_items: no more work
    int gi, tile_x, tile_y;   // homography group, tile position
    int f0, n_frames;         // frames [f0, f0 + n_frames) of the group's run
    int first, stride;        // the run: frame index = first + f * stride
};

// The consumers' loop state (kept small so that the frame loop fits 64 registers); the elected
// producer lane reads what it needs from the BoxPlan in shared memory.
struct LoopCtx {
    uint32_t ring, full0;          // shared-memory addresses of the ring / first barrier of the set
    uint32_t stride;               // bytes per stage (fps frames)
    uint32_t frame_bytes;          // bytes per staged frame
    uint32_t pitch;                // bytes per staged row
    int fps;                       // frames per stage
    int n_frames;
    const WarpFastMaps *maps;
    const BoxPlan *plan;           // in shared memory
};

// Elected thread: pull the stage that starts at frame i0 of the item towards the SM.  TO_SMEM
// starts the copies into its ring slot (`use` is the running count of stages this CTA has pushed
// through the barrier set); otherwise the boxes are only prefetched into L2 so that the later
// copy does not wait on HBM (used by depth-2 rings, whose copies run just one stage ahead).
// Deliberately not inlined: it runs in one lane of one warp per stage and must not cost the
// frame loop registers.
template <int SLOG, bool TO_SMEM>
__device__ __noinline__ void produce(const BoxPlan *plan, const WarpFastMaps *maps, int i0,
                                     uint32_t use)
{
    const BoxPlan &pl = *plan;
    const int nf = min(pl.fps, pl.n_frames - i0);
    const int nb = pl.n_boxes;
    int y = (pl.frame0 + i0 * pl.frame_step) * pl.src_h + pl.y0;
    const int y_step = pl.frame_step * pl.src_h;
    if (TO_SMEM) {
        const uint32_t slot = use & ((1u << SLOG) - 1u);
        const uint32_t fb = pl.full0 + 8 * slot;
        // k-th fill of a slot waits for the (k-1)-th release; the first passes at once
        mbar_wait(fb + (8u << SLOG), ((use >> SLOG) & 1u) ^ 1u);
        mbar_expect_tx(fb, pl.bytes * nf);
        uint32_t sdst = pl.ring + slot * pl.stride;
#pragma unroll 1
        for (int f = 0; f < nf; ++f, sdst += pl.frame_bytes, y += y_step)
#pragma unroll 1
            for (int b = 0; b < nb; ++b)
                tma_box_g2s(sdst + pl.row[b] * pl.pitch, &maps->m[pl.map_idx[b]], pl.x,
                            y + pl.row[b], fb);
    } else {
#pragma unroll 1
        for (int f = 0; f < nf; ++f, y += y_step)
#pragma unroll 1
            for (int b = 0; b < nb; ++b)
                tma_box_prefetch(&maps->m[pl.map_idx[b]], pl.x, y + pl.row[b]);
    }
}

// Keep the ring (and, for depth-2 rings, the L2 prefetch window) full.  `done` = frames of the
// item consumed up to and including the current stage.  The producer role rotates over the warps
// (one elected lane each) so that no warp is slower than the others -- a fixed producer warp
// paces the whole CTA, because every warp waits on the stages it issues.
template <int SLOG>
__device__ __forceinline__ void feed(const LoopCtx &c, int done, uint32_t use, int lane, int warp)
{
    constexpr int ahead = 1 << (SLOG - 1);
    if (lane != 0) return;
    const int turn = ((int)use - warp) & (kWarps - 1);
    const int i_load = done + (ahead - 1) * c.fps;  // first frame of the stage `ahead` stages on
    if (turn == 0 && i_load < c.n_frames) produce<SLOG, true>(c.plan, c.maps, i_load, use + ahead);
    if (SLOG == 1 && turn == kWarps / 2 && i_load + kPrefetchAhead * c.fps < c.n_frames)
        produce<SLOG, false>(c.plan, c.maps, i_load + kPrefetchAhead * c.fps, 0);
}

// The frame loop of a staged item.  Ring depth 2^SLOG stages of c.fps frames each.  Returns the
// advanced stage counter.  PX = pixel format policy (warp_u8c3.cuh / warp_f16c3.cuh).
template <typename PX, bool LINEAR, int SLOG, int SEGS>
__device__ __forceinline__ uint32_t frame_loop(const LoopCtx &c, const typename PX::Reg (&px)[4],
                                               uint32_t use, uint8_t *d, const uint32_t d_step,
                                               const uint32_t row_bytes, const bool (&seg_ok)[4],
                                               const typename PX::Store st, const int tid)
{
    constexpr uint32_t smask = (1u << SLOG) - 1u;
    // Stages in flight besides the one being consumed: half the ring.  The other half is slack
    // between the warps -- a slot is refilled S/2 stages after its last use, so the refilling lane
    // practically never waits for a slower warp to release it.
    constexpr int ahead = 1 << (SLOG - 1);
    constexpr uint32_t kLast = PX::template last_word_offset<LINEAR>();
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < ahead && s * c.fps < c.n_frames; ++s)
            produce<SLOG, true>(c.plan, c.maps, s * c.fps, use + s);
        if (SLOG == 1)
            for (int s = ahead; s < ahead + kPrefetchAhead && s * c.fps < c.n_frames; ++s)
                produce<SLOG, false>(c.plan, c.maps, s * c.fps, 0);
    }

    int done = 0;
#pragma unroll 1
    while (done < c.n_frames) {
        const uint32_t slot = use & smask;
        const uint32_t fb = c.full0 + 8 * slot;
        mbar_wait(fb, (use >> SLOG) & 1u);
        uint32_t sa = c.ring + slot * c.stride;  // row 0 of the windows, first frame of the stage
        int nf = min(c.fps, c.n_frames - done);
        done += nf;
#pragma unroll 1
        for (; nf > 0; --nf, sa += c.frame_bytes, d += d_step) {
            const uint32_t sb = sa + c.pitch;  // row 1
            typename PX::Out P[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t w[8];
                const uint32_t a = px[k].addr + sa, b = px[k].addr + sb;
                PX::template load<LINEAR>(px[k], a, b, a + kLast, b + kLast, w, [](uint32_t ad) { return lds32(ad); });
                if (k == 3 && nf == 1) {
                    // every shared-memory read of this stage is issued: hand the slot back
                    __syncwarp();
                    if (lane == 0) mbar_arrive(fb + (8u << SLOG));
                    feed<SLOG>(c, done, use, lane, warp);
                }
                P[k] = PX::template math<LINEAR>(px[k], w);
            }
            // pack the lanes' pixels into words and store coalesced row segments
#pragma unroll
            for (int k = 0; k < 4; ++k)
                PX::store(d + PX::kSegBytes * (k % SEGS) + (k / SEGS) * row_bytes, P[k], seg_ok[k], st, lane);
        }
        ++use;
    }
    return use;
}

// MINB = CTAs per SM the register allocation is bounded for (4 -> 64 registers, 3 -> 80);
// SEGS = tile shape (see tile_w / tile_h).
template <typename PX, bool LINEAR, int MINB, int SEGS>
__global__ void __launch_bounds__(kThreads, MINB)
warp_fast_kernel(const __grid_constant__ BevkWarpParams p,
                      const __grid_constant__ WarpFastMaps maps,
                      const __grid_constant__ ChunkPlan plan, const int tiles_x, const int tiles_y,
                      const int total_items, const int ring_bytes, int *const next_item)
{
    extern __shared__ __align__(128) uint8_t smem[];
    // three barrier sets, one per ring depth S = 2, 4, 8: full[S] then empty[S], at byte
    // offsets 0, 32 and 96.  Each set keeps its own running stage counter (s_use) across items,
    // so barrier phases never need a reset when consecutive items use different depths.
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + ring_bytes);
    __shared__ int s_box[4];
    __shared__ int s_any;
    __shared__ uint32_t s_use[3];
    __shared__ ItemDesc s_item[2];  // current / next work item, decoded by thread 0
    __shared__ double s_M[BEVK_MAX_GROUPS][9];
    __shared__ BoxPlan s_plan;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ring = smem_u32(smem);
    const uint32_t bar0 = smem_u32(bars);
    if (tid == 0) {
        for (int sl = 1; sl <= 3; ++sl) {
            const int S = 1 << sl, base = 16 * S - 32;
            for (int i = 0; i < S; ++i) {
                mbar_init(bar0 + base + 8 * i, 1);                // full: the producer's expect_tx
                mbar_init(bar0 + base + 8 * (S + i), kWarps);     // empty: one arrival per warp
            }
            s_use[sl - 1] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_tiles = tiles_x * tiles_y;
    const uint8_t *src = (const uint8_t *)p.src;
    uint8_t *dst = (uint8_t *)p.dst;
    constexpr int kBpp = PX::kBpp;
    const int src_row_bytes = p.src_w * kBpp;
    const long long src_frame_bytes = (long long)src_row_bytes * p.src_h;
    const long long dst_frame_bytes = (long long)p.dst_w * p.dst_h * kBpp;

    // Items are (chunk, group, tile) with the chunk index slowest: long chunks first.  The first
    // gridDim.x items are taken by block index, the rest are pulled from *next_item.  Thread 0
    // decodes an item (integer divisions, parameter reads) one item ahead of its use.
    const int per_chunk = p.n_groups * n_tiles;
    auto decode = [&](int item, ItemDesc &o) {
        o.item = item;
        if (item >= total_items) return;
        const int chunk = item / per_chunk, rem = item - chunk * per_chunk;
        const int gi = rem / n_tiles, tile = rem - gi * n_tiles;
        // column-major walk: concurrently running CTAs cover whole tile columns, i.e. both the
        // magnified far field (store-heavy) and the minified near field (load-heavy)
        o.gi = gi;
        o.tile_x = tile / tiles_y;
        o.tile_y = tile - o.tile_x * tiles_y;
        const int count = p.g[gi].count;
        o.f0 = (int)(((unsigned long long)count * plan.cum[chunk]) >> 16);
        o.n_frames = (int)(((unsigned long long)count * plan.cum[chunk + 1]) >> 16) - o.f0;
        o.first = p.g[gi].first;
        o.stride = p.g[gi].stride;
    };
    for (int i = tid; i < p.n_groups * 9; i += kThreads) s_M[i / 9][i % 9] = p.g[i / 9].M[i % 9];
    if (tid == 0) {
        decode(blockIdx.x, s_item[0]);
        s_box[0] = s_box[2] = 1 << 30;
        s_box[1] = s_box[3] = -1;
        s_any = 0;
    }
    __syncthreads();
    const bool bw0_pow2 = (p.bw0 & (p.bw0 - 1)) == 0;

    int par = 0;
#pragma unroll 1
    while (true) {
        const int item = s_item[par].item;
        if (item >= total_items) break;
        const int gi = s_item[par].gi, tile_x = s_item[par].tile_x, tile_y = s_item[par].tile_y;
        const int f0 = s_item[par].f0, n_frames = s_item[par].n_frames;
        const int g_first = s_item[par].first, g_stride = s_item[par].stride;
        constexpr int kRowsPerWarp = 4 / SEGS;
        const int x0 = tile_x * tile_w(SEGS), y0 = tile_y * tile_h(SEGS) + warp * kRowsPerWarp;
        const uint32_t row_bytes = (uint32_t)p.dst_w * (uint32_t)kBpp;
        par ^= 1;

        // ---- 1. set-up ------------------------------------------------------------------------
        int cs[4], rs[4], wc0[4], wc1[4], wr0[4], wr1[4];
        int bx0 = 1 << 30, bx1 = -1, by0 = 1 << 30, by1 = -1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = x0 + lane + 32 * (k % SEGS), y = y0 + k / SEGS;
            const bool in_dst = (x < p.dst_w) && (y < p.dst_h);
            int X, Y;
            const int xc = min(x, p.dst_w - 1);
            const int xb = bw0_pow2 ? (xc & ~(p.bw0 - 1)) : (xc / p.bw0) * p.bw0;
            bevk_map_pixel_xb(s_M[gi], xb, xc - xb, min(y, p.dst_h - 1), LINEAR ? 32.0 : 1.0, X, Y);
            if (LINEAR) {
                const int sx = bevk_sat16(X >> 5), sy = bevk_sat16(Y >> 5);
                window(sx, X & 31, p.src_w, cs[k], wc0[k], wc1[k]);
                window(sy, Y & 31, p.src_h, rs[k], wr0[k], wr1[k]);
            } else {
                const int sx = bevk_sat16(X), sy = bevk_sat16(Y);
                const bool in = sx >= 0 && sx < p.src_w && sy >= 0 && sy < p.src_h;
                cs[k] = min(max(sx, 0), p.src_w - 1);
                rs[k] = min(max(sy, 0), p.src_h - 1);
                wc0[k] = in ? 1 : 0;
                wc1[k] = 0;
                wr0[k] = in ? 1 : 0;
                wr1[k] = 0;
            }
            const bool active = in_dst && (wc0[k] | wc1[k]) != 0 && (wr0[k] | wr1[k]) != 0;
            if (!active) {
                wc0[k] = wc1[k] = wr0[k] = wr1[k] = 0;
            } else {
                bx0 = min(bx0, cs[k]);
                bx1 = max(bx1, cs[k] + (LINEAR ? 1 : 0));
                by0 = min(by0, rs[k]);
                by1 = max(by1, rs[k] + (LINEAR ? 1 : 0));
            }
        }
        bx0 = __reduce_min_sync(0xffffffffu, bx0);
        bx1 = __reduce_max_sync(0xffffffffu, bx1);
        by0 = __reduce_min_sync(0xffffffffu, by0);
        by1 = __reduce_max_sync(0xffffffffu, by1);
        if (lane == 0 && bx1 >= 0) {
            atomicMin(&s_box[0], bx0);
            atomicMax(&s_box[1], bx1);
            atomicMin(&s_box[2], by0);
            atomicMax(&s_box[3], by1);
            s_any = 1;
        }
        __syncthreads();
        const bool any = s_any != 0;
        bx0 = s_box[0];
        bx1 = s_box[1];
        by0 = s_box[2];
        by1 = s_box[3];

        // staged geometry: 16-byte aligned first column, box shapes from the tensor-map menu
        const int a0 = (kBpp * bx0) & ~15;
        const int need_w = ((kBpp * (bx1 + 1) + 15) & ~15) - a0;
        const int wi = map_width_index(need_w);
        const int pitch = map_width(wi);
        const int nrows = by1 - by0 + 1;
        bool staged = any && need_w <= kMaxBoxWidth && nrows <= kMaxBoxes * kMaxBoxHeight;
        // rows the boxes cover: at most kMaxBoxes boxes, each the smallest menu height that fits
        int box_rows = 0;
        if (staged)
            for (int rem = nrows; rem > 0;) {
                const int h = box_height(wi, min(rem, kMaxBoxHeight));
                box_rows += h;
                rem -= h;
            }
        const uint32_t frame_bytes = ((uint32_t)(box_rows * pitch) + 127u) & ~127u;
        staged = staged && 2 * frame_bytes <= (uint32_t)ring_bytes;
        // nearest reads one pixel per dst pixel: where the map minifies, most of the box would be
        // staged for nothing (and fetched from HBM for nothing) -- gather those tiles directly
        if (!LINEAR && frame_bytes > 3u * 1024u * (uint32_t)kBpp) staged = false;
        // frames per stage: as many as still leave a ring of 4 stages
        const int fps = (4 * kMaxStageFrames * frame_bytes <= (uint32_t)ring_bytes)
                            ? kMaxStageFrames
                            : (8 * frame_bytes <= (uint32_t)ring_bytes ? 2 : 1);
        const int stage_stride = fps * (int)frame_bytes;
        const int slog = (8 * stage_stride <= ring_bytes) ? 3 : (4 * stage_stride <= ring_bytes ? 2 : 1);
        const uint32_t full0 = bar0 + 16 * (1 << slog) - 32;  // the set's empty[] follow its full[]
        if (tid == 0) {
            if (staged) {
                int rem = nrows, row = 0, nb = 0;
                while (rem > 0) {
                    const int r = min(rem, kMaxBoxHeight);
                    s_plan.map_idx[nb] = box_map_index(wi, r);
                    s_plan.row[nb] = row;
                    row += box_height(wi, r);
                    rem -= box_height(wi, r);
                    ++nb;
                }
                s_plan.n_boxes = nb;
                s_plan.x = map_is_wide(wi) ? a0 >> 3 : a0 >> 2;  // in elements of the map
                s_plan.y0 = by0;
                s_plan.bytes = (uint32_t)(box_rows * pitch);
                s_plan.frame0 = g_first + f0 * g_stride;
                s_plan.frame_step = g_stride;
                s_plan.src_h = p.src_h;
                s_plan.n_frames = n_frames;
                s_plan.fps = fps;
                s_plan.ring = ring;
                s_plan.full0 = full0;
                s_plan.stride = stage_stride;
                s_plan.frame_bytes = frame_bytes;
                s_plan.pitch = pitch;
            }
        }
        __syncthreads();  // s_plan visible to the producer lanes; every thread has read s_box
        if (tid == 0) {
            // re-arm the box for the next item (its first atomics come after this item's last
            // barrier), then fetch + decode the next item: nobody waits for thread 0 until then
            s_box[0] = s_box[2] = 1 << 30;
            s_box[1] = s_box[3] = -1;
            s_any = 0;
            decode(next_item ? (int)gridDim.x + atomicAdd(next_item, 1) : item + (int)gridDim.x,
                   s_item[par]);
        }

        // ---- store geometry ---------------------------------------------------------------------
        // which words of a 32-pixel row segment this lane writes is the pixel format's business
        const typename PX::Store st = PX::store_setup(lane);
        bool seg_ok[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // pixels left in this 32-pixel segment: a multiple of 4 (dst_w % 4 == 0)
            const int valid_px = min(32, p.dst_w - (x0 + 32 * (k % SEGS)));
            seg_ok[k] = (y0 + k / SEGS < p.dst_h) && PX::lane_stores(lane, valid_px);
        }
        uint32_t d_step = (uint32_t)g_stride * (uint32_t)dst_frame_bytes;  // < 2^32 (host check)
        keep(d_step);
        uint8_t *d = dst + (long long)(g_first + f0 * g_stride) * dst_frame_bytes +
                     ((long long)y0 * p.dst_w + x0) * kBpp + PX::lane_offset(lane);

        if (!any) {
            // whole tile maps outside the source: constant border (0) for every frame
#pragma unroll 1
            for (int i = 0; i < n_frames; ++i, d += d_step)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    PX::store_zero(d + PX::kSegBytes * (k % SEGS) + (k / SEGS) * row_bytes, seg_ok[k], lane);
        } else if (staged) {
            typename PX::Reg px[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool act = (wc0[k] | wc1[k]) != 0;
                const int A = act ? (rs[k] - by0) * pitch + kBpp * cs[k] - a0 : 0;
                px[k] = PX::template make<LINEAR>(act, (uint32_t)A, wc0[k], wc1[k], wr0[k], wr1[k]);
                keep(px[k].addr);
                keep(px[k].sh);
            }
            LoopCtx c;
            c.ring = ring;
            c.full0 = full0;
            c.stride = stage_stride;
            c.frame_bytes = frame_bytes;
            c.pitch = pitch;
            c.fps = fps;
            c.n_frames = n_frames;
            c.maps = &maps;
            c.plan = &s_plan;
            keep(c.ring);
            keep(c.full0);
            // every thread tracks the counter in a register; thread 0 publishes it for the next item
            uint32_t use = s_use[slog - 1];
            if (slog == 3)
                use = frame_loop<PX, LINEAR, 3, SEGS>(c, px, use, d, d_step, row_bytes, seg_ok, st, tid);
            else if (slog == 2)
                use = frame_loop<PX, LINEAR, 2, SEGS>(c, px, use, d, d_step, row_bytes, seg_ok, st, tid);
            else
                use = frame_loop<PX, LINEAR, 1, SEGS>(c, px, use, d, d_step, row_bytes, seg_ok, st, tid);
            __syncthreads();  // every warp has read s_use and left the ring
            if (tid == 0) s_use[slog - 1] = use;
            continue;
        } else if (p.hard) {
            // split launch: leave the tile to the direct-gather kernel that follows on the stream
            if (tid == 0) p.hard[gi * p.hard_tiles + tile_x * tiles_y + tile_y] = 1;
        } else {
            // bounding box too large for the ring (strong minification) or too wide / tall for the
            // tensor-map menu: the same arithmetic on aligned 32-bit loads straight from global
            // memory (L1-cached, read-only path).  Frames are 16-byte aligned (src is, and a frame
            // is a multiple of 16 bytes), so the window words are addressed like the staged ones.
            constexpr uint32_t kLast = PX::template last_word_offset<LINEAR>();
            typename PX::Reg px[4];
            uint32_t last[4];  // last window word; clamped into the frame where it is not needed
            const uint32_t last_word = (uint32_t)src_frame_bytes - 4u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool act = (wc0[k] | wc1[k]) != 0;
                const uint32_t A = act ? (uint32_t)rs[k] * (uint32_t)src_row_bytes + (uint32_t)kBpp * (uint32_t)cs[k] : 0u;
                px[k] = PX::template make<LINEAR>(act, A, wc0[k], wc1[k], wr0[k], wr1[k]);
                // bilinear reads row 1 of the window at +src_row_bytes: clamp so that also that
                // read stays inside the frame (the alignments that need the word never clamp)
                last[k] = min(px[k].addr + kLast, last_word - (LINEAR ? (uint32_t)src_row_bytes : 0u));
            }
            const uint8_t *s = src + (long long)(g_first + f0 * g_stride) * src_frame_bytes;
            const long long s_step = (long long)g_stride * src_frame_bytes;
#pragma unroll 1
            for (int i = 0; i < n_frames; ++i, d += d_step, s += s_step) {
                typename PX::Out P[4];
                // all window words of the frame are requested before the first is used: this loop
                // lives on loads in flight
                uint32_t w[4][8];
#pragma unroll
                for (int k = 0; k < 4; ++k)  // "addresses" are byte offsets inside the frame here
                    PX::template load<LINEAR>(px[k], px[k].addr, px[k].addr + src_row_bytes, last[k],
                                              last[k] + src_row_bytes, w[k],
                                              [s](uint32_t off) { return ldg_sparse((const uint32_t *)(s + off)); });
#pragma unroll
                for (int k = 0; k < 4; ++k) P[k] = PX::template math<LINEAR>(px[k], w[k]);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    PX::store(d + PX::kSegBytes * (k % SEGS) + (k / SEGS) * row_bytes, P[k], seg_ok[k], st, lane);
            }
        }
        __syncthreads();  // the next item's descriptor (s_item) is complete and visible
    }
}

// ---- host side: tensor-map menu ---------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct MapCacheEntry {
    const void *base = nullptr;
    int row_bytes = 0;
    long long rows = 0;
    unsigned long long stamp = 0;
    WarpFastMaps maps;
};
constexpr int kMapCacheSize = 8;
MapCacheEntry g_map_cache[kMapCacheSize];
unsigned long long g_map_stamp = 0;
std::mutex g_map_mutex;
EncodeTiledFn g_encode = nullptr;

constexpr int kCounterSlots = 1024;
unsigned char *g_flags = nullptr;  // split launches: ring of flag slices
int g_flags_dev = -1;
unsigned g_flag_next = 0;
constexpr int kFlagSlices = 16, kFlagSliceBytes = 256 * 1024;
int *g_counters = nullptr;
int g_counters_dev = -1;
unsigned g_counter_next = 0;

// The batch viewed as a [rows][row_bytes / 4] uint32 matrix: one tensor map per box shape.
int get_maps(const void *base, int row_bytes, long long rows, WarpFastMaps &out)
{
    std::lock_guard<std::mutex> lock(g_map_mutex);
    if (!g_encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        BEVK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn)
            BEVK_FAIL(BEVK_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        g_encode = (EncodeTiledFn)fn;
    }
    MapCacheEntry *victim = &g_map_cache[0];
    for (int i = 0; i < kMapCacheSize; ++i) {
        MapCacheEntry &e = g_map_cache[i];
        if (e.base == base && e.row_bytes == row_bytes && e.rows == rows) {
            e.stamp = ++g_map_stamp;
            out = e.maps;
            return BEVK_OK;
        }
        if (e.stamp < victim->stamp) victim = &e;
    }
    const cuuint64_t gstride[1] = {(cuuint64_t)row_bytes};
    const cuuint32_t estride[2] = {1, 1};
    // every box row starts 16-byte aligned in global memory, which the TMA unit requires (a
    // start on another byte raises an illegal-instruction fault -- tried for a 2-byte shifted
    // second copy); 4-byte elements give boxes up to 1024 B wide, 8-byte elements up to 2048 B
    for (int wi = 0; wi < kNarrowW + kWideW; ++wi) {
        const bool wide = map_is_wide(wi);
        const int es = wide ? 8 : 4;
        const cuuint64_t gdim[2] = {(cuuint64_t)(row_bytes / es), (cuuint64_t)rows};
        for (int hi = 0; hi < (wide ? kWideH : kNarrowH); ++hi) {
            const int h = wide ? 4 * (hi + 1) : map_height(hi);
            const cuuint32_t box[2] = {(cuuint32_t)(map_width(wi) / es), (cuuint32_t)h};
            CUresult r = g_encode(&victim->maps.m[box_map_index(wi, h)],
                                  wide ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32,
                                  2, const_cast<void *>(base), gdim, gstride, box, estride,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                victim->base = nullptr;
                BEVK_FAIL(BEVK_E_CUDA, "cuTensorMapEncodeTiled failed (code %d) for a %ux%u box",
                          (int)r, box[0] * es, box[1]);
            }
        }
    }
    victim->base = base;
    victim->row_bytes = row_bytes;
    victim->rows = rows;
    victim->stamp = ++g_map_stamp;
    out = victim->maps;
    return BEVK_OK;
}

struct KernelConfig {
    bool ready = false;
    int ctas_per_sm = 0;
    int ring_bytes = 0;
};
// CTAs per SM the register allocation is bounded for: bilinear needs 80 registers per thread (a
// 64-register build spills in the frame loop and is slower), nearest fits 64 without spilling and
// gains 5 % from the fourth CTA.
constexpr int min_ctas(bool linear) { return linear ? 3 : 4; }
KernelConfig g_cfg[2][2][3];  // [pixel format: u8x3, f16x3][linear][tile shape: SEGS 4, 2, 1]
inline int segs_index(int segs) { return segs == 4 ? 0 : (segs == 2 ? 1 : 2); }

template <typename PX, bool LINEAR, int SEGS> int configure(KernelConfig &cfg)
{
    constexpr int kMinCtas = min_ctas(LINEAR);
    auto kern = warp_fast_kernel<PX, LINEAR, kMinCtas, SEGS>;
    // how many CTAs the register file allows, then split the shared memory evenly between them
    BEVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024));
    int by_regs = 0;
    BEVK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&by_regs, kern, kThreads, 16 * 1024));
    if (by_regs < 1) BEVK_FAIL(BEVK_E_CUDA, "staged warp kernel does not fit an SM");
    by_regs = by_regs > kMinCtas ? kMinCtas : by_regs;
    int dev = 0, smem_sm = 0;
    BEVK_CUDA(cudaGetDevice(&dev));
    BEVK_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    // per CTA: 1 KB reserved by the driver + static shared memory
    cudaFuncAttributes attr;
    BEVK_CUDA(cudaFuncGetAttributes(&attr, kern));
    int ring = smem_sm / by_regs - 1024 - (int)attr.sharedSizeBytes - kBarBytes - kTailSlack;
    ring &= ~127;
    if (ring > 200 * 1024) ring = 200 * 1024;
    BEVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   ring + kBarBytes + kTailSlack));
    cfg.ctas_per_sm = by_regs;
    cfg.ring_bytes = ring;
    cfg.ready = true;
    return BEVK_OK;
}

template <typename PX, bool LINEAR> int configure_segs(int segs, KernelConfig &cfg)
{
    return segs == 4 ? configure<PX, LINEAR, 4>(cfg)
                     : (segs == 2 ? configure<PX, LINEAR, 2>(cfg) : configure<PX, LINEAR, 1>(cfg));
}
template <typename PX> int configure_any(int linear, int segs, KernelConfig &cfg)
{
    return linear ? configure_segs<PX, true>(segs, cfg) : configure_segs<PX, false>(segs, cfg);
}

template <typename PX, bool LINEAR, int SEGS>
void launch(int grid, int smem, cudaStream_t stream, const BevkWarpParams &p, const WarpFastMaps &maps,
            const ChunkPlan &plan, int tiles_x, int tiles_y, int items, int ring_bytes, int *counter)
{
    warp_fast_kernel<PX, LINEAR, min_ctas(LINEAR), SEGS><<<grid, kThreads, smem, stream>>>(
        p, maps, plan, tiles_x, tiles_y, items, ring_bytes, counter);
}
template <typename PX>
void launch_any(int linear, int segs, int grid, int smem, cudaStream_t stream, const BevkWarpParams &p,
                const WarpFastMaps &maps, const ChunkPlan &plan, int tiles_x, int tiles_y, int items,
                int ring_bytes, int *counter)
{
#define BEVK_LAUNCH(LIN, SEGS) \
    launch<PX, LIN, SEGS>(grid, smem, stream, p, maps, plan, tiles_x, tiles_y, items, ring_bytes, counter)
    if (linear) {
        if (segs == 4) BEVK_LAUNCH(true, 4);
        else if (segs == 2) BEVK_LAUNCH(true, 2);
        else BEVK_LAUNCH(true, 1);
    } else {
        if (segs == 4) BEVK_LAUNCH(false, 4);
        else if (segs == 2) BEVK_LAUNCH(false, 2);
        else BEVK_LAUNCH(false, 1);
    }
#undef BEVK_LAUNCH
}

// ---- host side: which tile shape stages --------------------------------------------------------
// Fraction of the (sampled) tiles of one shape whose source box cannot be staged: wider than the
// widest tensor box, taller than kMaxBoxes boxes, or larger than half the ring.  The box of a tile
// is spanned by its four corner pixels (a projective map is monotone along lines as long as w
// keeps its sign); tiles where w changes sign are left out of the estimate (the kernel copes).
double unstaged_fraction(const BevkWarpParams &p, int linear, int bpp, int segs, int ring_bytes)
{
    const int tw = tile_w(segs), th = tile_h(segs);
    const int tiles_x = (p.dst_w + tw - 1) / tw, tiles_y = (p.dst_h + th - 1) / th;
    // up to 16 x 32 evenly spaced tiles, first and last row / column included
    const int nsx = tiles_x < 16 ? tiles_x : 16, nsy = tiles_y < 32 ? tiles_y : 32;
    const double scale = linear ? 32.0 : 1.0;
    long long active = 0, bad = 0;
    for (int gi = 0; gi < p.n_groups; ++gi) {
        const double *M = p.g[gi].M;
        for (int iy = 0; iy < nsy; ++iy)
            for (int ix = 0; ix < nsx; ++ix) {
                const int ty = nsy > 1 ? (int)((long long)iy * (tiles_y - 1) / (nsy - 1)) : 0;
                const int tx = nsx > 1 ? (int)((long long)ix * (tiles_x - 1) / (nsx - 1)) : 0;
                const int x0 = tx * tw, y0 = ty * th;
                const int x1 = (x0 + tw < p.dst_w ? x0 + tw : p.dst_w) - 1;
                const int y1 = (y0 + th < p.dst_h ? y0 + th : p.dst_h) - 1;
                const int cx[4] = {x0, x1, x0, x1}, cy[4] = {y0, y0, y1, y1};
                int lo_x = 1 << 30, hi_x = -(1 << 30), lo_y = 1 << 30, hi_y = -(1 << 30), pos = 0, neg = 0;
                for (int c = 0; c < 4; ++c) {
                    const double w = M[6] * cx[c] + M[7] * cy[c] + M[8];
                    pos += w > 0;
                    neg += w < 0;
                    int X, Y;
                    bevk_map_pixel(M, cx[c], cy[c], p.bw0, scale, X, Y);
                    const int sx = bevk_sat16(linear ? (X >> 5) : X), sy = bevk_sat16(linear ? (Y >> 5) : Y);
                    lo_x = sx < lo_x ? sx : lo_x;
                    hi_x = sx > hi_x ? sx : hi_x;
                    lo_y = sy < lo_y ? sy : lo_y;
                    hi_y = sy > hi_y ? sy : hi_y;
                }
                hi_x += linear;
                hi_y += linear;
                if (hi_x < 0 || hi_y < 0 || lo_x >= p.src_w || lo_y >= p.src_h) continue;  // border only
                if (pos != 4 && neg != 4) continue;  // horizon inside the tile: corners say nothing
                ++active;
                lo_x = lo_x < 0 ? 0 : lo_x;
                lo_y = lo_y < 0 ? 0 : lo_y;
                hi_x = hi_x > p.src_w - 1 ? p.src_w - 1 : hi_x;
                hi_y = hi_y > p.src_h - 1 ? p.src_h - 1 : hi_y;
                const int a0 = (bpp * lo_x) & ~15;
                const int need_w = ((bpp * (hi_x + 1) + 15) & ~15) - a0 + 16;  // one chunk of margin
                const int rows = hi_y - lo_y + 1 + 1;
                const int pitch = need_w <= kMaxBoxWidth ? map_width(map_width_index(need_w)) : need_w;
                if (need_w > kMaxBoxWidth || rows > kMaxBoxes * kMaxBoxHeight ||
                    2LL * pitch * (rows + 3) > ring_bytes)
                    ++bad;
            }
    }
    return active ? (double)bad / (double)active : 0.0;
}

struct ModeKey {
    double M[BEVK_MAX_GROUPS][9];
    int n_groups, src_h, src_w, dst_h, dst_w, linear, bpp;
};
struct ModeEntry {
    ModeKey key;
    int segs;  // 4, 2, 1, or 0: leave it to the direct-gather kernel
    int best;  // the shape with the fewest unstaged tiles (what a forced launch uses)
    bool split;
    unsigned long long stamp = 0;
    bool valid = false;
};
// Above this fraction of unstaged tiles the whole launch goes to the direct-gather kernel; between
// kNegligible and this the launch is split (staged kernel + direct kernel over the marked tiles).
constexpr double kMaxSplit = 0.35;
constexpr int kModeCacheSize = 8;
ModeEntry g_mode_cache[kModeCacheSize];
unsigned long long g_mode_stamp = 0;

// Largest tile shape whose boxes all stage; 0 if even the best shape leaves more than
// kMaxUnstaged of the tiles to the in-kernel fallback (strong minification: the boxes are mostly
// untouched pixels, the direct-gather kernel moves less data).
// *split is set when the chosen shape still leaves a noticeable fraction of the tiles unstaged:
// those tiles are cheaper in a second, direct-gather launch than in the staged kernel's fallback.
int pick_tile_shape(const BevkWarpParams &p, int linear, int bpp, int ring_bytes, bool force, bool *split)
{
    // nearest reads one tap per pixel, so its in-kernel fallback costs little: keep wide tiles
    const double kNegligible = linear ? 0.002 : 0.05;
    ModeKey key;
    memset(&key, 0, sizeof(key));
    for (int i = 0; i < p.n_groups; ++i) memcpy(key.M[i], p.g[i].M, sizeof(key.M[i]));
    key.n_groups = p.n_groups;
    key.src_h = p.src_h;
    key.src_w = p.src_w;
    key.dst_h = p.dst_h;
    key.dst_w = p.dst_w;
    key.linear = linear;
    key.bpp = bpp;
    std::lock_guard<std::mutex> lock(g_map_mutex);
    ModeEntry *victim = &g_mode_cache[0];  // an empty entry (stamp 0), else the least recently used
    for (int i = 0; i < kModeCacheSize; ++i) {
        ModeEntry &e = g_mode_cache[i];
        if (e.valid && memcmp(&e.key, &key, sizeof(key)) == 0) {
            e.stamp = ++g_mode_stamp;
            *split = e.split && e.segs != 0;
            return (e.segs == 0 && force) ? e.best : e.segs;
        }
        if (e.stamp < victim->stamp) victim = &e;
    }
    // the widest shape that stages (nearly) everything; else the shape that stages most
    int best = 0;
    double best_frac = 2.0;
    static const int shapes[3] = {4, 2, 1};
    const char *env = getenv("BEVK_FAST_SEGS");  // tuning aid: force a tile shape
    for (int si = 0; si < 3; ++si) {
        if (env && atoi(env) != shapes[si]) continue;
        const double f = unstaged_fraction(p, linear, bpp, shapes[si], ring_bytes);
        if (f < best_frac - 1e-9) {
            best_frac = f;
            best = shapes[si];
        }
        if (f <= kNegligible) break;
    }
    const int segs = best_frac <= kMaxSplit ? best : 0;
    victim->key = key;
    victim->segs = segs;
    victim->best = best ? best : 1;
    // up to ~2 % the in-kernel fallback is as fast as a second launch (measured on the cfg-4 cameras)
    victim->split = best_frac > (kNegligible > 0.02 ? kNegligible : 0.02);
    victim->stamp = ++g_mode_stamp;
    victim->valid = true;
    *split = victim->split && segs != 0;
    return (segs == 0 && force) ? victim->best : segs;
}

}  // namespace

int bevk_launch_warp_fast(const BevkWarpParams &p_in, int channels, int dtype, int linear, int force,
                          cudaStream_t stream)
{
    // qualification: uint8 x 3 or float16 x 3, zero border, rows the tensor maps / word stores
    // can address, per-frame dst pointer steps that fit 32 bits
    if (channels != 3 || (dtype != BEVK_U8 && dtype != BEVK_F16)) return 0;
    const int fmt = dtype == BEVK_F16 ? 1 : 0;
    const int bpp = fmt ? PxF16C3::kBpp : PxU8C3::kBpp;
    if (p_in.border[0] != 0.f || p_in.border[1] != 0.f || p_in.border[2] != 0.f) return 0;
    if (p_in.src_w < 2 || p_in.src_h < 2) return 0;
    if ((p_in.src_w * bpp) % 16 != 0 || (p_in.dst_w % 4) != 0) return 0;
    if (((uintptr_t)p_in.src % 16) != 0 || ((uintptr_t)p_in.dst % 4) != 0) return 0;

    BevkWarpParams p = p_in;
    int n_src_frames = 0, max_count = 0;
    for (int i = 0; i < p.n_groups; ++i)
        if ((long long)p.g[i].stride * p.dst_frame_elems * (fmt ? 2 : 1) > 0xffffffffLL || p.g[i].stride < 1)
            return 0;
    for (int i = 0; i < p.n_groups; ++i) {
        const int last = p.g[i].first + (p.g[i].count - 1) * p.g[i].stride;
        n_src_frames = n_src_frames > last + 1 ? n_src_frames : last + 1;
        max_count = max_count > p.g[i].count ? max_count : p.g[i].count;
    }
    // The staged kernel pays an FP64 set-up per (tile, chunk); with fewer than kMinFrames frames
    // per homography the direct-gather kernel is faster (measured: 27 vs 39 us for one 1080p frame)
    constexpr int kMinFrames = 4;
    if (!force && max_count < kMinFrames) return 0;
    const long long rows = (long long)n_src_frames * p.src_h;
    if (rows > 0x7fffffffLL) return 0;  // TMA coordinates are int32

    // tile shape: every shape's kernel has the same ring size, so configure the widest first
    KernelConfig &cfg0 = g_cfg[fmt][linear ? 1 : 0][0];
    if (!cfg0.ready) {
        int rc = fmt ? configure_any<PxF16C3>(linear, 4, cfg0) : configure_any<PxU8C3>(linear, 4, cfg0);
        if (rc) return rc;
    }
    bool split = false;
    const int segs = pick_tile_shape(p, linear, bpp, cfg0.ring_bytes, force != 0, &split);
    if (segs == 0) return 0;
    KernelConfig &cfg = g_cfg[fmt][linear ? 1 : 0][segs_index(segs)];
    if (!cfg.ready) {
        int rc = fmt ? configure_any<PxF16C3>(linear, segs, cfg) : configure_any<PxU8C3>(linear, segs, cfg);
        if (rc) return rc;
    }

    WarpFastMaps maps;
    int rc = get_maps(p.src, p.src_w * bpp, rows, maps);
    if (rc) return rc;

    const int tiles_x = (p.dst_w + tile_w(segs) - 1) / tile_w(segs);
    const int tiles_y = (p.dst_h + tile_h(segs) - 1) / tile_h(segs);
    const long long n_tiles = (long long)tiles_x * tiles_y;
    const int ctas = bevk_sm_count() * cfg.ctas_per_sm;

    // Frame chunks.  Every (tile, chunk) item pays one FP64 set-up, so chunks should be long; the
    // CTAs pull items from a shared counter, so the LAST items should be short.  With enough
    // frames the chunk lengths therefore decay (5/16, 4/16, 3/16, 1/8, then ever smaller);
    // short batches get fewer, equal chunks, just enough for ~3 items per CTA.
    ChunkPlan plan;
    memset(&plan, 0, sizeof(plan));
    const long long tile_groups = n_tiles * p.n_groups;
    if (max_count >= 128 && tile_groups * 5 >= 2LL * ctas) {
        static const uint32_t cum[8] = {0, 20480, 36864, 49152, 57344, 61952, 64512, 65536};
        plan.n_chunks = 7;
        memcpy(plan.cum, cum, sizeof(cum));
    } else {
        int k = (int)((3LL * ctas + tile_groups - 1) / tile_groups);
        k = k < 1 ? 1 : k;
        k = k > kMaxChunks ? kMaxChunks : k;
        k = k > (max_count + 7) / 8 ? (max_count + 7) / 8 : k;  // at least ~8 frames per chunk
        k = k < 1 ? 1 : k;
        plan.n_chunks = k;
        for (int i = 0; i <= k; ++i) plan.cum[i] = (uint32_t)(65536LL * i / k);
    }
    const long long items = tile_groups * plan.n_chunks;
    if (items > 0x7fffffffLL) return 0;
    const int grid = (int)(items < ctas ? items : ctas);

    // shared item counter: one zeroed slot per launch out of a ring (launches on different
    // streams may overlap).  Not needed when every CTA has exactly one item.
    int *counter = nullptr;
    if (items > grid) {
        std::lock_guard<std::mutex> lock(g_map_mutex);
        int dev = 0;
        BEVK_CUDA(cudaGetDevice(&dev));
        if (!g_counters || g_counters_dev != dev) {
            if (g_counters) cudaFree(g_counters);
            g_counters = nullptr;
            BEVK_CUDA(cudaMalloc(&g_counters, kCounterSlots * sizeof(int)));
            g_counters_dev = dev;
        }
        counter = g_counters + (g_counter_next++ % kCounterSlots);
    }
    if (counter) BEVK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), stream));

    // split launch: one byte per (group, tile), a zeroed slice of a ring of flag buffers
    const long long n_flags = n_tiles * p.n_groups;
    split = split && n_flags <= kFlagSliceBytes && !getenv("BEVK_NO_SPLIT");  // env: tuning aid
    if (split) {
        std::lock_guard<std::mutex> lock(g_map_mutex);
        int dev = 0;
        BEVK_CUDA(cudaGetDevice(&dev));
        if (!g_flags || g_flags_dev != dev) {
            if (g_flags) cudaFree(g_flags);
            g_flags = nullptr;
            BEVK_CUDA(cudaMalloc(&g_flags, (size_t)kFlagSlices * kFlagSliceBytes));
            g_flags_dev = dev;
        }
        p.hard = g_flags + (size_t)(g_flag_next++ % kFlagSlices) * kFlagSliceBytes;
        p.hard_tw = tile_w(segs);
        p.hard_th = tile_h(segs);
        p.hard_ty = tiles_y;
        p.hard_tiles = (int)n_tiles;
        BEVK_CUDA(cudaMemsetAsync(p.hard, 0, (size_t)n_flags, stream));
    }

    const int smem = cfg.ring_bytes + kBarBytes + kTailSlack;
    if (fmt)
        launch_any<PxF16C3>(linear, segs, grid, smem, stream, p, maps, plan, tiles_x, tiles_y, (int)items,
                            cfg.ring_bytes, counter);
    else
        launch_any<PxU8C3>(linear, segs, grid, smem, stream, p, maps, plan, tiles_x, tiles_y, (int)items,
                           cfg.ring_bytes, counter);
    BEVK_CUDA(cudaGetLastError());
    if (split) {
        // the tiles the staged kernel marked, through the direct-gather kernel (same stream)
        int rc2 = bevk_plan_generic_chunks(p, channels);
        if (rc2) return rc2;
        rc2 = bevk_launch_warp_generic(p, channels, dtype, linear, stream);
        if (rc2) return rc2;
    }
    return 1;
}
