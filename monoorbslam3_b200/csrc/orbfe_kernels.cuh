// orbfe_kernels.cuh — device code of the ORB extractor path, hand-written for sm_100a.
//
//   K1 k_resize    cv::resize INTER_LINEAR 8U, level l-1 -> l        (ORBExtractor.cpp:559-570)
//   K2 k_fast_planes  per-cell FAST-9/16 + NMS + threshold fallback  (ORBExtractor.cpp:592-617); k_fast = round 1's formulation, ORBFE_FAST_V1=1
//   K4 k_octree    DistributeOctree, deterministic, pointer-free     (ORBExtractor.cpp:367-413, 640-830)
//   K6 k_blur      GaussianBlur 7x7 sigma 2, 8.8 fixed point         (ORBExtractor.cpp:527-528)
//   K5/K7 k_describe  IC_Angle + fastAtan2 + rotated BRIEF-256        (ORBExtractor.cpp:18-97)
//
// Tiles are staged into shared memory with TMA (cp.async.bulk.tensor.3d over an (x, y, frame) tensor map,
// out-of-range elements zero-filled) or, with ORBFE_FLAG_NO_TMA, with 16-byte vector loads.
#pragma once
#include "orbfe_internal.cuh"

namespace orbfe {

// ------------------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + TMA tile load
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

// Stage a (kBoxW x box_h) byte box of frame `frame` into `tile` (pitch kBoxW).  TMA requires the box to start on a 16-byte
// boundary in global memory, so both paths load from ax = ox rounded down to 16 and return xo = ox - ax: the pixel (ox + c, oy + r)
// sits at tile[r * kBoxW + xo + c], and only kBoxW - xo columns are usable.
//   TMA:      out-of-range bytes are zero-filled by the hardware.
//   fallback: rows outside the image are zero; bytes right of the image width are whatever the (padded) row holds; neither is
//             ever consumed un-fixed.
// Must be called by all threads of the block; on return the tile is visible to all of them.
template <bool kTMA, int NT>
__device__ __forceinline__ int stage_box(uint8_t *tile, uint64_t *bar, const CUtensorMap *map, const uint8_t *frame_base,
                                         int pitch, int h, int ox, int oy, int frame, int box_h) {
    const int ax = ox & ~15;                           // floor to 16 (also for negative ox)
    if constexpr (kTMA) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            fence_barrier_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, (uint32_t) (kBoxW * box_h));
            tma_load_3d(tile, map, bar, ax, oy, frame);
        }
        mbar_wait(bar, 0);
    } else {
        constexpr int kVec = kBoxW / 16;
        for (int idx = threadIdx.x; idx < box_h * kVec; idx += NT) {
            const int r = idx / kVec, v = idx - r * kVec;
            const int gy = oy + r, gx = ax + 16 * v;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (gy >= 0 && gy < h && gx >= 0 && gx < pitch)
                val = __ldg(reinterpret_cast<const uint4 *>(frame_base + (size_t) gy * pitch + gx));
            *reinterpret_cast<uint4 *>(tile + r * kBoxW + 16 * v) = val;
        }
        __syncthreads();
    }
    return ox - ax;
}

template <bool kTMA> struct TilePitch { static constexpr int value = kBoxW; };

// ------------------------------------------------------------------------------------------------
// block-wide exclusive scan (NT threads, one value per thread)
// ------------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ int block_scan_excl(int v, int &total, int *s_warp) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    // every warp scans the NT / 32 warp totals itself (lane = warp): a handful of shuffles instead of a loop over all totals
    constexpr int NW = NT / 32;
    static_assert(NW <= 32, "one lane per warp total");
    const int x = lane < NW ? s_warp[lane] : 0;
    int xi = x;
#pragma unroll
    for (int o = 1; o < NW; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, xi, o);
        if (lane >= o) xi += t;
    }
    const int before = __shfl_sync(0xffffffffu, xi - x, wid);
    total = __shfl_sync(0xffffffffu, xi, NW - 1);
    __syncthreads();
    return before + inc - v;
}

#ifndef ORBFE_HELPERS_ONLY   // orbfe_match.cu only needs the helpers above
// ------------------------------------------------------------------------------------------------
// K1  resize (one launch per destination level; grid = (tiles, frames))
// ------------------------------------------------------------------------------------------------
struct ResizeArgs {
    const uint8_t *src; int sw, sh, spitch; unsigned long long sframe;
    uint8_t *dst; int dw, dh, dpitch; unsigned long long dframe;
    int tw, th, tiles_x;
    const int2 *xtab, *ytab;
};

// A thread owns 8 destination columns (two quads; their coefficient-table entries stay in registers) and walks down a run of
// destination rows, re-using the horizontal interpolation of a source row when consecutive destination rows share it (5 times out
// of 6 at scale 1.2).  96 threads = 24 column octets x 4 row runs; the per-row bookkeeping (row table entry, loop, reuse test) is
// paid once per 8 pixels.
// kPacked (every quad's source span fits 8 bytes, true for scale factors up to 2): a source row costs per quad three aligned word
// loads, two funnel shifts, and per column one PRMT (byte pair s0, s1) + one DP2A (s0 * a0 + s1 * a1); otherwise bytes are gathered.
// The horizontal results are kept pre-shifted (H >> 4), the vertical blend is two IMAD.HI with the row weights pre-shifted by 16.
constexpr int kRsQuads = 1;
constexpr int kRsThreads = (kRsMaxTW / (4 * kRsQuads)) * 4;

template <bool kTMA, bool kPacked>
__global__ void __launch_bounds__(kRsThreads) k_resize(const __grid_constant__ CUtensorMap tm_src, const ResizeArgs a) {
    constexpr int SP = TilePitch<kTMA>::value;
    constexpr int NC = 4 * kRsQuads;
    __shared__ __align__(128) uint8_t tile[kRsBoxH * SP + 16];                // + 16: the packed path may read one word past a row
    __shared__ __align__(8) uint64_t bar;
    const int frame = blockIdx.y;
    const int ty = blockIdx.x / a.tiles_x, tx = blockIdx.x - ty * a.tiles_x;
    const int dx0 = tx * a.tw, dy0 = ty * a.th;
    const int ox = __ldg(&a.xtab[dx0]).x & 0xffff, oy = __ldg(&a.ytab[dy0]).x & 0xffff;
    const int xo = stage_box<kTMA, kRsThreads>(tile, &bar, &tm_src, a.src + (size_t) frame * a.sframe, a.spitch, a.sh, ox, oy, frame, kRsBoxH);
    const int qx = threadIdx.x % (kRsMaxTW / NC), grp = threadIdx.x / (kRsMaxTW / NC);
    const int dxb = dx0 + NC * qx;
    if (NC * qx >= a.tw || dxb >= a.dw) return;
    int c0[NC], c1[NC];
    uint32_t wq[NC];                                      // a0 | a1 << 16
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int2 xe = __ldg(&a.xtab[min(dxb + i, a.dw - 1)]);
        c0[i] = (xe.x & 0xffff) - ox + xo; c1[i] = (xe.x >> 16) - ox + xo;
        wq[i] = (uint32_t) xe.y;
    }
    int cw[kRsQuads], sh[kRsQuads];
    uint32_t sel[NC];
#pragma unroll
    for (int q = 0; q < kRsQuads; ++q) {
        cw[q] = c0[4 * q] >> 2; sh[q] = (c0[4 * q] & 3) * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) sel[4 * q + i] = (uint32_t) (c0[4 * q + i] - c0[4 * q]) | ((uint32_t) (c1[4 * q + i] - c0[4 * q]) << 4);
    }
    // horizontal interpolation of one staged source row, >> 4
    auto hrow = [&](int r, uint32_t (&hs)[NC]) {
        if constexpr (kPacked) {
#pragma unroll
            for (int q = 0; q < kRsQuads; ++q) {
                const uint32_t *wp = reinterpret_cast<const uint32_t *>(tile + r * SP) + cw[q];
                const uint32_t wa = wp[0], wb = wp[1], wc = wp[2];
                const uint32_t lo = __funnelshift_r(wa, wb, sh[q]), hi = __funnelshift_r(wb, wc, sh[q]);
#pragma unroll
                for (int i = 0; i < 4; ++i) hs[4 * q + i] = __dp2a_lo(wq[4 * q + i], __byte_perm(lo, hi, sel[4 * q + i]), 0u) >> 4;
            }
        } else {
            const uint8_t *t = tile + r * SP;
#pragma unroll
            for (int i = 0; i < NC; ++i) hs[i] = (t[c0[i]] * (wq[i] & 0xffffu) + t[c1[i]] * (wq[i] >> 16)) >> 4;
        }
    };
    const int rows_per = (a.th + 3) >> 2;
    const int ry_end = min(min((grp + 1) * rows_per, a.th), a.dh - dy0);
    const int ry0 = grp * rows_per;
    if (ry0 >= ry_end) return;
    uint8_t *dst = a.dst + (size_t) frame * a.dframe + (size_t) (dy0 + ry0) * a.dpitch + dxb;
    const int2 *yt = a.ytab + dy0;
    int prev_r = -1;
    uint32_t hp[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) hp[i] = 0;
    for (int ry = ry0; ry < ry_end; ++ry) {
        const int2 ye = __ldg(&yt[ry]);
        const int r0 = (ye.x & 0xffff) - oy, r1 = (ye.x >> 16) - oy;
        const uint32_t b0 = (uint32_t) ye.y << 16, b1 = (uint32_t) ye.y & 0xffff0000u;        // row weights << 16
        uint32_t h0[NC], h1[NC];
        if (r0 == prev_r) {
#pragma unroll
            for (int i = 0; i < NC; ++i) h0[i] = hp[i];
        } else hrow(r0, h0);
        if (r1 == r0) {
#pragma unroll
            for (int i = 0; i < NC; ++i) h1[i] = h0[i];
        } else hrow(r1, h1);
        uint32_t v[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            v[i] = (__umulhi(b0, h0[i]) + __umulhi(b1, h1[i]) + 2u) >> 2;           // ((b0 * (H0 >> 4)) >> 16) + ...
            hp[i] = h1[i];
        }
        prev_r = r1;
        uint32_t out[kRsQuads];
#pragma unroll
        for (int q = 0; q < kRsQuads; ++q)          // v <= 255: byte 0 of each
            out[q] = __byte_perm(__byte_perm(v[4 * q], v[4 * q + 1], 0x0040), __byte_perm(v[4 * q + 2], v[4 * q + 3], 0x0040), 0x5410);
        if (kRsQuads == 2) *reinterpret_cast<uint2 *>(dst) = make_uint2(out[0], out[kRsQuads - 1]);
        else *reinterpret_cast<uint32_t *>(dst) = out[0];
        dst += a.dpitch;
    }
}

// ------------------------------------------------------------------------------------------------
// shared parameter block of the all-level kernels
// ------------------------------------------------------------------------------------------------
struct LevelSet {
    LevelGeom lv[ORBFE_MAX_LEVELS];
    const uint8_t *img[ORBFE_MAX_LEVELS];     // frame 0 of each level (level 0 may be the caller's buffer)
    int n_levels;
};
struct TmapSet { CUtensorMap m[ORBFE_MAX_LEVELS]; };

// ------------------------------------------------------------------------------------------------
// K2  FAST-9/16 per 30-px cell (cv::FAST on every cell, ORBExtractor.cpp:592-617).
// One CTA per strip of 8 cells (240 x 30 px, a 256 x 36 byte TMA box).  The corner measure m(x, y) does not depend on the cell
// grid, only NMS and the threshold fallback do, so the strip is processed in four block-wide stages:
//   A  rejection test on packed bytes, 8 pixels per thread: all eight opposite ring pairs (k, k+8) must hold a pixel with
//      |p - v| > t (VABSDIFF4 + SWAR compare; necessary for any 9-arc of either polarity); survivors -> shared queue
//   B  exact m for the queued pixels, one per thread, both polarities in one u16x2 register (IMAD packing, VIMNMX3.U16x2):
//      m = max over the 16 arcs of max(min_arc(p - v), min_arc(v - p)); corner iff m > t -> strip-wide score map
//   C  3x3 non-max suppression of the corners inside their own cell (outside = 0) -> per-cell row bit masks
//   D  one warp per cell: ordered (y, x) emission into the cell's slot, cell count
// Cells without a survivor at iniThFAST repeat A-D with minThFAST (the reference's second cv::FAST call).
// ------------------------------------------------------------------------------------------------
constexpr int kStripW = kCell * kCellsPerBlk;      // 240

__device__ __forceinline__ uint32_t bytes_gt(uint32_t x, uint32_t k7, bool t_ge_128) {
    // per byte: x > t ? 0x80 : 0, with k7 = (127 - (t & 127)) replicated
    const uint32_t low = (x & 0x7f7f7f7fu) + k7;
    return (t_ge_128 ? (low & x) : (low | x)) & 0x80808080u;
}

// exact corner measure of one pixel (centre c in the staged tile, pitch kBoxW).  Both polarities ride in one register:
// e_k = (256 + d_k) | (256 - d_k) << 16 with d_k = p_k - v, built by one IMAD per ring pixel (p * 0xFFFF0001 + bias; both halves
// stay in [1, 511], so nothing borrows across the halves).  m = max over the 16 arcs of min over the arc, per half (VIMNMX3.U16x2).
__device__ __forceinline__ int fast_measure1(const uint8_t *c) {
    constexpr int SP = kBoxW;
    constexpr int ofs[16] = {3 * SP, 3 * SP + 1, 2 * SP + 2, SP + 3, 3, -SP + 3, -2 * SP + 2, -3 * SP + 1,
                             -3 * SP, -3 * SP - 1, -2 * SP - 2, -SP - 3, -3, SP - 3, 2 * SP - 2, 3 * SP - 1};
    const uint32_t v = c[0];
    const uint32_t bias = (256u - v) | ((256u + v) << 16);
    uint32_t e[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) e[k] = (uint32_t) c[ofs[k]] * 0xFFFF0001u + bias;
    uint32_t mn[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) mn[k] = __vimin3_u16x2(e[k], e[(k + 1) & 15], e[(k + 2) & 15]);
    uint32_t best = 0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const uint32_t a = __vimin3_u16x2(mn[k], mn[(k + 3) & 15], mn[(k + 6) & 15]);
        const uint32_t b = __vimin3_u16x2(mn[k + 1], mn[(k + 4) & 15], mn[(k + 7) & 15]);
        best = __vimax3_u16x2(best, a, b);
    }
    return (int) max(best & 0xffffu, best >> 16) - 256;
}

// |a - b| > t per byte -> bit 7 of each byte of (x, y) OR-ed together (the low 7 bits are garbage).  k7 = (127 - (t & 127)) replicated.
//   exact form (any t):   ((x & 0x7f..) + k7)  |/&  x
//   fast form (t < 128):  (x + k7) | x  with the add issued as an IMAD (x * one + k7, `one` is a kernel argument equal to 1) so it
//   runs on the FMA pipe.  Without the 7-bit mask a byte >= 129 + t carries into its left neighbour, which can only turn a
//   neighbour byte equal to t into a pass: the test stays a superset of the exact one (stage B is exact), never a subset.
template <int kMode>      // 0: fast t < 128, 1: exact t < 128, 2: exact t >= 128
__device__ __forceinline__ uint32_t pair_gt(uint32_t v, uint32_t pa, uint32_t pb, uint32_t k7, uint32_t one, uint32_t acc) {
    const uint32_t xa = __vabsdiffu4(v, pa), xb = __vabsdiffu4(v, pb);
    if constexpr (kMode == 0) {
        const uint32_t ya = xa * one + k7, yb = xb * one + k7;
        return ((ya | xa | yb) | xb) & acc;
    } else {
        const uint32_t ya = (xa & 0x7f7f7f7fu) + k7, yb = (xb & 0x7f7f7f7fu) + k7;
        return kMode == 2 ? ((ya & xa) | (yb & xb)) & acc : ((ya | xa | yb) | xb) & acc;
    }
}

// Funnel shift right by 8 / 16 / 24 bits, (lo >> s) | (hi << (32 - s)).  kFma: on the FMA pipe as hi * 2^(32-s) + umulhi(lo, 2^(32-s))
// (IMAD + IMAD.HI; the two terms share no bit, so the add is an or; the multipliers are kernel arguments — as immediates the compiler
// turns the multiplies back into shifts).  An experiment kept behind ORBFE_FAST_FMA_SHIFT=1: k_fast executes 58 % of its instructions
// on the ALU pipe against 17 % on the FMA pipe, but it is as close to the issue limit as to the ALU limit, and two instructions for
// one cost more than the pipe change wins (measured: stage 1.456 -> 1.541 ms per 512 C1 frames).
template <bool kFma>
__device__ __forceinline__ uint32_t fsr(uint32_t lo, uint32_t hi, int s, uint32_t m) {
    if constexpr (kFma) return hi * m + __umulhi(lo, m);
    else return __funnelshift_r(lo, hi, s);
}

// Stage A of k_fast for one strip row y: a lane owns the aligned words 2*lane and 2*lane+1 (8 pixels).  A 9-arc of the 16-ring
// holds ring k or ring k+8 for every k, so a corner needs |p_k - v| > t or |p_(k+8) - v| > t for all eight opposite pairs.
// Returns the pass flags of the two words in bit 7 of each byte.
template <int kMode, bool kFma>
__device__ __forceinline__ void fast_pairs_row(const uint8_t *tile, int y, int lane, uint32_t k7, uint32_t one, uint3 sm, uint32_t &acc0, uint32_t &acc1) {
#define __funnelshift_r(a, b, s) fsr<kFma>((a), (b), (s), (s) == 8 ? sm.x : (s) == 16 ? sm.y : sm.z)
    constexpr int SP = kBoxW;
    const uint8_t *base = tile + y * SP + 8 * lane;
    const int lo = lane == 0 ? 0 : -4, hi = lane == 31 ? 4 : 8;          // clamp the neighbour words at the box edge (never candidates)
#define ORBFE_ROW(r, Wm, W0, W1, W2)                                                     \
    const uint2 c##r = *reinterpret_cast<const uint2 *>(base + (r) * SP);               \
    const uint32_t Wm = *reinterpret_cast<const uint32_t *>(base + (r) * SP + lo);      \
    const uint32_t W2 = *reinterpret_cast<const uint32_t *>(base + (r) * SP + hi);      \
    const uint32_t W0 = c##r.x, W1 = c##r.y;
#define ORBFE_PAIR(A0, A1, B0, B1)                                   \
    acc0 = pair_gt<kMode>(v0, A0, B0, k7, one, acc0);                \
    acc1 = pair_gt<kMode>(v1, A1, B1, k7, one, acc1);
    // centre row (ring 4 = (+3, 0), ring 12 = (-3, 0))
    ORBFE_ROW(3, m3, v0, v1, p3)
    acc0 = acc1 = 0xffffffffu;
    ORBFE_PAIR(__funnelshift_r(v0, v1, 24), __funnelshift_r(v1, p3, 24), __funnelshift_r(m3, v0, 8), __funnelshift_r(v0, v1, 8))
    {   // rows y+6 (dy = +3) and y (dy = -3): pairs (0, 8), (1, 9), (15, 7)
        ORBFE_ROW(6, dm, d0, d1, d2)
        ORBFE_ROW(0, um, u0, u1, u2)
        ORBFE_PAIR(d0, d1, u0, u1)
        ORBFE_PAIR(__funnelshift_r(d0, d1, 8), __funnelshift_r(d1, d2, 8), __funnelshift_r(um, u0, 24), __funnelshift_r(u0, u1, 24))
        ORBFE_PAIR(__funnelshift_r(dm, d0, 24), __funnelshift_r(d0, d1, 24), __funnelshift_r(u0, u1, 8), __funnelshift_r(u1, u2, 8))
    }
    {   // rows y+5 (dy = +2) and y+1 (dy = -2): pairs (2, 10), (14, 6)
        ORBFE_ROW(5, dm, d0, d1, d2)
        ORBFE_ROW(1, um, u0, u1, u2)
        const uint32_t dmid = __funnelshift_r(d0, d1, 16), umid = __funnelshift_r(u0, u1, 16);
        ORBFE_PAIR(dmid, __funnelshift_r(d1, d2, 16), __funnelshift_r(um, u0, 16), umid)
        ORBFE_PAIR(__funnelshift_r(dm, d0, 16), dmid, umid, __funnelshift_r(u1, u2, 16))
    }
    {   // rows y+4 (dy = +1) and y+2 (dy = -1): pairs (3, 11), (13, 5)
        ORBFE_ROW(4, dm, d0, d1, d2)
        ORBFE_ROW(2, um, u0, u1, u2)
        ORBFE_PAIR(__funnelshift_r(d0, d1, 24), __funnelshift_r(d1, d2, 24), __funnelshift_r(um, u0, 8), __funnelshift_r(u0, u1, 8))
        ORBFE_PAIR(__funnelshift_r(dm, d0, 8), __funnelshift_r(d0, d1, 8), __funnelshift_r(u0, u1, 24), __funnelshift_r(u1, u2, 24))
    }
#undef ORBFE_PAIR
#undef ORBFE_ROW
#undef __funnelshift_r
}

// bytes [0, n) set, n in [0, 4]
__device__ __forceinline__ uint32_t low_bytes(int n) { return n >= 4 ? 0xffffffffu : ((1u << (8 * n)) - 1u); }

struct FastArgs {
    uint32_t *slots; int *cell_cnt; const int *blk_tab;
    int cells_per_frame, t_ini, t_min;
    int one;        // 1: multiplier that keeps the stage-A adds on the FMA pipe (an immediate would be folded into an IADD)
    int flags;      // bit 0: exact (masked) stage-A compares, for A/B testing; bit 1: funnel shifts of stage A on the FMA pipe instead of SHF (experiment)
    uint3 shift_mul; // 2^24, 2^16, 2^8: the multipliers of the FMA-pipe funnel shifts by 8, 16, 24 bits
};

template <bool kTMA>
__global__ void __launch_bounds__(256, 6) k_fast(const __grid_constant__ LevelSet L, const __grid_constant__ TmapSet T, const FastArgs a) {
    constexpr int SP = kBoxW;
    __shared__ __align__(128) uint8_t tile[kFastBoxH * SP];
    // m of strip pixel (x, y) at [(y+1)*SP + x+1], zero elsewhere; between stages A and A2 rows 0..ch-1 hold the pass words of stage A
    __shared__ __align__(16) uint8_t mmap[32 * SP];
    __shared__ uint16_t queue[kStripW * kCell];                         // y << 8 | x  (strip coordinates)
    __shared__ uint32_t rowmask[kCellsPerBlk][32];                      // NMS survivors of cell row r, bit = cx
    __shared__ int s_has[kCellsPerBlk], s_scan[8];
    __shared__ __align__(8) uint64_t bar;

    const int frame = blockIdx.y, tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int packed = __ldg(&a.blk_tab[blockIdx.x]);                                   // level | cell row << 4 | strip << 16
    const int l = packed & 15, ci = (packed >> 4) & 0xfff, cg = packed >> 16;
    const LevelGeom &G = L.lv[l];
    const int px = kEdge - 3 + kStripW * cg, py = kEdge - 3 + kCell * ci;               // 16-byte aligned x origin
    stage_box<kTMA, 256>(tile, &bar, &T.m[l], L.img[l] + (size_t) frame * G.frame_stride, G.pitch, G.h, px, py, frame, kFastBoxH);

    const int strip_w = min(kStripW, G.w - kEdge - (kEdge + kStripW * cg));             // pixels of this strip inside maxBorderX
    const int ch = min(kCell, G.h - kEdge - (kEdge + kCell * ci));                      // rows inside maxBorderY
    const uint8_t *strip = tile + 3 * SP + 3;                                           // pixel (x, y) at strip[y * SP + x]
    unsigned open = (1u << ((strip_w + kCell - 1) / kCell)) - 1u;                       // cells that still need a result
    const uint32_t one = (uint32_t) a.one;

#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        const int t = round == 0 ? a.t_ini : a.t_min;
        rowmask[wid][lane] = 0;
        for (int i = tid; i < 32 * SP / 16; i += 256) reinterpret_cast<uint4 *>(mmap)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();

        // ---- A: opposite-pair rejection test, one warp per strip row, lane = aligned words 2*lane, 2*lane+1 (word j = pixels 4j-3 .. 4j)
        {
            const uint32_t k7 = (uint32_t) (127 - (t & 127)) * 0x01010101u;
            // candidate pixels of this lane: strip x = 8*lane + p - 3 in [0, strip_w), in a cell that is still open
            const int nv = strip_w + 3 - 8 * lane;
            uint32_t vm0 = low_bytes(max(nv, 0)) & 0x80808080u, vm1 = low_bytes(max(nv - 4, 0)) & 0x80808080u;
            if (lane == 0) vm0 &= 0xff000000u;
            if (round) {
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int x = max(8 * lane + p - 3, 0);
                    if (!((open >> ((x * 2185) >> 16)) & 1u)) { if (p < 4) vm0 &= ~(0x80u << (8 * p)); else vm1 &= ~(0x80u << (8 * (p - 4))); }
                }
            }
            const int mode = t >= 128 ? 2 : (a.flags & 1) ? 1 : (a.flags & 2) ? 0 : 3;
            for (int y = wid; y < ch; y += 8) {
                uint32_t a0, a1;
                if (mode == 0) fast_pairs_row<0, true>(tile, y, lane, k7, one, a.shift_mul, a0, a1);
                else if (mode == 3) fast_pairs_row<0, false>(tile, y, lane, k7, one, a.shift_mul, a0, a1);
                else if (mode == 1) fast_pairs_row<1, false>(tile, y, lane, k7, one, a.shift_mul, a0, a1);
                else fast_pairs_row<2, false>(tile, y, lane, k7, one, a.shift_mul, a0, a1);
                *reinterpret_cast<uint2 *>(mmap + y * SP + 8 * lane) = make_uint2(a0 & vm0, a1 & vm1);
            }
        }
        __syncthreads();
        // ---- A2: expand the pass words into the queue (order is irrelevant): thread = 8 consecutive words of one row; the words
        // are cleared on the way, which leaves the score map zeroed for stage B
        int n;
        {
            unsigned bits = 0;                                          // bit 8b + i = byte b of word i
            const int y = tid >> 3, seg = tid & 7;
            if (y < ch) {
                uint4 *w = reinterpret_cast<uint4 *>(mmap + y * SP + 32 * seg);
                const uint4 wa = w[0], wb = w[1];
                w[0] = make_uint4(0, 0, 0, 0); w[1] = make_uint4(0, 0, 0, 0);
                // the stored words only hold bit 7 of each byte
                bits = (wa.x >> 7) | (wa.y >> 6) | (wa.z >> 5) | (wa.w >> 4) | (wb.x >> 3) | (wb.y >> 2) | (wb.z >> 1) | wb.w;
            }
            int o = block_scan_excl<256>(__popc(bits), n, s_scan);
            const int xbase = (y << 8) + 32 * seg - 3;
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                queue[o++] = (uint16_t) (xbase + 4 * (b & 7) + (b >> 3));
            }
        }
        __syncthreads();

        // ---- B: exact measure of the queued pixels, one per thread; corners (m > t) go into the strip-wide score map
        for (int k = tid; k < n; k += 256) {
            const int pos = queue[k];
            const int m = fast_measure1(strip + (pos >> 8) * SP + (pos & 255));
            if (m > t) mmap[pos + SP + 1] = (uint8_t) m;
        }
        __syncthreads();

        // ---- C: non-max suppression inside the cell (strictly greater than the 8 neighbours; outside the cell = 0)
        for (int k = tid; k < n; k += 256) {
            const int pos = queue[k];
            const uint8_t *p = mmap + pos + SP + 1;
            const int m = p[0];
            if (m == 0) continue;                            // rejected by the exact measure
            const int x = pos & 255, y = pos >> 8, c = (x * 2185) >> 16, cx = x - kCell * c;
            int nb = max(p[-SP], p[SP]);
            if (cx > 0) nb = max(nb, max(max(p[-SP - 1], p[-1]), p[SP - 1]));
            if (cx < kCell - 1 && x + 1 < strip_w) nb = max(nb, max(max(p[-SP + 1], p[1]), p[SP + 1]));
            if (m > nb) atomicOr(&rowmask[c][y], 1u << cx);
        }
        __syncthreads();

        // ---- D: one warp per cell, lane = cell row: ordered emission (ORBExtractor.cpp:609-615)
        {
            unsigned mask = rowmask[wid][lane];
            const bool has = __any_sync(0xffffffffu, mask != 0);
            if (lane == 0) s_has[wid] = has;
            if (((open >> wid) & 1u) && (has || round == 1)) {
                const int cnt = __popc(mask);
                int inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
                const int total = __shfl_sync(0xffffffffu, inc, 31);
                const int cj = cg * kCellsPerBlk + wid;
                const int cell = G.cell_base + ci * G.n_cols + cj;
                uint32_t *slot = a.slots + ((size_t) frame * a.cells_per_frame + cell) * kSlotCap + (inc - cnt);
                while (mask) {
                    const int cx = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int m = mmap[(lane + 1) * SP + kCell * wid + cx + 1];
                    *slot++ = (uint32_t) (cj * kCell + cx) | ((uint32_t) (ci * kCell + lane) << 12) | ((uint32_t) (m - 1) << 24);
                }
                if (lane == 0) a.cell_cnt[(size_t) frame * a.cells_per_frame + cell] = total;
            }
        }
        __syncthreads();                                     // s_has is complete; stage D reads of mmap / rowmask precede the next round
        unsigned has = 0;
#pragma unroll
        for (int c = 0; c < kCellsPerBlk; ++c) has |= s_has[c] ? 1u << c : 0u;
        open &= ~has;
        if (open == 0) break;
    }
}

// ------------------------------------------------------------------------------------------------
// K2 (second formulation)  k_fast_planes: the same four stages, with stage A rebuilt around two observations.
//   1. |I(q) - I(p)| is symmetric: ring pixel k+8 of centre p is p - D_k, and its test against p is the test of ring pixel k of centre
//      p - D_k.  So only EIGHT difference planes F_D(p) = [|I(p + D) - I(p)| > t] exist (one per opposite pair), and the pair test of
//      (k, k+8) at p is F_D(p) | F_D(p - D).  Half the absolute differences and threshold adds of the per-pixel formulation.
//   2. Once a plane is packed one bit per pixel, the realignment by D costs one funnel shift per 32 pixels instead of one per 4,
//      and the AND over the eight pairs runs on 32 pixels per LOP3.
// Phase 1 (thread = 16 pixels of one row, all eight directions): directions are chosen with Dy >= 0, so a task reads rows r .. r+3;
// a direction with Dx > 0 displaces the ring row, one with Dx < 0 displaces the centre row instead (its plane is then stored |Dx| bits
// to the left; phase 2 compensates), which lets the displaced rows be shared: 24 byte funnel shifts per task.  Per 4 pixels and
// direction: VABSDIFF4, the threshold add as IMAD (FMA pipe), LOP3 (y | x) & 0x80808080 and one IDP4A (FMA pipe) that gathers the four
// bit-7 flags with the byte weights 1, 2, 4, 8 / 16, 32, 64, 128 into the row word: 2 ALU + 2 FMA instructions.
// Phase 2 (thread = 32 pixels): the eight pair terms from the planes of rows y, y-1, y-2, y-3 -> pass bits in pixel order, expanded
// into the queue by the same thread (queue offsets from a warp scan and one shared-memory atomic per warp; the order across warps is
// irrelevant).  The order INSIDE a warp matters: a thread's entries are adjacent pixels and a warp's lie in four adjacent rows, so the
// 17 byte loads per candidate of stage B fall into few words per warp.  (Spreading each thread over four distant bytes of pass bits
// halves the divergence of the expansion loop but makes stage B's loads hit 3.2 banks-conflicted wavefronts each — the kernel then
// sits on the shared-memory pipe at 0.9 wavefronts per clock: measured 1.42 ms against 1.27 ms per 512 C1 frames.)
// B / C / D are k_fast's, with the score map laid out cell by cell (32-byte cell stride, one zero column either side) so that the
// non-max test needs no cell-border predicates, and B leaves the score-map position of each corner in its queue entry (0 = rejected)
// so that C touches only corners.
// 288 threads = 33 plane rows x 16 tasks in two even passes; 5 CTAs per SM (41 KB, 40 registers).
// ------------------------------------------------------------------------------------------------
constexpr int kF2Threads = 288;
constexpr int kF2PlaneRows = kCell + 3;                 // strip rows -3 .. 29
constexpr int kF2Tile = kFastBoxH * kBoxW;              // bytes of the staged box
constexpr int kF2PlaneHalves = kF2PlaneRows * 16;       // u16 entries of one plane

// 16 pass bits of one direction: bit 4j + i = |c - r| > t for byte i of word j.  kMode as in pair_gt.
// The flag bytes (0x80 / 0) of a word are gathered by a dot product with the byte weights (1, 2, 4, 8) or (16, 32, 64, 128): IDP4A runs
// on the FMA pipe and accumulates, so two words cost two instructions there and nothing on the ALU pipe (as IMAD by 0x00204081 +
// funnel shift the gather took one FMA and one ALU instruction per word; the ALU pipe is the one that binds).
template <int kMode>
__device__ __forceinline__ uint32_t diff_flags(uint32_t c, uint32_t r, uint32_t k7, uint32_t one) {
    const uint32_t x = __vabsdiffu4(c, r);
    if constexpr (kMode == 0) return ((x * one + k7) | x) & 0x80808080u;
    else if constexpr (kMode == 1) return (((x & 0x7f7f7f7fu) + k7) | x) & 0x80808080u;
    else return (((x & 0x7f7f7f7fu) + k7) & x) & 0x80808080u;
}
template <int kMode>
__device__ __forceinline__ uint16_t diff_bits(const uint4 c, const uint4 r, uint32_t k7, uint32_t one) {
    constexpr uint32_t W0 = 0x08040201u, W1 = 0x80402010u;
    uint32_t lo = __dp4a(diff_flags<kMode>(c.x, r.x, k7, one), W0, 0u);          // flags * 128: bits 7 .. 14
    lo = __dp4a(diff_flags<kMode>(c.y, r.y, k7, one), W1, lo);
    uint32_t hi = __dp4a(diff_flags<kMode>(c.z, r.z, k7, one), W0, 0u);
    hi = __dp4a(diff_flags<kMode>(c.w, r.w, k7, one), W1, hi);
    return (uint16_t) ((lo >> 7) + hi * 2u);
}
// 16 bytes starting s bytes to the right of v (nx = the word after v)
__device__ __forceinline__ uint4 bytes_right(const uint4 v, uint32_t nx, int s) {
    return make_uint4(__funnelshift_r(v.x, v.y, 8 * s), __funnelshift_r(v.y, v.z, 8 * s), __funnelshift_r(v.z, v.w, 8 * s), __funnelshift_r(v.w, nx, 8 * s));
}

// phase 1 for the plane rows [0, kF2PlaneRows): tile = the staged box (pitch kBoxW), planes[d][row][16] u16
// kCompact (the minThFAST round): only the 16-pixel groups in `segs` (bit = group of a row) are computed — those that hold plane
// bits of a cell that is still open; tasks are (row, n-th listed group), so the lanes stay dense.
template <int kMode, bool kCompact>
__device__ __forceinline__ void fast_planes(const uint8_t *tile, uint16_t *planes, int tid, uint32_t k7, uint32_t one, uint32_t segs) {
    constexpr int SP = kBoxW, PH = kF2PlaneHalves;
    const int nseg = kCompact ? __popc(segs) : 16;
    for (int task = tid; task < kF2PlaneRows * nseg; task += kF2Threads) {
        int row, seg;
        if constexpr (kCompact) { row = task / nseg; seg = __fns(segs, 0, task - row * nseg + 1); }
        else { row = task >> 4; seg = task & 15; }
        const uint8_t *p = tile + row * SP + 16 * seg;                           // centre row
        uint16_t *out = planes + row * 16 + seg;
#define ORBFE_LD(off) (*reinterpret_cast<const uint4 *>(p + (off)))
        // the word after a 16-byte group is the first word of the next lane's group (the last group of a row gets a word that only
        // reaches columns >= 253, which are never candidates); compact tasks have no such neighbour and load it (the tile is padded)
        const unsigned am = __activemask();
#define ORBFE_NX(off, v) (kCompact ? *reinterpret_cast<const uint32_t *>(p + (off) + 16) : __shfl_down_sync(am, (v).x, 1))
        const uint4 c0 = ORBFE_LD(0); const uint32_t c0n = ORBFE_NX(0, c0);
        {
            const uint4 r3 = ORBFE_LD(3 * SP); const uint32_t r3n = ORBFE_NX(3 * SP, r3);
            out[0 * PH] = diff_bits<kMode>(c0, r3, k7, one);                               // ( 0, 3)  ring 0 / 8
            out[1 * PH] = diff_bits<kMode>(c0, bytes_right(r3, r3n, 1), k7, one);          // ( 1, 3)  ring 1 / 9
            out[7 * PH] = diff_bits<kMode>(bytes_right(c0, c0n, 1), r3, k7, one);          // (-1, 3)  ring 15 / 7, stored 1 bit left
        }
        {
            const uint4 r2 = ORBFE_LD(2 * SP); const uint32_t r2n = ORBFE_NX(2 * SP, r2);
            out[2 * PH] = diff_bits<kMode>(c0, bytes_right(r2, r2n, 2), k7, one);          // ( 2, 2)  ring 2 / 10
            out[6 * PH] = diff_bits<kMode>(bytes_right(c0, c0n, 2), r2, k7, one);          // (-2, 2)  ring 14 / 6, stored 2 bits left
        }
        {
            const uint4 r1 = ORBFE_LD(SP); const uint32_t r1n = ORBFE_NX(SP, r1);
            const uint4 c3 = bytes_right(c0, c0n, 3);
            out[3 * PH] = diff_bits<kMode>(c0, bytes_right(r1, r1n, 3), k7, one);          // ( 3, 1)  ring 3 / 11
            out[4 * PH] = diff_bits<kMode>(c0, c3, k7, one);                               // ( 3, 0)  ring 4 / 12
            out[5 * PH] = diff_bits<kMode>(c3, r1, k7, one);                               // (-3, 1)  ring 13 / 5, stored 3 bits left
        }
#undef ORBFE_LD
#undef ORBFE_NX
    }
}

// the most significant set bit of a non-zero word (bfind: one FLO)
__device__ __forceinline__ int top_bit(uint32_t v) { int b; asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(v)); return b; }

struct Fast2Args {
    uint32_t *slots; int *cell_cnt; const int *blk_tab;
    int cells_per_frame, t_ini, t_min;
    uint32_t one;       // 1: keeps the threshold add an IMAD (an immediate would be folded into an IADD on the ALU pipe)
    int exact;          // masked (exact) threshold adds also below 128, for A/B measurements
};

template <bool kTMA>
__global__ void __launch_bounds__(kF2Threads, 5) k_fast_planes(const __grid_constant__ LevelSet L, const __grid_constant__ TmapSet T, const Fast2Args a) {
    constexpr int SP = kBoxW, NT = kF2Threads;
    __shared__ __align__(128) uint8_t tile[kF2Tile + 16];                // + the word a compact phase-1 task reads after the last group
    __shared__ __align__(16) uint16_t planes[8 * kF2PlaneHalves];
    __shared__ __align__(16) uint8_t mmap[32 * SP];                      // score of cell c, column cx, strip row y at [(y+1)*SP + 32*c + cx + 1], zero elsewhere
    __shared__ uint16_t queue[kStripW * kCell];                          // y << 8 | tile column; after stage B: score-map position or 0
    __shared__ uint32_t rowmask[kCellsPerBlk][32];
    __shared__ int s_n, s_open[kCellsPerBlk];
    __shared__ __align__(8) uint64_t bar;

    const int frame = blockIdx.y, tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int packed = __ldg(&a.blk_tab[blockIdx.x]);                                   // level | cell row << 4 | strip << 16
    const int l = packed & 15, ci = (packed >> 4) & 0xfff, cg = packed >> 16;
    const LevelGeom &G = L.lv[l];
    const int px = kEdge - 3 + kStripW * cg, py = kEdge - 3 + kCell * ci;               // tile column c = strip x + 3, tile row = strip y + 3
    // the box is requested first; the per-round state of round 0 is cleared while it is in flight
    if constexpr (kTMA) {
        if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
        __syncthreads();
        if (tid == 0) { mbar_expect_tx(&bar, (uint32_t) kF2Tile); tma_load_3d(tile, &T.m[l], &bar, px, py, frame); }
    }
    const int strip_w = min(kStripW, G.w - kEdge - (kEdge + kStripW * cg));             // pixels of this strip inside maxBorderX
    const int ch = min(kCell, G.h - kEdge - (kEdge + kCell * ci));                      // rows inside maxBorderY
    unsigned open = (1u << ((strip_w + kCell - 1) / kCell)) - 1u;                       // cells that still need a result

#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        const int t = round == 0 ? a.t_ini : a.t_min;
        if (tid < kCellsPerBlk * 32) rowmask[wid][lane] = 0;
        if (tid == 0) s_n = 0;
        for (int i = tid; i < 32 * SP / 16; i += NT) reinterpret_cast<uint4 *>(mmap)[i] = make_uint4(0, 0, 0, 0);
        if (round == 0) {
            if constexpr (kTMA) mbar_wait(&bar, 0);
            else stage_box<false, NT>(tile, &bar, &T.m[l], L.img[l] + (size_t) frame * G.frame_stride, G.pitch, G.h, px, py, frame, kFastBoxH);
        }
        // ---- A, phase 1: the eight difference planes
        {
            const uint32_t k7 = (uint32_t) (127 - (t & 127)) * 0x01010101u;
            if (round == 0) {
                if (t >= 128) fast_planes<2, false>(tile, planes, tid, k7, a.one, 0xffffu);
                else if (a.exact) fast_planes<1, false>(tile, planes, tid, k7, a.one, 0xffffu);
                else fast_planes<0, false>(tile, planes, tid, k7, a.one, 0xffffu);
            } else {
                // plane columns read for the candidates of cell c (tile columns 30c+3 .. 30c+32): 30c .. 30c+32
                uint32_t segs = 0;
#pragma unroll
                for (int c = 0; c < kCellsPerBlk; ++c)
                    if ((open >> c) & 1u) segs |= (2u << ((kCell * c + 32) >> 4)) - (1u << ((kCell * c) >> 4));
                segs &= 0xffffu;
                if (t >= 128) fast_planes<2, true>(tile, planes, tid, k7, a.one, segs);
                else if (a.exact) fast_planes<1, true>(tile, planes, tid, k7, a.one, segs);
                else fast_planes<0, true>(tile, planes, tid, k7, a.one, segs);
            }
        }
        __syncthreads();
        // ---- A, phase 2: pair terms and their AND, 32 pixels per thread, and the pass bits -> queue.  Queue entries of a thread are
        // adjacent pixels and those of a warp lie in four adjacent rows: stage B's ring loads of a warp then fall into few words.
        {
            uint32_t pass = 0;
            const int y = tid >> 3, w = tid & 7;
            if (y < ch) {
                const uint32_t *P = reinterpret_cast<const uint32_t *>(planes) + (y + 3) * 8 + w;      // plane row of strip row y
                constexpr int PW = kF2PlaneRows * 8;                                                      // words per plane
#define ORBFE_AT(d, dy) P[(d) * PW - (dy) * 8]
#define ORBFE_SHL(d, dy, s) __funnelshift_l(w ? P[(d) * PW - (dy) * 8 - 1] : 0u, P[(d) * PW - (dy) * 8], s)
                pass = ORBFE_AT(0, 0) | ORBFE_AT(0, 3);
                pass &= ORBFE_AT(1, 0) | ORBFE_SHL(1, 3, 1);
                pass &= ORBFE_AT(2, 0) | ORBFE_SHL(2, 2, 2);
                pass &= ORBFE_AT(3, 0) | ORBFE_SHL(3, 1, 3);
                pass &= ORBFE_AT(4, 0) | ORBFE_SHL(4, 0, 3);
                pass &= ORBFE_SHL(5, 0, 3) | ORBFE_AT(5, 1);
                pass &= ORBFE_SHL(6, 0, 2) | ORBFE_AT(6, 2);
                pass &= ORBFE_SHL(7, 0, 1) | ORBFE_AT(7, 3);
#undef ORBFE_AT
#undef ORBFE_SHL
                // candidate columns: c - 3 in [0, strip_w), in a cell that is still open
                const int c0 = 32 * w - 3;                                                  // strip x of bit 0
                const int hi0 = strip_w - c0;
                uint32_t valid = hi0 <= 0 ? 0u : (hi0 >= 32 ? 0xffffffffu : (1u << hi0) - 1u);
                if (w == 0) valid &= ~7u;
                if (round) {
#pragma unroll 1
                    for (int c = 0; c < kCellsPerBlk; ++c) {
                        if (!((open >> c) & 1u)) {
                            const int lo = max(kCell * c - c0, 0), hi = min(kCell * (c + 1) - c0, 32);
                            if (hi > lo) valid &= ~((hi >= 32 ? 0xffffffffu : (1u << hi) - 1u) & ~((1u << lo) - 1u));
                        }
                    }
                }
                pass &= valid;
            }
            const int cnt = __popc(pass);
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
            int base = 0;
            if (lane == 31 && inc) base = atomicAdd(&s_n, inc);
            base = __shfl_sync(0xffffffffu, base, 31);
            uint16_t *q = queue + base + inc - cnt;
            const int pos0 = 32 * tid;                                                      // y << 8 | 32 * w
            while (pass) {
                const int b = top_bit(pass);
                pass ^= 1u << b;
                *q++ = (uint16_t) (pos0 + b);
            }
        }
        __syncthreads();
        const int n = s_n;

        // ---- B: exact measure of the queued pixels; corners (m > t) go into the score map, their map position into the queue entry
        for (int k = tid; k < n; k += NT) {
            const int pos = queue[k];
            const int m = fast_measure1(tile + 3 * SP + pos);
            int mp = 0;
            if (m > t) {
                const int x = (pos & 255) - 3;
                mp = (pos & 0xff00) + SP + x + 2 * ((x * 2185) >> 16) + 1;
                mmap[mp] = (uint8_t) m;
            }
            queue[k] = (uint16_t) mp;
        }
        __syncthreads();

        // ---- C: non-max suppression inside the cell (strictly greater than the 8 neighbours; the cell's border columns / rows are 0)
        for (int k = tid; k < n; k += NT) {
            const int mp = queue[k];
            if (mp == 0) continue;
            const uint8_t *p = mmap + mp;
            const uint32_t m = p[0];
            const uint32_t n0 = __vimax3_u32(p[-SP - 1], p[-SP], p[-SP + 1]), n1 = __vimax3_u32(p[-1], p[1], p[SP - 1]);
            const uint32_t nb = __vimax3_u32(n0, n1, max((uint32_t) p[SP], (uint32_t) p[SP + 1]));
            if (m > nb) atomicOr(&rowmask[(mp & 255) >> 5][(mp >> 8) - 1], 1u << ((mp & 31) - 1));
        }
        __syncthreads();

        // ---- D: one warp per cell, lane = cell row: ordered emission (ORBExtractor.cpp:609-615)
        bool still_open = false;
        if (wid < kCellsPerBlk) {
            unsigned mask = rowmask[wid][lane];
            const bool has = __any_sync(0xffffffffu, mask != 0);
            const bool mine = (open >> wid) & 1u;
            still_open = mine && !has;
            if (mine && (has || round == 1)) {
                const int cnt = __popc(mask);
                int inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
                const int total = __shfl_sync(0xffffffffu, inc, 31);
                const int cj = cg * kCellsPerBlk + wid;
                const int cell = G.cell_base + ci * G.n_cols + cj;
                uint32_t *slot = a.slots + ((size_t) frame * a.cells_per_frame + cell) * kSlotCap + (inc - cnt);
                while (mask) {
                    const int cx = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int m = mmap[(lane + 1) * SP + 32 * wid + cx + 1];
                    *slot++ = (uint32_t) (cj * kCell + cx) | ((uint32_t) (ci * kCell + lane) << 12) | ((uint32_t) (m - 1) << 24);
                }
                if (lane == 0) a.cell_cnt[(size_t) frame * a.cells_per_frame + cell] = total;
            }
        }
        if (round) break;
        // cells without a corner at iniThFAST go through a second round at minThFAST; the barrier also orders stage D's reads of the
        // score map / row masks before the next round's writes.  Each warp reports its own cell.
        if (__syncthreads_or(still_open) == 0) break;
        {
            // rebuild `open` identically in every thread: warp w's cell stays open iff it was open and had no corner; publish through s_open
            if (lane == 0 && wid < kCellsPerBlk) s_open[wid] = still_open;
            __syncthreads();
            unsigned o2 = 0;
#pragma unroll
            for (int c = 0; c < kCellsPerBlk; ++c) o2 |= s_open[c] ? 1u << c : 0u;
            open = o2;
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4  DistributeOctree.  One CTA per (level, frame).  See DESIGN.md "quadtree" for the derivation:
// a node is a path prefix, its key points are the candidates (in reference order) whose coordinates descend to it,
// so no per-node vectors are moved — only per-node counts; list order is recovered from the creation order of slots.
// ------------------------------------------------------------------------------------------------
struct OctArgs {
    const uint32_t *slots; const int *cell_cnt; int *cell_off;
    uint32_t *cand; int *cur; uint8_t *nodes; int *lists; uint32_t *kp; int *nkp; int *ncand; int *err;
    int cells_per_frame, cand_per_frame, nodes_per_frame, lists_per_frame, kp_per_frame;
    int smem_node_cap;      // nodes that fit the shared-memory pool
    int sort_cap;           // power of two
};

struct NodeBounds { short x0, x1, y0, y1; };

template <int NT>
__global__ void __launch_bounds__(NT, 2048 / NT) k_octree(const __grid_constant__ LevelSet L, const OctArgs a) {
    extern __shared__ __align__(16) uint8_t dyn[];
    __shared__ int s_warp[NT / 32];
    __shared__ int s_bcast[4];

    const int l = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const LevelGeom &G = L.lv[l];
    const int n_cells = G.n_cols * G.n_rows;
    const int *cell_cnt = a.cell_cnt + (size_t) frame * a.cells_per_frame + G.cell_base;
    int *cell_off = a.cell_off + (size_t) frame * a.cells_per_frame + G.cell_base;
    const uint32_t *slots = a.slots + ((size_t) frame * a.cells_per_frame + G.cell_base) * kSlotCap;
    uint32_t *cand = a.cand + (size_t) frame * a.cand_per_frame + G.cand_off;
    int *cur = a.cur + (size_t) frame * a.cand_per_frame + G.cand_off;
    int *E = a.lists + (size_t) frame * a.lists_per_frame + G.list_off;     // expandable nodes, creation order
    int *S = E + G.kp_cap;                                                  // split list of the current pass
    uint32_t *kp_out = a.kp + (size_t) frame * a.kp_per_frame + G.kp_off;

    // dynamic smem: [sort keys: sort_cap u64][node pool: smem_node_cap * 16 B].  The pool starts in shared memory (sized for the
    // slot counts frames produce in practice) and moves to the level's global pool — sized by the proven bound, see configure() —
    // the first time a pass would not fit, so a valid image can never exhaust it.
    unsigned long long *skey = reinterpret_cast<unsigned long long *>(dyn);
    // The pool is one base pointer and its capacity: [bounds 8 B x cap][count 4 B x cap][first child 4 B x cap]
    int cap_cur = min(a.smem_node_cap, G.node_cap);
    uint8_t *pool = dyn + (size_t) a.sort_cap * 8;
#define nbnd (reinterpret_cast<NodeBounds *>(pool))
#define ncnt (reinterpret_cast<int *>(pool + (size_t) cap_cur * 8))
#define nchild (reinterpret_cast<int *>(pool + (size_t) cap_cur * 12))
    auto use_global_pool = [&](int n_live) {            // block-uniform; copies the n_live slots in use
        uint8_t *gpool = a.nodes + ((size_t) frame * a.nodes_per_frame + G.node_off) * 16;
        NodeBounds *gb = reinterpret_cast<NodeBounds *>(gpool);
        int *gc = reinterpret_cast<int *>(gpool + (size_t) G.node_cap * 8), *gch = gc + G.node_cap;
        for (int j = tid; j < n_live; j += NT) { gb[j] = nbnd[j]; gc[j] = ncnt[j]; gch[j] = nchild[j]; }
        __syncthreads();
        pool = gpool; cap_cur = G.node_cap;
    };

    // ---- 1. candidates of this level in reference order: exclusive scan of the per-cell counts, then gather
    int n = 0;
    for (int base = 0; base < n_cells; base += NT) {
        const int c = base + tid;
        const int v = c < n_cells ? cell_cnt[c] : 0;
        int tot;
        const int ex = block_scan_excl<NT>(v, tot, s_warp);
        if (c < n_cells) cell_off[c] = n + ex;
        n += tot;
    }
    if (tid == 0) { a.ncand[frame * ORBFE_MAX_LEVELS + l] = n; a.nkp[frame * ORBFE_MAX_LEVELS + l] = 0; }
    if (n == 0) return;
    if (n > G.cand_cap || n >= (1 << 20)) { if (tid == 0) atomicExch(a.err, 1); return; }

    const int W = G.w - 2 * kEdge, H = G.h - 2 * kEdge;                    // maxX-minX, maxY-minY
    const int n_ini = (int) ceilf((float) W / (float) H);                   // ORBExtractor.cpp:645
    const int h_x = (int) ceilf((float) W / (float) n_ini);                 // :646
    if (n_ini * 5 > G.node_cap) { if (tid == 0) atomicExch(a.err, 2); return; }
    if (n_ini * 5 > cap_cur) use_global_pool(0);
    for (int r = tid; r < n_ini; r += NT) {                                 // :652-670 (last root ends at maxX, sic)
        NodeBounds b; b.x0 = (short) (h_x * r); b.x1 = (short) (r == n_ini - 1 ? G.w - kEdge : h_x * (r + 1)); b.y0 = 0; b.y1 = (short) H;
        nbnd[r] = b; ncnt[r] = 0; nchild[r] = 0;
    }
    __syncthreads();
    {   // 8 lanes per cell (a cell holds ~8 candidates on dense frames, up to kSlotCap), 4 cells per warp step
        const float inv_hx = 1.0f / (float) h_x;              // root = x / h_x (:673-675); x < 4096, so (x + 0.5) * (1 / h_x) truncates exactly
        const int sub = lane >> 3, kl = lane & 7;
        for (int c0 = 4 * wid; c0 < n_cells; c0 += 4 * (NT / 32)) {
            const int c = c0 + sub;
            const int cnt = c < n_cells ? cell_cnt[c] : 0, off = c < n_cells ? cell_off[c] : 0;
            for (int k = kl; k < cnt; k += 8) {
                const uint32_t v = slots[(size_t) c * kSlotCap + k];
                cand[off + k] = v;
                const int r = n_ini > 1 ? (int) (((float) (v & 0xfffu) + 0.5f) * inv_hx) : 0;
                cur[off + k] = r;
                const unsigned act = __activemask();
                const unsigned same = __match_any_sync(act, r);
                if (lane == __ffs(same) - 1) atomicAdd(&ncnt[r], __popc(same));
            }
        }
    }
    __syncthreads();

    // ---- 2. list bookkeeping.  len = list size; pool_top = next free slot; E = expandable nodes in creation order.
    int len = 0, ne = 0, pool_top = n_ini;
    {   // roots: empty ones are erased, single-point ones are final (:677-686); the first pass splits the rest in root order
        for (int base = 0; base < n_ini; base += NT) {
            const int r = base + tid;
            const int c = r < n_ini ? ncnt[r] : 0;
            int tot_ne, tot_ex;
            const int ex = block_scan_excl<NT>(c > 1 ? 1 : 0, tot_ex, s_warp);
            block_scan_excl<NT>(c > 0 ? 1 : 0, tot_ne, s_warp);
            if (c > 1) S[ne + ex] = r;
            ne += tot_ex; len += tot_ne;
        }
        __syncthreads();
    }
    int ns = ne;                 // S currently holds the split list of the first breadth pass
    bool careful = false;
    const int nF = G.quota;

    while (true) {
        const int prev = len;
        if (careful) {
            // stable sort of E by size ascending (canonical tie-break = creation order; ORBExtractor.cpp:757)
            int cap2 = 1; while (cap2 < ne) cap2 <<= 1;
            for (int j = tid; j < cap2; j += NT)
                skey[j] = j < ne ? (((unsigned long long) (unsigned) ncnt[E[j]]) << 32) | (unsigned) j : ~0ull;
            __syncthreads();
            for (int k = 2; k <= cap2; k <<= 1)
                for (int jj = k >> 1; jj > 0; jj >>= 1) {
                    for (int i = tid; i < cap2; i += NT) {
                        const int ixj = i ^ jj;
                        if (ixj > i) {
                            const unsigned long long x = skey[i], y = skey[ixj];
                            const bool up = (i & k) == 0;
                            if ((x > y) == up) { skey[i] = y; skey[ixj] = x; }
                        }
                    }
                    __syncthreads();
                }
            for (int j = tid; j < ne; j += NT) S[j] = E[(int) (skey[j] & 0xffffffffu)];
            ns = ne;
            __syncthreads();
        }
        if (ns == 0) break;                                    // nothing expandable: size == prevSize (:750 / :806)
        if (pool_top + 4 * ns > cap_cur) {
            if (pool_top + 4 * ns > G.node_cap) { if (tid == 0) atomicExch(a.err, 3); return; }
            use_global_pool(pool_top);
        }
        // ---- A. allocate 4 child slots per node of the split list (DivideNode geometry, :367-395)
        for (int j = tid; j < ns; j += NT) {
            const int nd = S[j];
            const NodeBounds b = nbnd[nd];
            const int base = pool_top + 4 * j;
            nchild[nd] = base;
            const short mx = (short) (b.x0 + (b.x1 - b.x0) / 2), my = (short) (b.y0 + (b.y1 - b.y0) / 2);
            NodeBounds c0 = {b.x0, mx, b.y0, my}, c1 = {mx, b.x1, b.y0, my}, c2 = {b.x0, mx, my, b.y1}, c3 = {mx, b.x1, my, b.y1};
            nbnd[base] = c0; nbnd[base + 1] = c1; nbnd[base + 2] = c2; nbnd[base + 3] = c3;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) { ncnt[base + qd] = 0; nchild[base + qd] = 0; }
        }
        __syncthreads();
        // ---- B. every candidate of a split node descends one level (:398-407); careful phase: count only (speculative).
        // Four candidates per thread and step, loads first, so the L2 latency of cur[] / cand[] is paid once per step.
        for (int i0 = tid; i0 < n; i0 += 4 * NT) {
            int nd[4]; uint32_t cv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + q * NT;
                nd[q] = i < n ? cur[i] : -1; cv[q] = i < n ? cand[i] : 0u;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int base = nd[q] >= 0 ? nchild[nd[q]] : 0;
                if (base > 0) {
                    const NodeBounds b = nbnd[base];           // TL child: x1 = midX, y1 = midY
                    const int x = (int) (cv[q] & 0xfffu), y = (int) ((cv[q] >> 12) & 0xfffu);
                    const int ch = base + (x < b.x1 ? 0 : 1) + (y < b.y1 ? 0 : 2);
                    if (!careful) cur[i0 + q * NT] = ch;
                    const unsigned act = __activemask();
                    const unsigned same = __match_any_sync(act, ch);
                    if (lane == __ffs(same) - 1) atomicAdd(&ncnt[ch], __popc(same));
                }
            }
        }
        __syncthreads();
        // ---- C. how many of the split list are committed (careful phase stops as soon as len >= nFeatures, :802)
        int n_commit = ns;
        if (careful) {
            int run = len, found = ns;
            for (int base = 0; base < ns; base += NT) {
                const int j = base + tid;
                int delta = 0;
                if (j < ns) {
                    const int b = pool_top + 4 * j;
                    delta = (ncnt[b] > 0) + (ncnt[b + 1] > 0) + (ncnt[b + 2] > 0) + (ncnt[b + 3] > 0) - 1;
                }
                int tot;
                const int ex = block_scan_excl<NT>(delta, tot, s_warp);
                const bool hit = j < ns && run + ex + delta >= nF;
                // first j reaching the quota
                const unsigned bal = __ballot_sync(0xffffffffu, hit);
                if (tid == 0) s_bcast[0] = 0x7fffffff;
                __syncthreads();
                if (bal && lane == 0) atomicMin(&s_bcast[0], base + wid * 32 + __ffs(bal) - 1);
                __syncthreads();
                const int first = s_bcast[0];
                __syncthreads();
                if (first != 0x7fffffff) { found = first; break; }
                run += tot;
            }
            n_commit = found < ns ? found + 1 : ns;
            // revert the uncommitted tail, move the committed candidates down
            for (int j = n_commit + tid; j < ns; j += NT) nchild[S[j]] = 0;
            __syncthreads();
            for (int i = tid; i < n; i += NT) {
                const int nd = cur[i];
                const int base = nchild[nd];
                if (base > 0) {
                    const NodeBounds b = nbnd[base];
                    const uint32_t v = cand[i];
                    const int x = (int) (v & 0xfffu), y = (int) ((v >> 12) & 0xfffu);
                    cur[i] = base + (x < b.x1 ? 0 : 1) + (y < b.y1 ? 0 : 2);
                }
            }
            __syncthreads();
        }
        // ---- D. new list size and the expandable children in creation order
        int new_ne = 0, added = 0;
        const int n_slots = 4 * n_commit;
        for (int base = 0; base < n_slots; base += NT) {
            const int sidx = base + tid;
            const int c = sidx < n_slots ? ncnt[pool_top + sidx] : 0;
            int tot_ex, tot_ne;
            const int ex = block_scan_excl<NT>(c > 1 ? 1 : 0, tot_ex, s_warp);
            block_scan_excl<NT>(c > 0 ? 1 : 0, tot_ne, s_warp);
            if (c > 1) {
                if (new_ne + ex < G.kp_cap) E[new_ne + ex] = pool_top + sidx;
            }
            new_ne += tot_ex; added += tot_ne;
        }
        __syncthreads();
        len = len - n_commit + added;
        pool_top += n_slots;
        if (new_ne > G.kp_cap) { if (tid == 0) atomicExch(a.err, 4); return; }
        ne = new_ne;
        if (careful) {
            if (len >= nF || len == prev) break;                           // :806
        } else {
            if (len > nF || len == prev) break;                            // :750
            if (len + 3 * ne > nF) careful = true;                         // :752
            else {                                                         // next breadth pass walks the list front to back =
                for (int j = tid; j < ne; j += NT) S[j] = E[ne - 1 - j];   // children in reverse creation order
                ns = ne;
                __syncthreads();
            }
        }
    }
    __syncthreads();
    if (len > G.kp_cap) { if (tid == 0) atomicExch(a.err, 5); return; }

    // ---- 3. list order: children slots newest first (push_front), then the surviving roots in order.
    // leaves get nchild = -1 - position; best[] accumulates (score, first index) per leaf.
    uint32_t *best = reinterpret_cast<uint32_t *>(S);      // S is free now (kp_cap entries)
    for (int j = tid; j < len; j += NT) best[j] = 0;
    int run = 0;
    for (int base = 0; base < pool_top; base += NT) {
        const int t = base + tid;
        int slot = -1, leaf = 0;
        if (t < pool_top) {
            slot = t < pool_top - n_ini ? pool_top - 1 - t : t - (pool_top - n_ini);
            leaf = ncnt[slot] > 0 && nchild[slot] == 0;
        }
        int tot;
        const int ex = block_scan_excl<NT>(leaf, tot, s_warp);
        if (leaf) nchild[slot] = -1 - (run + ex);
        run += tot;
    }
    __syncthreads();
    // ---- 4. best response per node, first one wins ties (:813-827)
    for (int i = tid; i < n; i += NT) {
        const int pos = -1 - nchild[cur[i]];
        const uint32_t key = ((cand[i] >> 24) << 20) | (uint32_t) (0xfffff - i);
        atomicMax(&best[pos], key);
    }
    __syncthreads();
    for (int j = tid; j < len; j += NT) {
        const int i = 0xfffff - (int) (best[j] & 0xfffffu);
        const uint32_t v = cand[i];
        kp_out[j] = ((v & 0xfffu) + kEdge) | ((((v >> 12) & 0xfffu) + kEdge) << 12) | (v & 0xff000000u);   // :625-632
    }
    if (tid == 0) a.nkp[frame * ORBFE_MAX_LEVELS + l] = len;
#undef nbnd
#undef ncnt
#undef nchild
}

// ------------------------------------------------------------------------------------------------
// K6  7x7 Gaussian blur, sigma 2, BORDER_REFLECT_101, OpenCV's 8.8 fixed-point path.
// Tile: 224 x 32 outputs from a 256 x 38 staged box (the box starts 16-byte aligned, 13 px left of the halo).  Horizontal pass -> u16 in smem, vertical pass sliding in registers.
// ------------------------------------------------------------------------------------------------
// kc: the packed 8.8 kernel taps, passed as arguments so they live in the constant bank / registers (as literals the compiler
// re-materialises them through uniform registers before every IDP):
//   [0] = {18,34,48,56} bytes, [1] = {48,34,18,0} bytes                    horizontal DP4A pair
//   [2..5] = even output row, [6..9] = odd output row                       vertical DP2A over row pairs (u16x2 . u8x2)
struct BlurArgs { uint8_t *blur; const int *blk_tab; uint32_t kc[10]; };      // blk_tab: block -> level | tile row << 4 | tile column << 16

__host__ __device__ inline void blur_taps(uint32_t (&kc)[10]) {
    kc[0] = 18u | (34u << 8) | (48u << 16) | (56u << 24); kc[1] = 48u | (34u << 8) | (18u << 16);
    kc[2] = 18u | (34u << 8); kc[3] = 48u | (56u << 8); kc[4] = 48u | (34u << 8); kc[5] = 18u;
    kc[6] = 18u << 8; kc[7] = 34u | (48u << 8); kc[8] = 56u | (48u << 8); kc[9] = 34u | (18u << 8);
}

__device__ __forceinline__ int refl101(int p, int n) { return p < 0 ? -p : (p >= n ? 2 * n - 2 - p : p); }

constexpr int kBlurThreads = (kBlurTileW / 4) * (kBlurTileH / 8);      // 224: 56 column quads x 4 row groups, in both passes

template <bool kTMA>
__global__ void __launch_bounds__(kBlurThreads, 6) k_blur(const __grid_constant__ LevelSet L, const __grid_constant__ TmapSet T, const __grid_constant__ BlurArgs a) {
    constexpr int SP = TilePitch<kTMA>::value;
    constexpr int XO = 13;                        // the box starts 16 px left of x0 (16-byte aligned), the 3-px halo at column 13
    constexpr int NT = kBlurThreads, QW = kBlurTileW / 4, NRP = kBlurBoxH / 2;
    static_assert(kBlurTileW % 16 == 0 && kBlurBoxH % 2 == 0, "tile origin must keep the halo at column 13 and rows must pair up");
    __shared__ __align__(128) uint8_t tile[kBlurBoxH * SP];
    __shared__ __align__(16) uint32_t hp[NRP * kBlurTileW];     // horizontal pass, rows paired: H[2p][x] | H[2p+1][x] << 16
    __shared__ __align__(8) uint64_t bar;
    const int frame = blockIdx.y, tid = threadIdx.x;
    const int packed = __ldg(&a.blk_tab[blockIdx.x]);
    const int l = packed & 15, ty = (packed >> 4) & 0xfff, tx = packed >> 16;
    const LevelGeom &G = L.lv[l];
    const int x0 = tx * kBlurTileW, y0 = ty * kBlurTileH;
    stage_box<kTMA, NT>(tile, &bar, &T.m[l], L.img[l] + (size_t) frame * G.frame_stride, G.pitch, G.h, x0 - 3, y0 - 3, frame, kBlurBoxH);
    uint8_t *t = tile + XO;                       // t[r*SP + c] = pixel (x0-3+c, y0-3+r)
    const int w = G.w, h = G.h;
    // reflect-101 fix-up of the columns / rows of the box that lie outside the image (only border tiles)
    const int need_w = min(kBlurTileW, w - x0) + 6, need_h = min(kBlurTileH, h - y0) + 6;
    if (x0 == 0 || x0 + kBlurTileW + 3 > w) {
        for (int idx = tid; idx < kBlurBoxH * 6; idx += NT) {
            const int r = idx / 6, k = idx - r * 6;
            const int c = k < 3 ? k : need_w - 6 + k;          // 3 columns left of x0, 3 right of the last needed column
            const int gx = x0 - 3 + c;
            if (gx < 0 || gx >= w) t[r * SP + c] = t[r * SP + refl101(gx, w) - (x0 - 3)];
        }
        __syncthreads();
    }
    if (y0 == 0 || y0 + kBlurTileH + 3 > h) {
        for (int idx = tid; idx < 6 * kBoxW; idx += NT) {
            const int k = idx >> 8, c = idx & 255;
            const int r = k < 3 ? k : need_h - 6 + k;
            const int gy = y0 - 3 + r;
            if ((gy < 0 || gy >= h) && c < need_w) t[r * SP + c] = t[(refl101(gy, h) - (y0 - 3)) * SP + c];
        }
        __syncthreads();
    }
    const int seg = tid / QW, qx = tid - seg * QW;      // column quad, row group (both passes)
    // horizontal pass on packed bytes: H[x] = dp4a(src[x-3..x], {18,34,48,56}) + dp4a(src[x+1..x+4], {48,34,18,0})  (exact, <= 65280).
    // One item = 4 outputs of two consecutive rows, stored as u16x2 pairs (row 2p | row 2p+1 << 16) for the dp2a vertical pass.
    // The first tap of output 4*qx sits at tile column 13 + 4*qx = byte 1 of the aligned word at column 12 + 4*qx.
    {
        const uint32_t K0123 = a.kc[0], K456 = a.kc[1];
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(tile + (XO - 1)) + qx + seg * (2 * SP / 4);
        uint32_t *hq = hp + seg * kBlurTileW + 4 * qx;
#pragma unroll
        for (int i = 0; i < (NRP + 3) / 4; ++i) {                   // row pairs seg, seg + 4, ...
            if (4 * i + 3 < NRP || seg + 4 * i < NRP) {
                uint32_t hh[2][4];
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const uint32_t wa = wp[(8 * i + rr) * (SP / 4)], wb = wp[(8 * i + rr) * (SP / 4) + 1], wc = wp[(8 * i + rr) * (SP / 4) + 2];
                    hh[rr][0] = __dp4a(__funnelshift_r(wa, wb, 8), K0123, __dp4a(__funnelshift_r(wb, wc, 8), K456, 0u));
                    hh[rr][1] = __dp4a(__funnelshift_r(wa, wb, 16), K0123, __dp4a(__funnelshift_r(wb, wc, 16), K456, 0u));
                    hh[rr][2] = __dp4a(__funnelshift_r(wa, wb, 24), K0123, __dp4a(__funnelshift_r(wb, wc, 24), K456, 0u));
                    hh[rr][3] = __dp4a(wb, K0123, __dp4a(wc, K456, 0u));
                }
                *reinterpret_cast<uint4 *>(hq + 4 * i * kBlurTileW) =
                    make_uint4(__byte_perm(hh[0][0], hh[1][0], 0x5410), __byte_perm(hh[0][1], hh[1][1], 0x5410),
                               __byte_perm(hh[0][2], hh[1][2], 0x5410), __byte_perm(hh[0][3], hh[1][3], 0x5410));
            }
        }
    }
    __syncthreads();
    // vertical pass: dst = (sum_j k[j] * H[y + j] + 32768) >> 16 with dp2a on the row pairs; a thread owns 4 columns x 8 rows
    const int gx = x0 + 4 * qx;
    if (gx >= w) return;
    uint4 pr[7];                                        // row pairs 4*seg .. 4*seg+6  (rows 8*seg .. 8*seg+13)
#pragma unroll
    for (int p = 0; p < 7; ++p) pr[p] = *reinterpret_cast<const uint4 *>(hp + (4 * seg + p) * kBlurTileW + 4 * qx);
    const int gy0 = y0 + seg * 8;
    uint8_t *dst = a.blur + G.img_off + (size_t) frame * G.frame_stride + (size_t) gy0 * G.pitch + gx;
    const int rows = min(8, h - gy0);
    const size_t pitch = (size_t) G.pitch;
    auto out_row = [&](int r) -> uint32_t {             // r is a compile-time constant after unrolling
        const int p0 = r >> 1;
        const uint32_t c0 = a.kc[(r & 1) ? 6 : 2], c1 = a.kc[(r & 1) ? 7 : 3], c2 = a.kc[(r & 1) ? 8 : 4], c3 = a.kc[(r & 1) ? 9 : 5];
        uint32_t sa[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t a0 = c == 0 ? pr[p0].x : c == 1 ? pr[p0].y : c == 2 ? pr[p0].z : pr[p0].w;
            const uint32_t a1 = c == 0 ? pr[p0 + 1].x : c == 1 ? pr[p0 + 1].y : c == 2 ? pr[p0 + 1].z : pr[p0 + 1].w;
            const uint32_t a2 = c == 0 ? pr[p0 + 2].x : c == 1 ? pr[p0 + 2].y : c == 2 ? pr[p0 + 2].z : pr[p0 + 2].w;
            const uint32_t a3 = c == 0 ? pr[p0 + 3].x : c == 1 ? pr[p0 + 3].y : c == 2 ? pr[p0 + 3].z : pr[p0 + 3].w;
            sa[c] = __dp2a_lo(a3, c3, __dp2a_lo(a2, c2, __dp2a_lo(a1, c1, __dp2a_lo(a0, c0, 32768u))));
        }
        // result of column c = byte 2 of sa[c] (<= 255: the kernel sums to 256)
        return __byte_perm(__byte_perm(sa[0], sa[1], 0x0062), __byte_perm(sa[2], sa[3], 0x0062), 0x5410);
    };
    if (rows == 8) {
#pragma unroll
        for (int r = 0; r < 8; ++r) *reinterpret_cast<uint32_t *>(dst + r * pitch) = out_row(r);
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) if (r < rows) *reinterpret_cast<uint32_t *>(dst + r * pitch) = out_row(r);
    }
}

// ------------------------------------------------------------------------------------------------
// K5 + K7  orientation (IC_Angle + fastAtan2) and rotated BRIEF.  One warp per key point, in output order.
// The two patches a key point needs — 31 rows of the un-blurred level (radius-15 disc) and 37 rows of the blurred level (the
// rotated pattern reaches 18 px) — are staged into the warp's own shared-memory buffers by two TMA box loads (48 x 31 and
// 64 x 37 bytes, x origin rounded down to 16), so every pixel access afterwards is an LDS with a compile-time row stride
// instead of a 64-bit global address computation, and the blurred patch arrives while the angle is being computed.
// ------------------------------------------------------------------------------------------------
constexpr int kPatchW = 48, kPatchH = 2 * kHalfPatch + 1;            // un-blurred patch box (x-15 .. x+15 after alignment)
constexpr int kBPatchR = 18, kBPatchW = 64, kBPatchH = 2 * kBPatchR + 1;   // blurred patch box
constexpr int kPatchBytes = 1536, kBPatchBytes = 2432;               // box sizes rounded up to 128 (TMA destination alignment)
static_assert(kPatchW * kPatchH <= kPatchBytes && kBPatchW * kBPatchH <= kBPatchBytes, "patch buffers");

struct PatchMaps { CUtensorMap img[ORBFE_MAX_LEVELS], blur[ORBFE_MAX_LEVELS]; };

struct DescArgs {
    const uint8_t *blur; const uint32_t *kp; const int *nkp; int kp_per_frame;
    orbfe_keypoint *out_kps; uint8_t *out_desc; int *out_n; int cap; int *err;
    const float4 *pattern;      // pair p = 8 * lane + j at [j * 32 + lane]: (x0, y0, x1, y1) as floats
    uint8_t u_max[kHalfPatch + 1];
};

// cv::fastAtan2: float32, no FMA contraction.
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float scale = (float) (180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float eps = (float) 2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a;
    if (ax >= ay) {
        const float c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        const float c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        const float c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        const float c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// u_max of the reference (ORBExtractor.cpp:458-474) for HALF_PATCH_SIZE = 15: a compile-time table; the host recomputes it with
// the reference's float32 formula at handle creation and refuses to run if the two ever disagree (DescArgs::u_max is that copy).
__device__ __forceinline__ constexpr int umax15(int v) {
    constexpr int t[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    return t[v < 0 ? -v : v];
}

// (float) cos((double) a), (float) sin((double) a) for a float angle in [0, 2 pi]: the reference evaluates cos/sin in double
// (ORBExtractor.cpp:53-54) and rounds to float.  Quadrant reduction with a two-part pi/2 and the fdlibm kernel polynomials
// (< 1 ulp in double), so the float results agree with glibc's unless the exact value lies within ~1e-16 of a float rounding
// boundary; the parity tests compare every descriptor bit.
__device__ __forceinline__ void sincos_to_float(float a, float &c, float &s) {
    const double x = (double) a;
    const double n = rint(x * 0.63661977236758134308);                       // 2 / pi
    double r = fma(-n, 1.57079632679489655800e+00, x);                        // pi/2 head (exact product for n <= 4)
    r = fma(-n, 6.12323399573676603587e-17, r);                               // pi/2 tail
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06); ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03); ps = fma(z, ps, -1.66666666666666324348e-01);
    const double sr = fma(z * r, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07); pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03); pc = fma(z, pc, 4.16666666666666019037e-02);
    const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
    const int q = (int) n & 3;
    const double sv = (q & 1) ? cr : sr, cv = (q & 1) ? sr : cr;
    s = (float) ((q & 2) ? -sv : sv);
    c = (float) (((q + 1) & 2) ? -cv : cv);
}

// Stage a (box_w x box_h) patch whose top-left pixel is (ax, ay) of frame `frame` into `dst` (pitch box_w); one warp.
template <bool kTMA>
__device__ __forceinline__ void stage_patch(uint8_t *dst, uint64_t *bar, const CUtensorMap *map, const uint8_t *frame_base, int pitch,
                                            int ax, int ay, int frame, int box_w, int box_h, int lane) {
    if constexpr (kTMA) {
        if (lane == 0) {
            mbar_expect_tx(bar, (uint32_t) (box_w * box_h));
            tma_load_3d(dst, map, bar, ax, ay, frame);
        }
    } else {
        const int wpr = box_w / 4;
        for (int idx = lane; idx < wpr * box_h; idx += 32) {
            const int r = idx / wpr, c = idx - r * wpr;
            const int gx = ax + 4 * c;
            uint32_t v = 0;
            if (gx + 4 <= pitch) v = __ldg(reinterpret_cast<const uint32_t *>(frame_base + (size_t) (ay + r) * pitch + gx));
            reinterpret_cast<uint32_t *>(dst)[idx] = v;
        }
    }
}

// grid = (ceil(kp_per_frame / 8), frames): warp = output slot (levels in order, quadtree order within a level)
template <bool kTMA>
__global__ void __launch_bounds__(256, 6) k_describe(const __grid_constant__ LevelSet L, const __grid_constant__ PatchMaps P, const DescArgs a) {
    __shared__ __align__(128) uint8_t sbuf[8][kPatchBytes + kBPatchBytes];
    __shared__ __align__(8) uint64_t bars[8][2];
    const int frame = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int out_idx = blockIdx.x * 8 + wid;
    // level of this output slot: the block's warps share the frame, so warp 0 turns the per-level counts into prefix sums once
    // (lane q = level q) and every warp finds its level with one compare + ballot
    static_assert(ORBFE_MAX_LEVELS <= 32, "one lane per level");
    __shared__ int s_pre[ORBFE_MAX_LEVELS + 1];
    if (wid == 0) {
        const int c = lane < L.n_levels ? __ldg(a.nkp + frame * ORBFE_MAX_LEVELS + lane) : 0;
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane < ORBFE_MAX_LEVELS) s_pre[lane] = inc - c;                                      // exclusive prefix: first slot of level `lane`
        if (lane == 31) s_pre[ORBFE_MAX_LEVELS] = inc;                                           // total
    }
    __syncthreads();
    const int tot = s_pre[ORBFE_MAX_LEVELS];
    const unsigned below = __ballot_sync(0xffffffffu, lane > 0 && lane < L.n_levels && out_idx >= s_pre[min(lane, ORBFE_MAX_LEVELS - 1)]);
    const int l = __popc(below);                                       // levels whose first slot is <= out_idx (empty levels included), minus level 0
    const int k = out_idx - s_pre[l];
    if (out_idx == 0 && lane == 0) a.out_n[frame] = tot;            // the first key point of the frame also publishes the total
    if (out_idx >= tot) return;
    if (out_idx >= a.cap) { if (lane == 0) atomicExch(a.err, 6); return; }
    const LevelGeom &G = L.lv[l];
    const uint32_t v = __ldg(&a.kp[(size_t) frame * a.kp_per_frame + G.kp_off + k]);
    const int x = (int) (v & 0xfffu), y = (int) ((v >> 12) & 0xfffu), score = (int) (v >> 24);

    uint8_t *ibuf = sbuf[wid], *bbuf = sbuf[wid] + kPatchBytes;
    const int ax = (x - kHalfPatch) & ~15, axb = (x - kBPatchR) & ~15;
    if (kTMA) {
        if (lane == 0) { mbar_init(&bars[wid][0], 1); mbar_init(&bars[wid][1], 1); fence_barrier_init(); }
        __syncwarp();
    }
    stage_patch<kTMA>(ibuf, &bars[wid][0], &P.img[l], L.img[l] + (size_t) frame * G.frame_stride, G.pitch, ax, y - kHalfPatch, frame, kPatchW, kPatchH, lane);
    stage_patch<kTMA>(bbuf, &bars[wid][1], &P.blur[l], a.blur + G.img_off + (size_t) frame * G.frame_stride, G.pitch, axb, y - kBPatchR, frame, kBPatchW, kBPatchH, lane);
    const float4 *pt = a.pattern + lane;
    if (kTMA) mbar_wait(&bars[wid][0], 0); else __syncwarp();

    // IC_Angle (ORBExtractor.cpp:18-42) on the un-blurred patch: lane = u + 15, rows +v and -v share u_max[v]
    const int u = lane - kHalfPatch, au = u < 0 ? -u : u;
    const uint8_t *c = ibuf + kHalfPatch * kPatchW + (x - ax) + u;           // lane 31 (u = 16) stays inside the 48-byte row and is masked
    int sum = au <= kHalfPatch ? (int) c[0] : 0, m01 = 0;
#pragma unroll
    for (int vv = 1; vv <= kHalfPatch; ++vv) {
        const int p = c[vv * kPatchW], m = c[-vv * kPatchW];
        if (au <= umax15(vv)) { sum += p + m; m01 += vv * (p - m); }
    }
    int m10 = u * sum;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float) m01, (float) m10);

    // computeOrbDescriptor (ORBExtractor.cpp:50-97) on the blurred patch
    const float factor_pi = (float) (3.1415926535897932384626433832795 / 180.f);
    float ca, sb;
    sincos_to_float(__fmul_rn(angle, factor_pi), ca, sb);
    if (kTMA) mbar_wait(&bars[wid][1], 0); else __syncwarp();
    const uint8_t *ctr = bbuf + kBPatchR * kBPatchW + (x - axb);
    unsigned desc_byte = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 q = __ldg(pt + j * 32);                                 // x0, y0, x1, y1 of pair 8 * lane + j
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(q.x, sb), __fmul_rn(q.y, ca)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(q.x, ca), __fmul_rn(q.y, sb)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(q.z, sb), __fmul_rn(q.w, ca)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(q.z, ca), __fmul_rn(q.w, sb)));
        const int t0 = ctr[r0 * kBPatchW + c0], t1 = ctr[r1 * kBPatchW + c1];
        desc_byte |= (unsigned) (t0 < t1) << j;
    }
    a.out_desc[((size_t) frame * a.cap + out_idx) * 32 + lane] = (uint8_t) desc_byte;
    if (lane < 7) {                                                           // cv::KeyPoint, one 32-bit field per lane (ORBExtractor.cpp:537-542)
        const float fx = l ? __fmul_rn((float) x, G.scale) : (float) x, fy = l ? __fmul_rn((float) y, G.scale) : (float) y;
        uint32_t w;
        switch (lane) {
            case 0: w = __float_as_uint(fx); break;
            case 1: w = __float_as_uint(fy); break;
            case 2: w = __float_as_uint(G.scale); break;
            case 3: w = __float_as_uint(angle); break;
            case 4: w = __float_as_uint((float) score); break;
            case 5: w = (uint32_t) l; break;
            default: w = 0xffffffffu; break;                                  // class_id = -1
        }
        reinterpret_cast<uint32_t *>(a.out_kps + (size_t) frame * a.cap + out_idx)[lane] = w;
    }
}

// frames with zero key points never reach k_describe's publisher: clear the counts first
__global__ void k_zero_counts(int *out_n, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out_n[i] = 0;
}

#endif  // ORBFE_HELPERS_ONLY

}  // namespace orbfe
