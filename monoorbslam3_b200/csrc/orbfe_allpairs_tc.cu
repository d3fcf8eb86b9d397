// orbfe_allpairs_tc.cu — brute-force best / second-best Hamming search on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// ORBMatcher::DescriptorDistance (ORBMatcher.cpp:17-31) over all pairs of two descriptor tables (BASELINE configs 4/5: key-frame
// window matching).  hamming(a, b) = (256 - <a', b'>) / 2 with a', b' the descriptors as +-1 int8 vectors, so the nq x nt distance
// matrix is an int8 GEMM with K = 256, exact in s32, followed by a (min, second-min, argmin) reduction along each row.
//
//   CTA            = 256 query rows (two M = 128 accumulator tiles) x a contiguous range of train tiles (N = 128 rows each)
//   warp 0         TMA producer: the CTA's query tile once (64 KB, resident), then train tiles through a 3-stage ring
//                  (cp.async.bulk.tensor.2d, 128-byte swizzle, two K halves of 128 bytes per tile)
//   warp 1         TMEM allocation (512 columns = 2 accumulator stages x 2 M tiles x 128 columns) and the MMA issue:
//                  one elected thread, 16 x tcgen05.mma.cta_group::1.kind::i8 (M 128, N 128, K 32) per train tile,
//                  tcgen05.commit releases the shared-memory stage and publishes the accumulator stage
//   warps 2..17    epilogue, 16 warps = 4 per scheduler: a thread owns one query row (= one TMEM lane) of one M tile for the even or
//                  for the odd train tiles (so one half of the warps computes while the other half waits for / loads its stage),
//                  reads the row's 128 accumulators of the stage with tcgen05.ld.x32,
//                  turns two neighbouring columns into one u16x2 word of 16-bit keys (distance * 128 + column within the tile:
//                  one multiply-add per column, on the FMA pipe) and keeps the two smallest keys of both halves with packed
//                  min / max (5 ALU instructions per 4 columns, two independent chains); per tile the best two of the four 16-bit
//                  candidates are merged into the row's 32-bit (distance << 22 | train index) pair, if they can improve it
// The first minimum wins ties, as in the sequential `if (dist < bestDist)` loops: keys order by (distance, index).
// Train ranges are split over blockIdx.y to fill the machine; k_allpairs_merge (orbfe_match.cu) merges the partial pairs.
// Optional per-row exclusion range [lo, hi) of train indices: a query never matches its own key frame's block.
#define ORBFE_HELPERS_ONLY
#include "orbfe_kernels.cuh"

#include <cudaTypedefs.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#ifndef ORBFE_TC_STAGES
#define ORBFE_TC_STAGES 5
#endif

namespace orbfe {

constexpr int kTcM = 128;                 // rows of one accumulator tile (TMEM lanes)
constexpr int kTcRows = 2 * kTcM;         // query rows per CTA
constexpr int kTcN = 128;                 // train rows per tile (accumulator columns)
constexpr int kTcStages = ORBFE_TC_STAGES;   // train tiles in flight in shared memory (5 x 32 KB + the 64 KB query tile = 224 KB)
constexpr int kTcHalf = 128 * 128;        // bytes of one (128 rows x 128 K-bytes) swizzled operand block
constexpr int kTcEpiWarps = 16;           // 2 M tiles x 4 lane quarters x 2 tile parities
constexpr int kTcThreads = 32 * (2 + kTcEpiWarps);
constexpr uint32_t kTcNone = (257u << 22) | 0x3fffffu;       // == kApNone of orbfe_match.cu
constexpr size_t kTcSmem = (size_t) (4 + 2 * kTcStages) * kTcHalf + 1024 /* alignment slack */ + 256 /* barriers */;

// ---- PTX wrappers -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {      // arrives on `bar` when every tcgen05.mma issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> s32, M 128, N 128, K 32
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major operand block stored the way TMA writes a (128-byte x rows) box with 128-byte swizzle:
// rows 128 bytes apart, groups of 8 rows 1024 bytes apart (stride byte offset), 16-byte chunks XOR-swizzled with the row number.
//   bits [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1), [32,46) stride byte offset >> 4,
//   [46,48) version = 1 (sm_100), [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    return (uint64_t) ((saddr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t) (1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor: [4,6) D format 2 = s32, [7,10) A format 1 = int8, [10,13) B format 1 = int8, [15] / [16] A / B major 0 = K,
// [17,23) N >> 3, [24,29) M >> 4
constexpr uint32_t kTcIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (kTcN >> 3) << 17) | ((uint32_t) (kTcM >> 4) << 24);

// +-1 expansion of the descriptor bits: row r, byte k = bit k of descriptor r (bit j of byte i is pair 8i + j) -> +1 / -1, 256 bytes per row
__global__ void k_expand_pm1_rows(const uint8_t *__restrict__ bits, int n, uint8_t *__restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * 32) return;
    const uint32_t b = bits[idx];
    uint2 v;
    v.x = ((b & 1u) ? 0x01u : 0xffu) | ((b & 2u) ? 0x0100u : 0xff00u) | ((b & 4u) ? 0x010000u : 0xff0000u) | ((b & 8u) ? 0x01000000u : 0xff000000u);
    v.y = ((b & 16u) ? 0x01u : 0xffu) | ((b & 32u) ? 0x0100u : 0xff00u) | ((b & 64u) ? 0x010000u : 0xff0000u) | ((b & 128u) ? 0x01000000u : 0xff000000u);
    reinterpret_cast<uint2 *>(out)[idx] = v;
}

// two smallest of {k1 <= k2} and {a, b}, per 16-bit half
__device__ __forceinline__ void two_smallest_u16x2(uint32_t &k1, uint32_t &k2, uint32_t a, uint32_t b) {
    const uint32_t lo = __vminu2(a, b), hi = __vmaxu2(a, b);
    const uint32_t x = __vmaxu2(k1, lo);
    k1 = __vminu2(k1, lo);
    k2 = __vimin3_u16x2(x, k2, hi);
}

// merge one 32-bit key into the (k1 <= k2) pair
__device__ __forceinline__ void merge_key(uint32_t &k1, uint32_t &k2, uint32_t key) {
    k2 = min(k2, max(key, k1));
    k1 = min(k1, key);
}

struct TcArgs {
    int nq, nt, tiles_per_split;
    const int2 *excl;        // per query row: train indices [x, y) are skipped; NULL = none
    uint2 *partial;          // [split][nq] (best key, second key)
    // slab mode (key-frame window straight from the extractor's output slabs): the table is n_frames blocks of `cap` rows of which
    // the first slab_n[f] are key points; the padding rows never match and are never matched, and a row skips its own block
    const int *slab_n; int cap;
};

__global__ void __launch_bounds__(kTcThreads, 1) k_allpairs_tc(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_t, const TcArgs a) {
    extern __shared__ uint8_t tc_dyn[];
    uint8_t *base = reinterpret_cast<uint8_t *>(((uintptr_t) tc_dyn + 1023) & ~(uintptr_t) 1023);      // the swizzle pattern is a function of the address
    uint8_t *s_q = base;                                   // [m tile][k half][128 x 128]
    uint8_t *s_t = base + 4 * kTcHalf;                     // [stage][k half][128 x 128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + (4 + 2 * kTcStages) * kTcHalf);
    uint64_t *b_full = bars, *b_empty = bars + kTcStages, *b_tfull = bars + 2 * kTcStages, *b_tempty = b_tfull + 2, *b_q = b_tempty + 2;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(b_q + 1);

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * kTcRows;
    const int n_tiles = (a.nt + kTcN - 1) / kTcN;
    const int t_begin = blockIdx.y * a.tiles_per_split, t_end = min(n_tiles, t_begin + a.tiles_per_split);
    const int my_tiles = max(t_end - t_begin, 0);

    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&b_tfull[s], 1); mbar_init(&b_tempty[s], kTcEpiWarps / 2); }
        mbar_init(b_q, 1);
        fence_barrier_init();
    }
    if (wid == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;

    if (wid == 0) {
        // ===== TMA producer
        if (lane == 0 && my_tiles > 0) {
            mbar_expect_tx(b_q, 4 * kTcHalf);
            for (int mt = 0; mt < 2; ++mt)
                for (int kh = 0; kh < 2; ++kh) tma_load_2d(s_q + (mt * 2 + kh) * kTcHalf, &tm_q, b_q, kh * 128, row0 + mt * kTcM);
            for (int i = 0; i < my_tiles; ++i) {
                const int s = i % kTcStages;
                mbar_wait(&b_empty[s], (uint32_t) (((i / kTcStages) & 1) ^ 1));          // passes at once on a fresh barrier
                mbar_expect_tx(&b_full[s], 2 * kTcHalf);
                for (int kh = 0; kh < 2; ++kh) tma_load_2d(s_t + (s * 2 + kh) * kTcHalf, &tm_t, &b_full[s], kh * 128, (t_begin + i) * kTcN);
            }
        }
    } else if (wid == 1) {
        // ===== MMA issue.  The descriptors of a tile differ only in the start-address field: the low words are precomputed and a k step
        // is one add (32 bytes = 2 in the >> 4 encoding; the second K half starts kTcHalf bytes further on)
        if (lane == 0 && my_tiles > 0) {
            const uint64_t desc_hi = (uint64_t) (1024 >> 4) << 32 | (1ull << 46) | (2ull << 61);
            auto desc_lo = [](uint32_t saddr) -> uint32_t { return ((saddr & 0x3ffffu) >> 4) | (1u << 16); };
            const uint32_t a_lo0 = desc_lo(smem_u32(s_q)), a_lo1 = desc_lo(smem_u32(s_q + 2 * kTcHalf)), b_lo0 = desc_lo(smem_u32(s_t));
            mbar_wait(b_q, 0);
            for (int i = 0; i < my_tiles; ++i) {
                const int s = i % kTcStages, ac = i & 1;
                mbar_wait(&b_tempty[ac], (uint32_t) (((i >> 1) & 1) ^ 1));
                mbar_wait(&b_full[s], (uint32_t) ((i / kTcStages) & 1));
                tc_fence_after();
                const uint32_t b_lo = b_lo0 + (uint32_t) (s * 2 * kTcHalf >> 4);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const uint32_t d = tmem + (uint32_t) (ac * 2 * kTcN + mt * kTcN);
                    const uint32_t a_lo = mt ? a_lo1 : a_lo0;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t ko = (uint32_t) (((ks >> 2) * kTcHalf + (ks & 3) * 32) >> 4);
                        tc_mma_i8(d, desc_hi | (a_lo + ko), desc_hi | (b_lo + ko), kTcIdesc, ks ? 1u : 0u);
                    }
                }
                tc_commit(&b_empty[s]);          // the stage's train tile may be overwritten once these MMAs have read it
                tc_commit(&b_tfull[ac]);         // ... and the accumulator stage is complete
            }
        }
    } else {
        // ===== epilogue: warps 2..17; TMEM lane quarter = wid % 4 (hardware rule); the four warps of a quarter take (M tile, tile parity):
        // the warps of the even tiles compute while those of the odd tiles wait for / load their accumulators, and vice versa
        const int lq = wid & 3, grp = (wid - 2) >> 2, mt = grp & 1, par = grp >> 1;
        const int row = row0 + mt * kTcM + lq * 32 + lane;
        int2 ex = make_int2(0, 0);
        if (a.excl && row < a.nq) ex = a.excl[row];
        bool row_live = row < a.nq;
        if (a.slab_n && row < a.nq) { const int f = row / a.cap; ex = make_int2(f * a.cap, (f + 1) * a.cap); row_live = row - f * a.cap < a.slab_n[f]; }
        // warp-wide hull of the exclusion ranges: tiles outside it take the fast path
        int w_lo = ex.x < ex.y ? ex.x : 0x7fffffff, w_hi = ex.x < ex.y ? ex.y : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { w_lo = min(w_lo, __shfl_xor_sync(0xffffffffu, w_lo, o)); w_hi = max(w_hi, __shfl_xor_sync(0xffffffffu, w_hi, o)); }
        // ... and their warp-wide core: a tile inside it is excluded for every row of the warp (a key-frame window matched against itself
        // spends cap / 128 train tiles per row inside the row's own key frame).  A warp whose rows are all padding has nothing to do at all.
        int c_lo = ex.x < ex.y ? ex.x : 0x7fffffff, c_hi = ex.x < ex.y ? ex.y : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { c_lo = max(c_lo, __shfl_xor_sync(0xffffffffu, c_lo, o)); c_hi = min(c_hi, __shfl_xor_sync(0xffffffffu, c_hi, o)); }
        const bool warp_live = __any_sync(0xffffffffu, row_live);
        uint32_t K1 = kTcNone, K2 = kTcNone;
        for (int i = par; i < my_tiles; i += 2) {
            const int col0 = (t_begin + i) * kTcN;
            mbar_wait(&b_tfull[par], (uint32_t) ((i >> 1) & 1));
            tc_fence_after();
            int f0 = 0, f_split = 0x7fffffff, n0 = 0x7fffffff, n1 = 0x7fffffff;              // slab mode: the (at most two) blocks the tile touches
            if (a.slab_n) {
                f0 = col0 / a.cap; f_split = (f0 + 1) * a.cap;
                n0 = f0 * a.cap + a.slab_n[f0];                                               // first padding row of block f0
                if (f_split < col0 + kTcN && f_split < a.nt) n1 = f_split + a.slab_n[f0 + 1];
            }
            if (!warp_live || (col0 >= c_lo && col0 + kTcN <= c_hi) || (col0 >= n0 && col0 + kTcN <= f_split)) {
                tc_fence_before();                                                           // nothing to read (also: a tile of padding rows
                __syncwarp();                                                                // only, when the slab capacity is a multiple of
                if (lane == 0) mbar_arrive(&b_tempty[par]);                                  // the tile height): hand the stage straight back
                continue;
            }
            const uint32_t taddr = tmem + ((uint32_t) (lq * 32) << 16) + (uint32_t) (par * 2 * kTcN + mt * kTcN);
            bool slow = col0 + kTcN > a.nt || (col0 < w_hi && col0 + kTcN > w_lo);            // ragged last tile / a row's own key-frame block
            if (a.slab_n) slow = slow || col0 + kTcN > n0;                                    // reaches block f0's padding (or the next block)
            // slow tiles: the dead columns of this row as a 128-bit mask, built from the (at most four) intervals with shifts instead of
            // six compares per column: beyond the table, the row's own key frame, the padding of block f0, the padding of block f0 + 1
            uint32_t dm[4] = {0u, 0u, 0u, 0u};
            if (slow) {
                auto ge = [](int t, int w) -> uint32_t {                                     // bits b of word w with 32 w + b >= t
                    const int sft = min(max(t - 32 * w, 0), 32);
                    return (uint32_t) (0xffffffffull << sft);
                };
                const int r_nt = a.nt - col0, r_e0 = ex.x - col0, r_e1 = ex.y - col0, r_n0 = n0 - col0, r_sp = f_split - col0, r_n1 = n1 - col0;
#pragma unroll
                for (int w = 0; w < 4; ++w)
                    dm[w] = ge(r_nt, w) | (ge(r_e0, w) & ~ge(r_e1, w)) | (ge(r_n0, w) & ~ge(r_sp, w)) | (ge(r_sp, w) & ge(r_n1, w));
            }
            // two independent (smallest, second) chains, per 16-bit half
            uint32_t a1 = 0xffffffffu, a2 = 0xffffffffu, b1 = 0xffffffffu, b2 = 0xffffffffu;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                uint32_t va[32], vb[32];
                tc_ld32(taddr + (uint32_t) (ch * 64), va);
                tc_ld32(taddr + (uint32_t) (ch * 64 + 32), vb);
                tc_ld_wait();
                if (ch == 1) {      // every accumulator this warp needs from the stage is in registers: hand the stage back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&b_tempty[par]);
                }
                // key16 = distance * 128 + column = 16384 - 64 * dot + column, two columns per word; neither half borrows or carries
                uint32_t pa[16], pb[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t ca = (uint32_t) (16384 + ch * 64 + 2 * j) | ((uint32_t) (16384 + ch * 64 + 2 * j + 1) << 16);
                    pa[j] = va[2 * j + 1] * 0xFFC00000u + (va[2 * j] * 0xFFFFFFC0u + ca);
                    pb[j] = vb[2 * j + 1] * 0xFFC00000u + (vb[2 * j] * 0xFFFFFFC0u + (ca + 0x00200020u));
                }
                if (slow) {                      // pa[j] holds the columns 64 ch + 2 j, + 1 and pb[j] those 32 further on: one test + one predicated OR each
                    const uint32_t ma = dm[2 * ch], mb = dm[2 * ch + 1];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (ma & (1u << (2 * j))) pa[j] |= 0x0000ffffu;
                        if (ma & (2u << (2 * j))) pa[j] |= 0xffff0000u;
                        if (mb & (1u << (2 * j))) pb[j] |= 0x0000ffffu;
                        if (mb & (2u << (2 * j))) pb[j] |= 0xffff0000u;
                    }
                }
#pragma unroll
                for (int j = 0; j < 16; j += 2) { two_smallest_u16x2(a1, a2, pa[j], pa[j + 1]); two_smallest_u16x2(b1, b2, pb[j], pb[j + 1]); }
            }
            const uint32_t k1 = __vminu2(a1, b1), k2 = __vimin3_u16x2(__vmaxu2(a1, b1), a2, b2);
            // across the two halves: smallest = min(lo, hi); second = min3(max(lo, hi), k2.lo, k2.hi)
            const uint32_t k1s = __byte_perm(k1, 0, 0x1032), k2s = __byte_perm(k2, 0, 0x1032);
            const uint32_t m1 = __vminu2(k1, k1s) & 0xffffu, m2 = __vimin3_u16x2(__vmaxu2(k1, k1s), k2, k2s) & 0xffffu;
            // 16-bit candidates -> 32-bit keys (distance << 22 | train index); most tiles cannot improve the row's pair
            const uint32_t key1 = ((m1 >> 7) << 22) | (uint32_t) (col0 + (int) (m1 & 127u));
            if (key1 < K2) {
                const uint32_t key2 = ((m2 >> 7) << 22) | (uint32_t) (col0 + (int) (m2 & 127u));
                merge_key(K1, K2, key1);
                merge_key(K1, K2, key2);
            }
        }
        if (row < a.nq) a.partial[(size_t) (blockIdx.y * 2 + par) * a.nq + row] = row_live ? make_uint2(K1, K2) : make_uint2(kTcNone, kTcNone);
    }
    tc_fence_before();
    __syncthreads();
    if (wid == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled tc_encode_fn() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled) p;
    }
    return fn;
}

// 2-D u8 map over a table of `rows` x 256 bytes; box = 128 bytes x 128 rows, 128-byte swizzle, rows beyond the table read as zero
static int tc_make_map(Handle *h, CUtensorMap *m, const uint8_t *table, int rows) {
    PFN_cuTensorMapEncodeTiled enc = tc_encode_fn();
    if (!enc) return set_error(h, ORBFE_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {256, (cuuint64_t) std::max(rows, 1)};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {128, 128}, estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *) table, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(h, ORBFE_E_CUDA, "cuTensorMapEncodeTiled (all-pairs table, %d rows) failed with CUresult %d", rows, (int) r);
    return ORBFE_OK;
}

int allpairs_tc_device_setup(Handle *h) {
    ORBFE_CUDA(h, cudaFuncSetAttribute(k_allpairs_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kTcSmem));
    return ORBFE_OK;
}

size_t allpairs_tc_scratch_bytes(int nq, int nt, int *n_split_out) {
    const int row_tiles = (nq + kTcRows - 1) / kTcRows, n_tiles = (nt + kTcN - 1) / kTcN;
    // split the train range so that the CTAs fill whole waves of the machine; a CTA pays about 6 tile times of fill and drain
    int best = 1; double best_t = 1e30;
    for (int s = 1; s <= std::max(1, n_tiles / 4); ++s) {
        const int per = (n_tiles + s - 1) / s;
        const int eff = (n_tiles + per - 1) / per;
        const double t = (double) ((row_tiles * eff + 147) / 148) * (per + 6);
        if (t < best_t - 1e-9) { best_t = t; best = eff; }
    }
    if (n_split_out) *n_split_out = best;
    auto up = [](size_t v) { return (v + 255) & ~(size_t) 255; };
    return up((size_t) nq * 256) + up((size_t) nt * 256) + up((size_t) 2 * best * nq * sizeof(uint2)) + 1024;    // even / odd tiles of a split report separately
}

// d_scratch: allpairs_tc_scratch_bytes(nq, nt) bytes.  The results go through `partial` and the caller's merge kernel.
int allpairs_tc_launch(Handle *h, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt, const int2 *d_excl, uint8_t *d_scratch,
                       uint2 **partial_out, int *n_split_out, cudaStream_t st, const int *d_slab_n, int slab_cap) {
    int n_split = 1;
    allpairs_tc_scratch_bytes(nq, nt, &n_split);
    auto up = [](size_t v) { return (v + 255) & ~(size_t) 255; };
    uint8_t *qe = d_scratch, *te = qe + up((size_t) nq * 256);
    uint2 *partial = reinterpret_cast<uint2 *>(te + up((size_t) nt * 256));
    k_expand_pm1_rows<<<(nq * 32 + 255) / 256, 256, 0, st>>>(d_q, nq, qe);
    if (d_t == d_q && nt == nq) te = qe;
    else k_expand_pm1_rows<<<(nt * 32 + 255) / 256, 256, 0, st>>>(d_t, nt, te);
    CUtensorMap mq, mt;
    int rc;
    if ((rc = tc_make_map(h, &mq, qe, nq)) || (rc = tc_make_map(h, &mt, te, nt))) return rc;
    const int n_tiles = (nt + kTcN - 1) / kTcN;
    TcArgs ta;
    ta.nq = nq; ta.nt = nt; ta.tiles_per_split = (n_tiles + n_split - 1) / n_split; ta.excl = d_excl; ta.partial = partial;
    ta.slab_n = d_slab_n; ta.cap = slab_cap;
    k_allpairs_tc<<<dim3((nq + kTcRows - 1) / kTcRows, n_split), kTcThreads, kTcSmem, st>>>(mq, mt, ta);
    h->launches += (te == qe) ? 2 : 3;
    ORBFE_CUDA(h, cudaGetLastError());
    *partial_out = partial; *n_split_out = 2 * n_split;          // the even and the odd tiles of every split report separately
    return ORBFE_OK;
}

}  // namespace orbfe
