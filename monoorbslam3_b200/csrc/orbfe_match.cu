// orbfe_match.cu — Hamming matching path of the B200-native ORB front-end.
//
//   K8  k_pair_distance      ORBMatcher::DescriptorDistance (ORBMatcher.cpp:17-31) over explicit pairs
//   K11 k_hamming_allpairs   brute-force best / second-best (BASELINE configs 4/5): train tiles staged with TMA bulk copies
//                            (cp.async.bulk + mbarrier, double buffered), uint4-packed descriptors, __popc, keyed min / second-min
//   K9  k_window             Frame::getFeaturesInArea windows (Frame.cpp:97-127) + distances, one warp per query (count, reserve, fill)
//   K10 k_resolve<variant>   the order-preserving greedy resolves of SearchForInitialization (:33-116), SearchByProjection
//                            (:203-348), SearchByProjection/local points (:350-415), SearchForTriangulation (:417-522) and
//                            SearchByBow (:118-201): warp 0 walks the queries in reference order on shared-memory state while
//                            the other warps stage the next batch of candidate lists
//       k_fuse               search half of the fuse SearchByProjection(KeyFrame, mapPoints) (:524-571), independent queries
//       k_compute_descriptors  MapPoint::computeDescriptor (MapPoint.cpp:103-152), one warp per map point
#define ORBFE_HELPERS_ONLY
#include "orbfe_kernels.cuh"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>

namespace orbfe {

constexpr int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30, GRID_SIZE = 40;   // ORBMatcher.cpp:13-15, Frame.h:18

int ensure_match_scratch(Handle *h, size_t bytes) {
    if (h->match_bytes >= bytes) return ORBFE_OK;
    ORBFE_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->d_match); h->d_match = nullptr; h->match_bytes = 0;
    bytes = bytes + bytes / 4 + 4096;
    ORBFE_CUDA(h, cudaMalloc(&h->d_match, bytes));
    h->match_bytes = bytes;
    return ORBFE_OK;
}

// bump allocator over the handle's matcher scratch
struct Bump {
    uint8_t *base; size_t off = 0;
    template <class T> T *take(size_t n) { off = (off + 255) & ~(size_t) 255; T *p = base ? reinterpret_cast<T *>(base + off) : nullptr; off += n * sizeof(T); return p; }
};

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ------------------------------------------------------------------------------------------------
// popc micro-benchmark: the roofline denominator of the matching kernels (SURVEY.md §8d asks for a measured figure).
// Every thread runs 8 independent xor+popc chains; 8 popc = one 256-bit "match".
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_popc_peak(unsigned *out, unsigned seed, int iters) {
    unsigned x[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u + blockIdx.x; acc[i] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i] += __popc(x[i] ^ acc[i]); }
    }
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= acc[i];
    if (r == 0xdeadbeefu) out[0] = r;          // keeps the chains alive without a store in practice
}

// ------------------------------------------------------------------------------------------------
// K8
// ------------------------------------------------------------------------------------------------
__global__ void k_pair_distance(const uint4 *a, const uint4 *b, const int *ia, const int *ib, int n, int *dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 *pa = a + 2 * (size_t) ia[i], *pb = b + 2 * (size_t) ib[i];
    dist[i] = hamming256(__ldg(pa), __ldg(pa + 1), __ldg(pb), __ldg(pb + 1));
}

// ------------------------------------------------------------------------------------------------
// K11 all-pairs.  CTA = 256 threads = 8 warps; lane <-> 2 queries (64 queries per CTA), warp <-> 1/8 of each train tile.
// key = dist << 22 | train index  (first minimum wins, like the sequential `if (d < best)` loop).
// ------------------------------------------------------------------------------------------------
constexpr int kApTile = 512;          // train descriptors per stage (16 KB)
constexpr int kApQ = 64;              // queries per CTA
constexpr uint32_t kApNone = (257u << 22) | 0x3fffffu;

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <bool kExcl>
__global__ void __launch_bounds__(256) k_hamming_allpairs(const uint4 *__restrict__ q, int nq, const uint4 *__restrict__ t, int nt, const int2 *__restrict__ excl,
                                                           int *best_idx, int *best_dist, int *second_dist) {
    __shared__ __align__(128) uint4 tile[2][kApTile * 2];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t part[8][kApQ][2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int q0 = blockIdx.x * kApQ;
    uint4 qa[2][2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int qi = min(q0 + lane + 32 * r, nq - 1);
        qa[r][0] = __ldg(q + 2 * (size_t) qi); qa[r][1] = __ldg(q + 2 * (size_t) qi + 1);
    }
    uint32_t k1[2] = {kApNone, kApNone}, k2[2] = {kApNone, kApNone};
    int2 ex[2] = {make_int2(0, 0), make_int2(0, 0)};            // train indices [x, y) this query skips (its own key frame's block)
    if (kExcl) {
#pragma unroll
        for (int r = 0; r < 2; ++r) ex[r] = __ldg(excl + min(q0 + lane + 32 * r, nq - 1));
    }
    const int n_tiles = (nt + kApTile - 1) / kApTile;
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    __syncthreads();
    if (tid == 0 && n_tiles > 0) {
        const uint32_t bytes = (uint32_t) min(kApTile, nt) * 32u;
        mbar_expect_tx(&bar[0], bytes);
        bulk_g2s(tile[0], t, bytes, &bar[0]);
    }
    for (int it = 0; it < n_tiles; ++it) {
        const int buf = it & 1;
        if (tid == 0 && it + 1 < n_tiles) {                    // prefetch the next tile into the other buffer (freed by the barrier below)
            const int j0 = (it + 1) * kApTile;
            const uint32_t bytes = (uint32_t) min(kApTile, nt - j0) * 32u;
            mbar_expect_tx(&bar[buf ^ 1], bytes);
            bulk_g2s(tile[buf ^ 1], t + 2 * (size_t) j0, bytes, &bar[buf ^ 1]);
        }
        mbar_wait(&bar[buf], (uint32_t) ((it >> 1) & 1));
        const int jbase = it * kApTile;
        const int cnt = min(kApTile, nt - jbase);
        const int lo = wid * (kApTile / 8), hi = min(lo + kApTile / 8, cnt);
#pragma unroll 4
        for (int j = lo; j < hi; ++j) {
            const uint4 b0 = tile[buf][2 * j], b1 = tile[buf][2 * j + 1];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                uint32_t key = ((uint32_t) hamming256(qa[r][0], qa[r][1], b0, b1) << 22) | (uint32_t) (jbase + j);
                if (kExcl && jbase + j >= ex[r].x && jbase + j < ex[r].y) key = kApNone;
                k2[r] = min(k2[r], max(key, k1[r]));
                k1[r] = min(k1[r], key);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) { part[wid][lane + 32 * r][0] = k1[r]; part[wid][lane + 32 * r][1] = k2[r]; }
    __syncthreads();
    if (tid < kApQ && q0 + tid < nq) {
        uint32_t a1 = kApNone, a2 = kApNone;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const uint32_t b1 = part[w][tid][0], b2 = part[w][tid][1];
            const uint32_t n2 = min(max(a1, b1), min(a2, b2));
            a1 = min(a1, b1); a2 = n2;
        }
        const int d1 = (int) (a1 >> 22), d2 = (int) (a2 >> 22);
        best_idx[q0 + tid] = d1 >= 257 ? -1 : (int) (a1 & 0x3fffffu);
        best_dist[q0 + tid] = d1; second_dist[q0 + tid] = d2;
    }
}

// ------------------------------------------------------------------------------------------------
// K11b all-pairs on the tensor cores.  hamming(a, b) = (256 - <a', b'>) / 2 with a', b' the descriptors as +-1 int8 vectors, so the
// N x M distance matrix is an int8 GEMM with K = 256 (exact in s32) followed by the same (min, second-min) key reduction as above.
// Descriptors are expanded once per call to 288-byte rows (256 of +-1, 32 of padding that makes the 64-bit fragment loads
// bank-conflict free).  CTA = 4 warps x 32 query rows; the A fragments of a warp's rows stay in registers for the whole kernel,
// train rows stream through shared memory in tiles of 64 (cp.async.bulk, double-buffered), the train range is split over
// blockIdx.y to fill the machine and the partial (min, second-min) pairs are merged by k_allpairs_merge.
// A thread loads 8 contiguous bytes of a row per k-step for both operands: the k order inside a 32-byte step is permuted the same
// way on both sides, which a dot product does not see.
// Legacy warp-level IMMA (mma.sync.m16n8k32.s8): measured 1.13 POPS = 2 200 GMatch/s equivalent on B200 (tools/micro/mma_b1.cu).
// ------------------------------------------------------------------------------------------------
constexpr int kImStride = 288;
constexpr int kImN = 64;
constexpr int kImM = 128;

__global__ void k_expand_pm1(const uint8_t *__restrict__ bits, int n, uint8_t *__restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * 36) return;
    const int row = idx / 36, g = idx - row * 36;
    uint2 v = make_uint2(0u, 0u);
    if (g < 32) {
        const uint32_t b = bits[(size_t) row * 32 + g];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v.x |= (((b >> j) & 1u) ? 0x01u : 0xffu) << (8 * j);
            v.y |= (((b >> (j + 4)) & 1u) ? 0x01u : 0xffu) << (8 * j);
        }
    }
    *reinterpret_cast<uint2 *>(out + (size_t) row * kImStride + g * 8) = v;
}

__device__ __forceinline__ void imma16832(int (&c)[4], const uint32_t (&a)[4], const uint2 b) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

// (min, second-min) update of a warp's 2 x 8 accumulator tiles; kRagged masks the columns beyond the last train row (their
// key base, which carries the column, is compared with `limit` = nt + (256 << 21))
template <bool kRagged>
__device__ __forceinline__ void imma_epilogue(const int (&acc)[2][8][4], uint32_t kbase, uint32_t limit, uint32_t (&k1)[4], uint32_t (&k2)[4]) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n8 = 0; n8 < 8; ++n8)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t cb = kbase + (uint32_t) (n8 * 8 + (c & 1));
                uint32_t key = (uint32_t) acc[mt][n8][c] * 0xFFE00000u + cb;
                if (kRagged && cb >= limit) key = kApNone;
                const int r = mt * 2 + (c >> 1);
                k2[r] = min(k2[r], max(key, k1[r]));
                k1[r] = min(k1[r], key);
            }
}

__global__ void __launch_bounds__(128) k_allpairs_imma(const uint8_t *__restrict__ qe, int nq, const uint8_t *__restrict__ te, int nt, int tiles_per_split,
                                                        uint2 *__restrict__ partial) {
    __shared__ __align__(128) uint8_t tile[2][kImN * kImStride];
    __shared__ __align__(8) uint64_t bar[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, q = lane & 3;
    const int row0 = blockIdx.x * kImM + wid * 32;
    uint32_t a[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int r0 = min(row0 + mt * 16 + g, nq - 1), r1 = min(row0 + mt * 16 + g + 8, nq - 1);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const uint2 v0 = __ldg(reinterpret_cast<const uint2 *>(qe + (size_t) r0 * kImStride + ks * 32 + q * 8));
            const uint2 v1 = __ldg(reinterpret_cast<const uint2 *>(qe + (size_t) r1 * kImStride + ks * 32 + q * 8));
            a[mt][ks][0] = v0.x; a[mt][ks][1] = v1.x; a[mt][ks][2] = v0.y; a[mt][ks][3] = v1.y;
        }
    }
    uint32_t k1[4] = {kApNone, kApNone, kApNone, kApNone}, k2[4] = {kApNone, kApNone, kApNone, kApNone};
    const int n_tiles = (nt + kImN - 1) / kImN;
    const int t_begin = blockIdx.y * tiles_per_split, t_end = min(n_tiles, t_begin + tiles_per_split);
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    __syncthreads();
    if (tid == 0 && t_begin < t_end) {
        const uint32_t bytes = (uint32_t) min(kImN, nt - t_begin * kImN) * kImStride;
        mbar_expect_tx(&bar[0], bytes);
        bulk_g2s(tile[0], te + (size_t) t_begin * kImN * kImStride, bytes, &bar[0]);
    }
    for (int it = t_begin; it < t_end; ++it) {
        const int li = it - t_begin, buf = li & 1;
        if (tid == 0 && it + 1 < t_end) {
            const uint32_t bytes = (uint32_t) min(kImN, nt - (it + 1) * kImN) * kImStride;
            mbar_expect_tx(&bar[buf ^ 1], bytes);
            bulk_g2s(tile[buf ^ 1], te + (size_t) (it + 1) * kImN * kImStride, bytes, &bar[buf ^ 1]);
        }
        mbar_wait(&bar[buf], (uint32_t) ((li >> 1) & 1));
        int acc[2][8][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int n8 = 0; n8 < 8; ++n8)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[mt][n8][c] = 0;
        const uint8_t *tb = tile[buf] + g * kImStride + q * 8;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
            for (int n8 = 0; n8 < 8; ++n8) {
                const uint2 b = *reinterpret_cast<const uint2 *>(tb + n8 * 8 * kImStride + ks * 32);
                imma16832(acc[0][n8], a[0][ks], b);
                imma16832(acc[1][n8], a[1][ks], b);
            }
        }
        const int j0 = it * kImN;
        // key = distance << 22 | column with (256 - dot) = 2 * distance: one multiply-add per pair, dot * -2^21 + ((256 << 21) + column)
        const uint32_t kbase = (256u << 21) + (uint32_t) (j0 + q * 2);
        if (nt - j0 < kImN) imma_epilogue<true>(acc, kbase, (uint32_t) nt + (256u << 21), k1, k2);
        else imma_epilogue<false>(acc, kbase, 0u, k1, k2);
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int m = 1; m < 4; m <<= 1) {
            const uint32_t b1 = __shfl_xor_sync(0xffffffffu, k1[r], m), b2 = __shfl_xor_sync(0xffffffffu, k2[r], m);
            const uint32_t n2 = min(max(k1[r], b1), min(k2[r], b2));
            k1[r] = min(k1[r], b1); k2[r] = n2;
        }
        const int row = row0 + (r >> 1) * 16 + g + (r & 1) * 8;
        if (q == 0 && row < nq) partial[(size_t) blockIdx.y * nq + row] = make_uint2(k1[r], k2[r]);
    }
}

__global__ void k_allpairs_merge(const uint2 *__restrict__ partial, int n_split, int nq, int *best_idx, int *best_dist, int *second_dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    uint32_t a1 = kApNone, a2 = kApNone;
    for (int s = 0; s < n_split; ++s) {
        const uint2 p = __ldg(partial + (size_t) s * nq + i);
        const uint32_t n2 = min(max(a1, p.x), min(a2, p.y));
        a1 = min(a1, p.x); a2 = n2;
    }
    const int d1 = (int) (a1 >> 22);
    best_idx[i] = d1 >= 257 ? -1 : (int) (a1 & 0x3fffffu);
    best_dist[i] = d1; second_dist[i] = (int) (a2 >> 22);
}

// IMMA micro-benchmark: the roofline denominator of k_allpairs_imma.  Register operands only, four independent accumulator chains
// per warp; one m16n8k32 instruction = 16 x 8 x 32 int8 multiply-adds = 1/8 of 128 descriptor pairs.
__global__ void __launch_bounds__(256) k_imma_peak(int *out, int iters) {
    uint32_t a[4] = {threadIdx.x * 2654435761u, threadIdx.x * 40503u + 1u, threadIdx.x ^ 0x9e3779b9u, threadIdx.x + 77u};
    const uint2 b = make_uint2(threadIdx.x * 2246822519u, threadIdx.x * 3266489917u);
    int c[4][4] = {};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) imma16832(c[u], a, b);
    }
    int s = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) s += c[u][0] + c[u][1] + c[u][2] + c[u][3];
    if (s == 0x7fffffff) out[0] = s;
}

// ------------------------------------------------------------------------------------------------
// K10 caller-supplied candidate lists (CSR): 8 lanes per query, lane l takes positions l, l+8, ... of the query's list.
// key = dist << 22 | position in the list, so the first minimum of the sequential `if (d < best)` scan wins.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_hamming_window(const uint4 *__restrict__ q, int nq, const uint4 *__restrict__ t, const int *__restrict__ off,
                                                        const int *__restrict__ idx, int *best_idx, int *best_dist, int *second_dist) {
    const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, sub = threadIdx.x & 7;
    const bool live = qi < nq;
    uint32_t k1 = kApNone, k2 = kApNone;
    int beg = 0;
    if (live) {
        const uint4 a0 = __ldg(q + 2 * (size_t) qi), a1 = __ldg(q + 2 * (size_t) qi + 1);
        beg = __ldg(off + qi);
        const int n = __ldg(off + qi + 1) - beg;
        for (int j = sub; j < n; j += 8) {
            const int ti = __ldg(idx + beg + j);
            const uint32_t key = ((uint32_t) hamming256(a0, a1, __ldg(t + 2 * (size_t) ti), __ldg(t + 2 * (size_t) ti + 1)) << 22) | (uint32_t) j;
            k2 = min(k2, max(key, k1));
            k1 = min(k1, key);
        }
    }
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) {
        const uint32_t b1 = __shfl_xor_sync(0xffffffffu, k1, m), b2 = __shfl_xor_sync(0xffffffffu, k2, m);
        const uint32_t n2 = min(max(k1, b1), min(k2, b2));
        k1 = min(k1, b1); k2 = n2;
    }
    if (live && sub == 0) {
        const int d1 = (int) (k1 >> 22);
        best_idx[qi] = d1 >= 257 ? -1 : __ldg(idx + beg + (int) (k1 & 0x3fffffu));
        best_dist[qi] = d1; second_dist[qi] = (int) (k2 >> 22);
    }
}

// ------------------------------------------------------------------------------------------------
// K9 windows.  Grid of frame 2 as CSR (cell = cx*rows + cy, members in key-point index order) built by the host mirror of
// Frame::Frame (Frame.cpp:32-51).  One warp per query; candidates come out in (cx, cy, insertion) order.
// ------------------------------------------------------------------------------------------------
struct WinArgs {
    const float *qx, *qy, *qr; const int *qmin, *qmax; const uint8_t *qvalid; const uint4 *qdesc; int nq;
    const orbfe_keypoint *kps2; const uint4 *desc2; const int *cell_off, *cell_idx; int cols, rows;
    int *q_beg, *q_end;              // candidate list of query i = [q_beg[i], q_end[i]) in c_idx / c_dist
    int *cursor, *overflow; int cand_cap;
    int *c_idx; int *c_dist;
};

__device__ __forceinline__ int floor_div_cell(float v) { return (int) floorf(v) / GRID_SIZE; }   // cvFloor(v) / GRID_SIZE, C division

// Frame::getFeaturesInArea (Frame.cpp:97-127) + DescriptorDistance for every candidate, one warp per query.  The warp walks its
// window twice: first it counts the candidates, reserves that many slots with one atomicAdd on the global cursor (the lists of
// different queries may land in any order, each list itself is in the reference's (cx, cy, insertion) order), then it fills them.
// If the reservation does not fit, the overflow flag is raised and the host retries with the exact total (the cursor).
__global__ void __launch_bounds__(256) k_window(const WinArgs a) {
    const int qi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (qi >= a.nq) return;
    int beg = 0, total = 0;
    if (a.qvalid[qi]) {
        const float x = a.qx[qi], y = a.qy[qi], r = a.qr[qi];
        const int min_l = a.qmin[qi], max_l = a.qmax[qi];
        const int min_cx = max(0, floor_div_cell(__fsub_rn(x, r))), max_cx = min(a.cols - 1, floor_div_cell(__fadd_rn(x, r)));
        const int min_cy = max(0, floor_div_cell(__fsub_rn(y, r))), max_cy = min(a.rows - 1, floor_div_cell(__fadd_rn(y, r)));
        const bool check_level = min_l > 0 || max_l >= 0;
        const uint4 d0 = __ldg(a.qdesc + 2 * (size_t) qi), d1 = __ldg(a.qdesc + 2 * (size_t) qi + 1);
        if (min_cx <= max_cx && min_cy <= max_cy) {
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                int run = 0;
                for (int cx = min_cx; cx <= max_cx; ++cx) {
                    // cells (cx, min_cy..max_cy) are contiguous in the CSR: walk their members as one run
                    const int cb = a.cell_off[cx * a.rows + min_cy], ce = a.cell_off[cx * a.rows + max_cy + 1];
                    for (int k0 = cb; k0 < ce; k0 += 32) {
                        const int k = k0 + lane;
                        bool ok = false; int idx = -1;
                        if (k < ce) {
                            idx = a.cell_idx[k];
                            const orbfe_keypoint kp = a.kps2[idx];
                            ok = true;
                            if (check_level) { if (kp.octave < min_l) ok = false; if (max_l >= 0 && kp.octave > max_l) ok = false; }
                            if (!(fabsf(__fsub_rn(kp.x, x)) <= r && fabsf(__fsub_rn(kp.y, y)) <= r)) ok = false;
                        }
                        const unsigned bal = __ballot_sync(0xffffffffu, ok);
                        if (pass && ok) {
                            const int o = beg + run + __popc(bal & ((1u << lane) - 1u));
                            a.c_idx[o] = idx;
                            a.c_dist[o] = hamming256(d0, d1, __ldg(a.desc2 + 2 * (size_t) idx), __ldg(a.desc2 + 2 * (size_t) idx + 1));
                        }
                        run += __popc(bal);
                    }
                }
                if (pass == 0) {
                    total = run;
                    if (total == 0) break;
                    if (lane == 0) beg = atomicAdd(a.cursor, total);
                    beg = __shfl_sync(0xffffffffu, beg, 0);
                    if (beg + total > a.cand_cap) { if (lane == 0) atomicExch(a.overflow, 1); total = 0; break; }
                }
            }
        }
    }
    if (lane == 0) { a.q_beg[qi] = beg; a.q_end[qi] = beg + total; }
}

// Search half of the fuse SearchByProjection(KeyFrame, mapPoints) (ORBMatcher.cpp:524-571): the queries are independent (the
// map-point bookkeeping that makes the reference loop sequential stays on the host), so one warp per projected map point walks
// KeyFrame::getFeaturesInArea's window (strict "< r", KeyFrame.cpp:204), applies the chi-square gate and keeps the first minimum.
struct FuseArgs {
    const float *qx, *qy, *qr; const int *qlevel; const uint8_t *qvalid; const uint4 *qdesc; int nq;
    const orbfe_keypoint *kps1; const uint4 *desc1; const int *cell_off, *cell_idx; int cols, rows;
    float sigma2[ORBFE_MAX_LEVELS]; int n_levels;
    int *best_idx, *best_dist, *n_matches;
};

__global__ void __launch_bounds__(256) k_fuse(const FuseArgs a) {
    const int qi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (qi >= a.nq) return;
    uint32_t key = 0xffffffffu; int my_idx = -1;
    if (a.qvalid[qi]) {
        const float x = a.qx[qi], y = a.qy[qi], r = a.qr[qi];
        const int min_l = a.qlevel[qi] - 1, max_l = a.qlevel[qi];                         // :553-554
        const int min_cx = max(0, floor_div_cell(__fsub_rn(x, r))), max_cx = min(a.cols - 1, floor_div_cell(__fadd_rn(x, r)));
        const int min_cy = max(0, floor_div_cell(__fsub_rn(y, r))), max_cy = min(a.rows - 1, floor_div_cell(__fadd_rn(y, r)));
        const bool check_level = min_l > 0 || max_l >= 0;
        const uint4 d0 = __ldg(a.qdesc + 2 * (size_t) qi), d1 = __ldg(a.qdesc + 2 * (size_t) qi + 1);
        int run = 0;
        if (min_cx <= max_cx && min_cy <= max_cy)
            for (int cx = min_cx; cx <= max_cx; ++cx) {
                const int cb = a.cell_off[cx * a.rows + min_cy], ce = a.cell_off[cx * a.rows + max_cy + 1];
                for (int k0 = cb; k0 < ce; k0 += 32) {
                    const int k = k0 + lane;
                    if (k < ce) {
                        const int idx = a.cell_idx[k];
                        const orbfe_keypoint kp = a.kps1[idx];
                        bool ok = true;
                        if (check_level) { if (kp.octave < min_l) ok = false; if (max_l >= 0 && kp.octave > max_l) ok = false; }
                        if (!(fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r)) ok = false;
                        if (ok) {
                            const float ex = __fsub_rn(x, kp.x), ey = __fsub_rn(y, kp.y);
                            const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                            const float s2 = a.sigma2[min(max(kp.octave, 0), a.n_levels - 1)];
                            if ((double) e2 > __dmul_rn(5.991, (double) s2)) ok = false;                          // :564
                        }
                        if (ok) {
                            const int d = hamming256(d0, d1, __ldg(a.desc1 + 2 * (size_t) idx), __ldg(a.desc1 + 2 * (size_t) idx + 1));
                            const uint32_t kk = ((uint32_t) d << 22) | (uint32_t) (run + k - k0);                 // position in enumeration order
                            if (kk < key) { key = kk; my_idx = idx; }
                        }
                    }
                    run += min(32, ce - k0);
                }
            }
    }
    const uint32_t best = __reduce_min_sync(0xffffffffu, key);
    const unsigned who = __ballot_sync(0xffffffffu, key == best && key != 0xffffffffu);
    int idx = -1, dist = TH_LOW + 1;
    if (who && (int) (best >> 22) < TH_LOW + 1) { idx = __shfl_sync(0xffffffffu, my_idx, __ffs(who) - 1); dist = (int) (best >> 22); }   // :560, 568
    if (lane == 0) {
        a.best_idx[qi] = idx; a.best_dist[qi] = dist;
        if (idx >= 0) atomicAdd(a.n_matches, 1);
    }
}

// MapPoint::computeDescriptor (MapPoint.cpp:103-152) for a batch of map points, one warp per map point: row i of the distance
// matrix goes to shared memory, its median (sorted row [(N-1)/2]) is found by bisection on the value with ballot counts, and the
// first row with the smallest median wins.
constexpr int kCdMaxObs = 512;
__global__ void __launch_bounds__(256) k_compute_descriptors(const uint4 *desc, const int *off, int n_groups, int *best, int *err) {
    __shared__ int s_row[8][kCdMaxObs];
    const int g = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (g >= n_groups) return;
    const int o = off[g], n = off[g + 1] - o;
    if (n <= 0) { if (lane == 0) best[g] = -1; return; }
    if (n > kCdMaxObs) { if (lane == 0) { best[g] = -1; atomicExch(err, 7); } return; }
    const int kth = (n - 1) / 2;
    int best_median = 256, best_idx = 0;
    for (int i = 0; i < n; ++i) {
        const uint4 a0 = __ldg(desc + 2 * (size_t) (o + i)), a1 = __ldg(desc + 2 * (size_t) (o + i) + 1);
        for (int j = lane; j < n; j += 32)
            s_row[w][j] = i == j ? 0 : hamming256(a0, a1, __ldg(desc + 2 * (size_t) (o + j)), __ldg(desc + 2 * (size_t) (o + j) + 1));
        __syncwarp();
        int lo = 0, hi = 256;                                   // smallest v with #{d <= v} >= kth + 1
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            int c = 0;
            for (int j = lane; j < n; j += 32) c += s_row[w][j] <= mid;
            c = __reduce_add_sync(0xffffffffu, c);
            if (c >= kth + 1) hi = mid; else lo = mid + 1;
        }
        if (lo < best_median) { best_median = lo; best_idx = i; }
        __syncwarp();
    }
    if (lane == 0) best[g] = best_idx;
}

// result / state arrays of the resolves in one launch
__global__ void k_window_init(int *bin_of, int n_bin, int *m12, int nq, int *m21, int *mdist, int *assigned, int n2, int *nmatch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_bin) bin_of[i] = -1;
    if (i < nq) m12[i] = -1;
    if (i < n2) { m21[i] = -1; mdist[i] = INT_MAX; assigned[i] = -1; }
    if (i < 4) nmatch[i] = 0;
}

// distances of CSR candidate lists given explicitly (SearchForTriangulation): one warp per query
__global__ void __launch_bounds__(256) k_csr_distance(const uint4 *qdesc, const int *q_desc_idx, const int *q_off, int nq,
                                                       const int *c_idx, const uint4 *desc2, int *c_dist) {
    const int qi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (qi >= nq) return;
    const int di = q_desc_idx[qi];
    const uint4 d0 = __ldg(qdesc + 2 * (size_t) di), d1 = __ldg(qdesc + 2 * (size_t) di + 1);
    for (int k = q_off[qi] + lane; k < q_off[qi + 1]; k += 32) {
        const int idx = c_idx[k];
        c_dist[k] = hamming256(d0, d1, __ldg(desc2 + 2 * (size_t) idx), __ldg(desc2 + 2 * (size_t) idx + 1));
    }
}

// ------------------------------------------------------------------------------------------------
// K10 resolves: one warp, queries in reference order, lanes over the query's candidates.
// Candidate key = dist << 22 | position: the two smallest keys are exactly the sequential loop's (best, second).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_two_smallest(uint32_t &k1, uint32_t &k2) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t o1 = __shfl_xor_sync(0xffffffffu, k1, o), o2 = __shfl_xor_sync(0xffffffffu, k2, o);
        const uint32_t n2 = min(max(k1, o1), min(k2, o2));
        k1 = min(k1, o1); k2 = n2;
    }
}

__device__ __forceinline__ int cv_round_sat(float v) {       // cvRound via cvtss2si: out-of-range -> 0x80000000
    if (!(v < 2147483648.f) || v < -2147483648.f) return INT_MIN;
    return __float2int_rn(v);
}

__device__ __forceinline__ int rot_bin(float a1, float a2) {  // ORBMatcher.cpp:85-88
    float rot = __fsub_rn(a1, a2);
    if (rot < 0) rot = __fadd_rn(rot, 360.f);
    int bin = __float2int_rn(__fmul_rn(rot, 1.f / HISTO_LENGTH));
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

// ComputeThreeMaxima (ORBMatcher.cpp:594-622) on the bin sizes
__device__ void three_maxima(const int *cnt, int &ind1, int &ind2, int &ind3) {
    int max1 = 0, max2 = -1, max3 = -2;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < HISTO_LENGTH; ++i) {
        const int n = cnt[i];
        if (n > max1) { max3 = max2; max2 = max1; max1 = n; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (n > max2) { max3 = max2; max2 = n; ind3 = ind2; ind2 = i; }
        else if (n > max3) { max3 = n; ind3 = i; }
    }
    if (max2 < max1 / 10) { ind2 = -1; ind3 = -1; }
    else if (max3 < max1 / 10) ind3 = -1;
}

struct ResolveArgs {
    int nq, n2;
    const int *q_beg, *q_end, *c_idx, *c_dist;   // candidate list of query i = [q_beg[i], q_end[i])
    const uint8_t *qvalid;
    const float *q_angle;            // angle of query i (kps1 angle / last key point angle)
    const orbfe_keypoint *kps2;
    const uint8_t *occupied;
    const int *q_out_idx;            // triangulation: key-point index of query i in frame 1
    const uint8_t *has_mp2;
    int *matches12;                  // init / triangulation: per frame-1 key point
    int *matches21, *matched_dist;   // init
    int *assigned;                   // projection / local points: per frame-2 key point
    int *bin_of;                     // rotation bin per entry (-1 = none)
    float *prematched;               // init: n1 x 2
    float nn_ratio; int check_orientation;
    int n_state;                     // variant 3: key points of frame 2 (n2 carries n1 there)
    const int *n_cand;               // device: total candidate entries in c_idx / c_dist (parallel resolve stages them in shared memory if they fit)
    int smem_entries;                // capacity of that staging area
    int acc_far;                     // variant 0: distances from here on only matter by existing (see k_resolve_init_par)
    int acc_limit;                   // variant 0: acceptors a slot may collect per round before the parallel resolve gives up (0 = kInitAcc)
    int *dbg;                        // optional: undecided queries after rounds 2 and 8, number of rounds (parallel resolve)
    int *n_matches;
};

// variant 0: SearchForInitialization; 1: SearchByProjection (frame/keyframe -> frame); 2: local map points; 3: triangulation;
// 4: SearchByBow
//
// The greedy loops of the reference are sequential over the queries (a query's decision depends on matchedDistance / the slots
// taken by earlier queries), so one warp walks the queries in reference order.  What makes that walk fast is that nothing on
// its critical path touches global memory: the per-key-point state (matched distance, owner, occupancy, angle / octave) lives in
// shared memory, and the other seven warps of the CTA stage the candidate lists and per-query fields of the next batch of
// queries into a double-buffered shared-memory area while warp 0 resolves the current batch.
constexpr int kResQB = 128;          // queries per batch (at most)
constexpr int kResE = 6144;          // staged candidate entries per batch (a single longer list reads its tail from global memory)

template <int kVariant>
static size_t resolve_smem_bytes(int n2) {
    return sizeof(int) * (size_t) n2 * (kVariant == 0 ? 3 : 2) + 2 * (sizeof(uint32_t) * kResE + sizeof(int) * (4 * kResQB + 4)) + 256;
}

__device__ __forceinline__ void loaders_sync() { asm volatile("bar.sync 1, 224;" ::: "memory"); }     // warps 1..7

template <int kVariant>
__global__ void __launch_bounds__(256) k_resolve(const ResolveArgs a, const int *run_if) {
    if (run_if && *run_if == 0) return;                       // launched behind the parallel resolve: only runs when that one gave up
    extern __shared__ __align__(16) uint8_t rs_dyn[];
    __shared__ int hist[HISTO_LENGTH];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n2s = kVariant == 3 ? a.n_state : a.n2;                     // key points of frame 2 (variant 3 passes n1 through n2)
    int *s_a = reinterpret_cast<int *>(rs_dyn);                           // v0: matched distance; v1/v2/v3: slot owner / taken flag
    int *s_b = s_a + n2s;                                                 // v0: matches21;        v1/v3: angle bits;  v2: octave
    int *s_c = s_b + n2s;                                                 // v0: angle bits
    // two batch buffers: [entries E][off QB+1][global start QB][length QB][angle QB][first query, query count]
    constexpr int kBufInts = kResE + 4 * kResQB + 4;
    int *s_buf = s_a + (size_t) n2s * (kVariant == 0 ? 3 : 2);
    if (tid < HISTO_LENGTH) hist[tid] = 0;
    for (int j = tid; j < n2s; j += 256) {
        const float ang = a.kps2[j].angle;
        if (kVariant == 0) { s_a[j] = INT_MAX; s_b[j] = -1; s_c[j] = __float_as_int(ang); }
        else if (kVariant == 1 || kVariant == 4) { s_a[j] = a.occupied[j] ? -2 : -1; s_b[j] = __float_as_int(ang); }
        else if (kVariant == 2) { s_a[j] = a.occupied[j] ? -2 : -1; s_b[j] = a.kps2[j].octave; }
        else { s_a[j] = a.has_mp2[j] ? -2 : -1; s_b[j] = __float_as_int(ang); }
    }
    // Stage the batch that starts at query q0 into buffer `buf` (called by warps 1..7 together; lt = 0..223).  A batch takes
    // consecutive queries until kResQB queries or kResE entries; all its global reads are two dependent round trips
    // (query fields, then the candidate entries), however many queries it holds.
    auto stage = [&](int q0, int buf) {
        int *bp = s_buf + (size_t) buf * kBufInts;
        uint32_t *pack = reinterpret_cast<uint32_t *>(bp);
        int *off = bp + kResE, *gs = off + kResQB + 1, *ln = gs + kResQB, *qa = ln + kResQB, *hdr = qa + kResQB;
        const int lt = tid - 32;
        for (int j = lt; j < kResQB; j += 224) {
            const int qi = q0 + j;
            int st = 0, n = 0;
            if (qi < a.nq && (kVariant == 3 || kVariant == 4 || a.qvalid[qi])) { st = a.q_beg[qi]; n = a.q_end[qi] - st; }
            gs[j] = st; ln[j] = n;
            qa[j] = (n && a.q_angle) ? __float_as_int(a.q_angle[qi]) : 0;
        }
        loaders_sync();
        if (wid == 1) {                                        // prefix of min(length, E) over the 128 slots, 4 per lane
            int c[4], run = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) { c[i] = min(ln[4 * lane + i], kResE); run += c[i]; }
            int inc = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            int ex = inc - run;
#pragma unroll
            for (int i = 0; i < 4; ++i) { off[4 * lane + i] = ex; ex += c[i]; }
            if (lane == 31) off[kResQB] = ex;
            __syncwarp();
            // queries in the batch: the longest prefix whose entries fit (at least one query), not past the last query
            int m = 0;
            for (int j = lane; j < kResQB; j += 32) if (off[j + 1] <= kResE) m = max(m, j + 1);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
            m = max(1, min(m, a.nq - q0));
            if (lane == 0) { hdr[0] = q0; hdr[1] = m; }
        }
        loaders_sync();
        const int m = hdr[1], total = min(off[m], kResE);
        for (int e = lt; e < total; e += 224) {
            int lo = 0, hi = m - 1;                            // last j with off[j] <= e
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (off[mid] <= e) lo = mid; else hi = mid - 1; }
            const int src = gs[lo] + (e - off[lo]);
            pack[e] = ((uint32_t) a.c_dist[src] << 16) | (uint32_t) a.c_idx[src];
        }
    };
    __syncthreads();
    if (wid > 0 && a.nq > 0) stage(0, 0);
    __syncthreads();
    int n_match = 0, q0 = 0;
    for (int b = 0; q0 < a.nq; ++b) {
        const int *bp = s_buf + (size_t) (b & 1) * kBufInts;
        const uint32_t *pack = reinterpret_cast<const uint32_t *>(bp);
        const int *off = bp + kResE, *gs = off + kResQB + 1, *ln = gs + kResQB, *qa = ln + kResQB, *hdr = qa + kResQB;
        const int m = hdr[1];
        if (wid > 0) {
            if (q0 + m < a.nq) stage(q0 + m, (b + 1) & 1);
        } else {
            for (int j0 = 0; j0 < m; j0 += 32) {
              // queries with candidates in this group of 32 (a single warp pays every latency in full: skip the empty ones in bulk)
              unsigned live = __ballot_sync(0xffffffffu, j0 + lane < m && ln[j0 + lane] != 0);
              while (live) {
                const int j = j0 + __ffs(live) - 1;
                live &= live - 1;
                const int n = ln[j];
                const int s = gs[j], qi = q0 + j, o = off[j], staged = min(n, kResE - o);
                auto entry = [&](int k) -> uint32_t { return k < staged ? pack[o + k] : (((uint32_t) a.c_dist[s + k] << 16) | (uint32_t) a.c_idx[s + k]); };
                uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu;
                for (int k = lane; k < n; k += 64) {           // two entries per step: their loads and state look-ups overlap
                    const bool has2 = k + 32 < n;
                    const uint32_t pa = entry(k), pb = has2 ? entry(k + 32) : 0u;
                    const int ia = (int) (pa & 0xffffu), da = (int) (pa >> 16), ib = (int) (pb & 0xffffu), db = (int) (pb >> 16);
                    const int sa = s_a[ia], sb = s_a[ib];
                    const bool skip_a = kVariant == 0 ? sa <= da : sa != -1;                 // :63 / occupied or already taken
                    const bool skip_b = !has2 || (kVariant == 0 ? sb <= db : sb != -1);
                    if (!skip_a) { const uint32_t key = ((uint32_t) da << 22) | (uint32_t) k; k2 = min(k2, max(key, k1)); k1 = min(k1, key); }
                    if (!skip_b) { const uint32_t key = ((uint32_t) db << 22) | (uint32_t) (k + 32); k2 = min(k2, max(key, k1)); k1 = min(k1, key); }
                }
                {   // two smallest keys of the warp with the hardware reductions (keys are unique: they carry the position)
                    const uint32_t g1 = __reduce_min_sync(0xffffffffu, k1);
                    const uint32_t g2 = __reduce_min_sync(0xffffffffu, k1 == g1 ? k2 : k1);
                    k1 = g1; k2 = g2;
                }
                if (k1 == 0xffffffffu) continue;               // every candidate skipped: best stays at its initial value -> no match
                const int best = (int) (k1 >> 22), best_idx2 = (int) (entry((int) (k1 & 0x3fffffu)) & 0xffffu);
                bool accept;
                if (kVariant == 0) {
                    const int best2 = k2 == 0xffffffffu ? INT_MAX : (int) (k2 >> 22);
                    accept = best <= TH_LOW && best < cv_round_sat(__fmul_rn((float) best2, a.nn_ratio));      // :74
                } else if (kVariant == 1) {
                    accept = best <= TH_HIGH;                                                                  // :245
                } else if (kVariant == 2) {
                    accept = best <= TH_HIGH;
                    if (accept && k2 != 0xffffffffu) {
                        const int second = (int) (k2 >> 22);
                        const int lvl1 = s_b[best_idx2], lvl2 = s_b[entry((int) (k2 & 0x3fffffu)) & 0xffffu];
                        if (lvl1 == lvl2 && (float) best > __fmul_rn(a.nn_ratio, (float) second)) accept = false;   // :401-405
                    }
                } else if (kVariant == 3) {
                    accept = best < TH_LOW && best_idx2 > 0;                                                   // :464-484 (sic: index 0 never accepted)
                } else {
                    const int second = k2 == 0xffffffffu ? 256 : (int) (k2 >> 22);                             // :149 initial 256
                    accept = best <= TH_LOW && (float) best < __fmul_rn(a.nn_ratio, (float) second);           // :164
                }
                if (accept) {
                    if (lane == 0) {
                        const float qang = __int_as_float(qa[j]);
                        if (kVariant == 0) {
                            const int old = s_b[best_idx2];
                            if (old >= 0) { a.matches12[old] = -1; n_match--; }
                            a.matches12[qi] = best_idx2; s_b[best_idx2] = qi; s_a[best_idx2] = best;
                            n_match++;
                            if (a.check_orientation) { const int bn = rot_bin(qang, __int_as_float(s_c[best_idx2])); hist[bn]++; a.bin_of[qi] = bn; }
                        } else if (kVariant == 3) {
                            const int idx1 = a.q_out_idx[qi];
                            a.matches12[idx1] = best_idx2; s_a[best_idx2] = 1; n_match++;
                            if (a.check_orientation) { const int bn = rot_bin(qang, __int_as_float(s_b[best_idx2])); hist[bn]++; a.bin_of[idx1] = bn; }
                        } else {
                            s_a[best_idx2] = kVariant == 4 ? a.q_out_idx[qi] : qi; n_match++;
                            if ((kVariant == 1 || kVariant == 4) && a.check_orientation) { const int bn = rot_bin(qang, __int_as_float(s_b[best_idx2])); hist[bn]++; a.bin_of[best_idx2] = bn; }
                        }
                    }
                    __syncwarp();
                }
              }
            }
        }
        q0 += m;
        __syncthreads();
    }
    // rotation consistency: keep the three dominant bins (ORBMatcher.cpp:95-108 and copies); warp 0 owns the counters
    __threadfence_block();
    __syncthreads();
    if (wid == 0) {
        n_match = __shfl_sync(0xffffffffu, n_match, 0);
        if (a.check_orientation && kVariant != 2) {
            int i1, i2, i3;
            three_maxima(hist, i1, i2, i3);
            // bin_of / matches12 are indexed by frame-1 key points for variants 0 and 3 (a.n2 carries n1 for variant 3)
            const int limit = kVariant == 0 ? a.nq : a.n2;
            int removed = 0;
            for (int i = lane; i < limit; i += 32) {
                const int bn = a.bin_of[i];
                if (bn < 0 || bn == i1 || bn == i2 || bn == i3) continue;
                if (kVariant == 0 || kVariant == 3) { if (a.matches12[i] >= 0) { a.matches12[i] = -1; removed++; } }
                else { s_a[i] = -1; removed++; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
            n_match -= removed;
        }
        if (lane == 0) *a.n_matches = n_match;
    }
    __threadfence_block();
    __syncthreads();
    if (kVariant == 0) {                                    // update previous match (:111-113)
        for (int i = tid; i < a.nq; i += 256) {
            const int m = a.matches12[i];
            if (m >= 0) { a.prematched[2 * i] = a.kps2[m].x; a.prematched[2 * i + 1] = a.kps2[m].y; }
        }
    } else if (kVariant == 1 || kVariant == 2 || kVariant == 4) {   // slot owners back to global memory (occupied slots report -1)
        for (int j = tid; j < a.n2; j += 256) a.assigned[j] = s_a[j] < -1 ? -1 : s_a[j];
    }
}

__global__ void k_fill_int(int *p, int v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// GRID_COLS / GRID_ROWS (Frame.cpp:33-41); the grid itself (CSR, cell = cx*rows + cy) is built on the device by
// frame_grid_launch (orbfe_frame.cu) from the uploaded key points
static void grid_dims(int img_w, int img_h, int &cols, int &rows) {
    cols = img_w % GRID_SIZE == 0 ? img_w / GRID_SIZE : img_w / GRID_SIZE + 1;
    rows = img_h % GRID_SIZE == 0 ? img_h / GRID_SIZE : img_h / GRID_SIZE + 1;
}

static void fill_int(Handle *h, int *p, int v, int n, cudaStream_t st) {
    if (n > 0) { k_fill_int<<<(n + 255) / 256, 256, 0, st>>>(p, v, n); h->launches++; }
}

// shared driver of the three grid-window searches
struct WindowProblem {
    const float *q_u, *q_v, *q_r; const int *q_min, *q_max; const uint8_t *q_valid; const uint8_t *q_desc; const float *q_angle; int nq;
    const orbfe_keypoint *kps2; const uint8_t *desc2; int n2; int img_w, img_h; const uint8_t *occupied;
    // device-resident operands (orbfe_frame): the searched frame with its grid, and — SearchForInitialization — the query frame's descriptors
    const orbfe_frame *frame2 = nullptr, *frame1 = nullptr;
};

// ------------------------------------------------------------------------------------------------
// Parallel resolve of the searches whose frame-2 slots are exclusive (variants 1, 2, 3, 4: a slot that holds a match is skipped by
// every later query).  The reference's loop is a serial dictatorship: the queries pick, in index order, their best still-free slot
// (or nothing, if the acceptance test fails).  Query i's decision is a function of the decisions of the queries before it only —
// dec[i] = f(dec[0 .. i-1]) — so the sequential result is the unique fixed point of evaluating all queries at once against the
// previous round's decisions (by induction over i: query 0 depends on nothing, query i is right once all earlier ones are):
//   1. every query that currently holds a slot marks it with its index (atomicMin: the earliest holder);
//   2. every query re-evaluates: a candidate is available unless it was occupied before the call or is held by an EARLIER query;
//      best / second-best available candidate (first minimum in list order), acceptance exactly as in the reference;
//   3. repeat until no decision changed.
// The number of rounds is the depth of the chains of actual displacements (6-9 on tracking-sized windows), not the number of queries
// that merely share a candidate.  One CTA of 1024 threads, thread per query, candidate lists staged in shared memory when they fit.
// ------------------------------------------------------------------------------------------------
template <int kVariant>
__global__ void __launch_bounds__(1024) k_resolve_par(const ResolveArgs a) {
    extern __shared__ __align__(16) uint8_t rp_dyn[];
    __shared__ int hist[HISTO_LENGTH];
    __shared__ int s_changed, s_nmatch;
    const int tid = threadIdx.x;
    const int n2s = kVariant == 3 ? a.n_state : a.n2;
    int *held = reinterpret_cast<int *>(rp_dyn);             // earliest query holding the slot; -1: occupied before the call; INT_MAX: free
    int *dec = held + n2s;                                   // slot query i takes, or -1
    int *s_beg = dec + a.nq, *s_end = s_beg + a.nq;          // candidate list of query i (empty for an invalid query): read every round
    uint8_t *s_occ = reinterpret_cast<uint8_t *>(s_end + a.nq);   // slots occupied before the call
    // candidate lists (dist << 16 | idx) staged once when they fit: the rounds then never touch global memory for them
    uint8_t *s_oct = s_occ + n2s;                            // variant 2: octave of every slot (the same-level ratio test reads two per query and round)
    uint32_t *s_pack = reinterpret_cast<uint32_t *>(s_oct + n2s + ((4 - ((2 * n2s) & 3)) & 3));
    const int total = a.n_cand ? *a.n_cand : 0;
    const bool staged = a.n_cand && total <= a.smem_entries;
    if (staged) for (int k = tid; k < total; k += 1024) s_pack[k] = ((uint32_t) a.c_dist[k] << 16) | (uint32_t) a.c_idx[k];
    auto cand_idx = [&](int k) -> int { return staged ? (int) (s_pack[k] & 0xffffu) : a.c_idx[k]; };
    auto cand_dist = [&](int k) -> int { return staged ? (int) (s_pack[k] >> 16) : a.c_dist[k]; };
    if (tid < HISTO_LENGTH) hist[tid] = 0;
    if (tid == 0) s_nmatch = 0;
    for (int qi = tid; qi < a.nq; qi += 1024) {
        dec[qi] = -1;
        const bool live = kVariant >= 3 || a.qvalid[qi];
        const int b = a.q_beg[qi];
        s_beg[qi] = b; s_end[qi] = live ? a.q_end[qi] : b;
    }
    for (int j = tid; j < n2s; j += 1024) {
        s_occ[j] = (kVariant == 3 ? a.has_mp2[j] : a.occupied[j]) ? 1 : 0;
        if (kVariant == 2) s_oct[j] = (uint8_t) a.kps2[j].octave;
    }
    __syncthreads();
    int round = 0;
    while (true) {
        for (int j = tid; j < n2s; j += 1024) held[j] = s_occ[j] ? -1 : INT_MAX;
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int qi = tid; qi < a.nq; qi += 1024) { const int s = dec[qi]; if (s >= 0) atomicMin(&held[s], qi); }
        __syncthreads();
        bool changed = false;
        for (int qi = tid; qi < a.nq; qi += 1024) {
            int nd = -1;
            const int beg = s_beg[qi], end = s_end[qi];
            if (end > beg) {
                uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu;
                for (int k = beg; k < end; ++k) {
                    if (held[cand_idx(k)] < qi) continue;                               // occupied, or taken by an earlier query
                    const uint32_t key = ((uint32_t) cand_dist(k) << 22) | (uint32_t) (k - beg);
                    k2 = min(k2, max(key, k1));
                    k1 = min(k1, key);
                }
                if (k1 != 0xffffffffu) {
                    const int best = (int) (k1 >> 22), best_idx2 = cand_idx(beg + (int) (k1 & 0x3fffffu));
                    const int s2 = k2 == 0xffffffffu ? -1 : cand_idx(beg + (int) (k2 & 0x3fffffu));
                    bool accept;
                    if (kVariant == 1) accept = best <= TH_HIGH;                                                  // :245
                    else if (kVariant == 2) {
                        accept = best <= TH_HIGH;
                        if (accept && s2 >= 0) {
                            const int second = (int) (k2 >> 22);
                            if (s_oct[best_idx2] == s_oct[s2] && (float) best > __fmul_rn(a.nn_ratio, (float) second)) accept = false;   // :401-405
                        }
                    } else if (kVariant == 3) accept = best < TH_LOW && best_idx2 > 0;                            // :464-484 (sic)
                    else {
                        const int second = s2 < 0 ? 256 : (int) (k2 >> 22);
                        accept = best <= TH_LOW && (float) best < __fmul_rn(a.nn_ratio, (float) second);          // :164
                    }
                    if (accept) nd = best_idx2;
                }
            }
            if (nd != dec[qi]) { dec[qi] = nd; changed = true; }
        }
        if (changed) s_changed = 1;
        __syncthreads();
        ++round;
        if (tid == 0 && a.dbg) a.dbg[2] = round;
        const bool more = s_changed != 0;
        __syncthreads();
        if (!more) break;
    }
    // fixed point reached: held[] is the unique holder of every taken slot.  Owners, match count, rotation histogram
    for (int qi = tid; qi < a.nq; qi += 1024) {
        const int s = dec[qi];
        if (s < 0) continue;
        const int owner = (kVariant == 3 || kVariant == 4) ? a.q_out_idx[qi] : qi;
        if (kVariant == 3) a.matches12[owner] = s;
        atomicAdd(&s_nmatch, 1);
        if (a.check_orientation && kVariant != 2) {
            const int bn = rot_bin(a.q_angle[qi], a.kps2[s].angle);
            atomicAdd(&hist[bn], 1);
            a.bin_of[kVariant == 3 ? owner : s] = bn;
        }
    }
    __syncthreads();
    // slot -> owner (the variants with q_out_idx report the owner's key-point index)
    int *taken = held;
    for (int j = tid; j < n2s; j += 1024) {
        const int q = held[j];
        taken[j] = (q >= 0 && q != INT_MAX) ? ((kVariant == 3 || kVariant == 4) ? a.q_out_idx[q] : q) : (q == -1 ? -2 : -1);
    }
    __syncthreads();
    // rotation consistency (ORBMatcher.cpp:95-108 and copies), then the slot owners go back to global memory
    if (a.check_orientation && kVariant != 2) {
        int i1, i2, i3;
        three_maxima(hist, i1, i2, i3);
        const int limit = a.n2;                               // variant 3: n2 carries n1 (bin_of / matches12 per frame-1 key point)
        for (int i = tid; i < limit; i += 1024) {
            const int bn = a.bin_of[i];
            if (bn < 0 || bn == i1 || bn == i2 || bn == i3) continue;
            if (kVariant == 3) { if (a.matches12[i] >= 0) { a.matches12[i] = -1; atomicSub(&s_nmatch, 1); } }
            else { taken[i] = -1; atomicSub(&s_nmatch, 1); }
        }
    }
    __syncthreads();
    if (kVariant != 3) for (int j = tid; j < a.n2; j += 1024) a.assigned[j] = taken[j] < -1 ? -1 : taken[j];
    if (tid == 0) *a.n_matches = s_nmatch;
}

// ------------------------------------------------------------------------------------------------
// Parallel resolve of SearchForInitialization (variant 0).  Here a slot is not exclusive: a later query takes it from its owner if its
// distance is smaller (ORBMatcher.cpp:63, 75-78), and a candidate is skipped when the slot's current matched distance is <= the
// candidate's.  The decision of query i is still a function of the decisions of the queries before it only, so the same fixed-point
// iteration applies: every round each slot collects its acceptors (query, distance) of the previous round; query i then sees, for a
// candidate slot, the smallest distance among the acceptors EARLIER than i — which is the matched distance the sequential loop would
// hold when it reaches i — and re-evaluates best / second-best / the ratio test.  At the fixed point the owner of a slot is its
// latest acceptor; the earlier ones are the stolen matches (they still count in the rotation histogram, as in the reference, where
// rotHist keeps their entries).  A slot keeps at most kInitAcc acceptors per round; if one ever needs more the kernel raises
// `fallback` and the sequential kernel k_resolve<0> runs instead.
// ------------------------------------------------------------------------------------------------
constexpr int kInitAcc = 4;

// Candidates at distance d_far or more never decide anything by their value: an accepted best is <= TH_LOW, so the ratio test
// best < cvRound(second * ratio) holds for every second >= d_far (d_far = the smallest distance with cvRound(d * ratio) > TH_LOW, computed
// on the host with the same float operations), and such a candidate can never be the best of an accepted match.  What matters is only
// whether an eligible one EXISTS when fewer than two near candidates are eligible (a lone candidate is tested against
// cvRound(INT_MAX * ratio), which saturates to INT_MIN for ratios from about 1.0 on, :74).  The kernel therefore keeps each query's near candidates (d < d_far; a handful) in
// shared memory with their list positions and scans the far ones in global memory only until the first eligible one.
__global__ void __launch_bounds__(1024) k_resolve_init_par(const ResolveArgs a, int *fallback) {
    extern __shared__ __align__(16) uint8_t ri_dyn[];
    __shared__ int hist[HISTO_LENGTH];
    __shared__ int s_changed, s_nmatch, s_overflow, s_near;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int *cnt = reinterpret_cast<int *>(ri_dyn);                              // acceptors of the slot in the previous round
    uint32_t *acc = reinterpret_cast<uint32_t *>(cnt + a.n2);                // [slot][kInitAcc]: query << 9 | distance
    int *dec = reinterpret_cast<int *>(acc + (size_t) a.n2 * kInitAcc);      // slot | distance << 16, or -1
    int *nbeg = dec + a.nq;                                                  // near candidates of query i: arena[nbeg[i] .. nbeg[i] + ncnt[i])
    int *ncnt = nbeg + a.nq;
    uint2 *arena = reinterpret_cast<uint2 *>(((uintptr_t) (ncnt + a.nq) + 7) & ~(uintptr_t) 7);      // (distance << 16 | slot, position in the query's list)
    const int arena_cap = a.smem_entries;
    const int d_far = a.acc_far;
    if (tid < HISTO_LENGTH) hist[tid] = 0;
    if (tid == 0) { s_nmatch = 0; s_overflow = 0; s_near = 0; }
    for (int qi = tid; qi < a.nq; qi += 1024) { dec[qi] = -1; ncnt[qi] = 0; nbeg[qi] = 0; }
    __syncthreads();
    // ---- near candidates -> shared memory, one warp per query (coalesced reads, ballot compaction, list order kept)
    for (int qi = wid; qi < a.nq; qi += 32) {
        if (!a.qvalid[qi]) continue;
        const int beg = a.q_beg[qi], end = a.q_end[qi];
        int n_near = 0;
        for (int k0 = beg; k0 < end; k0 += 32) n_near += __popc(__ballot_sync(0xffffffffu, k0 + lane < end && a.c_dist[k0 + lane] < d_far));
        if (n_near == 0) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_near, n_near);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base + n_near > arena_cap) { if (lane == 0) s_overflow = 1; continue; }
        if (lane == 0) { nbeg[qi] = base; ncnt[qi] = n_near; }
        int run = 0;
        for (int k0 = beg; k0 < end; k0 += 32) {
            const int k = k0 + lane;
            const int d = k < end ? a.c_dist[k] : INT_MAX;
            const unsigned bal = __ballot_sync(0xffffffffu, d < d_far);
            if (d < d_far) arena[base + run + __popc(bal & ((1u << lane) - 1u))] = make_uint2(((uint32_t) d << 16) | (uint32_t) a.c_idx[k], (uint32_t) (k - beg));
            run += __popc(bal);
        }
    }
    __syncthreads();
    if (s_overflow) { if (tid == 0) *fallback = 1; return; }
    const int acc_limit = a.acc_limit ? a.acc_limit : kInitAcc;
    // matchedDistance[c] <= d when the sequential loop reaches query qi (:63): some acceptor earlier than qi holds c at distance <= d
    auto blocked = [&](int c, int d, int qi) -> bool {
        const int m = min(cnt[c], kInitAcc);
        bool b = false;
        for (int t = 0; t < m; ++t) { const uint32_t ae = acc[(size_t) c * kInitAcc + t]; if ((int) (ae >> 9) < qi && (int) (ae & 0x1ffu) <= d) b = true; }
        return b;
    };
    while (true) {
        for (int j = tid; j < a.n2; j += 1024) cnt[j] = 0;
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int qi = tid; qi < a.nq; qi += 1024) {
            const int d = dec[qi];
            if (d < 0) continue;
            const int c = d & 0xffff, pos = atomicAdd(&cnt[c], 1);
            if (pos < acc_limit) acc[(size_t) c * kInitAcc + pos] = ((uint32_t) qi << 9) | (uint32_t) (d >> 16);
            else s_overflow = 1;
        }
        __syncthreads();
        if (s_overflow) { if (tid == 0) *fallback = 1; return; }
        bool changed = false;
        for (int qi = tid; qi < a.nq; qi += 1024) {
            int nd = -1;
            const int nn = ncnt[qi];
            if (nn) {                                                       // without a near candidate the best is > TH_LOW: no match
                uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu; int c1 = -1;
                const uint2 *e = arena + nbeg[qi];
                for (int t = 0; t < nn; ++t) {
                    const int c = (int) (e[t].x & 0xffffu), d = (int) (e[t].x >> 16);
                    if (blocked(c, d, qi)) continue;
                    const uint32_t key = ((uint32_t) d << 22) | e[t].y;
                    if (key < k1) { k2 = k1; k1 = key; c1 = c; } else if (key < k2) k2 = key;
                }
                if (k1 != 0xffffffffu && (int) (k1 >> 22) <= TH_LOW) {
                    const int best = (int) (k1 >> 22);
                    bool accept;
                    if (k2 != 0xffffffffu) accept = best < cv_round_sat(__fmul_rn((float) (int) (k2 >> 22), a.nn_ratio));        // :74, near second
                    else {                                                  // the second-best, if any, is a far candidate: it only has to exist
                        accept = false;
                        for (int k = a.q_beg[qi], end = a.q_end[qi]; k < end && !accept; ++k) {
                            const int d = a.c_dist[k];
                            if (d >= d_far && !blocked(a.c_idx[k], d, qi)) accept = true;
                        }
                        // a lone candidate: bestDist2 keeps its initial INT_MAX (saturates to INT_MIN for ratios from about 1.0 on)
                        if (!accept) accept = best < cv_round_sat(__fmul_rn((float) INT_MAX, a.nn_ratio));
                    }
                    if (accept) nd = c1 | (best << 16);
                }
            }
            if (nd != dec[qi]) { dec[qi] = nd; changed = true; }
        }
        if (changed) s_changed = 1;
        __syncthreads();
        const bool more = s_changed != 0;
        if (tid == 0 && a.dbg) a.dbg[2]++;
        __syncthreads();
        if (!more) break;
    }
    // fixed point: cnt / acc hold the acceptors of the final decisions.  Every acceptor enters the rotation histogram (:84-91); the
    // latest acceptor of a slot owns it (:75-78)
    for (int qi = tid; qi < a.nq; qi += 1024) {
        const int d = dec[qi];
        if (d < 0) continue;
        const int c = d & 0xffff;
        bool owner = true;
        const int m = min(cnt[c], kInitAcc);
        for (int t = 0; t < m; ++t) if ((int) (acc[(size_t) c * kInitAcc + t] >> 9) > qi) owner = false;
        if (owner) { a.matches12[qi] = c; atomicAdd(&s_nmatch, 1); }
        if (a.check_orientation) {
            const int bn = rot_bin(a.q_angle[qi], a.kps2[c].angle);
            atomicAdd(&hist[bn], 1);
            a.bin_of[qi] = bn;
        }
    }
    __syncthreads();
    if (a.check_orientation) {                                              // :95-108
        int i1, i2, i3;
        three_maxima(hist, i1, i2, i3);
        for (int i = tid; i < a.nq; i += 1024) {
            const int bn = a.bin_of[i];
            if (bn < 0 || bn == i1 || bn == i2 || bn == i3) continue;
            if (a.matches12[i] >= 0) { a.matches12[i] = -1; atomicSub(&s_nmatch, 1); }
        }
    }
    __syncthreads();
    for (int i = tid; i < a.nq; i += 1024) {                                // update previous match (:111-113)
        const int m = a.matches12[i];
        if (m >= 0) { a.prematched[2 * i] = a.kps2[m].x; a.prematched[2 * i + 1] = a.kps2[m].y; }
    }
    if (tid == 0) *a.n_matches = s_nmatch;
}

template <int kVariant>
static int launch_resolve(Handle *h, const ResolveArgs &ra, int n2, cudaStream_t st) {
    const char *serial_env = getenv("ORBFE_SERIAL_RESOLVE");                       // A/B aid, read per call
    const bool serial = serial_env && *serial_env == '1';
    if (n2 >= 65536) return set_error(h, ORBFE_E_ARG, "matcher: %d key points in the searched frame (limit 65535)", n2);
    if (kVariant != 0 && !serial) {
        const size_t state = sizeof(int) * ((size_t) n2 + 3 * (size_t) ra.nq) + 2 * (size_t) n2 + 64;  // holder, occupancy, octave per slot; decision + list bounds per query
        if (state <= 200 * 1024) {
            ResolveArgs rb = ra;
            rb.smem_entries = (int) ((200 * 1024 - state) / sizeof(uint32_t));
            k_resolve_par<kVariant><<<1, 1024, 200 * 1024, st>>>(rb);
            return ORBFE_OK;
        }
    }
    const size_t smem = resolve_smem_bytes<kVariant>(n2);
    if (smem > 200 * 1024)
        return set_error(h, ORBFE_E_ARG, "matcher: %d key points in the searched frame exceed the resolve kernel's shared-memory state (limit about %d)", n2,
                         kVariant == 0 ? 13000 : 20000);
    if (kVariant == 0 && !serial) {
        const size_t state = sizeof(int) * ((size_t) n2 * (1 + kInitAcc) + 3 * (size_t) ra.nq + 2) + 64;
        if (state + 8 * 1024 <= 200 * 1024) {
            ResolveArgs rb = ra;
            rb.smem_entries = (int) ((200 * 1024 - state) / sizeof(uint2));          // near-candidate arena
            rb.acc_far = 257;                                                        // smallest distance with cvRound(d * ratio) > TH_LOW
            for (int d = 256; d >= 0; --d) {
                const float v = (float) d * ra.nn_ratio;
                const int r = (!(v < 2147483648.f) || v < -2147483648.f) ? INT_MIN : (int) lrintf(v);
                if (r > TH_LOW) rb.acc_far = d; else break;
            }
            rb.acc_far = std::max(rb.acc_far, TH_LOW + 1);                          // a candidate that could be an accepted best is always "near"
            if (const char *e = getenv("ORBFE_INIT_ACC_LIMIT")) rb.acc_limit = std::min(std::max(atoi(e), 1), kInitAcc);    // tests: drive the fallback
            int *fallback = ra.n_matches + 3;                     // zeroed with the match counter; raised if a slot collects too many acceptors
            k_resolve_init_par<<<1, 1024, 200 * 1024, st>>>(rb, fallback);
            k_resolve<kVariant><<<1, 256, smem, st>>>(ra, fallback);
            h->launches++;
            return ORBFE_OK;
        }
    }
    k_resolve<kVariant><<<1, 256, smem, st>>>(ra, nullptr);
    return ORBFE_OK;
}

// The resolves use up to 200 KB of dynamic shared memory.  The opt-in is a per-device function attribute, so orbfe_create calls
// this for the handle's device (after cudaSetDevice) instead of a process-wide "done once" flag.
template <int kVariant>
static cudaError_t resolve_attrs() {
    cudaError_t e = cudaFuncSetAttribute(k_resolve<kVariant>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess && kVariant != 0) e = cudaFuncSetAttribute(k_resolve_par<kVariant>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    return e;
}
int match_device_setup(Handle *h) {
    ORBFE_CUDA(h, resolve_attrs<0>()); ORBFE_CUDA(h, resolve_attrs<1>()); ORBFE_CUDA(h, resolve_attrs<2>());
    ORBFE_CUDA(h, resolve_attrs<3>()); ORBFE_CUDA(h, resolve_attrs<4>());
    ORBFE_CUDA(h, cudaFuncSetAttribute(k_resolve_init_par, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return allpairs_tc_device_setup(h);
}

// pinned host staging of the matcher (one upload and one download per search instead of a dozen small copies)
static int ensure_match_pinned(Handle *h, size_t bytes) {
    if (h->mpin_bytes >= bytes) return ORBFE_OK;
    if (h->h_mpin) cudaFreeHost(h->h_mpin);
    h->h_mpin = nullptr; h->mpin_bytes = 0;
    bytes = bytes + bytes / 4 + 4096;
    ORBFE_CUDA(h, cudaMallocHost(&h->h_mpin, bytes));
    h->mpin_bytes = bytes;
    return ORBFE_OK;
}

template <int kVariant>
static int run_window_search(Handle *h, const WindowProblem &p, float nn_ratio, int check_orientation,
                             int *out_n2 /*assigned*/, int *matches12 /*init*/, float *prematched, int *n_matches) {
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int nq = p.nq, n2 = p.n2;
    int cols, rows;
    grid_dims(p.img_w, p.img_h, cols, rows);
    const size_t n_off = (size_t) cols * rows + 1, n_gidx = (size_t) std::max(n2, 1);
    static const bool trace = [] { const char *e = getenv("ORBFE_TRACE"); return e && *e == '1'; }();
    const auto t0 = std::chrono::steady_clock::now();
    auto us = [&] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); };
    // device layout: [upload block: hdr, queries, frame-2 key points and descriptors, prematched][download block: nmatch, result]
    // [device-only state].  The upload block is mirrored in pinned host memory at the same offsets.
    int *hdr; float *qx, *qy, *qr, *qang, *pre; int *qmin, *qmax, *coff, *cidx, *qbeg, *qend, *assigned, *m12, *m21, *mdist, *binof, *nmatch, *cand_idx, *cand_dist;
    uint8_t *qvalid, *occ; uint4 *qdesc, *desc2; orbfe_keypoint *kps2;
    size_t up_end = 0, down_beg = 0, down_end = 0;
    auto layout = [&](Bump &b, size_t cand_cap) {
        hdr = b.take<int>(4);
        qx = b.take<float>(nq); qy = b.take<float>(nq); qr = b.take<float>(nq); qmin = b.take<int>(nq); qmax = b.take<int>(nq);
        qvalid = b.take<uint8_t>(nq); qang = b.take<float>(nq); occ = b.take<uint8_t>(n2);
        // operands that live in a device-resident frame are neither staged nor uploaded
        qdesc = p.frame1 ? (uint4 *) p.frame1->d_desc : b.take<uint4>(2 * (size_t) nq);
        kps2 = p.frame2 ? (orbfe_keypoint *) p.frame2->d_kps : b.take<orbfe_keypoint>(n2);
        desc2 = p.frame2 ? (uint4 *) p.frame2->d_desc : b.take<uint4>(2 * (size_t) n2);
        const size_t pre_off = (b.off + 255) & ~(size_t) 255;
        pre = b.take<float>(2 * (size_t) nq);
        up_end = b.off;
        down_beg = kVariant == 0 ? pre_off : ((b.off + 255) & ~(size_t) 255);
        nmatch = b.take<int>(4);
        if (kVariant == 0) { m12 = b.take<int>(nq); down_end = b.off; assigned = b.take<int>(n2); }
        else { assigned = b.take<int>(n2); down_end = b.off; m12 = b.take<int>(nq); }
        coff = p.frame2 ? (int *) p.frame2->d_grid_off : b.take<int>(n_off); cidx = p.frame2 ? (int *) p.frame2->d_grid_idx : b.take<int>(n_gidx);
        qbeg = b.take<int>(nq); qend = b.take<int>(nq);
        m21 = b.take<int>(n2); mdist = b.take<int>(n2); binof = b.take<int>(std::max(nq, n2));
        cand_idx = b.take<int>(cand_cap); cand_dist = b.take<int>(cand_cap);
    };
    size_t cand_cap = (size_t) nq * 64 + 1024;      // first guess; an overflowing pass reports the exact total and is repeated
    for (int attempt = 0;; ++attempt) {
        Bump probe{nullptr}; layout(probe, cand_cap);
        int rc = ensure_match_scratch(h, probe.off + 1024);
        if (rc) return rc;
        if ((rc = ensure_match_pinned(h, std::max(up_end, down_end) + 1024))) return rc;
        Bump b{(uint8_t *) h->d_match}; layout(b, cand_cap);
        uint8_t *hp = (uint8_t *) h->h_mpin, *db = (uint8_t *) h->d_match;
        auto H = [&](const void *dev) { return hp + ((const uint8_t *) dev - db); };
        int hdr_h[4] = {n2, 0, 0, 0};
        memcpy(H(hdr), hdr_h, sizeof hdr_h);
        memcpy(H(qx), p.q_u, sizeof(float) * nq); memcpy(H(qy), p.q_v, sizeof(float) * nq); memcpy(H(qr), p.q_r, sizeof(float) * nq);
        memcpy(H(qmin), p.q_min, sizeof(int) * nq); memcpy(H(qmax), p.q_max, sizeof(int) * nq); memcpy(H(qvalid), p.q_valid, nq);
        if (p.q_angle) memcpy(H(qang), p.q_angle, sizeof(float) * nq); else memset(H(qang), 0, sizeof(float) * nq);
        if (p.occupied) memcpy(H(occ), p.occupied, n2); else memset(H(occ), 0, n2);
        if (!p.frame1) memcpy(H(qdesc), p.q_desc, 32 * (size_t) nq);
        if (!p.frame2) { memcpy(H(kps2), p.kps2, sizeof(orbfe_keypoint) * (size_t) n2); memcpy(H(desc2), p.desc2, 32 * (size_t) n2); }
        if (prematched) memcpy(H(pre), prematched, sizeof(float) * 2 * (size_t) nq);
        ORBFE_CUDA(h, cudaMemcpyAsync(db, hp, up_end, cudaMemcpyHostToDevice, st));
        if (!p.frame2 && (rc = frame_grid_launch(h, kps2, hdr, std::max(n2, 1), p.img_w, p.img_h, coff, cidx, st))) return rc;
        const int n_init = std::max(std::max(nq, n2), 4);
        k_window_init<<<(n_init + 255) / 256, 256, 0, st>>>(binof, std::max(nq, n2), m12, nq, m21, mdist, assigned, n2, nmatch);
        WinArgs wa;
        wa.qx = qx; wa.qy = qy; wa.qr = qr; wa.qmin = qmin; wa.qmax = qmax; wa.qvalid = qvalid; wa.qdesc = qdesc; wa.nq = nq;
        wa.kps2 = kps2; wa.desc2 = desc2; wa.cell_off = coff; wa.cell_idx = cidx; wa.cols = cols; wa.rows = rows;
        wa.q_beg = qbeg; wa.q_end = qend; wa.cursor = nmatch + 1; wa.overflow = nmatch + 2; wa.cand_cap = (int) std::min<size_t>(cand_cap, INT_MAX);
        wa.c_idx = cand_idx; wa.c_dist = cand_dist;
        k_window<<<(nq + 7) / 8, 256, 0, st>>>(wa);
        ResolveArgs ra; memset(&ra, 0, sizeof ra);
        ra.nq = nq; ra.n2 = n2; ra.q_beg = qbeg; ra.q_end = qend; ra.c_idx = cand_idx; ra.c_dist = cand_dist; ra.qvalid = qvalid; ra.q_angle = qang; ra.kps2 = kps2;
        ra.occupied = occ; ra.matches12 = m12; ra.matches21 = m21; ra.matched_dist = mdist; ra.assigned = assigned; ra.bin_of = binof;
        ra.prematched = pre; ra.nn_ratio = nn_ratio; ra.check_orientation = check_orientation; ra.n_matches = nmatch; ra.dbg = trace ? hdr + 1 : nullptr; ra.n_cand = nmatch + 1;
        if ((rc = launch_resolve<kVariant>(h, ra, n2, st))) return rc;
        h->launches += 3;
        ORBFE_CUDA(h, cudaGetLastError());
        const double t_issue = us();
        ORBFE_CUDA(h, cudaMemcpyAsync(hp + down_beg, db + down_beg, down_end - down_beg, cudaMemcpyDeviceToHost, st));
        ORBFE_CUDA(h, cudaStreamSynchronize(st));
        if (trace) {
            int dbg[4] = {0, 0, 0, 0};
            cudaMemcpy(dbg, hdr, sizeof dbg, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[orbfe trace] window search variant %d: nq %d n2 %d, issued %.1f us, results on host %.1f us; resolve rounds %d, %d candidate entries, sequential fallback %d\n",
                    kVariant, nq, n2, t_issue, us(), dbg[3], ((const int *) H(nmatch))[1], ((const int *) H(nmatch))[3]);
        }
        const int *nm_h = (const int *) H(nmatch);
        if (nm_h[2]) {                                    // candidate lists did not fit: nm_h[1] is the exact total
            if (attempt) return set_error(h, ORBFE_E_INTERNAL, "candidate count changed between passes");
            cand_cap = (size_t) nm_h[1] + 1024;
            continue;
        }
        *n_matches = nm_h[0];
        if (kVariant == 0) {
            memcpy(matches12, H(m12), sizeof(int) * nq);
            memcpy(prematched, H(pre), sizeof(float) * 2 * (size_t) nq);
        } else {
            memcpy(out_n2, H(assigned), sizeof(int) * n2);
        }
        return ORBFE_OK;
    }
}

}  // namespace orbfe

using namespace orbfe;

// Shared driver of the two vocabulary-node searches (SearchForTriangulation = variant 3, SearchByBow = variant 4): merge-join of
// the two DBoW2 feature vectors on the host (std::map order, lower_bound skips: ORBMatcher.cpp:131-187, 443-512) into the query
// list in reference order with the node members of frame 2 as candidate lists; distances and the greedy resolve on the device.
template <int kVariant>
static int run_node_search(orbfe_handle *h, const uint8_t *desc1, const float *angle1, const uint8_t *flag1, int n1, const int32_t *node_id1,
                           const int32_t *node_off1, const int32_t *node_idx1, int n_nodes1, const uint8_t *desc2, const float *angle2,
                           const uint8_t *flag2, int n2, const int32_t *node_id2, const int32_t *node_off2, const int32_t *node_idx2, int n_nodes2,
                           int32_t *out, float nn_ratio, int check_orientation, int *n_matches) {
    std::vector<int> q_idx1, q_off(1, 0), c_idx; std::vector<float> q_ang;
    int a = 0, b = 0;
    while (a < n_nodes1 && b < n_nodes2) {
        if (node_id1[a] == node_id2[b]) {
            for (int i = node_off1[a]; i < node_off1[a + 1]; ++i) {
                const int idx1 = node_idx1[i];
                if (idx1 < 0 || idx1 >= n1) return set_error(h, ORBFE_E_ARG, "feature vector 1 index out of range");
                if (kVariant == 3 ? flag1[idx1] != 0 : flag1[idx1] == 0) continue;      // :452 has a map point / :143-144 no good map point
                q_idx1.push_back(idx1); q_ang.push_back(angle1[idx1]);
                for (int k = node_off2[b]; k < node_off2[b + 1]; ++k) {
                    if (node_idx2[k] < 0 || node_idx2[k] >= n2) return set_error(h, ORBFE_E_ARG, "feature vector 2 index out of range");
                    c_idx.push_back(node_idx2[k]);
                }
                q_off.push_back((int) c_idx.size());
            }
            ++a; ++b;
        } else if (node_id1[a] < node_id2[b]) { while (a < n_nodes1 && node_id1[a] < node_id2[b]) ++a; }
        else { while (b < n_nodes2 && node_id2[b] < node_id1[a]) ++b; }
    }
    const int nq = (int) q_idx1.size(), total = (int) c_idx.size();
    if (nq == 0) return ORBFE_OK;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    std::vector<orbfe_keypoint> kps2(n2);
    for (int j = 0; j < n2; ++j) { memset(&kps2[j], 0, sizeof(orbfe_keypoint)); kps2[j].angle = angle2[j]; }
    const int n_bin = kVariant == 3 ? n1 : n2;                 // bin_of is indexed by frame-1 key points (3) / frame-2 slots (4)
    auto layout = [&](Bump &bp, uint4 *&d1, uint4 *&d2, int *&qi, int *&qo, int *&ci, int *&cd, float *&qa, orbfe_keypoint *&k2, uint8_t *&f2, int *&m12,
                      int *&asg, int *&binof, int *&nm) {
        d1 = bp.take<uint4>(2 * (size_t) n1); d2 = bp.take<uint4>(2 * (size_t) n2); qi = bp.take<int>(nq); qo = bp.take<int>(nq + 1);
        ci = bp.take<int>(std::max(total, 1)); cd = bp.take<int>(std::max(total, 1)); qa = bp.take<float>(nq); k2 = bp.take<orbfe_keypoint>(n2);
        f2 = bp.take<uint8_t>(n2); m12 = bp.take<int>(n1); asg = bp.take<int>(n2); binof = bp.take<int>(n_bin); nm = bp.take<int>(4);
    };
    uint4 *d1, *d2; int *qi, *qo, *ci, *cd, *m12, *asg, *binof, *nm; float *qa; orbfe_keypoint *k2; uint8_t *f2;
    Bump probe{nullptr}; layout(probe, d1, d2, qi, qo, ci, cd, qa, k2, f2, m12, asg, binof, nm);
    int rc = ensure_match_scratch(h, probe.off + 1024);
    if (rc) return rc;
    Bump bp{(uint8_t *) h->d_match}; layout(bp, d1, d2, qi, qo, ci, cd, qa, k2, f2, m12, asg, binof, nm);
#define UP(dst, src, bytes) ORBFE_CUDA(h, cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, st))
    UP(d1, desc1, 32 * (size_t) n1); UP(d2, desc2, 32 * (size_t) n2); UP(qi, q_idx1.data(), sizeof(int) * nq); UP(qo, q_off.data(), sizeof(int) * (nq + 1));
    if (total) UP(ci, c_idx.data(), sizeof(int) * total);
    UP(qa, q_ang.data(), sizeof(float) * nq); UP(k2, kps2.data(), sizeof(orbfe_keypoint) * (size_t) n2);
    if (flag2) UP(f2, flag2, n2); else ORBFE_CUDA(h, cudaMemsetAsync(f2, 0, n2, st));
#undef UP
    fill_int(h, m12, -1, n1, st); fill_int(h, asg, -1, n2, st); fill_int(h, binof, -1, n_bin, st);
    ORBFE_CUDA(h, cudaMemcpyAsync(nm + 1, &total, sizeof(int), cudaMemcpyHostToDevice, st));     // candidate count for the resolve's staging
    k_csr_distance<<<(nq + 7) / 8, 256, 0, st>>>(d1, qi, qo, nq, ci, d2, cd);
    ResolveArgs ra; memset(&ra, 0, sizeof ra);
    ra.nq = nq; ra.q_beg = qo; ra.q_end = qo + 1; ra.c_idx = ci; ra.c_dist = cd; ra.q_angle = qa; ra.kps2 = k2;
    ra.q_out_idx = qi; ra.matches12 = m12; ra.assigned = asg; ra.bin_of = binof; ra.check_orientation = check_orientation; ra.n_matches = nm;
    ra.nn_ratio = nn_ratio; ra.n_state = n2; ra.n_cand = nm + 1;
    if (kVariant == 3) { ra.n2 = n1 /* bin_of / matches12 are indexed by frame-1 key points */; ra.has_mp2 = f2; }
    else { ra.n2 = n2; ra.occupied = f2; }
    if ((rc = launch_resolve<kVariant>(h, ra, n2, st))) return rc;
    h->launches += 2;
    ORBFE_CUDA(h, cudaGetLastError());
    if (kVariant == 3) ORBFE_CUDA(h, cudaMemcpyAsync(out, m12, sizeof(int) * n1, cudaMemcpyDeviceToHost, st));
    else ORBFE_CUDA(h, cudaMemcpyAsync(out, asg, sizeof(int) * n2, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(n_matches, nm, sizeof(int), cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaStreamSynchronize(st));
    return ORBFE_OK;
}

extern "C" {

int orbfe_popc_peak(orbfe_handle *h, double *gpopc_per_s) {
    if (!h || !gpopc_per_s) return ORBFE_E_ARG;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_match_scratch(h, 4096);
    if (rc) return rc;
    cudaStream_t st = h->stream;
    cudaEvent_t e0, e1;
    ORBFE_CUDA(h, cudaEventCreate(&e0)); ORBFE_CUDA(h, cudaEventCreate(&e1));
    const int blocks = h->sm_count * 8, iters = 4096;
    k_popc_peak<<<blocks, 256, 0, st>>>((unsigned *) h->d_match, 12345u, 64);       // warm-up
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        ORBFE_CUDA(h, cudaEventRecord(e0, st));
        k_popc_peak<<<blocks, 256, 0, st>>>((unsigned *) h->d_match, 12345u + rep, iters);
        ORBFE_CUDA(h, cudaEventRecord(e1, st));
        ORBFE_CUDA(h, cudaEventSynchronize(e1));
        float ms = 0;
        ORBFE_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        best = std::max(best, (double) blocks * 256 * 8 * iters / (ms * 1e-3) / 1e9);
    }
    h->launches += 6;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *gpopc_per_s = best;
    return ORBFE_OK;
}

int orbfe_imma_peak(orbfe_handle *h, double *gmatch_per_s) {
    if (!h || !gmatch_per_s) return ORBFE_E_ARG;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_match_scratch(h, 4096);
    if (rc) return rc;
    cudaStream_t st = h->stream;
    cudaEvent_t e0, e1;
    ORBFE_CUDA(h, cudaEventCreate(&e0)); ORBFE_CUDA(h, cudaEventCreate(&e1));
    const int blocks = h->sm_count * 8, iters = 4096;
    k_imma_peak<<<blocks, 256, 0, st>>>((int *) h->d_match, 64);                     // warm-up
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        ORBFE_CUDA(h, cudaEventRecord(e0, st));
        k_imma_peak<<<blocks, 256, 0, st>>>((int *) h->d_match, iters);
        ORBFE_CUDA(h, cudaEventRecord(e1, st));
        ORBFE_CUDA(h, cudaEventSynchronize(e1));
        float ms = 0;
        ORBFE_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        // warps x iterations x 4 instructions, each 1/8 of a 16 x 8 tile of 256-bit pairs
        best = std::max(best, (double) blocks * 8 * iters * 4 * (128.0 / 8.0) / (ms * 1e-3) / 1e9);
    }
    h->launches += 6;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *gmatch_per_s = best;
    return ORBFE_OK;
}

int orbfe_descriptor_distance(orbfe_handle *h, const uint8_t *a, int na, const uint8_t *b, int nb, const int32_t *ia, const int32_t *ib, int n_pairs, int32_t *dist) {
    if (!h) return ORBFE_E_ARG;
    if (!a || !b || !ia || !ib || !dist || na < 0 || nb < 0 || n_pairs < 0) return set_error(h, ORBFE_E_ARG, "invalid argument");
    if (n_pairs == 0) return ORBFE_OK;
    for (int i = 0; i < n_pairs; ++i) if (ia[i] < 0 || ia[i] >= na || ib[i] < 0 || ib[i] >= nb) return set_error(h, ORBFE_E_ARG, "pair %d out of range", i);
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    Bump probe{nullptr};
    probe.take<uint4>(2 * (size_t) na); probe.take<uint4>(2 * (size_t) nb); probe.take<int>(n_pairs); probe.take<int>(n_pairs); probe.take<int>(n_pairs);
    int rc = ensure_match_scratch(h, probe.off + 1024);
    if (rc) return rc;
    Bump bp{(uint8_t *) h->d_match};
    uint4 *da = bp.take<uint4>(2 * (size_t) na), *db = bp.take<uint4>(2 * (size_t) nb);
    int *dia = bp.take<int>(n_pairs), *dib = bp.take<int>(n_pairs), *dd = bp.take<int>(n_pairs);
    ORBFE_CUDA(h, cudaMemcpyAsync(da, a, 32 * (size_t) na, cudaMemcpyHostToDevice, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(db, b, 32 * (size_t) nb, cudaMemcpyHostToDevice, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(dia, ia, sizeof(int) * n_pairs, cudaMemcpyHostToDevice, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(dib, ib, sizeof(int) * n_pairs, cudaMemcpyHostToDevice, st));
    k_pair_distance<<<(n_pairs + 255) / 256, 256, 0, st>>>(da, db, dia, dib, n_pairs, dd);
    h->launches++;
    ORBFE_CUDA(h, cudaGetLastError());
    ORBFE_CUDA(h, cudaMemcpyAsync(dist, dd, sizeof(int) * n_pairs, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaStreamSynchronize(st));
    return ORBFE_OK;
}

// Kernel choice of the brute-force search.  Problems of at least 256 x 512 run on the tensor cores — tcgen05 (k_allpairs_tc) by default,
// the legacy warp-level IMMA kernel with ORBFE_ALLPAIRS=imma — smaller ones and ORBFE_ALLPAIRS=popc (or ORBFE_ALLPAIRS_POPC=1) on the
// popc kernel.  The environment is read per call: bench.py and the tests time / compare the kernels in one process.
static int allpairs_dispatch(Handle *h, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt, const int2 *d_excl, int32_t *d_best_idx, int32_t *d_best_dist,
                             int32_t *d_second_dist, cudaStream_t st) {
    const char *popc_env = getenv("ORBFE_ALLPAIRS_POPC"), *kern_env = getenv("ORBFE_ALLPAIRS");
    int kern = 2;                                                      // 0 popc, 1 mma.sync IMMA, 2 tcgen05
    if (kern_env && !strcmp(kern_env, "popc")) kern = 0; else if (kern_env && !strcmp(kern_env, "imma")) kern = 1;
    if (popc_env && *popc_env == '1') kern = 0;
    if (nq < 2 * kImM || nt < 4 * kImN) kern = 0;
    if (kern == 1 && d_excl) kern = 2;                                 // the IMMA kernel has no exclusion path
    auto ensure_ap = [&](size_t need) -> int {
        if (h->ap_bytes >= need) return ORBFE_OK;
        ORBFE_CUDA(h, cudaDeviceSynchronize());
        cudaFree(h->d_ap); h->d_ap = nullptr; h->ap_bytes = 0;
        ORBFE_CUDA(h, cudaMalloc(&h->d_ap, need));
        h->ap_bytes = need;
        return ORBFE_OK;
    };
    if (kern == 2) {
        int rc = ensure_ap(allpairs_tc_scratch_bytes(nq, nt, nullptr));
        if (rc) return rc;
        uint2 *partial = nullptr; int n_split = 1;
        if ((rc = allpairs_tc_launch(h, d_q, nq, d_t, nt, d_excl, (uint8_t *) h->d_ap, &partial, &n_split, st))) return rc;
        k_allpairs_merge<<<(nq + 255) / 256, 256, 0, st>>>(partial, n_split, nq, d_best_idx, d_best_dist, d_second_dist);
        h->launches++;
    } else if (kern == 1) {
        const int n_mt = (nq + kImM - 1) / kImM, n_tiles = (nt + kImN - 1) / kImN;
        const int slots = std::max(h->sm_count, 1) * 3;                      // resident CTAs (about 160 registers x 128 threads, 37 KB)
        int n_split = 1; double best_eff = 0;
        for (int s = 1; s <= 32; s *= 2) {
            if (s > 1 && (n_tiles + s - 1) / s < 4) break;
            const long long total = (long long) n_mt * s;
            const double eff = (double) total / (double) (((total + slots - 1) / slots) * slots);
            if (eff > best_eff + 0.02) { best_eff = eff; n_split = s; }
        }
        const int tiles_per_split = (n_tiles + n_split - 1) / n_split;
        n_split = (n_tiles + tiles_per_split - 1) / tiles_per_split;
        const size_t qe_bytes = ((size_t) nq * kImStride + 255) & ~(size_t) 255, te_bytes = ((size_t) nt * kImStride + 255) & ~(size_t) 255;
        int rc = ensure_ap(qe_bytes + te_bytes + (size_t) n_split * nq * sizeof(uint2) + 256);
        if (rc) return rc;
        uint8_t *qe = (uint8_t *) h->d_ap, *te = qe + qe_bytes;
        uint2 *partial = reinterpret_cast<uint2 *>(te + te_bytes);
        k_expand_pm1<<<(nq * 36 + 255) / 256, 256, 0, st>>>(d_q, nq, qe);
        k_expand_pm1<<<(nt * 36 + 255) / 256, 256, 0, st>>>(d_t, nt, te);
        k_allpairs_imma<<<dim3(n_mt, n_split), 128, 0, st>>>(qe, nq, te, nt, tiles_per_split, partial);
        k_allpairs_merge<<<(nq + 255) / 256, 256, 0, st>>>(partial, n_split, nq, d_best_idx, d_best_dist, d_second_dist);
        h->launches += 4;
    } else {
        if (d_excl) k_hamming_allpairs<true><<<(nq + kApQ - 1) / kApQ, 256, 0, st>>>((const uint4 *) d_q, nq, (const uint4 *) d_t, nt, d_excl, d_best_idx, d_best_dist, d_second_dist);
        else k_hamming_allpairs<false><<<(nq + kApQ - 1) / kApQ, 256, 0, st>>>((const uint4 *) d_q, nq, (const uint4 *) d_t, nt, nullptr, d_best_idx, d_best_dist, d_second_dist);
        h->launches++;
    }
    ORBFE_CUDA(h, cudaGetLastError());
    return ORBFE_OK;
}

int orbfe_hamming_allpairs_excl_device(orbfe_handle *h, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt, const int32_t *d_excl, int32_t *d_best_idx,
                                       int32_t *d_best_dist, int32_t *d_second_dist, void *stream, int sync) {
    if (!h) return ORBFE_E_ARG;
    if (nq < 0 || nt < 0 || (nq && (!d_q || !d_best_idx || !d_best_dist || !d_second_dist)) || (nt && !d_t)) return set_error(h, ORBFE_E_ARG, "invalid argument");
    if (nt >= (1 << 22)) return set_error(h, ORBFE_E_ARG, "at most %d train descriptors per call", (1 << 22) - 1);
    if (((uintptr_t) d_q | (uintptr_t) d_t) & 15) return set_error(h, ORBFE_E_ARG, "descriptor arrays must be 16-byte aligned");
    if ((uintptr_t) d_excl & 7) return set_error(h, ORBFE_E_ARG, "exclusion ranges must be 8-byte aligned");
    if (nq == 0) return ORBFE_OK;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t) stream : h->stream;
    int rc = allpairs_dispatch(h, d_q, nq, d_t, nt, reinterpret_cast<const int2 *>(d_excl), d_best_idx, d_best_dist, d_second_dist, st);
    if (rc) return rc;
    if (sync) ORBFE_CUDA(h, cudaStreamSynchronize(st));
    return ORBFE_OK;
}

// Key-frame window straight from the extractor's output slabs: no compaction, no host synchronisation for the counts.
int orbfe_hamming_allpairs_slab_device(orbfe_handle *h, const uint8_t *d_desc, const int *d_n_per_frame, int n_frames, int cap, int32_t *d_best_idx,
                                       int32_t *d_best_dist, int32_t *d_second_dist, void *stream, int sync) {
    if (!h) return ORBFE_E_ARG;
    if (n_frames < 0 || cap < 1 || (n_frames && (!d_desc || !d_n_per_frame || !d_best_idx || !d_best_dist || !d_second_dist)))
        return set_error(h, ORBFE_E_ARG, "orbfe_hamming_allpairs_slab_device: invalid argument");
    if ((long long) n_frames * cap >= (1 << 22)) return set_error(h, ORBFE_E_ARG, "at most %d descriptor rows per window", (1 << 22) - 1);
    if ((uintptr_t) d_desc & 15) return set_error(h, ORBFE_E_ARG, "descriptor slab must be 16-byte aligned");
    if (n_frames == 0) return ORBFE_OK;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t) stream : h->stream;
    const int n = n_frames * cap;
    const size_t need = allpairs_tc_scratch_bytes(n, n, nullptr);
    if (h->ap_bytes < need) {
        ORBFE_CUDA(h, cudaDeviceSynchronize());
        cudaFree(h->d_ap); h->d_ap = nullptr; h->ap_bytes = 0;
        ORBFE_CUDA(h, cudaMalloc(&h->d_ap, need));
        h->ap_bytes = need;
    }
    uint2 *partial = nullptr; int n_split = 1;
    int rc = allpairs_tc_launch(h, d_desc, n, d_desc, n, nullptr, (uint8_t *) h->d_ap, &partial, &n_split, st, d_n_per_frame, cap);
    if (rc) return rc;
    k_allpairs_merge<<<(n + 255) / 256, 256, 0, st>>>(partial, n_split, n, d_best_idx, d_best_dist, d_second_dist);
    h->launches++;
    ORBFE_CUDA(h, cudaGetLastError());
    if (sync) ORBFE_CUDA(h, cudaStreamSynchronize(st));
    return ORBFE_OK;
}

int orbfe_hamming_allpairs_device(orbfe_handle *h, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt, int32_t *d_best_idx, int32_t *d_best_dist,
                                  int32_t *d_second_dist, void *stream, int sync) {
    return orbfe_hamming_allpairs_excl_device(h, d_q, nq, d_t, nt, nullptr, d_best_idx, d_best_dist, d_second_dist, stream, sync);
}

int orbfe_hamming_allpairs_excl(orbfe_handle *h, const uint8_t *q, int nq, const uint8_t *t, int nt, const int32_t *excl, int32_t *best_idx, int32_t *best_dist,
                                int32_t *second_dist) {
    if (!h) return ORBFE_E_ARG;
    if (nq < 0 || nt < 0 || (nq && (!q || !best_idx || !best_dist || !second_dist)) || (nt && !t)) return set_error(h, ORBFE_E_ARG, "invalid argument");
    if (nq == 0) return ORBFE_OK;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    Bump probe{nullptr};
    probe.take<uint4>(2 * (size_t) nq); probe.take<uint4>(2 * (size_t) std::max(nt, 1)); probe.take<int>(nq); probe.take<int>(nq); probe.take<int>(nq); probe.take<int2>(nq);
    int rc = ensure_match_scratch(h, probe.off + 1024);
    if (rc) return rc;
    Bump bp{(uint8_t *) h->d_match};
    uint4 *dq = bp.take<uint4>(2 * (size_t) nq), *dt = bp.take<uint4>(2 * (size_t) std::max(nt, 1));
    int *bi = bp.take<int>(nq), *bd = bp.take<int>(nq), *sd = bp.take<int>(nq);
    int2 *dex = bp.take<int2>(nq);
    ORBFE_CUDA(h, cudaMemcpyAsync(dq, q, 32 * (size_t) nq, cudaMemcpyHostToDevice, st));
    if (nt) ORBFE_CUDA(h, cudaMemcpyAsync(dt, t, 32 * (size_t) nt, cudaMemcpyHostToDevice, st));
    if (excl) ORBFE_CUDA(h, cudaMemcpyAsync(dex, excl, sizeof(int2) * (size_t) nq, cudaMemcpyHostToDevice, st));
    rc = orbfe_hamming_allpairs_excl_device(h, (const uint8_t *) dq, nq, (const uint8_t *) dt, nt, excl ? (const int32_t *) dex : nullptr, bi, bd, sd, st, 0);
    if (rc) return rc;
    ORBFE_CUDA(h, cudaMemcpyAsync(best_idx, bi, sizeof(int) * nq, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(best_dist, bd, sizeof(int) * nq, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(second_dist, sd, sizeof(int) * nq, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaStreamSynchronize(st));
    return ORBFE_OK;
}

int orbfe_hamming_allpairs(orbfe_handle *h, const uint8_t *q, int nq, const uint8_t *t, int nt, int32_t *best_idx, int32_t *best_dist, int32_t *second_dist) {
    return orbfe_hamming_allpairs_excl(h, q, nq, t, nt, nullptr, best_idx, best_dist, second_dist);
}

int orbfe_hamming_window(orbfe_handle *h, const uint8_t *q, int nq, const uint8_t *t, int nt, const int32_t *cand_offsets, const int32_t *cand_idx,
                         int32_t *best_idx, int32_t *best_dist, int32_t *second_dist) {
    if (!h) return ORBFE_E_ARG;
    if (nq < 0 || nt < 0 || (nq && (!q || !cand_offsets || !best_idx || !best_dist || !second_dist)) || (nt && !t)) return set_error(h, ORBFE_E_ARG, "invalid argument");
    if (nq == 0) return ORBFE_OK;
    const int total = cand_offsets[nq];
    if (cand_offsets[0] != 0 || total < 0 || (total && !cand_idx)) return set_error(h, ORBFE_E_ARG, "cand_offsets must start at 0 and be non-decreasing");
    for (int i = 0; i < nq; ++i) {
        if (cand_offsets[i + 1] < cand_offsets[i]) return set_error(h, ORBFE_E_ARG, "cand_offsets must start at 0 and be non-decreasing");
        if (cand_offsets[i + 1] - cand_offsets[i] >= (1 << 22)) return set_error(h, ORBFE_E_ARG, "query %d has 2^22 or more candidates", i);
    }
    for (int i = 0; i < total; ++i) if (cand_idx[i] < 0 || cand_idx[i] >= nt) return set_error(h, ORBFE_E_ARG, "candidate %d out of range", i);
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    Bump probe{nullptr};
    probe.take<uint4>(2 * (size_t) nq); probe.take<uint4>(2 * (size_t) std::max(nt, 1)); probe.take<int>(nq + 1); probe.take<int>(std::max(total, 1));
    probe.take<int>(3 * (size_t) nq);
    int rc = ensure_match_scratch(h, probe.off + 1024);
    if (rc) return rc;
    Bump bp{(uint8_t *) h->d_match};
    uint4 *dq = bp.take<uint4>(2 * (size_t) nq), *dt = bp.take<uint4>(2 * (size_t) std::max(nt, 1));
    int *doff = bp.take<int>(nq + 1), *didx = bp.take<int>(std::max(total, 1)), *res = bp.take<int>(3 * (size_t) nq);
    ORBFE_CUDA(h, cudaMemcpyAsync(dq, q, 32 * (size_t) nq, cudaMemcpyHostToDevice, st));
    if (nt) ORBFE_CUDA(h, cudaMemcpyAsync(dt, t, 32 * (size_t) nt, cudaMemcpyHostToDevice, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(doff, cand_offsets, sizeof(int) * ((size_t) nq + 1), cudaMemcpyHostToDevice, st));
    if (total) ORBFE_CUDA(h, cudaMemcpyAsync(didx, cand_idx, sizeof(int) * (size_t) total, cudaMemcpyHostToDevice, st));
    k_hamming_window<<<(nq * 8 + 255) / 256, 256, 0, st>>>(dq, nq, dt, doff, didx, res, res + nq, res + 2 * (size_t) nq);
    h->launches++;
    ORBFE_CUDA(h, cudaGetLastError());
    ORBFE_CUDA(h, cudaMemcpyAsync(best_idx, res, sizeof(int) * nq, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(best_dist, res + nq, sizeof(int) * nq, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(second_dist, res + 2 * (size_t) nq, sizeof(int) * nq, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaStreamSynchronize(st));
    return ORBFE_OK;
}

int orbfe_search_for_initialization(orbfe_handle *h, const orbfe_keypoint *kps1, const uint8_t *desc1, int n1, const orbfe_keypoint *kps2,
                                    const uint8_t *desc2, int n2, int img_w, int img_h, float *prematched_xy, int32_t *matches12, int window,
                                    float nn_ratio, int check_orientation, int *n_matches) {
    if (!h) return ORBFE_E_ARG;
    if (!n_matches || n1 < 0 || n2 < 0 || (n1 && (!kps1 || !desc1 || !prematched_xy || !matches12)) || (n2 && (!kps2 || !desc2)) || img_w <= 0 || img_h <= 0)
        return set_error(h, ORBFE_E_ARG, "invalid argument");
    *n_matches = 0;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (n1 == 0 || n2 == 0) return ORBFE_OK;
    // the adapter's part of the loop head (ORBMatcher.cpp:46-52): level-0 key points only, window centre = vecPreMatched
    std::vector<float> qx(n1), qy(n1), qr(n1, (float) window), qa(n1);
    std::vector<int> qmin(n1), qmax(n1); std::vector<uint8_t> qv(n1);
    for (int i = 0; i < n1; ++i) {
        qx[i] = prematched_xy[2 * i]; qy[i] = prematched_xy[2 * i + 1]; qa[i] = kps1[i].angle;
        qmin[i] = qmax[i] = kps1[i].octave; qv[i] = kps1[i].octave <= 0;
    }
    WindowProblem p{qx.data(), qy.data(), qr.data(), qmin.data(), qmax.data(), qv.data(), desc1, qa.data(), n1, kps2, desc2, n2, img_w, img_h, nullptr};
    return run_window_search<0>(h, p, nn_ratio, check_orientation, nullptr, matches12, prematched_xy, n_matches);
}

// ---------------------------------------------------------------- device-resident frames
int orbfe_frame_upload(orbfe_handle *h, const orbfe_keypoint *kps, const uint8_t *desc, int n, int img_w, int img_h, orbfe_frame **out) {
    if (!h) return ORBFE_E_ARG;
    if (!out || n < 0 || (n && (!kps || !desc)) || img_w <= 0 || img_h <= 0) return set_error(h, ORBFE_E_ARG, "orbfe_frame_upload: invalid argument");
    *out = nullptr;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    int cols, rows; grid_dims(img_w, img_h, cols, rows);
    Bump b{nullptr};
    auto layout = [&](Bump &bp, orbfe_keypoint *&dk, uint8_t *&dd, int *&off, int *&idx, int *&cnt) {
        dk = bp.take<orbfe_keypoint>(std::max(n, 1)); dd = bp.take<uint8_t>(32 * (size_t) std::max(n, 1)); off = bp.take<int>((size_t) cols * rows + 1);
        idx = bp.take<int>(std::max(n, 1)); cnt = bp.take<int>(4);
    };
    orbfe_keypoint *dk; uint8_t *dd; int *off, *idx, *cnt;
    layout(b, dk, dd, off, idx, cnt);
    void *mem = nullptr;
    ORBFE_CUDA(h, cudaMalloc(&mem, b.off + 256));
    Bump bp{(uint8_t *) mem}; layout(bp, dk, dd, off, idx, cnt);
    cudaStream_t st = h->stream;
    cudaError_t e = cudaSuccess;
    if (n) { e = cudaMemcpyAsync(dk, kps, sizeof(orbfe_keypoint) * (size_t) n, cudaMemcpyHostToDevice, st); if (e == cudaSuccess) e = cudaMemcpyAsync(dd, desc, 32 * (size_t) n, cudaMemcpyHostToDevice, st); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(cnt, &n, sizeof(int), cudaMemcpyHostToDevice, st);
    int rc = e == cudaSuccess ? frame_grid_launch(h, dk, cnt, std::max(n, 1), img_w, img_h, off, idx, st) : set_error(h, ORBFE_E_CUDA, "orbfe_frame_upload: %s", cudaGetErrorString(e));
    if (rc == ORBFE_OK && (e = cudaStreamSynchronize(st)) != cudaSuccess) rc = set_error(h, ORBFE_E_CUDA, "orbfe_frame_upload: %s", cudaGetErrorString(e));
    if (rc != ORBFE_OK) { cudaFree(mem); return rc; }
    orbfe_frame *f = new orbfe_frame();
    f->device = h->device; f->n = n; f->img_w = img_w; f->img_h = img_h; f->cols = cols; f->rows = rows;
    f->d_kps = dk; f->d_desc = dd; f->d_grid_off = off; f->d_grid_idx = idx; f->owned = mem;
    f->kps.assign(kps, kps + n);
    *out = f;
    return ORBFE_OK;
}

int orbfe_frame_wrap_device(orbfe_handle *h, const orbfe_keypoint *d_kps, const uint8_t *d_desc, int n, int img_w, int img_h, const int32_t *d_grid_off,
                            const int32_t *d_grid_idx, orbfe_frame **out) {
    if (!h) return ORBFE_E_ARG;
    if (!out || n < 0 || (n && (!d_kps || !d_desc)) || img_w <= 0 || img_h <= 0 || (!d_grid_off) != (!d_grid_idx) || ((uintptr_t) d_desc & 15))
        return set_error(h, ORBFE_E_ARG, "orbfe_frame_wrap_device: invalid argument (descriptors must be 16-byte aligned; grid offsets and indices come together)");
    *out = nullptr;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    int cols, rows; grid_dims(img_w, img_h, cols, rows);
    orbfe_frame *f = new orbfe_frame();
    f->device = h->device; f->n = n; f->img_w = img_w; f->img_h = img_h; f->cols = cols; f->rows = rows;
    f->d_kps = d_kps; f->d_desc = d_desc; f->d_grid_off = d_grid_off; f->d_grid_idx = d_grid_idx;
    f->kps.resize((size_t) n);
    cudaStream_t st = h->stream;
    cudaError_t e = n ? cudaMemcpyAsync(f->kps.data(), d_kps, sizeof(orbfe_keypoint) * (size_t) n, cudaMemcpyDeviceToHost, st) : cudaSuccess;
    int rc = e == cudaSuccess ? ORBFE_OK : set_error(h, ORBFE_E_CUDA, "orbfe_frame_wrap_device: %s", cudaGetErrorString(e));
    if (rc == ORBFE_OK && !d_grid_off) {            // no grid given: build one into buffers this object owns
        Bump b{nullptr}; b.take<int>((size_t) cols * rows + 1); b.take<int>(std::max(n, 1)); b.take<int>(4);
        void *mem = nullptr;
        if ((e = cudaMalloc(&mem, b.off + 256)) != cudaSuccess) rc = set_error(h, ORBFE_E_CUDA, "orbfe_frame_wrap_device: %s", cudaGetErrorString(e));
        else {
            Bump bp{(uint8_t *) mem};
            int *off = bp.take<int>((size_t) cols * rows + 1), *idx = bp.take<int>(std::max(n, 1)), *cnt = bp.take<int>(4);
            f->owned = mem; f->d_grid_off = off; f->d_grid_idx = idx;
            if ((e = cudaMemcpyAsync(cnt, &n, sizeof(int), cudaMemcpyHostToDevice, st)) != cudaSuccess) rc = set_error(h, ORBFE_E_CUDA, "orbfe_frame_wrap_device: %s", cudaGetErrorString(e));
            else rc = frame_grid_launch(h, d_kps, cnt, std::max(n, 1), img_w, img_h, off, idx, st);
        }
    }
    if (rc == ORBFE_OK && (e = cudaStreamSynchronize(st)) != cudaSuccess) rc = set_error(h, ORBFE_E_CUDA, "orbfe_frame_wrap_device: %s", cudaGetErrorString(e));
    if (rc != ORBFE_OK) { cudaFree(f->owned); delete f; return rc; }
    *out = f;
    return ORBFE_OK;
}

void orbfe_frame_destroy(orbfe_frame *f) {
    if (!f) return;
    if (f->owned) { cudaSetDevice(f->device); cudaFree(f->owned); }
    delete f;
}

int orbfe_frame_size(const orbfe_frame *f) { return f ? f->n : -1; }

static int check_frame(orbfe_handle *h, const orbfe_frame *f, const char *what) {
    if (!f) return set_error(h, ORBFE_E_ARG, "%s: frame is null", what);
    if (f->device != h->device) return set_error(h, ORBFE_E_ARG, "%s: the frame lives on device %d, the handle on device %d", what, f->device, h->device);
    return ORBFE_OK;
}

int orbfe_search_for_initialization_f(orbfe_handle *h, const orbfe_frame *frame1, const orbfe_frame *frame2, float *prematched_xy, int32_t *matches12,
                                      int window, float nn_ratio, int check_orientation, int *n_matches) {
    if (!h) return ORBFE_E_ARG;
    int rc;
    if ((rc = check_frame(h, frame1, "orbfe_search_for_initialization_f")) || (rc = check_frame(h, frame2, "orbfe_search_for_initialization_f"))) return rc;
    const int n1 = frame1->n, n2 = frame2->n;
    if (!n_matches || (n1 && (!prematched_xy || !matches12))) return set_error(h, ORBFE_E_ARG, "invalid argument");
    *n_matches = 0;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (n1 == 0 || n2 == 0) return ORBFE_OK;
    std::vector<float> qx(n1), qy(n1), qr(n1, (float) window), qa(n1);
    std::vector<int> qmin(n1), qmax(n1); std::vector<uint8_t> qv(n1);
    const orbfe_keypoint *kps1 = frame1->kps.data();
    for (int i = 0; i < n1; ++i) {                               // ORBMatcher.cpp:46-52
        qx[i] = prematched_xy[2 * i]; qy[i] = prematched_xy[2 * i + 1]; qa[i] = kps1[i].angle;
        qmin[i] = qmax[i] = kps1[i].octave; qv[i] = kps1[i].octave <= 0;
    }
    WindowProblem p{qx.data(), qy.data(), qr.data(), qmin.data(), qmax.data(), qv.data(), nullptr, qa.data(), n1, nullptr, nullptr, n2, frame2->img_w, frame2->img_h, nullptr};
    p.frame1 = frame1; p.frame2 = frame2;
    return run_window_search<0>(h, p, nn_ratio, check_orientation, nullptr, matches12, prematched_xy, n_matches);
}

int orbfe_search_by_projection_f(orbfe_handle *h, const float *q_u, const float *q_v, const float *q_radius, const int32_t *q_level, const float *q_angle,
                                 const uint8_t *q_desc, const uint8_t *q_valid, int nq, const orbfe_frame *frame2, const uint8_t *occupied, int32_t *assigned,
                                 int check_orientation, int *n_matches) {
    if (!h) return ORBFE_E_ARG;
    int rc;
    if ((rc = check_frame(h, frame2, "orbfe_search_by_projection_f"))) return rc;
    const int n2 = frame2->n;
    if (!n_matches || nq < 0 || (nq && (!q_u || !q_v || !q_radius || !q_level || !q_angle || !q_desc || !q_valid)) || (n2 && !assigned))
        return set_error(h, ORBFE_E_ARG, "invalid argument");
    *n_matches = 0;
    for (int j = 0; j < n2; ++j) assigned[j] = -1;
    if (nq == 0 || n2 == 0) return ORBFE_OK;
    std::vector<int> qmin(nq), qmax(nq);
    for (int i = 0; i < nq; ++i) { qmin[i] = q_level[i] - 1; qmax[i] = q_level[i] + 1; }             // ORBMatcher.cpp:229
    WindowProblem p{q_u, q_v, q_radius, qmin.data(), qmax.data(), q_valid, q_desc, q_angle, nq, nullptr, nullptr, n2, frame2->img_w, frame2->img_h, occupied};
    p.frame2 = frame2;
    return run_window_search<1>(h, p, 0.f, check_orientation, assigned, nullptr, nullptr, n_matches);
}

int orbfe_search_local_points_f(orbfe_handle *h, const float *q_u, const float *q_v, const float *q_radius, const int32_t *q_level, const uint8_t *q_desc,
                                const uint8_t *q_valid, int nq, const orbfe_frame *frame2, const uint8_t *occupied, int32_t *assigned, float nn_ratio,
                                int *n_matches) {
    if (!h) return ORBFE_E_ARG;
    int rc;
    if ((rc = check_frame(h, frame2, "orbfe_search_local_points_f"))) return rc;
    const int n2 = frame2->n;
    if (!n_matches || nq < 0 || (nq && (!q_u || !q_v || !q_radius || !q_level || !q_desc || !q_valid)) || (n2 && !assigned)) return set_error(h, ORBFE_E_ARG, "invalid argument");
    *n_matches = 0;
    for (int j = 0; j < n2; ++j) assigned[j] = -1;
    if (nq == 0 || n2 == 0) return ORBFE_OK;
    std::vector<int> qmin(nq), qmax(nq);
    for (int i = 0; i < nq; ++i) { qmin[i] = q_level[i] - 1; qmax[i] = q_level[i]; }                 // ORBMatcher.cpp:366-369
    WindowProblem p{q_u, q_v, q_radius, qmin.data(), qmax.data(), q_valid, q_desc, nullptr, nq, nullptr, nullptr, n2, frame2->img_w, frame2->img_h, occupied};
    p.frame2 = frame2;
    return run_window_search<2>(h, p, nn_ratio, 0, assigned, nullptr, nullptr, n_matches);
}

int orbfe_search_by_projection(orbfe_handle *h, const float *q_u, const float *q_v, const float *q_radius, const int32_t *q_level, const float *q_angle,
                               const uint8_t *q_desc, const uint8_t *q_valid, int nq, const orbfe_keypoint *kps2, const uint8_t *desc2, int n2,
                               int img_w, int img_h, const uint8_t *occupied, int32_t *assigned, int check_orientation, int *n_matches) {
    if (!h) return ORBFE_E_ARG;
    if (!n_matches || nq < 0 || n2 < 0 || (nq && (!q_u || !q_v || !q_radius || !q_level || !q_angle || !q_desc || !q_valid)) ||
        (n2 && (!kps2 || !desc2 || !assigned)) || img_w <= 0 || img_h <= 0)
        return set_error(h, ORBFE_E_ARG, "invalid argument");
    *n_matches = 0;
    for (int j = 0; j < n2; ++j) assigned[j] = -1;
    if (nq == 0 || n2 == 0) return ORBFE_OK;
    std::vector<int> qmin(nq), qmax(nq);
    for (int i = 0; i < nq; ++i) { qmin[i] = q_level[i] - 1; qmax[i] = q_level[i] + 1; }             // ORBMatcher.cpp:229
    WindowProblem p{q_u, q_v, q_radius, qmin.data(), qmax.data(), q_valid, q_desc, q_angle, nq, kps2, desc2, n2, img_w, img_h, occupied};
    return run_window_search<1>(h, p, 0.f, check_orientation, assigned, nullptr, nullptr, n_matches);
}

int orbfe_search_local_points(orbfe_handle *h, const float *q_u, const float *q_v, const float *q_radius, const int32_t *q_level, const uint8_t *q_desc,
                              const uint8_t *q_valid, int nq, const orbfe_keypoint *kps2, const uint8_t *desc2, int n2, int img_w, int img_h,
                              const uint8_t *occupied, int32_t *assigned, float nn_ratio, int *n_matches) {
    if (!h) return ORBFE_E_ARG;
    if (!n_matches || nq < 0 || n2 < 0 || (nq && (!q_u || !q_v || !q_radius || !q_level || !q_desc || !q_valid)) || (n2 && (!kps2 || !desc2 || !assigned)) ||
        img_w <= 0 || img_h <= 0)
        return set_error(h, ORBFE_E_ARG, "invalid argument");
    *n_matches = 0;
    for (int j = 0; j < n2; ++j) assigned[j] = -1;
    if (nq == 0 || n2 == 0) return ORBFE_OK;
    std::vector<int> qmin(nq), qmax(nq);
    for (int i = 0; i < nq; ++i) { qmin[i] = q_level[i] - 1; qmax[i] = q_level[i]; }                  // ORBMatcher.cpp:367-369
    WindowProblem p{q_u, q_v, q_radius, qmin.data(), qmax.data(), q_valid, q_desc, nullptr, nq, kps2, desc2, n2, img_w, img_h, occupied};
    return run_window_search<2>(h, p, nn_ratio, 0, assigned, nullptr, nullptr, n_matches);
}

int orbfe_search_fuse(orbfe_handle *h, const float *q_u, const float *q_v, const float *q_radius, const int32_t *q_level, const uint8_t *q_desc,
                      const uint8_t *q_valid, int nq, const orbfe_keypoint *kps1, const uint8_t *desc1, int n1, int img_w, int img_h,
                      int32_t *best_idx1, int32_t *best_dist, int *n_matches) {
    if (!h) return ORBFE_E_ARG;
    float s2[ORBFE_MAX_LEVELS];
    for (int l = 0; l < h->cfg.n_levels; ++l) s2[l] = h->scale[l] * h->scale[l];                 // square_sigmas (ORBExtractor.cpp:432-436)
    return orbfe_search_fuse_sigma(h, q_u, q_v, q_radius, q_level, q_desc, q_valid, nq, kps1, desc1, n1, img_w, img_h, s2, h->cfg.n_levels, best_idx1, best_dist,
                                   n_matches);
}

int orbfe_search_fuse_sigma(orbfe_handle *h, const float *q_u, const float *q_v, const float *q_radius, const int32_t *q_level, const uint8_t *q_desc,
                            const uint8_t *q_valid, int nq, const orbfe_keypoint *kps1, const uint8_t *desc1, int n1, int img_w, int img_h,
                            const float *square_sigmas, int n_levels, int32_t *best_idx1, int32_t *best_dist, int *n_matches) {
    if (!h) return ORBFE_E_ARG;
    if (!n_matches || nq < 0 || n1 < 0 || (nq && (!q_u || !q_v || !q_radius || !q_level || !q_desc || !q_valid || !best_idx1)) || (n1 && (!kps1 || !desc1)) ||
        img_w <= 0 || img_h <= 0 || !square_sigmas || n_levels < 1 || n_levels > ORBFE_MAX_LEVELS)
        return set_error(h, ORBFE_E_ARG, "invalid argument");
    // the chi-square gate indexes square_sigmas with the key point's octave (ORBMatcher.cpp:564): an octave outside the table is the caller's error
    for (int j = 0; j < n1; ++j)
        if (kps1[j].octave < 0 || kps1[j].octave >= n_levels)
            return set_error(h, ORBFE_E_ARG, "key point %d has octave %d but the sigma table has %d levels", j, kps1[j].octave, n_levels);
    *n_matches = 0;
    for (int i = 0; i < nq; ++i) { best_idx1[i] = -1; if (best_dist) best_dist[i] = TH_LOW + 1; }
    if (nq == 0 || n1 == 0) return ORBFE_OK;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    int cols, rows; grid_dims(img_w, img_h, cols, rows);
    const size_t n_off = (size_t) cols * rows + 1;
    float *qx, *qy, *qr; int *ql, *coff, *cidx, *bi, *bd, *nm; uint8_t *qv; uint4 *qd, *d1; orbfe_keypoint *k1;
    auto layout = [&](Bump &b) {
        nm = b.take<int>(4); qx = b.take<float>(nq); qy = b.take<float>(nq); qr = b.take<float>(nq); ql = b.take<int>(nq); qv = b.take<uint8_t>(nq);
        qd = b.take<uint4>(2 * (size_t) nq); k1 = b.take<orbfe_keypoint>(n1); d1 = b.take<uint4>(2 * (size_t) n1);
        coff = b.take<int>(n_off); cidx = b.take<int>(n1); bi = b.take<int>(nq); bd = b.take<int>(nq);
    };
    Bump probe{nullptr}; layout(probe);
    int rc = ensure_match_scratch(h, probe.off + 1024);
    if (rc) return rc;
    Bump b{(uint8_t *) h->d_match}; layout(b);
    const int hdr[4] = {0, n1, 0, 0};                            // nm[0] = match counter, nm[1] = key-point count for the grid kernel
#define UP(dst, src, bytes) ORBFE_CUDA(h, cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, st))
    UP(nm, hdr, sizeof hdr); UP(qx, q_u, sizeof(float) * nq); UP(qy, q_v, sizeof(float) * nq); UP(qr, q_radius, sizeof(float) * nq);
    UP(ql, q_level, sizeof(int) * nq); UP(qv, q_valid, nq); UP(qd, q_desc, 32 * (size_t) nq);
    UP(k1, kps1, sizeof(orbfe_keypoint) * (size_t) n1); UP(d1, desc1, 32 * (size_t) n1);
#undef UP
    if ((rc = frame_grid_launch(h, k1, nm + 1, n1, img_w, img_h, coff, cidx, st))) return rc;
    FuseArgs fa; memset(&fa, 0, sizeof fa);
    fa.qx = qx; fa.qy = qy; fa.qr = qr; fa.qlevel = ql; fa.qvalid = qv; fa.qdesc = qd; fa.nq = nq; fa.kps1 = k1; fa.desc1 = d1;
    fa.cell_off = coff; fa.cell_idx = cidx; fa.cols = cols; fa.rows = rows; fa.n_levels = n_levels;
    for (int l = 0; l < n_levels; ++l) fa.sigma2[l] = square_sigmas[l];
    fa.best_idx = bi; fa.best_dist = bd; fa.n_matches = nm;
    k_fuse<<<(nq + 7) / 8, 256, 0, st>>>(fa);
    h->launches++;
    ORBFE_CUDA(h, cudaGetLastError());
    ORBFE_CUDA(h, cudaMemcpyAsync(best_idx1, bi, sizeof(int) * nq, cudaMemcpyDeviceToHost, st));
    if (best_dist) ORBFE_CUDA(h, cudaMemcpyAsync(best_dist, bd, sizeof(int) * nq, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(n_matches, nm, sizeof(int), cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaStreamSynchronize(st));
    return ORBFE_OK;
}

int orbfe_compute_descriptors(orbfe_handle *h, const uint8_t *desc, const int32_t *group_off, int n_groups, int32_t *best) {
    if (!h) return ORBFE_E_ARG;
    if (n_groups < 0 || (n_groups && (!group_off || !best))) return set_error(h, ORBFE_E_ARG, "invalid argument");
    if (n_groups == 0) return ORBFE_OK;
    const int total = group_off[n_groups];
    if (group_off[0] != 0 || total < 0 || (total && !desc)) return set_error(h, ORBFE_E_ARG, "group offsets must start at 0 and be non-decreasing");
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    uint4 *dd; int *doff, *dbest, *derr;
    auto layout = [&](Bump &b) { dd = b.take<uint4>(2 * (size_t) std::max(total, 1)); doff = b.take<int>(n_groups + 1); dbest = b.take<int>(n_groups); derr = b.take<int>(4); };
    Bump probe{nullptr}; layout(probe);
    int rc = ensure_match_scratch(h, probe.off + 1024);
    if (rc) return rc;
    Bump b{(uint8_t *) h->d_match}; layout(b);
    if (total) ORBFE_CUDA(h, cudaMemcpyAsync(dd, desc, 32 * (size_t) total, cudaMemcpyHostToDevice, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(doff, group_off, sizeof(int) * (n_groups + 1), cudaMemcpyHostToDevice, st));
    ORBFE_CUDA(h, cudaMemsetAsync(derr, 0, sizeof(int), st));
    k_compute_descriptors<<<(n_groups + 7) / 8, 256, 0, st>>>(dd, doff, n_groups, dbest, derr);
    h->launches++;
    ORBFE_CUDA(h, cudaGetLastError());
    int e = 0;
    ORBFE_CUDA(h, cudaMemcpyAsync(best, dbest, sizeof(int) * n_groups, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(&e, derr, sizeof(int), cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaStreamSynchronize(st));
    if (e) return set_error(h, ORBFE_E_ARG, "orbfe_compute_descriptors: a map point has more than %d observations", kCdMaxObs);
    return ORBFE_OK;
}

int orbfe_search_for_triangulation(orbfe_handle *h, const uint8_t *desc1, const float *angle1, const uint8_t *has_mp1, int n1, const int32_t *node_id1,
                                   const int32_t *node_off1, const int32_t *node_idx1, int n_nodes1, const uint8_t *desc2, const float *angle2,
                                   const uint8_t *has_mp2, int n2, const int32_t *node_id2, const int32_t *node_off2, const int32_t *node_idx2,
                                   int n_nodes2, int32_t *matches12, int check_orientation, int *n_matches) {
    if (!h) return ORBFE_E_ARG;
    if (!n_matches || n1 < 0 || n2 < 0 || n_nodes1 < 0 || n_nodes2 < 0 || (n1 && (!desc1 || !angle1 || !has_mp1 || !matches12)) ||
        (n2 && (!desc2 || !angle2 || !has_mp2)) || (n_nodes1 && (!node_id1 || !node_off1 || !node_idx1)) || (n_nodes2 && (!node_id2 || !node_off2 || !node_idx2)))
        return set_error(h, ORBFE_E_ARG, "invalid argument");
    *n_matches = 0;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (n1 == 0 || n2 == 0 || n_nodes1 == 0 || n_nodes2 == 0) return ORBFE_OK;
    return run_node_search<3>(h, desc1, angle1, has_mp1, n1, node_id1, node_off1, node_idx1, n_nodes1, desc2, angle2, has_mp2, n2, node_id2, node_off2,
                              node_idx2, n_nodes2, matches12, 0.f, check_orientation, n_matches);
}

int orbfe_search_by_bow(orbfe_handle *h, const uint8_t *desc1, const float *angle1, const uint8_t *valid1, int n1, const int32_t *node_id1,
                        const int32_t *node_off1, const int32_t *node_idx1, int n_nodes1, const uint8_t *desc2, const float *angle2,
                        const uint8_t *occupied2, int n2, const int32_t *node_id2, const int32_t *node_off2, const int32_t *node_idx2, int n_nodes2,
                        int32_t *assigned, float nn_ratio, int check_orientation, int *n_matches) {
    if (!h) return ORBFE_E_ARG;
    if (!n_matches || n1 < 0 || n2 < 0 || n_nodes1 < 0 || n_nodes2 < 0 || (n1 && (!desc1 || !angle1 || !valid1)) ||
        (n2 && (!desc2 || !angle2 || !assigned)) || (n_nodes1 && (!node_id1 || !node_off1 || !node_idx1)) || (n_nodes2 && (!node_id2 || !node_off2 || !node_idx2)))
        return set_error(h, ORBFE_E_ARG, "invalid argument");
    *n_matches = 0;
    for (int j = 0; j < n2; ++j) assigned[j] = -1;
    if (n1 == 0 || n2 == 0 || n_nodes1 == 0 || n_nodes2 == 0) return ORBFE_OK;
    return run_node_search<4>(h, desc1, angle1, valid1, n1, node_id1, node_off1, node_idx1, n_nodes1, desc2, angle2, occupied2, n2, node_id2, node_off2,
                              node_idx2, n_nodes2, assigned, nn_ratio, check_orientation, n_matches);
}

}  // extern "C"
