// orbfe_internal.cuh — shared declarations of the B200-native ORB front-end (not part of the public C-ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/orbfe.h"

namespace orbfe {

constexpr int kEdge       = 19;   // EDGE_THRESHOLD   ORBExtractor.cpp:15
constexpr int kHalfPatch  = 15;   // HALF_PATCH_SIZE  ORBExtractor.cpp:14
constexpr int kCell       = 30;   // W                ORBExtractor.cpp:575
constexpr int kCellsPerBlk = 8;   // FAST: one warp per cell, 8 cells (240 px) per CTA
constexpr int kSlotCap    = 225;  // max NMS survivors of a 30x30 cell (one per 2x2 block)
constexpr int kBoxW       = 256;  // TMA box width in bytes (max box dimension)
constexpr int kBoxUsable  = kBoxW - 15;   // the box starts on a 16-byte boundary at or left of the wanted origin
constexpr int kFastBoxH   = 36;   // 30 + 2*3
constexpr int kBlurTileW  = 224;  // outputs per blur tile row (+6 halo +13 alignment slack <= 256)
constexpr int kBlurTileH  = 32;
constexpr int kBlurBoxH   = kBlurTileH + 6;
constexpr int kRsBoxH     = 80;   // source rows staged per resize tile
constexpr int kRsMaxTW    = 192;  // destination tile (chosen per level so the source span fits the box)
constexpr int kRsMaxTH    = 64;

// Per-level geometry, passed to kernels inside __grid_constant__ parameter blocks.
struct LevelGeom {
    int w, h, pitch;                 // level image size and row pitch (bytes, multiple of 64)
    int n_cols, n_rows, n_groups;    // FAST cells (30 px), groups of 8 cells per cell row
    int cell_base;                   // index of this level's first cell within a frame
    int fast_blk_base;               // first FAST block of this level (blocks = n_rows * n_groups)
    int blur_tx, blur_ty, blur_blk_base;
    int quota;                       // n_features_per_level
    int cand_off, cand_cap;          // per-frame candidate scratch (compact, reference order)
    int node_off, node_cap;          // per-frame quadtree node pool
    int kp_off, kp_cap;              // per-frame per-level selected key points
    int list_off;                    // per-frame offset of the E/S work lists (2 * kp_cap ints)
    float scale;                     // scale_factors[level]
    unsigned long long img_off;      // byte offset of frame 0 of this level inside the image / blur arenas
    unsigned long long frame_stride; // bytes between consecutive frames of this level
};

struct ResizeLevel {                 // destination level l (source = l-1)
    int tw, th, tiles_x, tiles_y;
    bool packed;                     // every 4-column quad reads a source span of at most 8 bytes (k_resize packed path)
    const int2 *xtab;                // per dst x: {sx0 | sx1<<16, a0 | a1<<16}
    const int2 *ytab;                // per dst y: {sy0 | sy1<<16, b0 | b1<<16}
};

struct Geometry {
    int w = 0, h = 0, n_levels = 0;
    LevelGeom lv[ORBFE_MAX_LEVELS];
    ResizeLevel rs[ORBFE_MAX_LEVELS];
    int cells_per_frame = 0, fast_blocks = 0, blur_blocks = 0;
    int cand_per_frame = 0, nodes_per_frame = 0, kp_per_frame = 0, lists_per_frame = 0;
    size_t img_bytes_per_frame = 0;  // sum over levels of pitch*h (levels 0..n-1)
    int oct_smem_bytes = 0;          // dynamic smem of the quadtree kernel
    int typ_node_cap = 0;            // largest per-level node count frames need in practice (sizes the shared-memory pool)
    int sort_cap = 0;                // power of two >= max kp_cap
};

struct Handle {
    orbfe_config cfg{};
    int device = 0;
    cudaStream_t stream = nullptr;
    bool use_tma = true, keep_stages = false;
    bool fast_fma_shift = false;     // ORBFE_FAST_FMA_SHIFT=1: FAST stage-A funnel shifts as IMAD + IMAD.HI on the FMA pipe instead of SHF (experiment, slower)
    bool fast_v1 = false;            // ORBFE_FAST_V1=1: k_fast (per-pixel pair tests) instead of k_fast_planes (difference planes), for A/B measurements
    bool fast_exact_cmp = false;     // ORBFE_FAST_EXACT=1: masked (exact) byte compares in FAST stage A, for A/B measurements
    float scale[ORBFE_MAX_LEVELS], inv_scale[ORBFE_MAX_LEVELS];
    int quota[ORBFE_MAX_LEVELS];
    int max_kp = 0;
    int sm_count = 148;
    long long launches = 0;

    // geometry-dependent state (rebuilt when the frame size changes)
    Geometry g;
    int batch_cap = 0;               // frames the arena holds
    uint8_t *d_img = nullptr;        // [level][frame][h][pitch]
    uint8_t *d_blur = nullptr;
    uint32_t *d_slots = nullptr;     // [frame][cell][kSlotCap] packed candidates
    int *d_cell_cnt = nullptr;       // [frame][cell]
    int *d_cell_off = nullptr;       // [frame][cell] exclusive prefix within the level
    uint32_t *d_cand = nullptr;      // [frame][cand_per_frame]
    int *d_cur = nullptr;            // [frame][cand_per_frame] current node of each candidate
    uint8_t *d_nodes = nullptr;      // [frame][nodes_per_frame] * 16 B (global fallback of the node pool)
    int *d_lists = nullptr;          // [frame][lists_per_frame]
    uint32_t *d_kp = nullptr;        // [frame][kp_per_frame] packed selected key points
    int *d_nkp = nullptr;            // [frame][ORBFE_MAX_LEVELS]
    int *d_ncand = nullptr;          // [frame][ORBFE_MAX_LEVELS]
    int *d_err = nullptr;            // device error flag
    void *d_tables = nullptr;        // resize tables
    int *d_fast_tab = nullptr;       // FAST block -> (level, cell row, strip)
    // staging for the host entry points
    orbfe_keypoint *d_out_kps = nullptr; uint8_t *d_out_desc = nullptr; int *d_out_n = nullptr; int out_cap = 0;
    int out_frames = 0;              // frames the output staging holds
    static constexpr int kStageSlots = 3;   // staging / output slots of the host pipeline (uploads run up to two chunks ahead of the passes)
    uint8_t *d_stage[kStageSlots] = {}; size_t stage_bytes = 0;          // dense H2D staging of the pipelined host path
    cudaStream_t s_up = nullptr, s_down = nullptr, s_aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_up[kStageSlots] = {}, ev_done[kStageSlots] = {}, ev_down[kStageSlots] = {};
    int last_batch = 0;              // frames of the last pass (for the stage getters)
    // state of the host pipeline across calls (orbfe_extract_batch_submit keeps batches in flight)
    long long pipe_seq = 0;          // chunks enqueued since creation: slot = seq % kStageSlots, arena = seq & 1
    bool pipe_pending = false, pipe_peer_used = false;
    int pipe_w = 0, pipe_h = 0, pipe_cap = 0, pipe_chunk = 0, pipe_cn = 0;
    static constexpr int kTickets = 8;
    long long pipe_ticket = 0;       // last ticket handed out
    cudaEvent_t ev_ticket[kTickets] = {};
    int *h_ticket_err = nullptr;     // pinned [kTickets][2]: overflow flags of both arenas as of the ticket's last download
    // Second arena + stream for the pipelined host path: odd chunks run their pass on the peer so that the tail of one pass
    // (quadtree, descriptors) overlaps with the head of the next (pyramid, FAST).  A complete handle, created lazily.
    orbfe_handle *peer = nullptr;
    cudaEvent_t ev_split[2] = {};    // fork / join of the two half-batch passes of orbfe_extract_batch_device
    // CUDA graph of the single-frame pass (orbfe_extract): captured once per (arena, output staging), replayed per frame
    cudaGraphExec_t graph1 = nullptr;
    const void *graph1_key[4] = {nullptr, nullptr, nullptr, nullptr};
    int graph1_cap = 0;
    long long graph1_launches = 0;
    // tensor maps (img arena; level 0 may be rebuilt for an in-place user buffer)
    CUtensorMap tm_fast[ORBFE_MAX_LEVELS], tm_blur[ORBFE_MAX_LEVELS], tm_rs[ORBFE_MAX_LEVELS];
    CUtensorMap tm_pimg[ORBFE_MAX_LEVELS], tm_pblur[ORBFE_MAX_LEVELS];   // per-key-point patch boxes of k_describe (image / blurred arena)
    float4 *d_pattern = nullptr;     // rotated-BRIEF pattern as floats, lane-major (DescArgs::pattern)
    // Fisheye uncertainty map (scale_mat) cached on the device by host pointer (orbfe_frame.cu)
    float *d_unc = nullptr; const float *unc_host = nullptr; size_t unc_bytes = 0;
    // matcher scratch
    void *d_match = nullptr; size_t match_bytes = 0;
    void *d_ap = nullptr; size_t ap_bytes = 0;        // all-pairs on the tensor cores: +-1 int8 rows of both sides + partial results
    void *h_pinned = nullptr; size_t pinned_bytes = 0;
    void *h_mpin = nullptr; size_t mpin_bytes = 0;      // pinned staging of the matcher's window searches

    // optional per-stage timing (orbfe_profile)
    bool prof = false, prof_pending = false;
    cudaEvent_t prof_ev[ORBFE_N_STAGES + 1] = {};
    double prof_ms[ORBFE_N_STAGES] = {};
    int prof_passes = 0;

    std::string err;
};

int set_error(Handle *h, int code, const char *fmt, ...);
int ensure_match_scratch(Handle *h, size_t bytes);
// per-device kernel attributes (dynamic shared-memory opt-ins) of the matcher and frame kernels; called by orbfe_create
int match_device_setup(Handle *h);
int frame_device_setup(Handle *h);
// all-pairs search on tcgen05 (orbfe_allpairs_tc.cu): scratch size for a problem, launch (partial results for k_allpairs_merge)
int allpairs_tc_device_setup(Handle *h);
size_t allpairs_tc_scratch_bytes(int nq, int nt, int *n_split_out);
int allpairs_tc_launch(Handle *h, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt, const int2 *d_excl, uint8_t *d_scratch,
                       uint2 **partial_out, int *n_split_out, cudaStream_t st, const int *d_slab_n = nullptr, int slab_cap = 0);
// Frame grid (CSR) of one device-resident key-point array; d_n holds the count (orbfe_frame.cu)
int frame_grid_launch(Handle *h, const orbfe_keypoint *d_kps, const int *d_n, int cap, int img_w, int img_h, int *d_grid_off, int *d_grid_idx, cudaStream_t st);

#define ORBFE_CUDA(h, call)                                                                     \
    do { cudaError_t e__ = (call);                                                              \
         if (e__ != cudaSuccess)                                                                \
             return orbfe::set_error((h), ORBFE_E_CUDA, "%s failed: %s (%s:%d)", #call,       \
                                     cudaGetErrorString(e__), __FILE__, __LINE__); } while (0)

}  // namespace orbfe

struct orbfe_handle : orbfe::Handle {};

// Device-resident frame (include/orbfe.h): what Frame::Frame leaves behind for the matchers.  Buffers are owned (orbfe_frame_upload)
// or borrowed from the caller (orbfe_frame_wrap_device); the host keeps a copy of the key points (octaves and angles drive the
// query lists of SearchForInitialization).
struct orbfe_frame {
    int device = 0, n = 0, img_w = 0, img_h = 0, cols = 0, rows = 0;
    const orbfe_keypoint *d_kps = nullptr; const uint8_t *d_desc = nullptr; const int *d_grid_off = nullptr, *d_grid_idx = nullptr;
    void *owned = nullptr;                       // one allocation behind every buffer this object owns
    std::vector<orbfe_keypoint> kps;             // host mirror
};
