// orbfe_extract.cu — host side of the extractor path: handle, geometry, tensor maps, launches, C-ABI.
// Replaces ORBExtractor (modules/ORB/ORBExtractor.{h,cpp}) behind include/orbfe.h.  No CPU fallback.
#include "orbfe_kernels.cuh"

#include <cudaTypedefs.h>
#include <chrono>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

namespace orbfe {

static thread_local std::string g_create_error;

int set_error(Handle *h, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

// ------------------------------------------------------------------------------------------------
// constructor tables — ORBExtractor::ORBExtractor, ORBExtractor.cpp:424-475 (float32 semantics as written there)
// ------------------------------------------------------------------------------------------------
static int round_half_even(float v) { return (int) lrintf(v); }          // cvRound
static int floor_f(float v) { int i = (int) v; return i - (v < (float) i); }
static int ceil_f(float v) { int i = (int) v; return i + (v > (float) i); }

// bit_pattern_31_ (ORBExtractor.cpp:108-365): 256 pairs (x0, y0, x1, y1)
static const int8_t kBriefPattern[1024] = {
#include "brief_pattern.inc"
};

static void build_ctor_tables(Handle *h) {
    const orbfe_config &c = h->cfg;
    h->scale[0] = 1.f; h->inv_scale[0] = 1.f;
    for (int i = 1; i < c.n_levels; ++i) {                                  // :434-439
        h->scale[i] = h->scale[i - 1] * c.scale_factor;
        h->inv_scale[i] = 1.f / h->scale[i];
    }
    const float inv2 = 1.0f / (c.scale_factor * c.scale_factor);           // :443-452
    float desired = (float) ((double) ((float) c.n_features * (1 - inv2)) / (1 - std::pow((double) inv2, (double) c.n_levels)));
    int sum = 0;
    for (int l = 0; l < c.n_levels - 1; ++l) {
        h->quota[l] = round_half_even(desired);
        sum += h->quota[l];
        desired *= inv2;
    }
    h->quota[c.n_levels - 1] = std::max(c.n_features - sum, 1);
}

static void build_u_max(uint8_t *u_max) {                                   // :458-474
    int um[kHalfPatch + 2] = {0};
    const int v_max = floor_f(kHalfPatch * sqrtf(2.f) / 2 + 1);
    const int v_min = ceil_f(kHalfPatch * sqrtf(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= v_max; ++v) um[v] = (int) lrint(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= v_min; --v) {
        while (um[v0] == um[v0 + 1]) ++v0;
        um[v] = v0;
        ++v0;
    }
    for (int v = 0; v <= kHalfPatch; ++v) u_max[v] = (uint8_t) um[v];
}

// cv::resize INTER_LINEAR coefficient tables: per destination index the two source indices and the
// two 11-bit weights.  The x axis zeroes the fraction at the borders, the y axis clamps the rows.
static void build_axis_table(int dn, int sn, bool is_x, std::vector<int2> &tab) {
    tab.resize(dn);
    const double scale = 1.0 / ((double) dn / (double) sn);
    for (int d = 0; d < dn; ++d) {
        float f = (float) ((d + 0.5) * scale - 0.5);
        int s = floor_f(f);
        f -= (float) s;
        int s0, s1;
        if (is_x) {
            if (s < 0) { f = 0.f; s = 0; }
            if (s >= sn - 1) { f = 0.f; s = sn - 1; }
            s0 = s; s1 = std::min(s + 1, sn - 1);
        } else {
            s0 = std::min(std::max(s, 0), sn - 1);
            s1 = std::min(std::max(s + 1, 0), sn - 1);
        }
        const int c0 = (short) round_half_even((1.f - f) * 2048.f), c1 = (short) round_half_even(f * 2048.f);
        tab[d] = make_int2(s0 | (s1 << 16), (c0 & 0xffff) | (c1 << 16));
    }
}

// ------------------------------------------------------------------------------------------------
// tensor maps
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled) p;
    }
    return fn;
}

// 3-D u8 tensor (x, y, frame) over `frames` images of w x h with the given pitch / frame stride; box = 256 x box_h x 1.
static int make_tmap(Handle *h, CUtensorMap *m, const uint8_t *base, int w, int ht, size_t pitch, size_t frame_stride, int frames, int box_h, int box_w = kBoxW) {
    PFN_cuTensorMapEncodeTiled enc = get_encode_fn();
    if (!enc) return set_error(h, ORBFE_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t) w, (cuuint64_t) ht, (cuuint64_t) std::max(frames, 1)};
    cuuint64_t strides[2] = {(cuuint64_t) pitch, (cuuint64_t) frame_stride};
    cuuint32_t box[3] = {(cuuint32_t) box_w, (cuuint32_t) box_h, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *) base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(h, ORBFE_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (w=%d h=%d pitch=%zu)", (int) r, w, ht, pitch);
    return ORBFE_OK;
}

// ------------------------------------------------------------------------------------------------
// geometry + arena
// ------------------------------------------------------------------------------------------------
static void drop_graph(Handle *h) { if (h->graph1) { cudaGraphExecDestroy(h->graph1); h->graph1 = nullptr; } }

static void free_arena(Handle *h) {
    drop_graph(h);
    cudaFree(h->d_img); cudaFree(h->d_blur); cudaFree(h->d_slots); cudaFree(h->d_cell_cnt); cudaFree(h->d_cell_off);
    cudaFree(h->d_cand); cudaFree(h->d_cur); cudaFree(h->d_nodes); cudaFree(h->d_lists); cudaFree(h->d_kp); cudaFree(h->d_nkp);
    cudaFree(h->d_ncand); cudaFree(h->d_tables); cudaFree(h->d_fast_tab); h->d_fast_tab = nullptr; cudaFree(h->d_out_kps); cudaFree(h->d_out_desc); cudaFree(h->d_out_n);
    h->d_img = h->d_blur = nullptr; h->d_slots = nullptr; h->d_cell_cnt = h->d_cell_off = nullptr; h->d_cand = nullptr; h->d_cur = nullptr;
    h->d_nodes = nullptr; h->d_lists = nullptr; h->d_kp = nullptr; h->d_nkp = h->d_ncand = nullptr; h->d_tables = nullptr;
    h->d_out_kps = nullptr; h->d_out_desc = nullptr; h->d_out_n = nullptr; h->out_cap = 0; h->out_frames = 0;
    for (int i = 0; i < Handle::kStageSlots; ++i) { cudaFree(h->d_stage[i]); h->d_stage[i] = nullptr; }
    h->stage_bytes = 0;
    h->batch_cap = 0; h->g.w = h->g.h = 0;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int configure(Handle *h, int w, int ht, int batch) {
    if (h->g.w == w && h->g.h == ht && h->batch_cap >= batch) return ORBFE_OK;
    ORBFE_CUDA(h, cudaStreamSynchronize(h->stream));
    const int keep_batch = std::max(batch, (h->g.w == w && h->g.h == ht) ? h->batch_cap : 0);
    free_arena(h);
    Geometry &g = h->g;
    g = Geometry();
    const int nl = h->cfg.n_levels;
    g.n_levels = nl;
    std::vector<std::vector<int2>> xt(nl), yt(nl);
    size_t img_off = 0, table_elems = 0;
    int max_kp_cap = 1, max_node_cap = 1;
    for (int l = 0; l < nl; ++l) {
        LevelGeom &L = g.lv[l];
        L.w = l ? round_half_even((float) w * h->inv_scale[l]) : w;          // ORBExtractor.cpp:563-564
        L.h = l ? round_half_even((float) ht * h->inv_scale[l]) : ht;
        if (L.w < 2 * kEdge + 1 || L.h < 2 * kEdge + 1 || L.w > 4000 || L.h > 4000)
            return set_error(h, ORBFE_E_ARG, "level %d of a %dx%d image is %dx%d: every level must be within [39, 4000] px", l, w, ht, L.w, L.h);
        L.pitch = (int) align_up((size_t) L.w, 64);
        L.frame_stride = (unsigned long long) L.pitch * L.h;
        L.img_off = img_off;
        img_off += align_up((size_t) L.frame_stride * keep_batch, 256);
        const int bw = L.w - 2 * kEdge, bh = L.h - 2 * kEdge;
        L.n_cols = (bw + kCell - 1) / kCell; L.n_rows = (bh + kCell - 1) / kCell;          // :589-590
        L.n_groups = (L.n_cols + kCellsPerBlk - 1) / kCellsPerBlk;
        L.cell_base = g.cells_per_frame; g.cells_per_frame += L.n_cols * L.n_rows;
        L.fast_blk_base = g.fast_blocks; g.fast_blocks += L.n_rows * L.n_groups;
        L.blur_tx = (L.w + kBlurTileW - 1) / kBlurTileW; L.blur_ty = (L.h + kBlurTileH - 1) / kBlurTileH;
        L.blur_blk_base = g.blur_blocks; g.blur_blocks += L.blur_tx * L.blur_ty;
        L.quota = h->quota[l];
        const int n_ini = ceil_f((float) bw / (float) bh);
        L.cand_off = g.cand_per_frame; L.cand_cap = L.n_cols * L.n_rows * kSlotCap; g.cand_per_frame += L.cand_cap;
        L.kp_cap = std::max(L.quota + 4, 4 * n_ini + 4); L.kp_off = g.kp_per_frame; g.kp_per_frame += L.kp_cap;
        // Quadtree node pool.  A split takes 4 slots and slots are never reclaimed.  Every pass that does not end the loop splits
        // all expandable nodes, so after p passes they all sit at depth p and are at most ceil(size / 2^p) wide; key points have
        // distinct pixel coordinates, so a 1 x 1 node is never expandable: at most ceil(log2(max(W, H))) + 1 passes, each over at
        // most max(quota, n_ini) nodes (the loop only continues while the list is within the quota).  That is the size of the
        // global pool; the shared-memory pool keeps the size frames need in practice (typ_node_cap) and k_octree moves to the
        // global one when a pass would overflow it.
        int depth = 1; while ((1 << depth) < std::max(bw, bh)) ++depth;
        L.node_cap = 4 * L.kp_cap * (depth + 2) + 5 * n_ini + 64; L.node_off = g.nodes_per_frame; g.nodes_per_frame += L.node_cap;
        const int typ_node_cap = 8 * (L.quota + 4) + 5 * n_ini + 64;
        L.list_off = g.lists_per_frame; g.lists_per_frame += 2 * L.kp_cap;
        L.scale = h->scale[l];
        max_kp_cap = std::max(max_kp_cap, L.kp_cap); max_node_cap = std::max(max_node_cap, typ_node_cap);
        if (l) {
            build_axis_table(L.w, g.lv[l - 1].w, true, xt[l]);
            build_axis_table(L.h, g.lv[l - 1].h, false, yt[l]);
            // destination tile so that the staged source box (256 x kRsBoxH) covers it
            ResizeLevel &R = g.rs[l];
            auto span_ok = [](const std::vector<int2> &tab, int tile, int limit) {
                const int n = (int) tab.size();
                for (int d0 = 0; d0 < n; d0 += tile) {
                    const int o = tab[d0].x & 0xffff;
                    int mx = 0;
                    for (int d = d0; d < std::min(n, d0 + tile); ++d) mx = std::max(mx, (tab[d].x >> 16) - o);
                    if (mx >= limit) return false;
                }
                return true;
            };
            R.tw = kRsMaxTW; while (R.tw > 16 && !span_ok(xt[l], R.tw, kBoxUsable)) R.tw -= 16;
            R.th = kRsMaxTH; while (R.th > 1 && !span_ok(yt[l], R.th, kRsBoxH)) R.th -= 1;
            if (!span_ok(xt[l], R.tw, kBoxUsable) || !span_ok(yt[l], R.th, kRsBoxH))
                return set_error(h, ORBFE_E_ARG, "scale factor %f too large for the resize tile", (double) h->cfg.scale_factor);
            R.tiles_x = (L.w + R.tw - 1) / R.tw; R.tiles_y = (L.h + R.th - 1) / R.th;
            R.packed = true;                              // quads start at multiples of 4 (tw is a multiple of 16)
            for (int d0 = 0; d0 < L.w && R.packed; d0 += 4)
                for (int i = 0; i < 4; ++i)
                    if ((xt[l][std::min(d0 + i, L.w - 1)].x >> 16) - (xt[l][d0].x & 0xffff) > 7) R.packed = false;
            table_elems += xt[l].size() + yt[l].size();
        }
    }
    g.w = w; g.h = ht;
    g.img_bytes_per_frame = 0;
    for (int l = 0; l < nl; ++l) g.img_bytes_per_frame += g.lv[l].frame_stride;
    g.sort_cap = 1; while (g.sort_cap < max_kp_cap) g.sort_cap <<= 1;
    if (g.sort_cap > 8192) return set_error(h, ORBFE_E_ARG, "n_features too large: %d key points on one level (limit 8188)", max_kp_cap);
    {   // quadtree smem: sort keys + node pool (levels whose pool does not fit use the global pool)
        const int budget = 100 * 1024 - g.sort_cap * 8;
        int cap = std::min(max_node_cap, budget / 16);
        g.oct_smem_bytes = g.sort_cap * 8 + cap * 16;
        g.typ_node_cap = max_node_cap;
        // stored in Geometry via oct_smem_bytes; smem_node_cap is recomputed at launch
    }
    h->max_kp = g.kp_per_frame;

    const size_t B = (size_t) keep_batch;
    ORBFE_CUDA(h, cudaMalloc(&h->d_img, img_off + 1024));
    ORBFE_CUDA(h, cudaMalloc(&h->d_blur, img_off + 1024));
    ORBFE_CUDA(h, cudaMemsetAsync(h->d_img, 0, img_off + 1024, h->stream));
    ORBFE_CUDA(h, cudaMemsetAsync(h->d_blur, 0, img_off + 1024, h->stream));
    ORBFE_CUDA(h, cudaMalloc(&h->d_slots, B * g.cells_per_frame * kSlotCap * sizeof(uint32_t)));
    ORBFE_CUDA(h, cudaMalloc(&h->d_cell_cnt, B * g.cells_per_frame * sizeof(int)));
    ORBFE_CUDA(h, cudaMalloc(&h->d_cell_off, B * g.cells_per_frame * sizeof(int)));
    ORBFE_CUDA(h, cudaMalloc(&h->d_cand, B * g.cand_per_frame * sizeof(uint32_t)));
    ORBFE_CUDA(h, cudaMalloc(&h->d_cur, B * g.cand_per_frame * sizeof(int)));
    ORBFE_CUDA(h, cudaMalloc(&h->d_nodes, B * g.nodes_per_frame * 16));
    ORBFE_CUDA(h, cudaMalloc(&h->d_lists, B * g.lists_per_frame * sizeof(int)));
    ORBFE_CUDA(h, cudaMalloc(&h->d_kp, B * g.kp_per_frame * sizeof(uint32_t)));
    ORBFE_CUDA(h, cudaMalloc(&h->d_nkp, B * ORBFE_MAX_LEVELS * sizeof(int)));
    ORBFE_CUDA(h, cudaMalloc(&h->d_ncand, B * ORBFE_MAX_LEVELS * sizeof(int)));
    ORBFE_CUDA(h, cudaMemsetAsync(h->d_nkp, 0, B * ORBFE_MAX_LEVELS * sizeof(int), h->stream));
    ORBFE_CUDA(h, cudaMemsetAsync(h->d_ncand, 0, B * ORBFE_MAX_LEVELS * sizeof(int), h->stream));
    // resize tables
    ORBFE_CUDA(h, cudaMalloc(&h->d_tables, std::max<size_t>(table_elems, 1) * sizeof(int2)));
    {
        int2 *p = (int2 *) h->d_tables;
        for (int l = 1; l < nl; ++l) {
            ORBFE_CUDA(h, cudaMemcpyAsync(p, xt[l].data(), xt[l].size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
            g.rs[l].xtab = p; p += xt[l].size();
            ORBFE_CUDA(h, cudaMemcpyAsync(p, yt[l].data(), yt[l].size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
            g.rs[l].ytab = p; p += yt[l].size();
        }
        // FAST block table: block -> (level, cell row, strip of 8 cells)
        // followed by the blur block table: block -> (level, tile row, tile column)
        std::vector<int> tab((size_t) g.fast_blocks + g.blur_blocks);
        for (int l = 0; l < nl; ++l) {
            for (int ci = 0; ci < g.lv[l].n_rows; ++ci)
                for (int cg = 0; cg < g.lv[l].n_groups; ++cg) tab[(size_t) g.lv[l].fast_blk_base + ci * g.lv[l].n_groups + cg] = l | (ci << 4) | (cg << 16);
            for (int ty = 0; ty < g.lv[l].blur_ty; ++ty)
                for (int tx = 0; tx < g.lv[l].blur_tx; ++tx) tab[(size_t) g.fast_blocks + g.lv[l].blur_blk_base + ty * g.lv[l].blur_tx + tx] = l | (ty << 4) | (tx << 16);
        }
        ORBFE_CUDA(h, cudaMalloc(&h->d_fast_tab, tab.size() * sizeof(int)));
        ORBFE_CUDA(h, cudaMemcpyAsync(h->d_fast_tab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        ORBFE_CUDA(h, cudaStreamSynchronize(h->stream));      // the host vectors die with this scope
    }
    h->batch_cap = keep_batch;
    if (h->use_tma) {
        for (int l = 0; l < nl; ++l) {
            const LevelGeom &L = g.lv[l];
            int rc;
            if ((rc = make_tmap(h, &h->tm_fast[l], h->d_img + L.img_off, L.w, L.h, L.pitch, L.frame_stride, keep_batch, kFastBoxH))) return rc;
            if ((rc = make_tmap(h, &h->tm_blur[l], h->d_img + L.img_off, L.w, L.h, L.pitch, L.frame_stride, keep_batch, kBlurBoxH))) return rc;
            if ((rc = make_tmap(h, &h->tm_pimg[l], h->d_img + L.img_off, L.w, L.h, L.pitch, L.frame_stride, keep_batch, kPatchH, kPatchW))) return rc;
            if ((rc = make_tmap(h, &h->tm_pblur[l], h->d_blur + L.img_off, L.w, L.h, L.pitch, L.frame_stride, keep_batch, kBPatchH, kBPatchW))) return rc;
            if ((rc = make_tmap(h, &h->tm_rs[l], h->d_img + L.img_off, L.w, L.h, L.pitch, L.frame_stride, keep_batch, kRsBoxH))) return rc;
        }
    }
    return ORBFE_OK;
}

static int ensure_out_staging(Handle *h, int cap) {
    if (h->d_out_kps && h->out_cap >= cap && h->out_frames >= h->batch_cap) return ORBFE_OK;
    ORBFE_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->d_out_kps); cudaFree(h->d_out_desc); cudaFree(h->d_out_n);
    h->d_out_kps = nullptr; h->d_out_desc = nullptr; h->d_out_n = nullptr;
    const size_t B = (size_t) h->batch_cap;
    ORBFE_CUDA(h, cudaMalloc(&h->d_out_kps, B * cap * sizeof(orbfe_keypoint)));
    ORBFE_CUDA(h, cudaMalloc(&h->d_out_desc, B * cap * 32));
    ORBFE_CUDA(h, cudaMalloc(&h->d_out_n, B * sizeof(int)));
    h->out_cap = cap; h->out_frames = (int) B;
    return ORBFE_OK;
}

// staging of the pipelined host entry point: kStageSlots dense input slots and output slots, copy streams and events
static int ensure_pipeline(Handle *h, size_t stage_bytes, int chunk, int cap) {
    if (!h->s_up) {
        ORBFE_CUDA(h, cudaStreamCreateWithFlags(&h->s_up, cudaStreamNonBlocking));
        ORBFE_CUDA(h, cudaStreamCreateWithFlags(&h->s_down, cudaStreamNonBlocking));
        for (int i = 0; i < Handle::kStageSlots; ++i) {
            ORBFE_CUDA(h, cudaEventCreateWithFlags(&h->ev_up[i], cudaEventDisableTiming));
            ORBFE_CUDA(h, cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
            ORBFE_CUDA(h, cudaEventCreateWithFlags(&h->ev_down[i], cudaEventDisableTiming));
        }
    }
    if (h->stage_bytes < stage_bytes) {
        ORBFE_CUDA(h, cudaDeviceSynchronize());
        for (int i = 0; i < Handle::kStageSlots; ++i) { cudaFree(h->d_stage[i]); h->d_stage[i] = nullptr; }
        for (int i = 0; i < Handle::kStageSlots; ++i) ORBFE_CUDA(h, cudaMalloc(&h->d_stage[i], stage_bytes + 256));
        h->stage_bytes = stage_bytes;
    }
    if (!h->d_out_kps || h->out_cap < cap || h->out_frames < Handle::kStageSlots * chunk) {
        ORBFE_CUDA(h, cudaDeviceSynchronize());
        cudaFree(h->d_out_kps); cudaFree(h->d_out_desc); cudaFree(h->d_out_n);
        h->d_out_kps = nullptr; h->d_out_desc = nullptr; h->d_out_n = nullptr;
        const size_t B = (size_t) std::max(Handle::kStageSlots * chunk, h->batch_cap);
        ORBFE_CUDA(h, cudaMalloc(&h->d_out_kps, B * cap * sizeof(orbfe_keypoint)));
        ORBFE_CUDA(h, cudaMalloc(&h->d_out_desc, B * cap * 32));
        ORBFE_CUDA(h, cudaMalloc(&h->d_out_n, B * sizeof(int)));
        h->out_cap = cap; h->out_frames = (int) B;
    }
    return ORBFE_OK;
}

// ------------------------------------------------------------------------------------------------
// one device pass over nb <= batch_cap frames.  level-0 images are read from (l0, l0_pitch, l0_fstride), which is either
// the arena (after an H2D / D2D copy) or the caller's device buffer used in place.
// ------------------------------------------------------------------------------------------------
// ORBFE_DEBUG_SYNC=1 synchronises after every launch and names the kernel that faulted (debugging aid, off by default)
static bool no_graph() { static int v = -1; if (v < 0) { const char *e = getenv("ORBFE_NO_GRAPH"); v = e && *e == '1'; } return v == 1; }
static bool debug_sync() { static int v = -1; if (v < 0) { const char *e = getenv("ORBFE_DEBUG_SYNC"); v = e && *e == '1'; } return v == 1; }
#define ORBFE_AFTER_LAUNCH(h, st, name)                                                                          \
    do { (h)->launches++;                                                                                        \
         if (debug_sync()) { cudaError_t e__ = cudaStreamSynchronize(st);                                        \
             if (e__ == cudaSuccess) e__ = cudaGetLastError();                                                   \
             if (e__ != cudaSuccess) return set_error((h), ORBFE_E_CUDA, "kernel %s failed: %s", name, cudaGetErrorString(e__)); } } while (0)

static int prof_collect(Handle *h) {
    if (!h->prof_pending) return ORBFE_OK;
    ORBFE_CUDA(h, cudaEventSynchronize(h->prof_ev[ORBFE_N_STAGES]));
    for (int s = 0; s < ORBFE_N_STAGES; ++s) {
        float ms = 0.f;
        ORBFE_CUDA(h, cudaEventElapsedTime(&ms, h->prof_ev[s], h->prof_ev[s + 1]));
        h->prof_ms[s] += ms;
    }
    h->prof_passes++;
    h->prof_pending = false;
    return ORBFE_OK;
}
#define ORBFE_PROF_MARK(h, st, i) do { if ((h)->prof) ORBFE_CUDA(h, cudaEventRecord((h)->prof_ev[i], st)); } while (0)

template <bool kTMA>
static int run_pass_t(Handle *h, int nb, const uint8_t *l0, size_t l0_pitch, size_t l0_fstride,
                      orbfe_keypoint *d_kps, uint8_t *d_desc, int *d_n, int cap, cudaStream_t st) {
    const Geometry &g = h->g;
    const int nl = g.n_levels;
    LevelSet LS; memset(&LS, 0, sizeof LS);
    LS.n_levels = nl;
    for (int l = 0; l < nl; ++l) { LS.lv[l] = g.lv[l]; LS.img[l] = h->d_img + g.lv[l].img_off; }
    TmapSet TF, TB; CUtensorMap tm_rs0; PatchMaps PM;
    memset(&TF, 0, sizeof TF); memset(&TB, 0, sizeof TB); memset(&tm_rs0, 0, sizeof tm_rs0); memset(&PM, 0, sizeof PM);
    if (kTMA) {
        for (int l = 0; l < nl; ++l) { TF.m[l] = h->tm_fast[l]; TB.m[l] = h->tm_blur[l]; PM.img[l] = h->tm_pimg[l]; PM.blur[l] = h->tm_pblur[l]; }
        tm_rs0 = h->tm_rs[0];
    }
    const bool inplace = l0 != h->d_img + g.lv[0].img_off;
    if (inplace) {
        LS.img[0] = l0; LS.lv[0].pitch = (int) l0_pitch; LS.lv[0].frame_stride = l0_fstride;
        if (kTMA) {
            int rc;
            if ((rc = make_tmap(h, &TF.m[0], l0, g.w, g.h, l0_pitch, l0_fstride, nb, kFastBoxH))) return rc;
            if ((rc = make_tmap(h, &TB.m[0], l0, g.w, g.h, l0_pitch, l0_fstride, nb, kBlurBoxH))) return rc;
            if ((rc = make_tmap(h, &tm_rs0, l0, g.w, g.h, l0_pitch, l0_fstride, nb, kRsBoxH))) return rc;
            if ((rc = make_tmap(h, &PM.img[0], l0, g.w, g.h, l0_pitch, l0_fstride, nb, kPatchH, kPatchW))) return rc;
            // the blurred level 0 is written with the caller's strides as well (the kernels use one LevelGeom for both arenas)
            if ((rc = make_tmap(h, &PM.blur[0], h->d_blur + g.lv[0].img_off, g.w, g.h, l0_pitch, l0_fstride, nb, kBPatchH, kBPatchW))) return rc;
        }
    }
    if (h->prof) { int rc = prof_collect(h); if (rc) return rc; }
    ORBFE_PROF_MARK(h, st, 0);
    // K1 pyramid: level l from level l-1 (ComputePyramid, ORBExtractor.cpp:559-570)
    for (int l = 1; l < nl; ++l) {
        const LevelGeom &S = LS.lv[l - 1], &D = LS.lv[l];
        const ResizeLevel &R = g.rs[l];
        ResizeArgs ra;
        ra.src = LS.img[l - 1]; ra.sw = S.w; ra.sh = S.h; ra.spitch = S.pitch; ra.sframe = S.frame_stride;
        ra.dst = h->d_img + D.img_off; ra.dw = D.w; ra.dh = D.h; ra.dpitch = D.pitch; ra.dframe = D.frame_stride;
        ra.tw = R.tw; ra.th = R.th; ra.tiles_x = R.tiles_x; ra.xtab = R.xtab; ra.ytab = R.ytab;
        if (R.packed) k_resize<kTMA, true><<<dim3(R.tiles_x * R.tiles_y, nb), kRsThreads, 0, st>>>(l == 1 ? tm_rs0 : h->tm_rs[l - 1], ra);
        else k_resize<kTMA, false><<<dim3(R.tiles_x * R.tiles_y, nb), kRsThreads, 0, st>>>(l == 1 ? tm_rs0 : h->tm_rs[l - 1], ra);
        ORBFE_AFTER_LAUNCH(h, st, "k_resize");
    }
    ORBFE_PROF_MARK(h, st, 1);
    // The blur only depends on the pyramid: it runs on an auxiliary stream next to FAST + quadtree (which are issue- and
    // latency-bound) and is joined before the descriptors.
    BlurArgs ba; ba.blur = h->d_blur; ba.blk_tab = h->d_fast_tab + g.fast_blocks; blur_taps(ba.kc);
    const bool fork_blur = !h->prof && !debug_sync();
    if (fork_blur) {
        if (!h->s_aux) {
            ORBFE_CUDA(h, cudaStreamCreateWithFlags(&h->s_aux, cudaStreamNonBlocking));
            ORBFE_CUDA(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
            ORBFE_CUDA(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        }
        ORBFE_CUDA(h, cudaEventRecord(h->ev_fork, st));
        ORBFE_CUDA(h, cudaStreamWaitEvent(h->s_aux, h->ev_fork, 0));
        k_blur<kTMA><<<dim3(g.blur_blocks, nb), kBlurThreads, 0, h->s_aux>>>(LS, TB, ba);
        h->launches++;
        ORBFE_CUDA(h, cudaEventRecord(h->ev_join, h->s_aux));
    }
    // K2 FAST + per-cell NMS
    FastArgs fa; fa.slots = h->d_slots; fa.cell_cnt = h->d_cell_cnt; fa.blk_tab = h->d_fast_tab; fa.cells_per_frame = g.cells_per_frame;
    fa.t_ini = h->cfg.ini_th_fast; fa.t_min = h->cfg.min_th_fast;
    fa.one = 1; fa.flags = (h->fast_exact_cmp ? 1 : 0) | (h->fast_fma_shift ? 2 : 0);
    fa.shift_mul = make_uint3(1u << 24, 1u << 16, 1u << 8);
    if (h->fast_v1) {
        k_fast<kTMA><<<dim3(g.fast_blocks, nb), 256, 0, st>>>(LS, TF, fa);
        ORBFE_AFTER_LAUNCH(h, st, "k_fast");
    } else {
        Fast2Args f2; f2.slots = fa.slots; f2.cell_cnt = fa.cell_cnt; f2.blk_tab = fa.blk_tab; f2.cells_per_frame = fa.cells_per_frame;
        f2.t_ini = fa.t_ini; f2.t_min = fa.t_min; f2.one = 1u; f2.exact = h->fast_exact_cmp ? 1 : 0;
        k_fast_planes<kTMA><<<dim3(g.fast_blocks, nb), kF2Threads, 0, st>>>(LS, TF, f2);
        ORBFE_AFTER_LAUNCH(h, st, "k_fast_planes");
    }
    ORBFE_PROF_MARK(h, st, 2);
    // K4 quadtree
    OctArgs oa;
    oa.slots = h->d_slots; oa.cell_cnt = h->d_cell_cnt; oa.cell_off = h->d_cell_off; oa.cand = h->d_cand; oa.cur = h->d_cur;
    oa.nodes = h->d_nodes; oa.lists = h->d_lists; oa.kp = h->d_kp; oa.nkp = h->d_nkp; oa.ncand = h->d_ncand; oa.err = h->d_err;
    oa.cells_per_frame = g.cells_per_frame; oa.cand_per_frame = g.cand_per_frame; oa.nodes_per_frame = g.nodes_per_frame;
    oa.lists_per_frame = g.lists_per_frame; oa.kp_per_frame = g.kp_per_frame;
    // one CTA per (level, frame).  Few frames: the level-0 CTA is the critical path and most SMs are idle -> 1024 threads and a node
    // pool of up to 200 KB in shared memory (1 CTA per SM; measured 0.161 -> 0.129 ms for one 1920x1080 / 4000-feature frame);
    // batches: 512 threads and the 100 KB budget (measured 0.286 ms vs 0.312 ms with 256 and 0.53 ms with 1024 threads per 512 C1
    // frames; the large pool would cost occupancy there).  ORBFE_OCT_NT overrides the thread count.
    static const int oct_nt = [] { const char *e = getenv("ORBFE_OCT_NT"); return e ? atoi(e) : 0; }();
    const bool few = nl * nb <= h->sm_count;
    const int nt = oct_nt ? oct_nt : (few ? 1024 : 512);
    int oct_smem = g.oct_smem_bytes;
    if (few) {
        oct_smem = std::max(oct_smem, std::min(200 * 1024, g.sort_cap * 8 + g.typ_node_cap * 16));
    }
    oa.sort_cap = g.sort_cap; oa.smem_node_cap = (oct_smem - g.sort_cap * 8) / 16;
    // ORBFE_OCT_SMEM_NODES=n shrinks the shared-memory pool so that the tests can drive the move to the global pool
    if (const char *e = getenv("ORBFE_OCT_SMEM_NODES")) { const int v = atoi(e); if (v > 0) oa.smem_node_cap = std::min(oa.smem_node_cap, v); }
    if (nt == 1024) k_octree<1024><<<dim3(nl, nb), 1024, oct_smem, st>>>(LS, oa);
    else if (nt == 512) k_octree<512><<<dim3(nl, nb), 512, oct_smem, st>>>(LS, oa);
    else k_octree<256><<<dim3(nl, nb), 256, oct_smem, st>>>(LS, oa);
    ORBFE_AFTER_LAUNCH(h, st, "k_octree");
    ORBFE_PROF_MARK(h, st, 3);
    // K6 blur (launched above on the auxiliary stream unless profiling / debugging serialises the stages)
    if (!fork_blur) {
        k_blur<kTMA><<<dim3(g.blur_blocks, nb), kBlurThreads, 0, st>>>(LS, TB, ba);
        ORBFE_AFTER_LAUNCH(h, st, "k_blur");
    } else {
        ORBFE_CUDA(h, cudaStreamWaitEvent(st, h->ev_join, 0));
    }
    ORBFE_PROF_MARK(h, st, 4);
    // K5 + K7 orientation and descriptors, final assembly
    k_zero_counts<<<(nb + 255) / 256, 256, 0, st>>>(d_n, nb);
    ORBFE_AFTER_LAUNCH(h, st, "k_zero_counts");
    DescArgs da;
    da.blur = h->d_blur; da.kp = h->d_kp; da.nkp = h->d_nkp; da.kp_per_frame = g.kp_per_frame;
    da.out_kps = d_kps; da.out_desc = d_desc; da.out_n = d_n; da.cap = cap; da.err = h->d_err;
    build_u_max(da.u_max);
    {   // the kernel's compile-time u_max table must be what the reference's formula gives
        static const uint8_t expect[kHalfPatch + 1] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
        if (memcmp(da.u_max, expect, sizeof expect) != 0) return set_error(h, ORBFE_E_INTERNAL, "u_max table mismatch");
    }
    da.pattern = h->d_pattern;
    k_describe<kTMA><<<dim3((g.kp_per_frame + 7) / 8, nb), 256, 0, st>>>(LS, PM, da);
    ORBFE_AFTER_LAUNCH(h, st, "k_describe");
    ORBFE_PROF_MARK(h, st, 5);
    if (h->prof) h->prof_pending = true;
    ORBFE_CUDA(h, cudaGetLastError());
    h->last_batch = nb;
    return ORBFE_OK;
}

static int run_pass(Handle *h, int nb, const uint8_t *l0, size_t l0_pitch, size_t l0_fstride,
                    orbfe_keypoint *d_kps, uint8_t *d_desc, int *d_n, int cap, cudaStream_t st) {
    return h->use_tma ? run_pass_t<true>(h, nb, l0, l0_pitch, l0_fstride, d_kps, d_desc, d_n, cap, st)
                      : run_pass_t<false>(h, nb, l0, l0_pitch, l0_fstride, d_kps, d_desc, d_n, cap, st);
}

static int check_device_error(Handle *h, cudaStream_t st) {
    int e = 0;
    ORBFE_CUDA(h, cudaMemcpyAsync(&e, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaStreamSynchronize(st));
    if (e) {
        cudaMemsetAsync(h->d_err, 0, sizeof(int), st);
        static const char *what[] = {"", "candidate capacity", "root nodes exceed the node pool", "quadtree node pool", "expandable-node list",
                                     "key points per level", "caller key-point capacity (cap) too small"};
        return set_error(h, e == 6 ? ORBFE_E_CAPACITY : ORBFE_E_INTERNAL, "device reported overflow: %s (code %d)", e >= 1 && e <= 6 ? what[e] : "?", e);
    }
    return ORBFE_OK;
}

// Waits for every batch submitted to the host pipeline and reports (and clears) the overflow flags of both arenas.
static int pipeline_drain(Handle *h) {
    if (!h->pipe_pending) return ORBFE_OK;
    h->pipe_pending = false;
    ORBFE_CUDA(h, cudaStreamSynchronize(h->s_down));
    ORBFE_CUDA(h, cudaStreamSynchronize(h->s_up));
    int rc = check_device_error(h, h->stream);
    if (h->peer && h->pipe_peer_used) {
        h->pipe_peer_used = false;
        const int rp = check_device_error(h->peer, h->peer->stream);
        if (!rc && rp) rc = set_error(h, rp, "pipeline peer: %s", orbfe_last_error(static_cast<orbfe_handle *>(h->peer)));
    }
    return rc;
}

}  // namespace orbfe

using namespace orbfe;

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" {

const char *orbfe_version(void) { return "orbfe 0.1 (sm_100a)"; }

int orbfe_create(const orbfe_config *cfg, orbfe_handle **out) {
    if (!cfg || !out) return set_error(nullptr, ORBFE_E_ARG, "null argument");
    *out = nullptr;
    if (cfg->n_levels < 1 || cfg->n_levels > ORBFE_MAX_LEVELS || cfg->n_features < 1 || !(cfg->scale_factor > 1.0f) || cfg->max_batch < 1 ||
        cfg->ini_th_fast < 0 || cfg->min_th_fast < 0 || cfg->ini_th_fast > 254 || cfg->min_th_fast > 254)
        return set_error(nullptr, ORBFE_E_ARG, "invalid config (n_levels 1..%d, n_features >= 1, scale_factor > 1, thresholds 0..254, max_batch >= 1)", ORBFE_MAX_LEVELS);
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return set_error(nullptr, ORBFE_E_CUDA, "no CUDA device: %s (this library has no CPU fallback)", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (cfg->device < 0 || cfg->device >= n_dev) return set_error(nullptr, ORBFE_E_ARG, "device %d out of range (%d devices)", cfg->device, n_dev);
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return set_error(nullptr, ORBFE_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) return set_error(nullptr, ORBFE_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10) return set_error(nullptr, ORBFE_E_CUDA, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
    orbfe_handle *h = new orbfe_handle();
    h->cfg = *cfg; h->device = cfg->device; h->sm_count = prop.multiProcessorCount;
    h->use_tma = !(cfg->flags & ORBFE_FLAG_NO_TMA);
    h->keep_stages = (cfg->flags & ORBFE_FLAG_KEEP_STAGES) != 0;
    if (const char *ev = getenv("ORBFE_FAST_EXACT")) h->fast_exact_cmp = *ev == '1';
    if (const char *ev = getenv("ORBFE_FAST_FMA_SHIFT")) h->fast_fma_shift = *ev == '1';
    if (const char *ev = getenv("ORBFE_FAST_V1")) h->fast_v1 = *ev == '1';
    build_ctor_tables(h);
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaMalloc(&h->d_err, sizeof(int))) != cudaSuccess || (e = cudaMemset(h->d_err, 0, sizeof(int))) != cudaSuccess) {
        set_error(nullptr, ORBFE_E_CUDA, "handle setup: %s", cudaGetErrorString(e));
        delete h; return ORBFE_E_CUDA;
    }
    {   // pattern table of k_describe: pair p = 8 * lane + j (bit j of descriptor byte `lane`) at float4 slot j * 32 + lane
        std::vector<float> pat(1024);
        for (int pair = 0; pair < 256; ++pair)
            for (int c = 0; c < 4; ++c) pat[(size_t) (((pair & 7) * 32) + (pair >> 3)) * 4 + c] = (float) kBriefPattern[pair * 4 + c];
        if ((e = cudaMalloc(&h->d_pattern, pat.size() * sizeof(float))) != cudaSuccess ||
            (e = cudaMemcpy(h->d_pattern, pat.data(), pat.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) {
            set_error(nullptr, ORBFE_E_CUDA, "pattern table: %s", cudaGetErrorString(e));
            orbfe_destroy(h); return ORBFE_E_CUDA;
        }
    }
    // dynamic shared-memory opt-ins are per device: set them for this handle's device (current after cudaSetDevice above)
    if ((e = cudaFuncSetAttribute(k_octree<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_octree<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_octree<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) {
        set_error(nullptr, ORBFE_E_CUDA, "kernel attributes: %s", cudaGetErrorString(e));
        orbfe_destroy(h); return ORBFE_E_CUDA;
    }
    if (match_device_setup(h) != ORBFE_OK || frame_device_setup(h) != ORBFE_OK) {
        set_error(nullptr, ORBFE_E_CUDA, "kernel attributes: %s", h->err.c_str());
        orbfe_destroy(h); return ORBFE_E_CUDA;
    }
    *out = h;
    return ORBFE_OK;
}

void orbfe_destroy(orbfe_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->s_up) cudaStreamSynchronize(h->s_up);                 // batches still in flight (orbfe_extract_batch_submit)
    if (h->s_down) cudaStreamSynchronize(h->s_down);
    if (h->peer) { orbfe_destroy(h->peer); h->peer = nullptr; }
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_arena(h);
    cudaFree(h->d_err); cudaFree(h->d_match); cudaFree(h->d_ap); cudaFree(h->d_pattern); cudaFree(h->d_unc);
    for (int i = 0; i <= ORBFE_N_STAGES; ++i) if (h->prof_ev[i]) cudaEventDestroy(h->prof_ev[i]);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    if (h->h_mpin) cudaFreeHost(h->h_mpin);
    if (h->h_ticket_err) cudaFreeHost(h->h_ticket_err);
    for (int i = 0; i < Handle::kTickets; ++i) if (h->ev_ticket[i]) cudaEventDestroy(h->ev_ticket[i]);
    for (auto &e : h->ev_split) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->s_aux) { cudaStreamDestroy(h->s_aux); cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join); }
    if (h->s_up) cudaStreamDestroy(h->s_up);
    if (h->s_down) cudaStreamDestroy(h->s_down);
    for (int i = 0; i < Handle::kStageSlots; ++i) {
        if (h->ev_up[i]) cudaEventDestroy(h->ev_up[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
        if (h->ev_down[i]) cudaEventDestroy(h->ev_down[i]);
    }
    delete h;
}

const char *orbfe_last_error(const orbfe_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

float orbfe_scale_factor(const orbfe_handle *h, int level) { return (h && level >= 0 && level < h->cfg.n_levels) ? h->scale[level] : 0.f; }
int orbfe_features_per_level(const orbfe_handle *h, int level) { return (h && level >= 0 && level < h->cfg.n_levels) ? h->quota[level] : -1; }
int orbfe_max_keypoints(const orbfe_handle *h) {
    if (!h) return -1;
    if (h->max_kp) return h->max_kp;
    int s = 0;                                  // before the first frame fixes the geometry: a bound that holds for aspect ratios up to 64:1
    for (int l = 0; l < h->cfg.n_levels; ++l) s += std::max(h->quota[l] + 40, 260);
    return s;
}
long long orbfe_launch_count(const orbfe_handle *h) { return h ? h->launches + (h->peer ? h->peer->launches : 0) : 0; }

int orbfe_profile(orbfe_handle *h, int enable) {
    if (!h) return ORBFE_E_ARG;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    if (enable && !h->prof_ev[0])
        for (int i = 0; i <= ORBFE_N_STAGES; ++i) ORBFE_CUDA(h, cudaEventCreate(&h->prof_ev[i]));
    if (!enable) { int rc = prof_collect(h); if (rc) return rc; }
    h->prof = enable != 0;
    return ORBFE_OK;
}

int orbfe_profile_read(orbfe_handle *h, float *stage_ms, int *n_passes, int reset) {
    if (!h || !stage_ms || !n_passes) return ORBFE_E_ARG;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    int rc = prof_collect(h);
    if (rc) return rc;
    for (int s = 0; s < ORBFE_N_STAGES; ++s) stage_ms[s] = (float) h->prof_ms[s];
    *n_passes = h->prof_passes;
    if (reset) { for (int s = 0; s < ORBFE_N_STAGES; ++s) h->prof_ms[s] = 0; h->prof_passes = 0; }
    return ORBFE_OK;
}

int orbfe_host_alloc(void **ptr, size_t bytes) { return cudaHostAlloc(ptr, bytes, cudaHostAllocDefault) == cudaSuccess ? ORBFE_OK : ORBFE_E_CUDA; }
void orbfe_host_free(void *ptr) { if (ptr) cudaFreeHost(ptr); }

int orbfe_extract_batch_device(orbfe_handle *h, const uint8_t *d_frames, int n_frames, int width, int height, size_t row_stride,
                               size_t frame_stride, orbfe_keypoint *d_kps, uint8_t *d_desc, int cap, int *d_n, void *stream, int sync) {
    if (!h) return ORBFE_E_ARG;
    if (!d_frames || !d_kps || !d_desc || !d_n || n_frames < 0 || cap < 1 || row_stride < (size_t) width)
        return set_error(h, ORBFE_E_ARG, "orbfe_extract_batch_device: invalid argument (frames=%p kps=%p desc=%p n=%p n_frames=%d cap=%d row_stride=%zu width=%d)",
                         (const void *) d_frames, (void *) d_kps, (void *) d_desc, (void *) d_n, n_frames, cap, row_stride, width);
    if (n_frames == 0 || width <= 0 || height <= 0) return ORBFE_OK;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t) stream : h->stream;
    int rc = pipeline_drain(h);                                  // submitted host batches share the arena
    if (rc) return rc;
    // A batch of 64 frames or more runs as two half-size passes on two arenas / streams (this handle's and the peer's, forked from
    // and joined back into `st`): the latency-bound tail of one half (quadtree, descriptors) fills the issue slots the other half's
    // FAST leaves idle.  512 C1 frames: 2.99 -> 2.89 ms (tools/split_probe.py).  Stage profiling keeps the single pass.
    static const bool no_peer = [] { const char *e = getenv("ORBFE_NO_PEER"); return e && *e == '1'; }();
    const int whole = std::min(n_frames, h->cfg.max_batch);
    const bool split = !no_peer && !h->prof && !debug_sync() && n_frames >= 64;
    const int pass = split ? (whole + 1) / 2 : whole;
    if ((rc = configure(h, width, height, pass))) return rc;
    Handle *hp = h;
    if (split) {
        if (!h->peer) {
            orbfe_config pc = h->cfg; pc.max_batch = pass;
            if ((rc = orbfe_create(&pc, &h->peer))) return set_error(h, rc, "peer arena: %s", orbfe_last_error(nullptr));
        }
        if ((rc = configure(h->peer, width, height, pass))) return set_error(h, rc, "peer arena: %s", orbfe_last_error(h->peer));
        h->peer->use_tma = h->use_tma;
        hp = h->peer;
        if (!h->ev_split[0]) for (auto &e : h->ev_split) ORBFE_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ORBFE_CUDA(h, cudaEventRecord(h->ev_split[0], st));
        ORBFE_CUDA(h, cudaStreamWaitEvent(hp->stream, h->ev_split[0], 0));
    }
    const int step = std::min(pass, std::min(h->batch_cap, hp->batch_cap));
    const bool inplace = ((uintptr_t) d_frames % 16 == 0) && row_stride % 16 == 0 && frame_stride % 16 == 0;
    int i = 0;
    for (int b0 = 0; b0 < n_frames; b0 += step, ++i) {
        const int nb = std::min(step, n_frames - b0);
        Handle *hc = (i & 1) ? hp : h;
        cudaStream_t sc = hc == h ? st : hc->stream;
        const uint8_t *src = d_frames + (size_t) b0 * frame_stride;
        const LevelGeom &L0 = hc->g.lv[0];
        if (inplace) {
            rc = run_pass(hc, nb, src, row_stride, frame_stride, d_kps + (size_t) b0 * cap, d_desc + (size_t) b0 * cap * 32, d_n + b0, cap, sc);
        } else {
            // image.clone() into the pitched arena (ORBExtractor.cpp:567)
            if (frame_stride == row_stride * (size_t) height)
                ORBFE_CUDA(h, cudaMemcpy2DAsync(hc->d_img + L0.img_off, L0.pitch, src, row_stride, width, (size_t) height * nb, cudaMemcpyDeviceToDevice, sc));
            else
                for (int b = 0; b < nb; ++b)
                    ORBFE_CUDA(h, cudaMemcpy2DAsync(hc->d_img + L0.img_off + (size_t) b * L0.frame_stride, L0.pitch, src + (size_t) b * frame_stride,
                                                    row_stride, width, height, cudaMemcpyDeviceToDevice, sc));
            rc = run_pass(hc, nb, hc->d_img + L0.img_off, L0.pitch, L0.frame_stride, d_kps + (size_t) b0 * cap, d_desc + (size_t) b0 * cap * 32, d_n + b0, cap, sc);
        }
        if (rc) return hc == h ? rc : set_error(h, rc, "peer arena: %s", orbfe_last_error(static_cast<orbfe_handle *>(hc)));
    }
    if (split) {
        ORBFE_CUDA(h, cudaEventRecord(h->ev_split[1], hp->stream));
        ORBFE_CUDA(h, cudaStreamWaitEvent(st, h->ev_split[1], 0));
    }
    if (sync) {
        rc = check_device_error(h, st);
        if (split) {
            const int rp = check_device_error(hp, hp->stream);
            if (!rc && rp) rc = set_error(h, rp, "peer arena: %s", orbfe_last_error(static_cast<orbfe_handle *>(hp)));
        }
        return rc;
    }
    return ORBFE_OK;
}

// Host batch entry point: a 3-stream software pipeline over chunks of frames.  Chunk c is copied host->device into one of two
// dense staging buffers on the copy stream while chunk c-1 runs on the compute stream and the results of chunk c-2 go back to the
// host on the download stream; level 0 is read in place from the staging buffer when the row size allows TMA (multiple of 16).
//
// The pipeline state (slot rotation, arena parity, slot events) lives in the handle, so a batch submitted while the previous one is
// still in flight continues the rotation: its first uploads run under the last passes of the previous batch.
static int batch_enqueue(Handle *h, const uint8_t *frames, int n_frames, int width, int height, size_t row_stride, size_t frame_stride,
                         orbfe_keypoint *kps, uint8_t *desc, int cap, int *n_per_frame, bool blocking) {
    if (!frames || !kps || !desc || !n_per_frame || n_frames < 0 || cap < 1 || row_stride < (size_t) width)
        return set_error(h, ORBFE_E_ARG, "orbfe_extract_batch: invalid argument (frames=%p kps=%p desc=%p n=%p n_frames=%d cap=%d row_stride=%zu width=%d)",
                         (const void *) frames, (void *) kps, (void *) desc, (void *) n_per_frame, n_frames, cap, row_stride, width);
    for (int b = 0; b < n_frames; ++b) n_per_frame[b] = 0;
    if (n_frames == 0 || width <= 0 || height <= 0) return ORBFE_OK;                 // ORBExtractor.cpp:497
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    int chunk_target = 128;                               // frames per pipeline stage (ORBFE_CHUNK overrides, for tuning)
    if (const char *e = getenv("ORBFE_CHUNK")) chunk_target = std::max(1, atoi(e));
    const int chunk = std::min(h->cfg.max_batch, std::max(std::min(16, chunk_target), std::min(chunk_target, (n_frames + 3) / 4)));
    // batches in flight share the output slots: a change of frame size, chunk size or capacity drains the pipeline first
    if (h->pipe_pending && (h->pipe_w != width || h->pipe_h != height || h->pipe_cap != cap || h->pipe_chunk != chunk)) {
        int rcw = pipeline_drain(h);
        if (rcw) return rcw;
    }
    int rc = configure(h, width, height, std::min(n_frames, chunk));
    if (rc) return rc;
    const int cn = std::min(h->batch_cap, chunk);                                   // frames per pipeline stage
    if ((rc = ensure_pipeline(h, (size_t) cn * width * height, cn, cap))) return rc;
    if (h->pipe_pending && h->pipe_cn != cn) { if ((rc = pipeline_drain(h))) return rc; }
    h->pipe_w = width; h->pipe_h = height; h->pipe_cap = cap; h->pipe_chunk = chunk; h->pipe_cn = cn;
    const size_t frame_bytes = (size_t) width * height;
    const bool dense_rows = row_stride == (size_t) width, dense_frames = dense_rows && frame_stride == frame_bytes;
    const bool inplace = width % 16 == 0;
    const LevelGeom &L0 = h->g.lv[0];
    cudaStream_t su = h->s_up, sd = h->s_down;
    // Two passes in flight: even chunks run on this handle's arena and stream, odd chunks on the peer's (ORBFE_NO_PEER=1 disables)
    static const bool no_peer = [] { const char *e = getenv("ORBFE_NO_PEER"); return e && *e == '1'; }();
    if (!no_peer && n_frames > cn) {
        if (!h->peer) {
            orbfe_config pc = h->cfg; pc.max_batch = cn;
            if ((rc = orbfe_create(&pc, &h->peer))) return set_error(h, rc, "pipeline peer: %s", orbfe_last_error(nullptr));
        }
        if ((rc = configure(h->peer, width, height, cn))) return set_error(h, rc, "pipeline peer: %s", orbfe_last_error(h->peer));
        h->peer->use_tma = h->use_tma;
    }
    Handle *const hp[2] = {h, (!no_peer && n_frames > cn) ? (Handle *) h->peer : h};
    // Chunk schedule.  A pass costs about 0.13 ms + 5.9 us per frame and an upload 6.5 us per frame (measured, 752x480 on PCIe 5):
    // uploads and passes are balanced, so the call takes about the upload time of the batch plus the first upload and the last pass.
    std::vector<int> sizes;
    if (const char *e = getenv("ORBFE_SCHED")) {              // explicit chunk sizes "a,b,c,..." (tuning aid; the rest is filled with cn)
        int left = n_frames;
        for (const char *p = e; *p && left > 0;) {
            const int v = std::min(std::min(atoi(p), cn), left);
            if (v > 0) { sizes.push_back(v); left -= v; }
            while (*p && *p != ',') ++p;
            if (*p == ',') ++p;
        }
        while (left > 0) { const int v = std::min(cn, left); sizes.push_back(v); left -= v; }
    } else {
        // short chunks at both ends (the first upload and the last pass + download overlap with nothing), full ones in between;
        // measured on 512 frames with 128-frame chunks: 64,128,128,96,64,32 (two passes in flight, three staging slots) is the
        // best of the variants tried (tools/e2e_probe.py with ORBFE_SCHED)
        const int head[1] = {std::max(1, cn / 2)}, tail[3] = {std::max(1, 3 * cn / 4), std::max(1, cn / 2), std::max(1, cn / 4)};
        // A submitted batch (orbfe_extract_batch_submit) is one of a stream: its ends overlap with its neighbours, and full chunks
        // throughout are best (512 frames, two batches in flight: 147.7 k frames/s against 137.9 k with the short ends)
        const bool short_ends = blocking && n_frames >= 3 * cn;
        int left = n_frames;
        if (short_ends) {
            for (int v : head) { sizes.push_back(v); left -= v; }
            left -= tail[0] + tail[1] + tail[2];
        }
        while (left > 0) { const int v = std::min(cn, left); sizes.push_back(v); left -= v; }
        if (short_ends) for (int v : tail) sizes.push_back(v);
    }
    // ORBFE_TRACE=1: per-chunk completion times of upload / pass / download (ms since the first upload was issued), on stderr
    static const bool trace_env = [] { const char *e = getenv("ORBFE_TRACE"); return e && *e == '1'; }();
    const bool trace = trace_env && blocking;
    std::vector<cudaEvent_t> tev;
    if (trace) {
        tev.resize(3 * sizes.size() + 1);
        for (auto &e : tev) cudaEventCreate(&e);
        cudaEventRecord(tev[0], su);
    }
    h->pipe_pending = true;
    h->pipe_peer_used = h->pipe_peer_used || hp[1] != h;
    int c = 0;
    for (int b0 = 0, nb = 0; b0 < n_frames; b0 += nb, ++c) {
        nb = sizes[(size_t) c];
        const long long seq = h->pipe_seq++;                   // chunk number since the handle was created
        const int slot = (int) (seq % Handle::kStageSlots);    // staging + output slot; the arena / stream alternates with the chunk parity
        Handle *const hc = hp[seq & 1];
        cudaStream_t sc = hc->stream;
        const uint8_t *src = frames + (size_t) b0 * frame_stride;
        uint8_t *stage = h->d_stage[slot];
        if (seq >= Handle::kStageSlots) ORBFE_CUDA(h, cudaStreamWaitEvent(su, h->ev_done[slot], 0));    // the pass that read this slot has finished
        if (dense_frames) ORBFE_CUDA(h, cudaMemcpyAsync(stage, src, frame_bytes * nb, cudaMemcpyHostToDevice, su));
        else
            for (int b = 0; b < nb; ++b)
                ORBFE_CUDA(h, cudaMemcpy2DAsync(stage + (size_t) b * frame_bytes, width, src + (size_t) b * frame_stride, row_stride, width, height,
                                                cudaMemcpyHostToDevice, su));
        ORBFE_CUDA(h, cudaEventRecord(h->ev_up[slot], su));
        if (trace) cudaEventRecord(tev[1 + 3 * c], su);
        ORBFE_CUDA(h, cudaStreamWaitEvent(sc, h->ev_up[slot], 0));
        if (seq >= Handle::kStageSlots) ORBFE_CUDA(h, cudaStreamWaitEvent(sc, h->ev_down[slot], 0));    // the output slot has been downloaded
        orbfe_keypoint *okps = h->d_out_kps + (size_t) slot * cn * cap;
        uint8_t *odesc = h->d_out_desc + (size_t) slot * cn * cap * 32;
        int *on = h->d_out_n + (size_t) slot * cn;
        if (inplace) rc = run_pass(hc, nb, stage, width, frame_bytes, okps, odesc, on, cap, sc);
        else {
            ORBFE_CUDA(h, cudaMemcpy2DAsync(hc->d_img + L0.img_off, L0.pitch, stage, width, width, (size_t) height * nb, cudaMemcpyDeviceToDevice, sc));
            rc = run_pass(hc, nb, hc->d_img + L0.img_off, L0.pitch, L0.frame_stride, okps, odesc, on, cap, sc);
        }
        if (rc) return hc == h ? rc : set_error(h, rc, "pipeline peer: %s", orbfe_last_error(static_cast<orbfe_handle *>(hc)));
        ORBFE_CUDA(h, cudaEventRecord(h->ev_done[slot], sc));
        if (trace) cudaEventRecord(tev[2 + 3 * c], sc);
        ORBFE_CUDA(h, cudaStreamWaitEvent(sd, h->ev_done[slot], 0));
        ORBFE_CUDA(h, cudaMemcpyAsync(n_per_frame + b0, on, sizeof(int) * nb, cudaMemcpyDeviceToHost, sd));
        ORBFE_CUDA(h, cudaMemcpyAsync(kps + (size_t) b0 * cap, okps, sizeof(orbfe_keypoint) * (size_t) nb * cap, cudaMemcpyDeviceToHost, sd));
        ORBFE_CUDA(h, cudaMemcpyAsync(desc + (size_t) b0 * cap * 32, odesc, (size_t) nb * cap * 32, cudaMemcpyDeviceToHost, sd));
        ORBFE_CUDA(h, cudaEventRecord(h->ev_down[slot], sd));
        if (trace) cudaEventRecord(tev[3 + 3 * c], sd);
    }
    if (trace) {
        ORBFE_CUDA(h, cudaStreamSynchronize(sd));
        ORBFE_CUDA(h, cudaStreamSynchronize(su));
        for (size_t i = 0; i < sizes.size(); ++i) {
            float a = 0, b = 0, d = 0;
            cudaEventElapsedTime(&a, tev[0], tev[1 + 3 * i]); cudaEventElapsedTime(&b, tev[0], tev[2 + 3 * i]); cudaEventElapsedTime(&d, tev[0], tev[3 + 3 * i]);
            fprintf(stderr, "[orbfe trace] chunk %zu (%d frames): uploaded %.3f  pass done %.3f  downloaded %.3f ms\n", i, sizes[i], a, b, d);
        }
        for (auto &e : tev) cudaEventDestroy(e);
    }
    return ORBFE_OK;
}

int orbfe_extract_batch_submit(orbfe_handle *h, const uint8_t *frames, int n_frames, int width, int height, size_t row_stride, size_t frame_stride,
                               orbfe_keypoint *kps, uint8_t *desc, int cap, int *n_per_frame, long long *ticket) {
    if (!h) return ORBFE_E_ARG;
    if (!ticket) return set_error(h, ORBFE_E_ARG, "orbfe_extract_batch_submit: ticket is null");
    *ticket = -1;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    int rc = batch_enqueue(h, frames, n_frames, width, height, row_stride, frame_stride, kps, desc, cap, n_per_frame, false);
    if (rc) return rc;
    if (!h->s_down) { *ticket = h->pipe_ticket; return ORBFE_OK; }                    // nothing was enqueued (empty batch / empty image)
    // the ticket's event follows the batch's last download on the download stream; the overflow flags of both arenas ride along
    const long long t = ++h->pipe_ticket;
    const int r = (int) (t % Handle::kTickets);
    if (!h->ev_ticket[r]) {
        ORBFE_CUDA(h, cudaEventCreateWithFlags(&h->ev_ticket[r], cudaEventDisableTiming));
        if (!h->h_ticket_err) ORBFE_CUDA(h, cudaMallocHost(&h->h_ticket_err, sizeof(int) * 2 * Handle::kTickets));
    } else {
        ORBFE_CUDA(h, cudaEventSynchronize(h->ev_ticket[r]));                         // ring slot of a batch submitted kTickets ago
    }
    int *flags = h->h_ticket_err + 2 * r;
    flags[0] = flags[1] = 0;
    ORBFE_CUDA(h, cudaMemcpyAsync(flags, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, h->s_down));
    if (h->peer && h->pipe_peer_used)                                                 // (every pass is ahead of its chunk's download on s_down)
        ORBFE_CUDA(h, cudaMemcpyAsync(flags + 1, h->peer->d_err, sizeof(int), cudaMemcpyDeviceToHost, h->s_down));
    ORBFE_CUDA(h, cudaEventRecord(h->ev_ticket[r], h->s_down));
    *ticket = t;
    return ORBFE_OK;
}

int orbfe_extract_batch_wait(orbfe_handle *h, long long ticket) {
    if (!h) return ORBFE_E_ARG;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    if (ticket < 0 || ticket >= h->pipe_ticket) return pipeline_drain(h);            // the newest batch (or "everything"): drain
    if (ticket + Handle::kTickets <= h->pipe_ticket) return ORBFE_OK;                // its ring slot was synchronised when it was reused
    const int r = (int) (ticket % Handle::kTickets);
    if (!h->ev_ticket[r]) return ORBFE_OK;
    ORBFE_CUDA(h, cudaEventSynchronize(h->ev_ticket[r]));
    if (h->h_ticket_err[2 * r] || h->h_ticket_err[2 * r + 1]) return pipeline_drain(h);   // reports and clears the device flag
    return ORBFE_OK;
}

int orbfe_extract_batch(orbfe_handle *h, const uint8_t *frames, int n_frames, int width, int height, size_t row_stride, size_t frame_stride,
                        orbfe_keypoint *kps, uint8_t *desc, int cap, int *n_per_frame) {
    if (!h) return ORBFE_E_ARG;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    int rc = batch_enqueue(h, frames, n_frames, width, height, row_stride, frame_stride, kps, desc, cap, n_per_frame, true);
    if (rc) return rc;
    return pipeline_drain(h);
}

int orbfe_extract(orbfe_handle *h, const uint8_t *gray, int width, int height, size_t stride, orbfe_keypoint *kps, uint8_t *desc, int cap, int *n_out) {
    if (!h) return ORBFE_E_ARG;
    if (!n_out) return set_error(h, ORBFE_E_ARG, "n_out is null");
    *n_out = 0;
    if (!gray || width <= 0 || height <= 0) return ORBFE_OK;                          // image.empty(): outputs untouched
    if (!kps || !desc || cap < 1 || stride < (size_t) width) return set_error(h, ORBFE_E_ARG, "orbfe_extract: invalid argument (kps/desc null, cap < 1 or stride < width)");
    static const bool trace = [] { const char *e = getenv("ORBFE_TRACE"); return e && *e == '1'; }();
    const auto t0 = std::chrono::steady_clock::now();
    auto us = [&] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); };
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    int rc = pipeline_drain(h);                                  // submitted host batches share the arena and the output staging
    if (rc) return rc;
    if ((rc = configure(h, width, height, 1))) return rc;
    // device-side capacity is the handle's bound, so a too-small caller capacity is detected on the host without clobbering kps/desc
    const int dcap = h->max_kp;
    if ((rc = ensure_out_staging(h, dcap))) return rc;
    // pinned mirror of the outputs: [n, err][kps dcap][desc dcap]; one download + one synchronisation per frame
    const size_t kp_bytes = sizeof(orbfe_keypoint) * (size_t) dcap, desc_bytes = (size_t) dcap * 32;
    // ... followed by a pinned copy of a small input frame: the driver's pageable 2-D upload costs 0.1 ms more than a host
    // memcpy into pinned memory plus one DMA when the rows are not 16-byte multiples (1241x376: 0.37 -> 0.26 ms per call);
    // above 512 KB the driver's pipelined staging wins (1920x1080: 0.47 vs 0.54 ms) and is kept
    static const bool pageable_upload = [] { const char *e = getenv("ORBFE_PAGEABLE_UPLOAD"); return e && *e == '1'; }();
    const bool stage_in = !pageable_upload && (size_t) width * height <= 512 * 1024;
    const size_t in_bytes = stage_in ? (size_t) width * height : 0;
    if (h->pinned_bytes < 16 + kp_bytes + desc_bytes + in_bytes) {
        if (h->h_pinned) cudaFreeHost(h->h_pinned);
        h->h_pinned = nullptr; h->pinned_bytes = 0;
        ORBFE_CUDA(h, cudaMallocHost(&h->h_pinned, 16 + kp_bytes + desc_bytes + in_bytes));
        h->pinned_bytes = 16 + kp_bytes + desc_bytes + in_bytes;
    }
    int *p_n = (int *) h->h_pinned;
    uint8_t *p_kps = (uint8_t *) h->h_pinned + 16, *p_desc = p_kps + kp_bytes, *p_in = p_desc + desc_bytes;
    cudaStream_t st = h->stream;
    const LevelGeom &L0 = h->g.lv[0];
    if (stage_in) {
        if (stride == (size_t) width) memcpy(p_in, gray, in_bytes);
        else for (int y = 0; y < height; ++y) memcpy(p_in + (size_t) y * width, gray + (size_t) y * stride, width);
        ORBFE_CUDA(h, cudaMemcpy2DAsync(h->d_img + L0.img_off, L0.pitch, p_in, width, width, height, cudaMemcpyHostToDevice, st));
    } else {
        ORBFE_CUDA(h, cudaMemcpy2DAsync(h->d_img + L0.img_off, L0.pitch, gray, stride, width, height, cudaMemcpyHostToDevice, st));
    }
    const double t_up = us();
    // the 13 launches of a one-frame pass are replayed from a CUDA graph (their host-side issue time and the gaps between the
    // short kernels are a third of the frame's latency); stage profiling and ORBFE_DEBUG_SYNC use the plain launches
    if (h->prof || debug_sync() || no_graph()) {
        if ((rc = run_pass(h, 1, h->d_img + L0.img_off, L0.pitch, L0.frame_stride, h->d_out_kps, h->d_out_desc, h->d_out_n, h->out_cap, st))) return rc;
    } else {
        const void *key[4] = {h->d_img, h->d_blur, h->d_out_kps, h->d_out_desc};
        if (h->graph1 && (memcmp(key, h->graph1_key, sizeof key) != 0 || h->graph1_cap != h->out_cap)) drop_graph(h);
        if (!h->graph1) {
            const long long l0 = h->launches;
            cudaGraph_t gr = nullptr;
            ORBFE_CUDA(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
            rc = run_pass(h, 1, h->d_img + L0.img_off, L0.pitch, L0.frame_stride, h->d_out_kps, h->d_out_desc, h->d_out_n, h->out_cap, st);
            cudaError_t ce = cudaStreamEndCapture(st, &gr);
            if (rc) { if (gr) cudaGraphDestroy(gr); return rc; }
            if (ce != cudaSuccess) return set_error(h, ORBFE_E_CUDA, "graph capture of the one-frame pass failed: %s", cudaGetErrorString(ce));
            ce = cudaGraphInstantiate(&h->graph1, gr, 0);
            cudaGraphDestroy(gr);
            if (ce != cudaSuccess) { h->graph1 = nullptr; return set_error(h, ORBFE_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ce)); }
            memcpy(h->graph1_key, key, sizeof key); h->graph1_cap = h->out_cap;
            h->graph1_launches = h->launches - l0; h->launches = l0;
        }
        ORBFE_CUDA(h, cudaGraphLaunch(h->graph1, st));
        h->launches += h->graph1_launches;
        h->last_batch = 1;
    }
    const double t_launch = us();
    // a frame holds about n_features key points, far fewer than the staging capacity: download the count and the first
    // `guess` rows in one go, and fetch the rest only if the frame has more
    const int guess = std::min(dcap, h->cfg.n_features + 4 * h->g.n_levels + 16);
    ORBFE_CUDA(h, cudaMemcpyAsync(p_n, h->d_out_n, sizeof(int), cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(p_n + 1, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(p_kps, h->d_out_kps, sizeof(orbfe_keypoint) * (size_t) guess, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(p_desc, h->d_out_desc, (size_t) guess * 32, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaStreamSynchronize(st));
    const double t_sync = us();
    if (p_n[1] != 0) return check_device_error(h, st);                               // reports and clears the device error flag
    const int n = p_n[0];
    if (n == 0) return ORBFE_OK;                                                       // ORBExtractor.cpp:512
    if (n > cap) return set_error(h, ORBFE_E_CAPACITY, "%d key points but capacity %d", n, cap);
    if (n > guess) {
        ORBFE_CUDA(h, cudaMemcpyAsync(p_kps + sizeof(orbfe_keypoint) * (size_t) guess, h->d_out_kps + guess, sizeof(orbfe_keypoint) * (size_t) (n - guess), cudaMemcpyDeviceToHost, st));
        ORBFE_CUDA(h, cudaMemcpyAsync(p_desc + (size_t) guess * 32, h->d_out_desc + (size_t) guess * 32, (size_t) (n - guess) * 32, cudaMemcpyDeviceToHost, st));
        ORBFE_CUDA(h, cudaStreamSynchronize(st));
    }
    memcpy(kps, p_kps, sizeof(orbfe_keypoint) * (size_t) n);
    memcpy(desc, p_desc, (size_t) n * 32);
    *n_out = n;
    if (trace) fprintf(stderr, "[orbfe trace] extract: upload issued %.1f us, launches issued %.1f us, results on host %.1f us, done %.1f us\n", t_up, t_launch, t_sync, us());
    return ORBFE_OK;
}

// ---------------------------------------------------------------- stage getters (parity tests)
int orbfe_level_size(orbfe_handle *h, int level, int *w, int *ht) {
    if (!h || level < 0 || level >= h->g.n_levels) return h ? set_error(h, ORBFE_E_ARG, "no geometry / bad level") : ORBFE_E_ARG;
    if (w) *w = h->g.lv[level].w;
    if (ht) *ht = h->g.lv[level].h;
    return ORBFE_OK;
}

static int get_image(orbfe_handle *h, const uint8_t *arena, int frame, int level, uint8_t *out) {
    if (!h || !out || level < 0 || level >= h->g.n_levels || frame < 0 || frame >= h->last_batch) return h ? set_error(h, ORBFE_E_ARG, "bad frame/level") : ORBFE_E_ARG;
    const LevelGeom &L = h->g.lv[level];
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    ORBFE_CUDA(h, cudaStreamSynchronize(h->stream));
    ORBFE_CUDA(h, cudaMemcpy2D(out, L.w, arena + L.img_off + (size_t) frame * L.frame_stride, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    return ORBFE_OK;
}
int orbfe_get_level_image(orbfe_handle *h, int frame, int level, uint8_t *out) { return get_image(h, h ? h->d_img : nullptr, frame, level, out); }
int orbfe_get_level_blurred(orbfe_handle *h, int frame, int level, uint8_t *out) { return get_image(h, h ? h->d_blur : nullptr, frame, level, out); }

static int get_packed(orbfe_handle *h, const uint32_t *d_src, const int *d_cnt, size_t per_frame, int off, int frame, int level, int32_t *xys, int cap, int *n) {
    if (!h || !n || level < 0 || level >= h->g.n_levels || frame < 0 || frame >= h->last_batch) return h ? set_error(h, ORBFE_E_ARG, "bad frame/level") : ORBFE_E_ARG;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    ORBFE_CUDA(h, cudaStreamSynchronize(h->stream));
    int cnt = 0;
    ORBFE_CUDA(h, cudaMemcpy(&cnt, d_cnt + frame * ORBFE_MAX_LEVELS + level, sizeof(int), cudaMemcpyDeviceToHost));
    *n = cnt;
    if (!xys || cnt == 0) return ORBFE_OK;
    if (cnt > cap) return set_error(h, ORBFE_E_CAPACITY, "%d entries but capacity %d", cnt, cap);
    std::vector<uint32_t> tmp(cnt);
    ORBFE_CUDA(h, cudaMemcpy(tmp.data(), d_src + (size_t) frame * per_frame + off, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost));
    for (int i = 0; i < cnt; ++i) { xys[3 * i] = tmp[i] & 0xfff; xys[3 * i + 1] = (tmp[i] >> 12) & 0xfff; xys[3 * i + 2] = tmp[i] >> 24; }
    return ORBFE_OK;
}
int orbfe_get_level_candidates(orbfe_handle *h, int frame, int level, int32_t *xys, int cap, int *n) {
    if (!h || level < 0 || level >= h->g.n_levels) return ORBFE_E_ARG;
    return get_packed(h, h->d_cand, h->d_ncand, h->g.cand_per_frame, h->g.lv[level].cand_off, frame, level, xys, cap, n);
}
int orbfe_get_level_keypoints(orbfe_handle *h, int frame, int level, int32_t *xys, int cap, int *n) {
    if (!h || level < 0 || level >= h->g.n_levels) return ORBFE_E_ARG;
    return get_packed(h, h->d_kp, h->d_nkp, h->g.kp_per_frame, h->g.lv[level].kp_off, frame, level, xys, cap, n);
}

}  // extern "C"
