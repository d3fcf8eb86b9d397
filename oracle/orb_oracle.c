#define _POSIX_C_SOURCE 200809L      /* clock_gettime (stage timers of orc_extract) under -std=c99 */
#include <time.h>
/*
 * orb_oracle.c — CPU oracle (plain C99) for the ORB front-end hot path.  TEST INFRASTRUCTURE ONLY:
 * see the header of orb_oracle.h for who may call this and how parity is pinned.
 *
 * Every function cites the reference lines (relative to /root/reference) it restates, or the
 * OpenCV behaviour it models.  Build with -O3 -ffp-contract=off: the float
 * expressions below must not be fused into FMAs (the reference builds without -march flags).
 */
#include "orb_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_EDGE 19        /* EDGE_THRESHOLD, ORBExtractor.cpp:15 */
#define ORC_HALF_PATCH 15  /* HALF_PATCH_SIZE, ORBExtractor.cpp:14 */
#define ORC_CELL 30        /* W, ORBExtractor.cpp:575 */
#define ORC_MAX_LEVELS 32

static const int8_t k_brief_pattern[1024] = {
#include "brief_pattern.inc"
};

/* ------------------------------------------------------------------------------------------
 * rounding helpers: cvRound = round-half-to-even (SSE cvtss2si / cvtsd2si), cvFloor, cvCeil
 * ------------------------------------------------------------------------------------------ */
int orc_round(double v) { return (int) lrint(v); }
int orc_roundf(float v) { return (int) lrintf(v); }
static int orc_floorf(float v) { int i = (int) v; return i - (v < (float) i); }
static int orc_ceilf(float v) { int i = (int) v; return i + (v > (float) i); }

/* ------------------------------------------------------------------------------------------
 * cv::resize(src, dst, Size(dw,dh), 0, 0, INTER_LINEAR) for CV_8UC1 — Appendix A1.
 * Called by the reference at ORBExtractor.cpp:565.  11-bit coefficients, horizontal pass in int32,
 * vertical pass ((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2 >> 2.
 * ------------------------------------------------------------------------------------------ */
static void resize_axis_tables(int dn, int sn, int clamp_frac, int *ofs, short *coef) {
    const double scale = 1.0 / ((double) dn / (double) sn);
    for (int d = 0; d < dn; ++d) {
        float f = (float) ((d + 0.5) * scale - 0.5);
        int s = orc_floorf(f);
        f -= (float) s;
        if (clamp_frac) {              /* the x axis zeroes the fraction at the borders */
            if (s < 0) { f = 0.f; s = 0; }
            if (s >= sn - 1) { f = 0.f; s = sn - 1; }
        }
        ofs[d] = s;
        coef[2 * d] = (short) orc_roundf((1.f - f) * 2048.f);
        coef[2 * d + 1] = (short) orc_roundf(f * 2048.f);
    }
}

void orc_resize_linear_u8(const uint8_t *src, int sw, int sh, size_t sstride,
                          uint8_t *dst, int dw, int dh, size_t dstride) {
    int *xofs = (int *) malloc(sizeof(int) * (size_t) dw);
    int *yofs = (int *) malloc(sizeof(int) * (size_t) dh);
    short *xc = (short *) malloc(sizeof(short) * 2 * (size_t) dw);
    short *yc = (short *) malloc(sizeof(short) * 2 * (size_t) dh);
    int *row0 = (int *) malloc(sizeof(int) * (size_t) dw);
    int *row1 = (int *) malloc(sizeof(int) * (size_t) dw);
    resize_axis_tables(dw, sw, 1, xofs, xc);
    resize_axis_tables(dh, sh, 0, yofs, yc);   /* the y axis clamps rows instead */
    for (int dy = 0; dy < dh; ++dy) {
        int sy0 = yofs[dy], sy1 = yofs[dy] + 1;
        if (sy0 < 0) sy0 = 0; if (sy0 > sh - 1) sy0 = sh - 1;
        if (sy1 < 0) sy1 = 0; if (sy1 > sh - 1) sy1 = sh - 1;
        const uint8_t *s0 = src + (size_t) sy0 * sstride, *s1 = src + (size_t) sy1 * sstride;
        for (int dx = 0; dx < dw; ++dx) {
            const int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            const int a0 = xc[2 * dx], a1 = xc[2 * dx + 1];
            row0[dx] = s0[sx] * a0 + s0[sx1] * a1;
            row1[dx] = s1[sx] * a0 + s1[sx1] * a1;
        }
        const int b0 = yc[2 * dy], b1 = yc[2 * dy + 1];
        uint8_t *d = dst + (size_t) dy * dstride;
        for (int dx = 0; dx < dw; ++dx)
            d[dx] = (uint8_t) ((((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2);
    }
    free(xofs); free(yofs); free(xc); free(yc); free(row0); free(row1);
}

/* ------------------------------------------------------------------------------------------
 * cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1 — Appendix A2.
 * Called at ORBExtractor.cpp:528.  8.8 fixed-point kernel, exact u16 horizontal pass,
 * vertical pass rounded with (+32768)>>16.
 * ------------------------------------------------------------------------------------------ */
static int refl101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; }
    return p;
}

void orc_gaussian_blur7_u8(const uint8_t *src, int w, int h, size_t sstride, uint8_t *dst, size_t dstride) {
    static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
    uint16_t *hb = (uint16_t *) malloc(sizeof(uint16_t) * (size_t) w * (size_t) h);
    for (int y = 0; y < h; ++y) {
        const uint8_t *s = src + (size_t) y * sstride;
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            if (x >= 3 && x < w - 3) for (int i = 0; i < 7; ++i) acc += k[i] * s[x - 3 + i];
            else for (int i = 0; i < 7; ++i) acc += k[i] * s[refl101(x - 3 + i, w)];
            hb[(size_t) y * w + x] = (uint16_t) acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t *r[7];
        for (int j = 0; j < 7; ++j) r[j] = hb + (size_t) refl101(y - 3 + j, h) * w;
        uint8_t *d = dst + (size_t) y * dstride;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 32768u;
            for (int j = 0; j < 7; ++j) acc += (uint32_t) k[j] * r[j][x];
            acc >>= 16;
            d[x] = (uint8_t) (acc > 255u ? 255u : acc);
        }
    }
    free(hb);
}

/* ------------------------------------------------------------------------------------------
 * cv::FAST(img, kps, threshold, nms) TYPE_9_16 — Appendix A3.  Called per 30-px cell at
 * ORBExtractor.cpp:601-606.  score = (max over the 16 arcs of 9 ring pixels of
 * max(min_arc(v-p), min_arc(p-v))) - 1 ; corner iff that maximum exceeds the threshold.
 * ------------------------------------------------------------------------------------------ */
static const int k_ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int k_ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

static int has_arc9(unsigned m16) {             /* 9 contiguous set bits on the 16-ring? */
    unsigned m = m16 | (m16 << 16);
    m &= m >> 1; m &= m >> 2; m &= m >> 4;     /* runs of 8 */
    m &= (m16 | (m16 << 16)) >> 8;             /* runs of 9 */
    return (m & 0xffffu) != 0;
}

/* returns m (see above) if m > t, else 0 */
static int fast_m(const uint8_t *p, const ptrdiff_t *ofs, int t) {
    const int v = p[0];
    int d0 = v - p[ofs[0]], d8 = v - p[ofs[8]];
    if (d0 <= t && d0 >= -t && d8 <= t && d8 >= -t) return 0;     /* every 9-arc holds ring 0 or ring 8 */
    int d4 = v - p[ofs[4]], d12 = v - p[ofs[12]];
    if (d4 <= t && d4 >= -t && d12 <= t && d12 >= -t) return 0;
    int d[25];
    unsigned dark = 0, bright = 0;
    for (int k = 0; k < 16; ++k) {
        d[k] = v - p[ofs[k]];
        dark |= (unsigned) (d[k] > t) << k;
        bright |= (unsigned) (d[k] < -t) << k;
    }
    if (!has_arc9(dark) && !has_arc9(bright)) return 0;
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = INT_MIN;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) { if (d[k + j] < mn) mn = d[k + j]; if (d[k + j] > mx) mx = d[k + j]; }
        if (mn > best) best = mn;
        if (-mx > best) best = -mx;
    }
    return best > t ? best : 0;
}

int orc_fast9_16(const uint8_t *img, int w, int h, size_t stride, int threshold, int nms,
                 orc_corner *out, int cap) {
    if (w < 7 || h < 7) return 0;
    ptrdiff_t ofs[16];
    for (int k = 0; k < 16; ++k) ofs[k] = (ptrdiff_t) k_ring_dy[k] * (ptrdiff_t) stride + k_ring_dx[k];
    const int iw = w - 6, ih = h - 6;                /* evaluated interior */
    /* score map with a 1-px zero frame so that NMS neighbours outside the interior read 0 */
    const int sw = iw + 2;
    int *score = (int *) calloc((size_t) sw * (size_t) (ih + 2), sizeof(int));   /* holds m (0 = not a corner) */
    for (int y = 0; y < ih; ++y) {
        const uint8_t *row = img + (size_t) (y + 3) * stride + 3;
        int *srow = score + (size_t) (y + 1) * sw + 1;
        for (int x = 0; x < iw; ++x) srow[x] = fast_m(row + x, ofs, threshold);
    }
#define SC(m) ((m) > 0 ? (m) - 1 : 0)     /* cornerScore of a corner is m-1; non-corners count as 0 */
    int n = 0;
    for (int y = 0; y < ih; ++y) {
        const int *s = score + (size_t) (y + 1) * sw + 1;
        for (int x = 0; x < iw; ++x) {
            if (s[x] <= 0) continue;
            const int c = SC(s[x]);
            if (nms) {
                if (!(c > SC(s[x - 1]) && c > SC(s[x + 1]) && c > SC(s[x - sw - 1]) && c > SC(s[x - sw]) && c > SC(s[x - sw + 1]) &&
                      c > SC(s[x + sw - 1]) && c > SC(s[x + sw]) && c > SC(s[x + sw + 1]))) continue;
            }
            if (n >= cap) { free(score); return -1; }
            out[n].x = x + 3; out[n].y = y + 3; out[n].score = c; ++n;
        }
    }
#undef SC
    free(score);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * cv::fastAtan2(y, x) — Appendix A4.  Called at ORBExtractor.cpp:41.  float32, no FMA.
 * ------------------------------------------------------------------------------------------ */
float orc_fast_atan2(float y, float x) {
    const float scale = (float) (180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float eps = (float) 2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* ------------------------------------------------------------------------------------------
 * extractor state — ORBExtractor::ORBExtractor, ORBExtractor.cpp:424-475
 * ------------------------------------------------------------------------------------------ */
struct orc_extractor {
    int n_features, n_levels, ini_th, min_th;
    float scale_factor;
    float scale[ORC_MAX_LEVELS], inv_scale[ORC_MAX_LEVELS];
    int quota[ORC_MAX_LEVELS];
    int u_max[ORC_HALF_PATCH + 1];
    /* stage outputs of the last call */
    uint8_t *img[ORC_MAX_LEVELS], *blur[ORC_MAX_LEVELS];
    int lw[ORC_MAX_LEVELS], lh[ORC_MAX_LEVELS];
    orc_corner *cand[ORC_MAX_LEVELS]; int n_cand[ORC_MAX_LEVELS];
    orc_keypoint *kp[ORC_MAX_LEVELS]; int n_kp[ORC_MAX_LEVELS];
};

static void compute_u_max(int *u_max) {   /* ORBExtractor.cpp:458-474 */
    int v, v0;
    const int v_max = orc_floorf(ORC_HALF_PATCH * sqrtf(2.f) / 2 + 1);
    const int v_min = orc_ceilf(ORC_HALF_PATCH * sqrtf(2.f) / 2);
    const double hp2 = ORC_HALF_PATCH * ORC_HALF_PATCH;
    for (v = 0; v <= v_max; ++v) u_max[v] = orc_round(sqrt(hp2 - v * v));
    for (v = ORC_HALF_PATCH, v0 = 0; v >= v_min; --v) {
        while (u_max[v0] == u_max[v0 + 1]) ++v0;
        u_max[v] = v0;
        ++v0;
    }
}

orc_extractor *orc_extractor_create(int n_features, float scale_factor, int n_levels, int ini_th_fast, int min_th_fast) {
    if (n_levels < 1 || n_levels > ORC_MAX_LEVELS) return NULL;
    orc_extractor *ex = (orc_extractor *) calloc(1, sizeof(*ex));
    ex->n_features = n_features; ex->n_levels = n_levels; ex->ini_th = ini_th_fast; ex->min_th = min_th_fast;
    ex->scale_factor = scale_factor;
    ex->scale[0] = 1.f; ex->inv_scale[0] = 1.f;
    for (int i = 1; i < n_levels; ++i) {                       /* :434-439 */
        ex->scale[i] = ex->scale[i - 1] * scale_factor;
        ex->inv_scale[i] = 1.f / ex->scale[i];
    }
    /* :443-452 — float / double mix as written in the reference */
    const float inv2 = 1.0f / (scale_factor * scale_factor);
    float desired = (float) ((double) ((float) n_features * (1 - inv2)) / (1 - pow((double) inv2, (double) n_levels)));
    int sum = 0;
    for (int l = 0; l < n_levels - 1; ++l) {
        ex->quota[l] = orc_roundf(desired);
        sum += ex->quota[l];
        desired *= inv2;
    }
    ex->quota[n_levels - 1] = n_features - sum > 1 ? n_features - sum : 1;
    compute_u_max(ex->u_max);
    return ex;
}

static void free_stage(orc_extractor *ex) {
    for (int l = 0; l < ORC_MAX_LEVELS; ++l) {
        free(ex->img[l]); free(ex->blur[l]); free(ex->cand[l]); free(ex->kp[l]);
        ex->img[l] = ex->blur[l] = NULL; ex->cand[l] = NULL; ex->kp[l] = NULL;
        ex->n_cand[l] = ex->n_kp[l] = 0;
    }
}

void orc_extractor_destroy(orc_extractor *ex) { if (ex) { free_stage(ex); free(ex); } }
int orc_extractor_quota(const orc_extractor *ex, int level) { return ex->quota[level]; }
float orc_extractor_scale(const orc_extractor *ex, int level) { return ex->scale[level]; }

/* ------------------------------------------------------------------------------------------
 * DistributeOctree + ExtractorNode::DivideNode — ORBExtractor.cpp:367-413, 640-830, written as the
 * flat procedure of SURVEY.md Appendix E.  Tie-break of the careful-phase sort (:757) is CANONICAL:
 * stable by node size (the reference's pointer tie-break is heap-dependent).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int x0, x1, y0, y1;   /* UL.x, UR.x, UL.y, BL.y */
    int *idx; int n;      /* candidate indices, in parent order */
    int no_more;
    int prev, next;       /* list links */
} onode;

typedef struct {
    onode *nodes; int n_nodes, cap_nodes;
    int head, tail, len;
} olist;

static int ol_new(olist *L) {
    if (L->n_nodes == L->cap_nodes) {
        L->cap_nodes = L->cap_nodes ? 2 * L->cap_nodes : 256;
        L->nodes = (onode *) realloc(L->nodes, sizeof(onode) * (size_t) L->cap_nodes);
    }
    onode *nd = &L->nodes[L->n_nodes];
    memset(nd, 0, sizeof(*nd));
    nd->prev = nd->next = -1;
    return L->n_nodes++;
}
static void ol_push_back(olist *L, int id) {
    L->nodes[id].prev = L->tail; L->nodes[id].next = -1;
    if (L->tail >= 0) L->nodes[L->tail].next = id; else L->head = id;
    L->tail = id; L->len++;
}
static void ol_push_front(olist *L, int id) {
    L->nodes[id].next = L->head; L->nodes[id].prev = -1;
    if (L->head >= 0) L->nodes[L->head].prev = id; else L->tail = id;
    L->head = id; L->len++;
}
static void ol_erase(olist *L, int id) {
    const int p = L->nodes[id].prev, n = L->nodes[id].next;
    if (p >= 0) L->nodes[p].next = n; else L->head = n;
    if (n >= 0) L->nodes[n].prev = p; else L->tail = p;
    L->len--;
}

typedef struct { int size, node; } size_node;

static void stable_sort_by_size(size_node *a, size_node *tmp, int n) {
    for (int width = 1; width < n; width *= 2) {
        for (int lo = 0; lo < n; lo += 2 * width) {
            int mid = lo + width < n ? lo + width : n, hi = lo + 2 * width < n ? lo + 2 * width : n;
            int i = lo, j = mid, k = lo;
            while (i < mid && j < hi) tmp[k++] = (a[j].size < a[i].size) ? a[j++] : a[i++];
            while (i < mid) tmp[k++] = a[i++];
            while (j < hi) tmp[k++] = a[j++];
        }
        memcpy(a, tmp, sizeof(size_node) * (size_t) n);
    }
}

/* split node `id` (DivideNode), push the non-empty children to the front in TL,TR,BL,BR order and
 * record the expandable ones (more than one candidate) — the block repeated at :697-731 and :765-799 */
static void divide_and_push(olist *L, int id, const orc_corner *c, size_node **vec, int *n_vec, int *cap_vec, int *n_expand) {
    const onode P = L->nodes[id];
    const int mid_x = P.x0 + (P.x1 - P.x0) / 2, mid_y = P.y0 + (P.y1 - P.y0) / 2;
    int cnt[4] = {0, 0, 0, 0};
    for (int i = 0; i < P.n; ++i) {
        const orc_corner *k = &c[P.idx[i]];
        cnt[(k->x < mid_x ? 0 : 1) + (k->y < mid_y ? 0 : 2)]++;
    }
    int child[4];
    for (int q = 0; q < 4; ++q) {
        child[q] = -1;
        if (!cnt[q]) continue;
        child[q] = ol_new(L);
        onode *ch = &L->nodes[child[q]];
        ch->x0 = (q & 1) ? mid_x : P.x0; ch->x1 = (q & 1) ? P.x1 : mid_x;
        ch->y0 = (q & 2) ? mid_y : P.y0; ch->y1 = (q & 2) ? P.y1 : mid_y;
        ch->idx = (int *) malloc(sizeof(int) * (size_t) cnt[q]);
        ch->n = 0;
    }
    for (int i = 0; i < P.n; ++i) {
        const orc_corner *k = &c[P.idx[i]];
        onode *ch = &L->nodes[child[(k->x < mid_x ? 0 : 1) + (k->y < mid_y ? 0 : 2)]];
        ch->idx[ch->n++] = P.idx[i];
    }
    for (int q = 0; q < 4; ++q) {
        if (child[q] < 0) continue;
        onode *ch = &L->nodes[child[q]];
        ch->no_more = ch->n == 1;
        ol_push_front(L, child[q]);
        if (ch->n > 1) {
            (*n_expand)++;
            if (*n_vec == *cap_vec) { *cap_vec = *cap_vec ? 2 * *cap_vec : 256; *vec = (size_node *) realloc(*vec, sizeof(size_node) * (size_t) *cap_vec); }
            (*vec)[*n_vec].size = ch->n; (*vec)[*n_vec].node = child[q]; (*n_vec)++;
        }
    }
}

int orc_distribute_octree(const orc_corner *c, int n, int min_x, int max_x, int min_y, int max_y,
                          int n_features, int *out_idx, int cap) {
    olist L; memset(&L, 0, sizeof(L)); L.head = L.tail = -1;
    const int n_ini = orc_ceilf((float) (max_x - min_x) / (float) (max_y - min_y));   /* :645 */
    const int h_x = orc_ceilf((float) (max_x - min_x) / (float) n_ini);               /* :646 */
    int *root_cnt = (int *) calloc((size_t) (n_ini > 0 ? n_ini : 1), sizeof(int));
    if (n_ini <= 0) { free(root_cnt); return 0; }
    for (int i = 0; i < n; ++i) root_cnt[c[i].x / h_x]++;
    int *roots = (int *) malloc(sizeof(int) * (size_t) n_ini);
    for (int i = 0; i < n_ini; ++i) {                                                  /* :652-670 */
        roots[i] = ol_new(&L);
        onode *nd = &L.nodes[roots[i]];
        nd->x0 = h_x * i; nd->x1 = (i == n_ini - 1) ? max_x : h_x * (i + 1);
        nd->y0 = 0; nd->y1 = max_y - min_y;
        nd->idx = (int *) malloc(sizeof(int) * (size_t) (root_cnt[i] ? root_cnt[i] : 1));
        nd->n = 0;
        ol_push_back(&L, roots[i]);
    }
    for (int i = 0; i < n; ++i) { onode *nd = &L.nodes[roots[c[i].x / h_x]]; nd->idx[nd->n++] = i; }   /* :673-675 */
    for (int i = 0; i < n_ini; ++i) {                                                  /* :677-686 */
        onode *nd = &L.nodes[roots[i]];
        if (nd->n == 1) nd->no_more = 1; else if (nd->n == 0) ol_erase(&L, roots[i]);
    }
    free(root_cnt); free(roots);

    size_node *vec = NULL, *pre = NULL, *tmp = NULL; int n_vec = 0, cap_vec = 0, cap_pre = 0;
    int finish = 0;
    while (!finish) {                                                                  /* :692-810 */
        int prev = L.len, n_expand = 0;
        n_vec = 0;
        for (int it = L.head; it >= 0;) {                       /* breadth pass; new children sit before `it` */
            if (L.nodes[it].no_more) { it = L.nodes[it].next; continue; }
            divide_and_push(&L, it, c, &vec, &n_vec, &cap_vec, &n_expand);
            const int nxt = L.nodes[it].next;
            ol_erase(&L, it);
            it = nxt;
        }
        if (L.len > n_features || L.len == prev) finish = 1;                           /* :750 */
        else if (L.len + n_expand * 3 > n_features) {                                  /* :752 */
            while (!finish) {
                prev = L.len;
                if (n_vec > cap_pre) { cap_pre = n_vec; pre = (size_node *) realloc(pre, sizeof(size_node) * (size_t) cap_pre); tmp = (size_node *) realloc(tmp, sizeof(size_node) * (size_t) cap_pre); }
                const int n_pre = n_vec;
                memcpy(pre, vec, sizeof(size_node) * (size_t) n_pre);
                n_vec = 0;
                stable_sort_by_size(pre, tmp, n_pre);                                  /* :757, canonical ties */
                for (int j = 0; j < n_pre; ++j) {                                      /* smallest first */
                    divide_and_push(&L, pre[j].node, c, &vec, &n_vec, &cap_vec, &n_expand);
                    ol_erase(&L, pre[j].node);
                    if (L.len >= n_features) break;                                    /* :802 */
                }
                if (L.len >= n_features || L.len == prev) finish = 1;                  /* :806 */
            }
        }
    }
    int n_out = 0;
    for (int it = L.head; it >= 0; it = L.nodes[it].next) {                            /* :813-827 */
        const onode *nd = &L.nodes[it];
        int best = nd->idx[0];
        for (int k = 1; k < nd->n; ++k) if (c[nd->idx[k]].score > c[best].score) best = nd->idx[k];
        if (n_out < cap) out_idx[n_out] = best;
        ++n_out;
    }
    for (int i = 0; i < L.n_nodes; ++i) free(L.nodes[i].idx);
    free(L.nodes); free(vec); free(pre); free(tmp);
    return n_out;
}

/* ------------------------------------------------------------------------------------------
 * IC_Angle — ORBExtractor.cpp:18-42
 * ------------------------------------------------------------------------------------------ */
static float ic_angle_umax(const uint8_t *img, size_t stride, int x, int y, const int *u_max) {
    int m_01 = 0, m_10 = 0;
    const uint8_t *center = img + (size_t) y * stride + x;
    const ptrdiff_t step = (ptrdiff_t) stride;
    for (int u = -ORC_HALF_PATCH; u <= ORC_HALF_PATCH; ++u) m_10 += u * center[u];
    for (int v = 1; v <= ORC_HALF_PATCH; ++v) {
        int v_sum = 0;
        const int d = u_max[v];
        for (int u = -d; u <= d; ++u) {
            const int plus = center[u + v * step], minus = center[u - v * step];
            v_sum += plus - minus;
            m_10 += u * (plus + minus);
        }
        m_01 += v * v_sum;
    }
    return orc_fast_atan2((float) m_01, (float) m_10);
}

float orc_ic_angle(const uint8_t *img, size_t stride, int x, int y) {
    int u_max[ORC_HALF_PATCH + 1];
    compute_u_max(u_max);
    return ic_angle_umax(img, stride, x, y, u_max);
}

/* ------------------------------------------------------------------------------------------
 * computeOrbDescriptor — ORBExtractor.cpp:50-97.  cos/sin are libm's float functions, exactly
 * what the reference calls (`using namespace std` + float argument).
 * ------------------------------------------------------------------------------------------ */
void orc_brief_descriptor(const uint8_t *blurred, size_t stride, int x, int y, float angle_deg, uint8_t *desc32) {
    const float factor_pi = (float) (3.1415926535897932384626433832795 / 180.f);
    const float angle = angle_deg * factor_pi;
    const float a = cosf(angle), b = sinf(angle);
    const uint8_t *center = blurred + (size_t) y * stride + x;
    const int step = (int) stride;
    const int8_t *p = k_brief_pattern;
    for (int i = 0; i < 32; ++i) {
        int val = 0;
        for (int j = 0; j < 8; ++j, p += 4) {
            const int t0 = center[orc_roundf(p[0] * b + p[1] * a) * step + orc_roundf(p[0] * a - p[1] * b)];
            const int t1 = center[orc_roundf(p[2] * b + p[3] * a) * step + orc_roundf(p[2] * a - p[3] * b)];
            val |= (t0 < t1) << j;
        }
        desc32[i] = (uint8_t) val;
    }
}

/* ------------------------------------------------------------------------------------------
 * operator() — ORBExtractor.cpp:495-547 with ComputePyramid :559-570 and
 * ComputeKeyPointsOctTree :572-638
 * ------------------------------------------------------------------------------------------ */
/* Wall-clock split of the last orc_extract call on this thread: [0] pyramid (resize), [1] FAST per cell, [2] Gaussian blur,
 * [3] everything else (quadtree, orientation, descriptors, copies).  bench.py's cv2 composite uses [3] as the part of a frame that
 * OpenCV's SIMD primitives do not speed up. */
static __thread double g_stage_ms[4];
static double now_ms(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; }
void orc_last_stage_ms(double out[4]) { for (int i = 0; i < 4; ++i) out[i] = g_stage_ms[i]; }

int orc_extract(orc_extractor *ex, const uint8_t *img, int w, int h, size_t stride,
                orc_keypoint *kps, uint8_t *desc, int cap) {
    if (!img || w <= 0 || h <= 0) return 0;                                          /* :497 */
    const double t_begin = now_ms();
    double t_pyr = 0, t_fast = 0, t_blur = 0, t0;
    g_stage_ms[0] = g_stage_ms[1] = g_stage_ms[2] = g_stage_ms[3] = 0;
    free_stage(ex);
    const int nl = ex->n_levels;
    /* pyramid: level l is resized from level l-1 to a size derived from the ORIGINAL size (:563-565) */
    for (int l = 0; l < nl; ++l) {
        if (l == 0) { ex->lw[0] = w; ex->lh[0] = h; }
        else { ex->lw[l] = orc_roundf((float) w * ex->inv_scale[l]); ex->lh[l] = orc_roundf((float) h * ex->inv_scale[l]); }
        ex->img[l] = (uint8_t *) malloc((size_t) ex->lw[l] * (size_t) ex->lh[l]);
        if (l == 0) for (int y = 0; y < h; ++y) memcpy(ex->img[0] + (size_t) y * w, img + (size_t) y * stride, (size_t) w);
        else { t0 = now_ms(); orc_resize_linear_u8(ex->img[l - 1], ex->lw[l - 1], ex->lh[l - 1], (size_t) ex->lw[l - 1], ex->img[l], ex->lw[l], ex->lh[l], (size_t) ex->lw[l]); t_pyr += now_ms() - t0; }
    }
    int total = 0;
    for (int l = 0; l < nl; ++l) {
        const int lw = ex->lw[l], lh = ex->lh[l];
        const int min_bx = ORC_EDGE, min_by = ORC_EDGE, max_bx = lw - ORC_EDGE, max_by = lh - ORC_EDGE;   /* :578-581 */
        const int width = max_bx - min_bx, height = max_by - min_by;
        if (width <= 0 || height <= 0) continue;
        const int n_cols = width % ORC_CELL == 0 ? width / ORC_CELL : width / ORC_CELL + 1;              /* :589-590 */
        const int n_rows = height % ORC_CELL == 0 ? height / ORC_CELL : height / ORC_CELL + 1;
        int cap_c = 1024, n_c = 0;
        orc_corner *cand = (orc_corner *) malloc(sizeof(orc_corner) * (size_t) cap_c);
        orc_corner cell[36 * 36];
        t0 = now_ms();
        for (int i = 0; i < n_rows; ++i) {
            const int ini_y = min_by + i * ORC_CELL, max_y = ini_y + ORC_CELL < max_by ? ini_y + ORC_CELL : max_by;
            for (int j = 0; j < n_cols; ++j) {
                const int ini_x = min_bx + j * ORC_CELL, max_x = ini_x + ORC_CELL < max_bx ? ini_x + ORC_CELL : max_bx;
                const uint8_t *sub = ex->img[l] + (size_t) (ini_y - 3) * lw + (ini_x - 3);
                int nc = orc_fast9_16(sub, max_x - ini_x + 6, max_y - ini_y + 6, (size_t) lw, ex->ini_th, 1, cell, 36 * 36);   /* :601 */
                if (nc == 0) nc = orc_fast9_16(sub, max_x - ini_x + 6, max_y - ini_y + 6, (size_t) lw, ex->min_th, 1, cell, 36 * 36); /* :604-607 */
                for (int k = 0; k < nc; ++k) {                                                          /* :609-615 */
                    if (n_c == cap_c) { cap_c *= 2; cand = (orc_corner *) realloc(cand, sizeof(orc_corner) * (size_t) cap_c); }
                    cand[n_c].x = cell[k].x + j * ORC_CELL - 3;
                    cand[n_c].y = cell[k].y + i * ORC_CELL - 3;
                    cand[n_c].score = cell[k].score;
                    ++n_c;
                }
            }
        }
        t_fast += now_ms() - t0;
        ex->cand[l] = cand; ex->n_cand[l] = n_c;
        int *sel = (int *) malloc(sizeof(int) * (size_t) (n_c > 0 ? n_c : 1));
        const int n_sel = n_c ? orc_distribute_octree(cand, n_c, min_bx, max_bx, min_by, max_by, ex->quota[l], sel, n_c) : 0;  /* :622 */
        ex->kp[l] = (orc_keypoint *) malloc(sizeof(orc_keypoint) * (size_t) (n_sel > 0 ? n_sel : 1));
        ex->n_kp[l] = n_sel;
        for (int k = 0; k < n_sel; ++k) {                                                               /* :625-632 */
            orc_keypoint *kp = &ex->kp[l][k];
            kp->x = (float) cand[sel[k]].x + (float) min_bx;
            kp->y = (float) cand[sel[k]].y + (float) min_by;
            kp->size = ex->scale[l];
            kp->angle = -1.f;
            kp->response = (float) cand[sel[k]].score;
            kp->octave = l; kp->class_id = -1;
        }
        free(sel);
        total += n_sel;
    }
    for (int l = 0; l < nl; ++l)                                                                        /* :636-637 */
        for (int k = 0; k < ex->n_kp[l]; ++k)
            ex->kp[l][k].angle = ic_angle_umax(ex->img[l], (size_t) ex->lw[l], orc_roundf(ex->kp[l][k].x), orc_roundf(ex->kp[l][k].y), ex->u_max);
    g_stage_ms[0] = t_pyr; g_stage_ms[1] = t_fast; g_stage_ms[3] = now_ms() - t_begin - t_pyr - t_fast;
    if (total == 0) return 0;                                                                           /* :512 */
    if (total > cap) return -1;
    int off = 0;
    for (int l = 0; l < nl; ++l) {                                                                      /* :520-546 */
        if (ex->n_kp[l] == 0) continue;
        const int lw = ex->lw[l], lh = ex->lh[l];
        ex->blur[l] = (uint8_t *) malloc((size_t) lw * (size_t) lh);
        t0 = now_ms(); orc_gaussian_blur7_u8(ex->img[l], lw, lh, (size_t) lw, ex->blur[l], (size_t) lw); t_blur += now_ms() - t0;
        for (int k = 0; k < ex->n_kp[l]; ++k) {
            orc_keypoint kp = ex->kp[l][k];
            orc_brief_descriptor(ex->blur[l], (size_t) lw, orc_roundf(kp.x), orc_roundf(kp.y), kp.angle, desc + (size_t) (off + k) * 32);
            if (l != 0) { kp.x *= ex->scale[l]; kp.y *= ex->scale[l]; }
            kps[off + k] = kp;
        }
        off += ex->n_kp[l];
    }
    g_stage_ms[2] = t_blur; g_stage_ms[3] = now_ms() - t_begin - t_pyr - t_fast - t_blur;
    return total;
}

const uint8_t *orc_level_image(const orc_extractor *ex, int level, int *w, int *h, size_t *stride) {
    if (w) *w = ex->lw[level]; if (h) *h = ex->lh[level]; if (stride) *stride = (size_t) ex->lw[level];
    return ex->img[level];
}
const uint8_t *orc_level_blurred(const orc_extractor *ex, int level, int *w, int *h, size_t *stride) {
    if (w) *w = ex->lw[level]; if (h) *h = ex->lh[level]; if (stride) *stride = (size_t) ex->lw[level];
    return ex->blur[level];
}
int orc_level_candidates(const orc_extractor *ex, int level, orc_corner *out, int cap) {
    const int n = ex->n_cand[level];
    if (out) memcpy(out, ex->cand[level], sizeof(orc_corner) * (size_t) (n < cap ? n : cap));
    return n;
}
int orc_level_keypoints(const orc_extractor *ex, int level, orc_keypoint *out, int cap) {
    const int n = ex->n_kp[level];
    if (out) memcpy(out, ex->kp[level], sizeof(orc_keypoint) * (size_t) (n < cap ? n : cap));
    return n;
}

/* ==========================================================================================
 * matcher — ORBMatcher.cpp
 * ========================================================================================== */
#define TH_LOW 50          /* ORBMatcher.cpp:13 */
#define TH_HIGH 100        /* :14 */
#define HISTO_LENGTH 30    /* :15 */
#define GRID_SIZE 40       /* BasicObject/Frame.h:18 */

int orc_descriptor_distance(const uint8_t *a, const uint8_t *b) {   /* ORBMatcher.cpp:17-31, the SWAR popcount */
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t pa, pb;
        memcpy(&pa, a + 4 * i, 4); memcpy(&pb, b + 4 * i, 4);
        uint32_t v = pa ^ pb;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int) ((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

struct orc_grid {   /* grid[cx][cy] of BasicObject/Frame.cpp:32-51 as CSR, cell id = cx*rows + cy */
    int cols, rows, *off, *idx;
};

orc_grid *orc_grid_build(const orc_keypoint *kps, int n, int img_w, int img_h) {
    orc_grid *g = (orc_grid *) calloc(1, sizeof(*g));
    g->cols = img_w % GRID_SIZE == 0 ? img_w / GRID_SIZE : img_w / GRID_SIZE + 1;
    g->rows = img_h % GRID_SIZE == 0 ? img_h / GRID_SIZE : img_h / GRID_SIZE + 1;
    const int nc = g->cols * g->rows;
    g->off = (int *) calloc((size_t) nc + 1, sizeof(int));
    g->idx = (int *) malloc(sizeof(int) * (size_t) (n > 0 ? n : 1));
    int *cell = (int *) malloc(sizeof(int) * (size_t) (n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {                     /* PosInGrid, Frame.cpp:90-95 */
        const int x = orc_floorf(kps[i].x), y = orc_floorf(kps[i].y);
        if (x < 0 || x >= img_w || y < 0 || y >= img_h) { cell[i] = -1; continue; }
        cell[i] = (x / GRID_SIZE) * g->rows + y / GRID_SIZE;
        g->off[cell[i] + 1]++;
    }
    for (int c = 0; c < nc; ++c) g->off[c + 1] += g->off[c];
    int *fill = (int *) calloc((size_t) nc, sizeof(int));
    for (int i = 0; i < n; ++i) if (cell[i] >= 0) g->idx[g->off[cell[i]] + fill[cell[i]]++] = i;
    free(fill); free(cell);
    return g;
}
void orc_grid_destroy(orc_grid *g) { if (g) { free(g->off); free(g->idx); free(g); } }

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

int orc_features_in_area(const orc_grid *g, const orc_keypoint *kps, float x, float y, float r,
                         int min_level, int max_level, int strict, int *out, int cap) {
    const int min_cx = imax(0, orc_floorf(x - r) / GRID_SIZE);
    const int max_cx = imin(g->cols - 1, orc_floorf(x + r) / GRID_SIZE);
    if (min_cx > max_cx) return 0;
    const int min_cy = imax(0, orc_floorf(y - r) / GRID_SIZE);
    const int max_cy = imin(g->rows - 1, orc_floorf(y + r) / GRID_SIZE);
    if (min_cy > max_cy) return 0;
    const int check_level = min_level > 0 || max_level >= 0;
    int n = 0;
    for (int cx = min_cx; cx <= max_cx; ++cx)
        for (int cy = min_cy; cy <= max_cy; ++cy) {
            const int c = cx * g->rows + cy;
            for (int k = g->off[c]; k < g->off[c + 1]; ++k) {
                const int idx = g->idx[k];
                const orc_keypoint *kp = &kps[idx];
                if (check_level) {
                    if (kp->octave < min_level) continue;
                    if (max_level >= 0 && kp->octave > max_level) continue;
                }
                const float dx = fabsf(kp->x - x), dy = fabsf(kp->y - y);
                if (strict ? (dx < r && dy < r) : (dx <= r && dy <= r)) { if (n < cap) out[n] = idx; ++n; }
            }
        }
    return n;
}

void orc_compute_three_maxima(const int *cnt, int n_bins, int *ind1, int *ind2, int *ind3) {   /* :594-622 */
    int max1 = 0, max2 = -1, max3 = -2;
    for (int i = 0; i < n_bins; ++i) {
        const int n = cnt[i];
        if (n > max1) { max3 = max2; max2 = max1; max1 = n; *ind3 = *ind2; *ind2 = *ind1; *ind1 = i; }
        else if (n > max2) { max3 = max2; max2 = n; *ind3 = *ind2; *ind2 = i; }
        else if (n > max3) { max3 = n; *ind3 = i; }
    }
    if (max2 < max1 / 10) { *ind2 = -1; *ind3 = -1; }
    else if (max3 < max1 / 10) { *ind3 = -1; }
}

typedef struct { int *v[HISTO_LENGTH]; int n[HISTO_LENGTH], cap[HISTO_LENGTH]; } rot_hist;
static void rh_push(rot_hist *h, int bin, int val) {
    if (h->n[bin] == h->cap[bin]) { h->cap[bin] = h->cap[bin] ? 2 * h->cap[bin] : 64; h->v[bin] = (int *) realloc(h->v[bin], sizeof(int) * (size_t) h->cap[bin]); }
    h->v[bin][h->n[bin]++] = val;
}
static void rh_free(rot_hist *h) { for (int i = 0; i < HISTO_LENGTH; ++i) free(h->v[i]); }
static int rot_bin(float a1, float a2) {              /* :85-88 and its copies */
    float rot = a1 - a2;
    if (rot < 0) rot += 360;
    int bin = orc_roundf(rot * (1.f / HISTO_LENGTH));
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

int orc_search_for_initialization(const orc_keypoint *kps1, const uint8_t *desc1, int n1,
                                  const orc_keypoint *kps2, const uint8_t *desc2, int n2,
                                  int img_w, int img_h, float *pre, int *matches12,
                                  int window, float nn_ratio, int check_orientation) {   /* :33-116 */
    int n_matches = 0;
    orc_grid *g = orc_grid_build(kps2, n2, img_w, img_h);
    int *matches21 = (int *) malloc(sizeof(int) * (size_t) (n2 > 0 ? n2 : 1));
    int *matched_dist = (int *) malloc(sizeof(int) * (size_t) (n2 > 0 ? n2 : 1));
    int *cand = (int *) malloc(sizeof(int) * (size_t) (n2 > 0 ? n2 : 1));
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    for (int i = 0; i < n2; ++i) { matches21[i] = -1; matched_dist[i] = INT_MAX; }
    rot_hist rh; memset(&rh, 0, sizeof(rh));
    for (int idx1 = 0; idx1 < n1; ++idx1) {
        const int level1 = kps1[idx1].octave;
        if (level1 > 0) continue;
        const int nc = orc_features_in_area(g, kps2, pre[2 * idx1], pre[2 * idx1 + 1], (float) window, level1, level1, 0, cand, n2);
        if (nc == 0) continue;
        int best = INT_MAX - 1, best2 = INT_MAX, best_idx2 = -1;
        for (int k = 0; k < nc; ++k) {
            const int idx2 = cand[k];
            const int dist = orc_descriptor_distance(desc1 + 32 * (size_t) idx1, desc2 + 32 * (size_t) idx2);
            if (matched_dist[idx2] <= dist) continue;
            if (dist < best) { best2 = best; best = dist; best_idx2 = idx2; }
            else if (dist < best2) best2 = dist;
        }
        if (best <= TH_LOW && best < (int) lrintf((float) best2 * nn_ratio)) {   /* :74 — (int) of the long wraps like cvtss2si */
            if (matches21[best_idx2] >= 0) { matches12[matches21[best_idx2]] = -1; n_matches--; }
            matches12[idx1] = best_idx2; matches21[best_idx2] = idx1; matched_dist[best_idx2] = best;
            n_matches++;
            if (check_orientation) rh_push(&rh, rot_bin(kps1[idx1].angle, kps2[best_idx2].angle), idx1);
        }
    }
    if (check_orientation) {
        int i1 = -1, i2 = -1, i3 = -1;
        orc_compute_three_maxima(rh.n, HISTO_LENGTH, &i1, &i2, &i3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int k = 0; k < rh.n[i]; ++k) if (matches12[rh.v[i][k]] >= 0) { matches12[rh.v[i][k]] = -1; n_matches--; }
        }
    }
    for (int idx1 = 0; idx1 < n1; ++idx1)
        if (matches12[idx1] >= 0) { pre[2 * idx1] = kps2[matches12[idx1]].x; pre[2 * idx1 + 1] = kps2[matches12[idx1]].y; }
    rh_free(&rh); free(matches21); free(matched_dist); free(cand); orc_grid_destroy(g);
    return n_matches;
}

int orc_search_by_projection(const float *q_u, const float *q_v, const float *q_radius, const int *q_level,
                             const float *q_angle, const uint8_t *q_desc, const uint8_t *q_valid, int nq,
                             const orc_keypoint *kps2, const uint8_t *desc2, int n2, int img_w, int img_h,
                             const uint8_t *occupied, int *assigned, int check_orientation) {   /* :203-274, :276-348 */
    orc_grid *g = orc_grid_build(kps2, n2, img_w, img_h);
    int *cand = (int *) malloc(sizeof(int) * (size_t) (n2 > 0 ? n2 : 1));
    for (int j = 0; j < n2; ++j) assigned[j] = -1;
    rot_hist rh; memset(&rh, 0, sizeof(rh));
    int n_match = 0;
    for (int i = 0; i < nq; ++i) {
        if (!q_valid[i]) continue;
        const int nc = orc_features_in_area(g, kps2, q_u[i], q_v[i], q_radius[i], q_level[i] - 1, q_level[i] + 1, 0, cand, n2);
        if (nc == 0) continue;
        int best = TH_HIGH + 1, best_idx2 = -1;
        for (int k = 0; k < nc; ++k) {
            const int idx2 = cand[k];
            if (occupied[idx2] || assigned[idx2] >= 0) continue;          /* curFrame->map_points[idx2] != nullptr */
            const int dist = orc_descriptor_distance(q_desc + 32 * (size_t) i, desc2 + 32 * (size_t) idx2);
            if (dist < best) { best = dist; best_idx2 = idx2; }
        }
        if (best <= TH_HIGH) {
            assigned[best_idx2] = i; n_match++;
            if (check_orientation) rh_push(&rh, rot_bin(q_angle[i], kps2[best_idx2].angle), best_idx2);
        }
    }
    if (check_orientation) {
        int i1 = -1, i2 = -1, i3 = -1;
        orc_compute_three_maxima(rh.n, HISTO_LENGTH, &i1, &i2, &i3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int k = 0; k < rh.n[i]; ++k) { assigned[rh.v[i][k]] = -1; n_match--; }
        }
    }
    rh_free(&rh); free(cand); orc_grid_destroy(g);
    return n_match;
}

int orc_search_local_points(const float *q_u, const float *q_v, const float *q_radius, const int *q_level,
                            const uint8_t *q_desc, const uint8_t *q_valid, int nq,
                            const orc_keypoint *kps2, const uint8_t *desc2, int n2, int img_w, int img_h,
                            const uint8_t *occupied, int *assigned, float nn_ratio) {              /* :350-415 */
    orc_grid *g = orc_grid_build(kps2, n2, img_w, img_h);
    int *cand = (int *) malloc(sizeof(int) * (size_t) (n2 > 0 ? n2 : 1));
    for (int j = 0; j < n2; ++j) assigned[j] = -1;
    int n_match = 0;
    for (int i = 0; i < nq; ++i) {
        if (!q_valid[i]) continue;
        const int nc = orc_features_in_area(g, kps2, q_u[i], q_v[i], q_radius[i], q_level[i] - 1, q_level[i], 0, cand, n2);
        if (nc == 0) continue;
        int best = 256, best_level = -1, second = 257, second_level = -1, best_idx = -1;
        for (int k = 0; k < nc; ++k) {
            const int idx = cand[k];
            if (occupied[idx] || assigned[idx] >= 0) continue;
            const int dist = orc_descriptor_distance(q_desc + 32 * (size_t) i, desc2 + 32 * (size_t) idx);
            if (dist < best) { second = best; best = dist; second_level = best_level; best_level = kps2[idx].octave; best_idx = idx; }
            else if (dist < second) { second = dist; second_level = kps2[idx].octave; }
        }
        if (best <= TH_HIGH) {
            if (best_level == second_level && (float) best > nn_ratio * (float) second) continue;
            assigned[best_idx] = i; n_match++;
        }
    }
    free(cand); orc_grid_destroy(g);
    return n_match;
}

int orc_search_for_triangulation(const uint8_t *desc1, const float *angle1, const uint8_t *has_mp1, int n1,
                                 const int *node_id1, const int *node_off1, const int *node_idx1, int n_nodes1,
                                 const uint8_t *desc2, const float *angle2, const uint8_t *has_mp2, int n2,
                                 const int *node_id2, const int *node_off2, const int *node_idx2, int n_nodes2,
                                 int *matches12, int check_orientation) {                            /* :417-522 */
    int n_match = 0;
    uint8_t *matched2 = (uint8_t *) calloc((size_t) (n2 > 0 ? n2 : 1), 1);
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    rot_hist rh; memset(&rh, 0, sizeof(rh));
    int a = 0, b = 0;
    while (a < n_nodes1 && b < n_nodes2) {
        if (node_id1[a] == node_id2[b]) {
            for (int i = node_off1[a]; i < node_off1[a + 1]; ++i) {
                const int idx1 = node_idx1[i];
                if (has_mp1[idx1]) continue;
                int best = TH_LOW, best_idx2 = -1;
                for (int k = node_off2[b]; k < node_off2[b + 1]; ++k) {
                    const int idx2 = node_idx2[k];
                    if (matched2[idx2] || has_mp2[idx2]) continue;
                    const int dist = orc_descriptor_distance(desc1 + 32 * (size_t) idx1, desc2 + 32 * (size_t) idx2);
                    if (dist < best) { best_idx2 = idx2; best = dist; }
                }
                if (best_idx2 > 0) {                                      /* sic: index 0 is never accepted (:484) */
                    matches12[idx1] = best_idx2; matched2[best_idx2] = 1; n_match++;
                    if (check_orientation) rh_push(&rh, rot_bin(angle1[idx1], angle2[best_idx2]), idx1);
                }
            }
            ++a; ++b;
        } else if (node_id1[a] < node_id2[b]) { while (a < n_nodes1 && node_id1[a] < node_id2[b]) ++a; }   /* lower_bound */
        else { while (b < n_nodes2 && node_id2[b] < node_id1[a]) ++b; }
    }
    if (check_orientation) {
        int i1 = -1, i2 = -1, i3 = -1;
        orc_compute_three_maxima(rh.n, HISTO_LENGTH, &i1, &i2, &i3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int k = 0; k < rh.n[i]; ++k) { matches12[rh.v[i][k]] = -1; n_match--; }
        }
    }
    rh_free(&rh); free(matched2);
    (void) n1; (void) n2;
    return n_match;
}

int orc_search_by_bow(const uint8_t *desc1, const float *angle1, const uint8_t *valid1, int n1,
                      const int *node_id1, const int *node_off1, const int *node_idx1, int n_nodes1,
                      const uint8_t *desc2, const float *angle2, const uint8_t *occupied2, int n2,
                      const int *node_id2, const int *node_off2, const int *node_idx2, int n_nodes2,
                      int *assigned, float nn_ratio, int check_orientation) {                          /* ORBMatcher.cpp:118-201 */
    /* valid1[i]: key frame key point i has a map point that is not bad (:143-144); occupied2[j]: frame->map_points[j] != nullptr;
       assigned[j] = key-frame index whose map point the call writes into frame->map_points[j], or -1 */
    int n_match = 0;
    uint8_t *taken = (uint8_t *) calloc((size_t) (n2 > 0 ? n2 : 1), 1);
    for (int j = 0; j < n2; ++j) { assigned[j] = -1; taken[j] = occupied2 ? occupied2[j] : 0; }
    rot_hist rh; memset(&rh, 0, sizeof(rh));
    int a = 0, b = 0;
    while (a < n_nodes1 && b < n_nodes2) {
        if (node_id1[a] == node_id2[b]) {
            for (int i = node_off1[a]; i < node_off1[a + 1]; ++i) {
                const int idx1 = node_idx1[i];
                if (!valid1[idx1]) continue;
                int best = 256, second = 256, best_idx2 = -1;                                     /* :149 */
                for (int k = node_off2[b]; k < node_off2[b + 1]; ++k) {
                    const int idx2 = node_idx2[k];
                    if (taken[idx2]) continue;                                                    /* :151 */
                    const int dist = orc_descriptor_distance(desc1 + 32 * (size_t) idx1, desc2 + 32 * (size_t) idx2);
                    if (dist < best) { second = best; best = dist; best_idx2 = idx2; }
                    else if (dist < second) second = dist;
                }
                if (best <= TH_LOW && (float) best < nn_ratio * (float) second) {                 /* :164 */
                    taken[best_idx2] = 1; assigned[best_idx2] = idx1; n_match++;
                    if (check_orientation) rh_push(&rh, rot_bin(angle1[idx1], angle2[best_idx2]), best_idx2);
                }
            }
            ++a; ++b;
        } else if (node_id1[a] < node_id2[b]) { while (a < n_nodes1 && node_id1[a] < node_id2[b]) ++a; }   /* lower_bound */
        else { while (b < n_nodes2 && node_id2[b] < node_id1[a]) ++b; }
    }
    if (check_orientation) {
        int i1 = -1, i2 = -1, i3 = -1;
        orc_compute_three_maxima(rh.n, HISTO_LENGTH, &i1, &i2, &i3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int k = 0; k < rh.n[i]; ++k) { assigned[rh.v[i][k]] = -1; n_match--; }
        }
    }
    rh_free(&rh); free(taken);
    (void) n1;
    return n_match;
}

/* The search half of the fuse ORBMatcher::SearchByProjection(KeyFrame, mapPoints, Map*, th) (ORBMatcher.cpp:524-592): for every map
   point the adapter projected into the key frame (u, v, radius = th * scale[predictLevel], level window [predict-1, predict]) the best
   key point of KeyFrame::getFeaturesInArea (strict "< r", KeyFrame.cpp:181-211) that passes the chi-square gate (:563-564) with
   dist < TH_LOW + 1 (:560, 568).  best_idx1[i] = -1 if none.  The map-point bookkeeping that follows (:573-586) stays on the host. */
int orc_search_fuse(const float *q_u, const float *q_v, const float *q_radius, const int *q_level, const uint8_t *q_desc, const uint8_t *q_valid, int nq,
                    const orc_keypoint *kps1, const uint8_t *desc1, int n1, int img_w, int img_h, const float *square_sigmas,
                    int *best_idx1, int *best_dist) {
    orc_grid *g = orc_grid_build(kps1, n1, img_w, img_h);
    int *cand = (int *) malloc(sizeof(int) * (size_t) (n1 > 0 ? n1 : 1));
    int n_match = 0;
    for (int i = 0; i < nq; ++i) {
        best_idx1[i] = -1; best_dist[i] = TH_LOW + 1;
        if (!q_valid[i]) continue;
        const int nc = orc_features_in_area(g, kps1, q_u[i], q_v[i], q_radius[i], q_level[i] - 1, q_level[i], 1, cand, n1);
        int best = TH_LOW + 1, bi = -1;
        for (int k = 0; k < nc; ++k) {
            const orc_keypoint *kp = &kps1[cand[k]];
            const float e2 = (q_u[i] - kp->x) * (q_u[i] - kp->x) + (q_v[i] - kp->y) * (q_v[i] - kp->y);
            if ((double) e2 > 5.991 * (double) square_sigmas[kp->octave]) continue;                    /* :564 */
            const int dist = orc_descriptor_distance(desc1 + 32 * (size_t) cand[k], q_desc + 32 * (size_t) i);
            if (dist < best) { best = dist; bi = cand[k]; }
        }
        best_idx1[i] = bi; best_dist[i] = best;
        if (bi != -1) n_match++;
    }
    free(cand); orc_grid_destroy(g);
    return n_match;
}

/* MapPoint::computeDescriptor (BasicObject/MapPoint.cpp:103-152) for a batch of map points: group g owns the descriptor rows
   [off[g], off[g+1]); best[g] = index within the group of the descriptor with the least median distance to the others
   (median = sorted row [(N-1)/2], first minimum wins, bestMedian starts at 256); -1 for an empty group. */
static int cmp_int(const void *a, const void *b) { return *(const int *) a - *(const int *) b; }
void orc_compute_descriptors(const uint8_t *desc, const int *off, int n_groups, int *best) {
    for (int g = 0; g < n_groups; ++g) {
        const int n = off[g + 1] - off[g];
        best[g] = -1;
        if (n <= 0) continue;
        int *row = (int *) malloc(sizeof(int) * (size_t) n);
        int best_median = 256, best_idx = 0;
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < n; ++j) row[j] = i == j ? 0 : orc_descriptor_distance(desc + 32 * (size_t) (off[g] + i), desc + 32 * (size_t) (off[g] + j));
            qsort(row, (size_t) n, sizeof(int), cmp_int);
            const int median = row[(n - 1) / 2];
            if (median < best_median) { best_median = median; best_idx = i; }
        }
        best[g] = best_idx;
        free(row);
    }
}

void orc_hamming_allpairs(const uint8_t *q, int nq, const uint8_t *t, int nt,
                          int *best_idx, int *best_dist, int *second_dist) {
    for (int i = 0; i < nq; ++i) {
        int best = 257, second = 257, bi = -1;
        for (int j = 0; j < nt; ++j) {
            const int d = orc_descriptor_distance(q + 32 * (size_t) i, t + 32 * (size_t) j);
            if (d < best) { second = best; best = d; bi = j; } else if (d < second) second = d;
        }
        best_idx[i] = bi; best_dist[i] = best; second_dist[i] = second;
    }
}
