/*
 * orbfe.h — C-ABI of the B200-native ORB front-end (liborbfe.so).
 *
 * This is the drop-in boundary for the one data-parallel hot path of Whitby-Li/monoORBSLAM3:
 *   - ORBExtractor::operator()            modules/ORB/ORBExtractor.h:38-39, ORBExtractor.cpp:495-547
 *   - ORBMatcher::DescriptorDistance      modules/ORB/ORBMatcher.h:18,      ORBMatcher.cpp:17-31
 *   - ORBMatcher::SearchForInitialization modules/ORB/ORBMatcher.h:21-23,   ORBMatcher.cpp:33-116
 *   - ORBMatcher::SearchByProjection      modules/ORB/ORBMatcher.h:28-37,   ORBMatcher.cpp:203-415
 *   - ORBMatcher::SearchForTriangulation  modules/ORB/ORBMatcher.h:40-42,   ORBMatcher.cpp:417-522
 *   - ORBMatcher::SearchByBow             modules/ORB/ORBMatcher.h:26,      ORBMatcher.cpp:118-201
 *   - ORBMatcher::SearchByProjection(KeyFrame, mapPoints) search half   ORBMatcher.h:44-45, ORBMatcher.cpp:524-571
 *   - MapPoint::computeDescriptor         modules/BasicObject/MapPoint.cpp:103-152
 *   - Frame::Frame post-processing        modules/BasicObject/Frame.cpp:22-51 (size *= uncertainty, undistortKeyPoints, 40-px grid)
 * Plain pointers and sizes only; no C++/torch types cross this boundary.  Every entry point returns an
 * int status (ORBFE_OK or a negative ORBFE_E_*), never throws, and records a message retrievable with
 * orbfe_last_error().  One handle per host thread (a handle owns its CUDA stream, device arena and tensor
 * maps); entry points are re-entrant across handles.  There is NO CPU fallback: without a CUDA device
 * orbfe_create() fails with ORBFE_E_CUDA.
 *
 * INTEGRATION.md shows the reference-side binding (the C++ adapter classes with the reference's own
 * signatures, and the ctypes stub used by the Python host mirror).
 */
#ifndef ORBFE_H
#define ORBFE_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

#define ORBFE_OK            0
#define ORBFE_E_ARG        -1   /* invalid argument / unsupported geometry                */
#define ORBFE_E_CUDA       -2   /* CUDA runtime/driver error (see orbfe_last_error)        */
#define ORBFE_E_CAPACITY   -3   /* caller-provided output capacity too small               */
#define ORBFE_E_INTERNAL   -4   /* device-side scratch overflow (reported, never silent)   */

#define ORBFE_MAX_LEVELS   16

/* flags for orbfe_config.flags */
#define ORBFE_FLAG_NO_TMA      1u   /* stage tiles with 16-byte vector loads instead of TMA (A/B + debugging) */
#define ORBFE_FLAG_KEEP_STAGES 2u   /* keep per-stage outputs readable through orbfe_get_* (parity tests)     */

/* Layout-compatible with cv::KeyPoint (7 x 4 bytes) — what ORBExtractor::operator() fills (ORBExtractor.cpp:537-545). */
typedef struct orbfe_keypoint {
    float   x, y;       /* pt, already multiplied by scale_factors[octave] (ORBExtractor.cpp:537-542) */
    float   size;       /* = scale_factors[octave] (ORBExtractor.cpp:631; sic, not 31*scale)          */
    float   angle;      /* IC_Angle / fastAtan2 in degrees [0,360)                                    */
    float   response;   /* FAST score                                                                 */
    int32_t octave;
    int32_t class_id;   /* -1 */
} orbfe_keypoint;

/* Constructor arguments of ORBExtractor (ORBExtractor.h:29-30) plus device selection. */
typedef struct orbfe_config {
    int32_t  n_features;     /* nFeatures   (default 1000) */
    float    scale_factor;   /* scaleFactor (default 1.2f) */
    int32_t  n_levels;       /* nLevels     (default 8, <= ORBFE_MAX_LEVELS) */
    int32_t  ini_th_fast;    /* iniThFast   (default 20)   */
    int32_t  min_th_fast;    /* minThFast   (default 10; the yaml files use 7) */
    int32_t  device;         /* CUDA device ordinal */
    int32_t  max_batch;      /* frames processed per device pass (arena is sized for this); >=1 */
    uint32_t flags;          /* ORBFE_FLAG_* */
} orbfe_config;

typedef struct orbfe_handle orbfe_handle;

/* ---------------------------------------------------------------- lifecycle */
int         orbfe_create(const orbfe_config *cfg, orbfe_handle **out);
void        orbfe_destroy(orbfe_handle *h);
const char *orbfe_last_error(const orbfe_handle *h);   /* h may be NULL: last error of a failed create on this thread */
const char *orbfe_version(void);

/* Static tables of ORBExtractor (ORBExtractor.h:44-86): scale_factors[level], features per level. */
float orbfe_scale_factor(const orbfe_handle *h, int level);
int   orbfe_features_per_level(const orbfe_handle *h, int level);
/* Upper bound on keypoints returned per frame (sum over levels of quota+3). */
int   orbfe_max_keypoints(const orbfe_handle *h);

/* Pinned host memory helpers (optional; any host pointer is accepted by the entry points below). */
int   orbfe_host_alloc(void **ptr, size_t bytes);
void  orbfe_host_free(void *ptr);

/* ---------------------------------------------------------------- extractor
 * orbfe_extract == ORBExtractor::operator()(image, keyPoints, descriptors) for one CV_8UC1 image.
 *   gray/stride : host pointer to row 0 and row pitch in bytes
 *   kps, desc   : caller-allocated, capacity `cap` keypoints / cap*32 bytes
 *   n_out       : number of keypoints written.  As in the reference, an empty image or zero keypoints
 *                 returns ORBFE_OK with *n_out = 0 and leaves kps/desc untouched (ORBExtractor.cpp:497,512).
 */
int orbfe_extract(orbfe_handle *h, const uint8_t *gray, int width, int height, size_t stride,
                  orbfe_keypoint *kps, uint8_t *desc, int cap, int *n_out);

/* The same over a batch of equally-sized frames in HOST memory (frame b at frames + b*frame_stride).
 * Outputs are slabs: kps[b*cap .. b*cap+n_per_frame[b]), desc[(b*cap+i)*32 ..].  Host<->device copies
 * are part of the call (this is what bench.py's e2e times). */
int orbfe_extract_batch(orbfe_handle *h, const uint8_t *frames, int n_frames, int width, int height,
                        size_t row_stride, size_t frame_stride,
                        orbfe_keypoint *kps, uint8_t *desc, int cap, int *n_per_frame);

/* Asynchronous form for streams of batches: orbfe_extract_batch_submit enqueues the uploads, passes and downloads of one batch
 * and returns a ticket; orbfe_extract_batch_wait(ticket) blocks until that batch's outputs are in the caller's buffers and reports
 * its device errors (ticket < 0 waits for everything submitted).  A batch submitted while the previous one is in flight continues
 * the pipeline: its first uploads run under the previous batch's last passes, so a stream of batches runs at the PCIe rate without
 * the fill / drain cost each synchronous call pays.  Input and output buffers must be pinned (orbfe_host_alloc) for the copies to
 * be asynchronous and must stay valid until the batch has been waited for; at most 8 tickets are outstanding (submit blocks on
 * the oldest).  orbfe_extract_batch == submit + wait. */
int orbfe_extract_batch_submit(orbfe_handle *h, const uint8_t *frames, int n_frames, int width, int height,
                               size_t row_stride, size_t frame_stride,
                               orbfe_keypoint *kps, uint8_t *desc, int cap, int *n_per_frame, long long *ticket);
int orbfe_extract_batch_wait(orbfe_handle *h, long long ticket);

/* The same with frames and outputs already resident in DEVICE memory (what bench.py's `value` times).
 * `stream` is a cudaStream_t (NULL = the handle's own stream); the call is asynchronous on that stream
 * unless `sync` != 0.  n_frames may exceed max_batch (processed in passes). */
int orbfe_extract_batch_device(orbfe_handle *h, const uint8_t *d_frames, int n_frames, int width, int height,
                               size_t row_stride, size_t frame_stride,
                               orbfe_keypoint *d_kps, uint8_t *d_desc, int cap, int *d_n_per_frame,
                               void *stream, int sync);

/* Stage outputs of frame `frame` of the last pass (requires ORBFE_FLAG_KEEP_STAGES); used by the parity tests.
 * All copy to host memory.  Images are written densely (w bytes per row). */
int orbfe_level_size(orbfe_handle *h, int level, int *w, int *ht);
int orbfe_get_level_image(orbfe_handle *h, int frame, int level, uint8_t *out);
int orbfe_get_level_blurred(orbfe_handle *h, int frame, int level, uint8_t *out);
/* FAST candidates after per-cell NMS/threshold fallback, reference order (cellRow, cellCol, y, x):
 * xys[3*i] = x, y (relative to the 19-px border, as fed to DistributeOctree), score. */
int orbfe_get_level_candidates(orbfe_handle *h, int frame, int level, int32_t *xys, int cap, int *n);
/* keypoints selected by the quadtree, list order: x, y (level pixel coords), score. */
int orbfe_get_level_keypoints(orbfe_handle *h, int frame, int level, int32_t *xys, int cap, int *n);

/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
long long orbfe_launch_count(const orbfe_handle *h);

/* Per-stage device timing with CUDA events recorded on the launching stream inside every extractor pass.
 * Stages: 0 pyramid (n_levels-1 launches), 1 FAST+NMS, 2 quadtree, 3 blur, 4 orientation+descriptors. */
#define ORBFE_N_STAGES 5
int orbfe_profile(orbfe_handle *h, int enable);
/* Accumulated milliseconds per stage and the number of passes since the last reset. */
int orbfe_profile_read(orbfe_handle *h, float *stage_ms, int *n_passes, int reset);

/* ---------------------------------------------------------------- Frame post-processing (Frame.cpp:22-51)
 * What Frame::Frame does with the extractor output before any matcher sees it:
 *   kp.size *= camera->uncertainty(kp.pt)                 Frame.cpp:24-26; Pinhole.cpp:55-57 (1.f), Fisheye.cpp:110-112 (scale_mat lookup)
 *   camera->undistortKeyPoints(raw, key_points)           Frame.cpp:28; Pinhole.cpp:59-84 = cv::undistortPoints(pts, K, dist, noArray(), K),
 *                                                         skipped when dist[0] == 0; Fisheye.cpp:114-117 = plain copy
 *   grid[x/40][y/40].push_back(i) if PosInGrid(kp)        Frame.cpp:31-51, 90-95
 * The grid comes back as CSR: cell = cx * rows + cy (the reference's grid[cx][cy]), grid_off[cols*rows + 1], grid_idx in
 * (cell, insertion) order — exactly the order Frame::getFeaturesInArea (Frame.cpp:97-127) enumerates.
 * Limits: at most 16384 key points per frame (the reference's extractors return 1000-8000). */
#define ORBFE_CAMERA_PINHOLE 0   /* DistortionModel "radtan"      (Camera.cpp:43-44) */
#define ORBFE_CAMERA_FISHEYE 1   /* DistortionModel "equidistant" (Camera.cpp:45-46) */
typedef struct orbfe_camera {
    int32_t model;
    float   fx, fy, cx, cy;            /* CameraMatrix (CV_32F, Camera.cpp:20-21) */
    float   dist[12];                  /* Distortion, OpenCV order k1 k2 p1 p2 [k3 k4 k5 k6 s1 s2 s3 s4]; the yaml files give 4 */
    int32_t n_dist;
    const float *uncertainty_map;      /* Fisheye::scale_mat, row-major height x width floats in HOST memory, or NULL (= 1.f) */
    int32_t uncertainty_w, uncertainty_h;
} orbfe_camera;

/* GRID_COLS / GRID_ROWS of an image (Frame.cpp:33-41). */
int orbfe_grid_size(int img_w, int img_h, int *cols, int *rows);

/* One frame, host memory.  kps_raw (n, in/out): size is multiplied in place; kps_un (n, out): undistorted copy;
 * grid_off (cols*rows + 1, out), grid_idx (n, out; the first *n_in_grid entries are valid). */
int orbfe_frame_postprocess(orbfe_handle *h, const orbfe_camera *cam, orbfe_keypoint *kps_raw, int n, int img_w, int img_h,
                            orbfe_keypoint *kps_un, int32_t *grid_off, int32_t *grid_idx, int *n_in_grid);

/* A batch of frames already in DEVICE memory, laid out like the outputs of orbfe_extract_batch_device (slabs of `cap`
 * key points per frame, counts in d_n_per_frame): chained after the extractor there is no host round trip between the
 * extractor and the matcher.  d_grid_off: n_frames x (cols*rows + 1); d_grid_idx: n_frames x cap; d_n_in_grid: n_frames or NULL. */
int orbfe_frame_postprocess_device(orbfe_handle *h, const orbfe_camera *cam, orbfe_keypoint *d_kps_raw, orbfe_keypoint *d_kps_un,
                                   const int *d_n_per_frame, int n_frames, int cap, int img_w, int img_h,
                                   int32_t *d_grid_off, int32_t *d_grid_idx, int32_t *d_n_in_grid, void *stream, int sync);

/* ---------------------------------------------------------------- matcher
 * All descriptor arrays are n x 32 bytes, row-major (cv::Mat N x 32 CV_8U as produced by the extractor).
 * Limits of the window / node searches: fewer than 65536 key points in the searched frame (the greedy state of a search lives in
 * the shared memory of one CTA); SearchForInitialization about 13000.  Larger inputs return ORBFE_E_ARG, never a wrong result.
 */

/* Measured popc throughput of the device (10^9 32-bit popc per second; 8 popc = one 256-bit match): the roofline
 * denominator bench.py quotes the matching kernels against. */
int orbfe_popc_peak(orbfe_handle *h, double *gpopc_per_s);

/* Measured throughput of the int8 tensor-core instruction the large all-pairs searches run on (mma.sync m16n8k32.s8), expressed in
 * 10^9 descriptor pairs per second (256 int8 multiply-adds = one pair): the roofline denominator of that path. */
int orbfe_imma_peak(orbfe_handle *h, double *gmatch_per_s);

/* ORBMatcher::DescriptorDistance over explicit pairs: dist[i] = hamming(a[ia[i]], b[ib[i]]). */
int orbfe_descriptor_distance(orbfe_handle *h, const uint8_t *a, int na, const uint8_t *b, int nb,
                              const int32_t *ia, const int32_t *ib, int n_pairs, int32_t *dist);

/* Brute-force best / second-best over all pairs (BASELINE configs 4/5): for each query row the train index of the
 * minimum distance (first minimum wins), that distance and the second-smallest distance (257 if none). Host memory.
 * Problems of at least 256 x 256 run as an int8 GEMM on the tensor cores (hamming = (256 - <a, b>) / 2 over +-1 vectors, exact):
 * tcgen05.mma kind::i8 with TMA-staged operands and accumulators in tensor memory (csrc/orbfe_allpairs_tc.cu); smaller ones on
 * the popc kernel; the results are identical.  ORBFE_ALLPAIRS=popc|imma|tc selects a kernel for A/B runs (imma = the warp-level
 * mma.sync kernel of round 1); ORBFE_ALLPAIRS_POPC=1 is the older spelling of popc. */
int orbfe_hamming_allpairs(orbfe_handle *h, const uint8_t *q, int nq, const uint8_t *t, int nt,
                           int32_t *best_idx, int32_t *best_dist, int32_t *second_dist);
/* Device-resident variant (asynchronous on `stream` unless sync). */
int orbfe_hamming_allpairs_device(orbfe_handle *h, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt,
                                  int32_t *d_best_idx, int32_t *d_best_dist, int32_t *d_second_dist,
                                  void *stream, int sync);

/* The same with a per-query exclusion range: excl[2*i], excl[2*i+1] = train indices [lo, hi) that query i skips (lo >= hi: none).
 * This is how a key-frame window is matched against itself (BASELINE configs 4/5): the table holds the window's key frames back to
 * back and every descriptor skips the block of its own key frame, so best / second-best are its nearest neighbours in the OTHER
 * key frames (without it every row finds itself at distance 0).  excl == NULL is orbfe_hamming_allpairs. */
int orbfe_hamming_allpairs_excl(orbfe_handle *h, const uint8_t *q, int nq, const uint8_t *t, int nt, const int32_t *excl,
                                int32_t *best_idx, int32_t *best_dist, int32_t *second_dist);
int orbfe_hamming_allpairs_excl_device(orbfe_handle *h, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt, const int32_t *d_excl,
                                       int32_t *d_best_idx, int32_t *d_best_dist, int32_t *d_second_dist, void *stream, int sync);

/* A key-frame window matched against itself straight from the extractor's output slabs (orbfe_extract_batch_device): d_desc holds
 * n_frames blocks of `cap` descriptor rows of which the first d_n_per_frame[f] are key points.  Every key point's best / second-best
 * among the key points of the OTHER frames of the window; padding rows neither match nor are matched (their outputs are -1 / 257 /
 * 257).  Outputs have n_frames * cap entries and best_idx is a slab row (frame * cap + key point).  No compaction pass and no host
 * synchronisation for the counts: the call chains behind the extraction on the same stream.  Runs on the tensor cores for every size.
 * Any cap works; with cap a multiple of 128 (the kernel's train-tile height) every block starts on a tile boundary, the tiles of a row's
 * own key frame and the tiles of padding rows are skipped whole, and only one tile per block needs per-column masking. */
int orbfe_hamming_allpairs_slab_device(orbfe_handle *h, const uint8_t *d_desc, const int *d_n_per_frame, int n_frames, int cap,
                                       int32_t *d_best_idx, int32_t *d_best_dist, int32_t *d_second_dist, void *stream, int sync);

/* Best / second-best over caller-supplied candidate lists (SURVEY.md section 8b `orbfe_hamming_window`): the candidates of
 * query i are t[cand_idx[cand_offsets[i] .. cand_offsets[i+1])], scanned in list order like the `if (dist < bestDist)` loops of
 * ORBMatcher.cpp:60-72 / :237-248 (first minimum wins).  best_idx is the train index of the minimum (-1 and 257 for an
 * empty list), second_dist the second-smallest distance of the list (257 if fewer than two).  Host memory. */
int orbfe_hamming_window(orbfe_handle *h, const uint8_t *q, int nq, const uint8_t *t, int nt,
                         const int32_t *cand_offsets, const int32_t *cand_idx,
                         int32_t *best_idx, int32_t *best_dist, int32_t *second_dist);

/* ORBMatcher::SearchForInitialization (ORBMatcher.cpp:33-116).  kps are the Frames' (undistorted) key points,
 * prematched_xy (n1 x 2, in/out) is vecPreMatched, matches12 (n1, out) the result; returns the match count in
 * *n_matches.  The candidate windows follow Frame::getFeaturesInArea on the 40-px grid (Frame.cpp:97-127). */
int orbfe_search_for_initialization(orbfe_handle *h,
                                    const orbfe_keypoint *kps1, const uint8_t *desc1, int n1,
                                    const orbfe_keypoint *kps2, const uint8_t *desc2, int n2,
                                    int img_w, int img_h, float *prematched_xy, int32_t *matches12,
                                    int window, float nn_ratio, int check_orientation, int *n_matches);

/* ---------------------------------------------------------------- device-resident frames
 * What Frame::Frame leaves behind for the matchers (Frame.cpp:20-51) kept in HBM: undistorted key points, descriptors and the 40-px
 * grid.  Tracking calls the matcher several times per frame on the same frames (Tracking.cpp:284-296 calls SearchByProjection twice
 * on one frame pair, :412-425 searches the local map in the current frame): with a frame object the key points and descriptors are
 * uploaded and the grid is built once, and the orbfe_search_*_f entry points only move the queries and the result.
 *   orbfe_frame_upload       copies host arrays (n key points, n x 32 descriptor bytes) to the device and builds the grid
 *   orbfe_frame_wrap_device  borrows buffers that are already on the device — the outputs of orbfe_extract_batch_device /
 *                            orbfe_frame_postprocess_device for one frame (pass its grid, or NULL, NULL to have one built); the
 *                            buffers must outlive the frame object.  n is the frame's key-point count (d_n_per_frame[b] read back).
 * A frame belongs to the device of the handle that made it; the searches require a handle on the same device. */
typedef struct orbfe_frame orbfe_frame;
int  orbfe_frame_upload(orbfe_handle *h, const orbfe_keypoint *kps, const uint8_t *desc, int n, int img_w, int img_h, orbfe_frame **out);
int  orbfe_frame_wrap_device(orbfe_handle *h, const orbfe_keypoint *d_kps, const uint8_t *d_desc, int n, int img_w, int img_h,
                             const int32_t *d_grid_off, const int32_t *d_grid_idx, orbfe_frame **out);
void orbfe_frame_destroy(orbfe_frame *f);
int  orbfe_frame_size(const orbfe_frame *f);

/* The window searches on device-resident frames: same semantics and results as the host-array forms below. */
int orbfe_search_for_initialization_f(orbfe_handle *h, const orbfe_frame *frame1, const orbfe_frame *frame2, float *prematched_xy,
                                      int32_t *matches12, int window, float nn_ratio, int check_orientation, int *n_matches);
int orbfe_search_by_projection_f(orbfe_handle *h,
                                 const float *q_u, const float *q_v, const float *q_radius, const int32_t *q_level,
                                 const float *q_angle, const uint8_t *q_desc, const uint8_t *q_valid, int nq,
                                 const orbfe_frame *frame2, const uint8_t *occupied, int32_t *assigned, int check_orientation, int *n_matches);
int orbfe_search_local_points_f(orbfe_handle *h,
                                const float *q_u, const float *q_v, const float *q_radius, const int32_t *q_level,
                                const uint8_t *q_desc, const uint8_t *q_valid, int nq,
                                const orbfe_frame *frame2, const uint8_t *occupied, int32_t *assigned, float nn_ratio, int *n_matches);

/* ORBMatcher::SearchByProjection(Frame|KeyFrame -> Frame) (ORBMatcher.cpp:203-348) after the adapter projected the
 * map points: query i is considered iff q_valid[i]; window centre (q_u,q_v), radius q_radius (= th * kp.size),
 * octave window [q_level-1, q_level+1]; occupied[j] != 0 marks current-frame slots that already hold a map point;
 * assigned[j] (n2, out) = index of the query written into slot j, or -1. */
int orbfe_search_by_projection(orbfe_handle *h,
                               const float *q_u, const float *q_v, const float *q_radius, const int32_t *q_level,
                               const float *q_angle, const uint8_t *q_desc, const uint8_t *q_valid, int nq,
                               const orbfe_keypoint *kps2, const uint8_t *desc2, int n2, int img_w, int img_h,
                               const uint8_t *occupied, int32_t *assigned, int check_orientation, int *n_matches);

/* ORBMatcher::SearchByProjection(Frame, local map points) (ORBMatcher.cpp:350-415): octave window
 * [q_level-1, q_level], best/second-best with the same-level ratio test. */
int orbfe_search_local_points(orbfe_handle *h,
                              const float *q_u, const float *q_v, const float *q_radius, const int32_t *q_level,
                              const uint8_t *q_desc, const uint8_t *q_valid, int nq,
                              const orbfe_keypoint *kps2, const uint8_t *desc2, int n2, int img_w, int img_h,
                              const uint8_t *occupied, int32_t *assigned, float nn_ratio, int *n_matches);

/* ORBMatcher::SearchForTriangulation (ORBMatcher.cpp:417-522).  DBoW2 feature vectors are passed as CSR: ascending
 * node ids, offsets (n_nodes+1) and member key-point indices. */
int orbfe_search_for_triangulation(orbfe_handle *h,
                                   const uint8_t *desc1, const float *angle1, const uint8_t *has_mp1, int n1,
                                   const int32_t *node_id1, const int32_t *node_off1, const int32_t *node_idx1, int n_nodes1,
                                   const uint8_t *desc2, const float *angle2, const uint8_t *has_mp2, int n2,
                                   const int32_t *node_id2, const int32_t *node_off2, const int32_t *node_idx2, int n_nodes2,
                                   int32_t *matches12, int check_orientation, int *n_matches);

/* ORBMatcher::SearchByBow(KeyFrame, Frame) (ORBMatcher.cpp:118-201), SURVEY.md 8f rank 3.  Feature vectors as CSR like above.
 * valid1[i] != 0: key-frame key point i has a map point that is not bad (:143-144); occupied2[j] != 0: frame->map_points[j] is
 * already set (:151).  assigned[j] (n2, out) = key-frame key-point index whose map point the call puts into frame slot j, or -1.
 * Best / second-best with the float ratio test bestDist < nn_ratio * secondDist (:164), rotation histogram as in the reference. */
int orbfe_search_by_bow(orbfe_handle *h,
                        const uint8_t *desc1, const float *angle1, const uint8_t *valid1, int n1,
                        const int32_t *node_id1, const int32_t *node_off1, const int32_t *node_idx1, int n_nodes1,
                        const uint8_t *desc2, const float *angle2, const uint8_t *occupied2, int n2,
                        const int32_t *node_id2, const int32_t *node_off2, const int32_t *node_idx2, int n_nodes2,
                        int32_t *assigned, float nn_ratio, int check_orientation, int *n_matches);

/* Search half of the fuse ORBMatcher::SearchByProjection(KeyFrame, mapPoints, Map*, th) (ORBMatcher.cpp:524-571), SURVEY.md 8f rank 3:
 * for every map point the adapter projected into the key frame (q_valid, u, v, radius = th * scale[predictLevel], predictLevel) the best
 * key point within KeyFrame::getFeaturesInArea (strict "< r", KeyFrame.cpp:181-211; levels [predict-1, predict]) that passes the
 * chi-square gate (:563-564, square_sigmas of the handle's own scale table: scale_factor / n_levels of its orbfe_config) with dist <= TH_LOW.  best_idx1[i] = -1 if none; best_dist may
 * be NULL.  The observation / replace bookkeeping of :573-586 is the adapter's. */
int orbfe_search_fuse(orbfe_handle *h, const float *q_u, const float *q_v, const float *q_radius, const int32_t *q_level,
                      const uint8_t *q_desc, const uint8_t *q_valid, int nq,
                      const orbfe_keypoint *kps1, const uint8_t *desc1, int n1, int img_w, int img_h,
                      int32_t *best_idx1, int32_t *best_dist, int *n_matches);

/* The same with the gate's sigma table passed explicitly (ORBExtractor::getSquareSigma(octave), n_levels entries): use this form when the
 * matcher's handle was not created with the extractor's scaleFactor / nLevels.  A key point whose octave lies outside the table is
 * rejected with ORBFE_E_ARG (the reference would index past its vector). */
int orbfe_search_fuse_sigma(orbfe_handle *h, const float *q_u, const float *q_v, const float *q_radius, const int32_t *q_level,
                            const uint8_t *q_desc, const uint8_t *q_valid, int nq,
                            const orbfe_keypoint *kps1, const uint8_t *desc1, int n1, int img_w, int img_h,
                            const float *square_sigmas, int n_levels, int32_t *best_idx1, int32_t *best_dist, int *n_matches);

/* MapPoint::computeDescriptor (BasicObject/MapPoint.cpp:103-152) for a batch of map points: group g owns the descriptor rows
 * [group_off[g], group_off[g+1]) (its observations in std::map order); best[g] = index within the group of the descriptor with the
 * least median Hamming distance to the others (first minimum wins), -1 for an empty group.  At most 512 observations per point. */
int orbfe_compute_descriptors(orbfe_handle *h, const uint8_t *desc, const int32_t *group_off, int n_groups, int32_t *best);

/* ---------------------------------------------------------------- DBoW2 vocabulary descent
 * Frame::computeBow / KeyFrame::computeBow (BasicObject/Frame.cpp:168-178, KeyFrame.cpp:213-223) call
 * ORBVocabulary::transform(descriptors, bow_vector, feature_vector, 4) of the vendored DBoW2
 * (thirdParty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1172, 1217-1259).  The vocabulary is handed over the way loadFromTextFile
 * (:1338-1420) reads ORBvoc.txt: node i = 1..n_nodes-1 in file order with its parent id, leaf flag, 32-byte descriptor and
 * weight (index 0 = root, ignored).  orbfe_vocab_transform returns, per feature, the word id, the node reached at level
 * L - levelsup and the word weight, and (if the fv_* pointers are given) the FeatureVector as CSR in std::map order with
 * stopped words (weight 0) left out; the BowVector is the adapter's sum of `weight` per `word_id` in feature order followed by
 * the scoring object's normalisation (:1150-1166). */
typedef struct orbfe_vocab orbfe_vocab;
int  orbfe_vocab_create(orbfe_handle *h, int k, int L, int n_nodes, const int32_t *parent, const uint8_t *is_leaf, const uint8_t *desc,
                        const double *weight, orbfe_vocab **out);
void orbfe_vocab_destroy(orbfe_vocab *v);
int  orbfe_vocab_words(const orbfe_vocab *v);
int  orbfe_vocab_transform(orbfe_vocab *v, const uint8_t *desc, int n, int levelsup, int32_t *word_id, int32_t *node_id, double *weight,
                           int32_t *fv_node_id, int32_t *fv_off /* n + 1 */, int32_t *fv_idx /* n */, int *fv_n_nodes);

/* ---------------------------------------------------------------- RANSAC hypothesis scoring
 * TwoViewReconstruction::CheckHomography / CheckFundamental (Frontend/TwoViewReconstruction.cpp:226-288, 290-345) for all
 * hypotheses of FindHomography / FindFundamental (:86-160) at once.  Matrices are 3x3 row-major floats, n_hyp of them; H12 is the
 * caller's H21.inverse() (:227); pts1 / pts2 are the matched key points' (x, y) in match order; sigma as in the constructor
 * (TwoViewReconstruction.h:18-20).  scores[n_hyp] and inliers[n_hyp x n_matches] (may be NULL) are bit-identical to the
 * reference's float results: same operation order, no FMA, scores summed in match order. */
int orbfe_check_homography(orbfe_handle *h, const float *H21, const float *H12, int n_hyp, const float *pts1, const float *pts2, int n_matches,
                           float sigma, float *scores, uint8_t *inliers);
int orbfe_check_fundamental(orbfe_handle *h, const float *F21, int n_hyp, const float *pts1, const float *pts2, int n_matches,
                            float sigma, float *scores, uint8_t *inliers);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* ORBFE_H */
