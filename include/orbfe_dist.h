/*
 * orbfe_dist.h — multi-GPU entry points of the B200-native ORB front-end (liborbfe_dist.so, on top of liborbfe.so + NCCL).
 *
 * The reference has no counterpart: it is a single-sequence, CPU-only tracker.  These are the two workloads of the path that shard
 * (BASELINE.json config 5): per-frame extraction over a batch (frames are independent units) and all-pairs
 * matching of a key-frame window (query rows are independent units).  One process, one host thread and one NCCL communicator per
 * GPU (ncclCommInitAll); units are split in contiguous blocks, sizes differing by at most one, in rank order.  No reduction ever
 * crosses GPUs: the only exchange steps are a broadcast of the descriptor tables before matching and a gather of the result slabs
 * to the root GPU (NCCL send / recv over NVLink).  The single-sequence tracking loop does not shard (replicas only).
 * Plain pointers and sizes; every call returns ORBFE_OK or a negative ORBFE_E_* and records a message for orbfe_dist_last_error().
 */
#ifndef ORBFE_DIST_H
#define ORBFE_DIST_H

#include "orbfe.h"

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orbfe_dist orbfe_dist;

/* n_gpus handles created from *cfg (cfg->device is ignored: GPU i of the group is devices[i], or CUDA device i when devices == NULL). */
int         orbfe_dist_init(const orbfe_config *cfg, int n_gpus, const int *devices, orbfe_dist **out);
void        orbfe_dist_destroy(orbfe_dist *d);
int         orbfe_dist_size(const orbfe_dist *d);
orbfe_handle *orbfe_dist_handle(orbfe_dist *d, int rank);           /* the rank's own handle (stream, arena) */
const char *orbfe_dist_last_error(const orbfe_dist *d);             /* d may be NULL: last error of a failed init on this thread */
/* Block of `n` units owned by `rank` of `world`: [*lo, *hi). */
void        orbfe_dist_shard(int n, int rank, int world, int *lo, int *hi);

/* orbfe_extract_batch over all GPUs of the group: frames and results in HOST memory; every GPU uploads its own block of frames and
 * downloads its own block of results (no data-path collective: nothing needs to cross GPUs). */
int orbfe_extract_batch_sharded(orbfe_dist *d, const uint8_t *frames, int n_frames, int width, int height, size_t row_stride, size_t frame_stride,
                                orbfe_keypoint *kps, uint8_t *desc, int cap, int *n_per_frame);

/* The same with the frames already resident on the GPUs: rank r holds n_frames[r] dense frames at d_frames[r] on its own device.
 * The fixed-capacity result slabs (count + cap x 28 B key points + cap x 32 B descriptors per frame) of all ranks are gathered over
 * NVLink into the root GPU's buffers in global frame order (rank 0's frames first): d_kps_root / d_desc_root / d_n_root live on
 * device `root` of the group and hold sum(n_frames) frames. */
int orbfe_extract_batch_sharded_device(orbfe_dist *d, const uint8_t *const *d_frames, const int *n_frames, int width, int height,
                                       size_t row_stride, size_t frame_stride, int root,
                                       orbfe_keypoint *d_kps_root, uint8_t *d_desc_root, int cap, int *d_n_root);

/* orbfe_hamming_allpairs_excl over all GPUs of the group (host tables in, host results out): rank 0 uploads the tables once and
 * broadcasts them with NCCL, every GPU searches its own block of query rows against the whole train table, the row-local results are
 * gathered at rank 0 with NCCL send / recv and downloaded.  excl may be NULL. */
int orbfe_allpairs_sharded(orbfe_dist *d, const uint8_t *q, int nq, const uint8_t *t, int nt, const int32_t *excl,
                           int32_t *best_idx, int32_t *best_dist, int32_t *second_dist);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* ORBFE_DIST_H */
