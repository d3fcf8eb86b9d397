// orbfe_dist.cpp — multi-GPU entry points (include/orbfe_dist.h): one host thread and one NCCL communicator per GPU of the group, on
// top of the single-GPU C-ABI of liborbfe.so.  Built into liborbfe_dist.so (links libnccl and liborbfe.so).
#include "../../include/orbfe_dist.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_init_error;

struct Rank {
    int device = 0;
    orbfe_handle *h = nullptr;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;          // collectives and the copies around them
    // device scratch of this rank (grown on demand)
    void *scratch = nullptr; size_t scratch_bytes = 0;
    int rc = ORBFE_OK; std::string err;
};

}  // namespace

struct orbfe_dist {
    std::vector<Rank> ranks;
    std::string err;
    // host-side agreement before every collective: a rank that failed must not leave the others waiting inside NCCL
    std::atomic<int> arrived{0}, failed{0};
    std::atomic<long long> epoch{0};
};

namespace {

int fail(orbfe_dist *d, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (d) d->err = buf; else g_init_error = buf;
    return code;
}

int rank_fail(Rank &r, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    r.rc = code; r.err = buf;
    return code;
}

#define RK_CUDA(r, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return rank_fail((r), ORBFE_E_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); } while (0)
#define RK_NCCL(r, call) do { ncclResult_t e__ = (call); if (e__ != ncclSuccess) return rank_fail((r), ORBFE_E_CUDA, "%s: %s", #call, ncclGetErrorString(e__)); } while (0)
#define RK_ORBFE(r, call) do { int e__ = (call); if (e__ != ORBFE_OK) return rank_fail((r), e__, "%s", orbfe_last_error((r).h)); } while (0)

int ensure_scratch(Rank &r, size_t bytes) {
    if (r.scratch_bytes >= bytes) return ORBFE_OK;
    RK_CUDA(r, cudaStreamSynchronize(r.stream));
    cudaFree(r.scratch); r.scratch = nullptr; r.scratch_bytes = 0;
    bytes += bytes / 4 + 4096;
    RK_CUDA(r, cudaMalloc(&r.scratch, bytes));
    r.scratch_bytes = bytes;
    return ORBFE_OK;
}

// All ranks meet here before a collective; returns true when every rank is healthy (so all of them enter the collective) and false
// when any rank has failed (so none does).  Sense-reversing counter barrier between the group's host threads.
bool all_ok(orbfe_dist *d, const Rank &r) {
    const int n = (int) d->ranks.size();
    if (n == 1) return r.rc == ORBFE_OK;
    if (r.rc != ORBFE_OK) d->failed.store(1);
    const long long e = d->epoch.load();
    if (d->arrived.fetch_add(1) + 1 == n) { d->arrived.store(0); d->epoch.fetch_add(1); }
    else while (d->epoch.load() == e) std::this_thread::yield();
    return d->failed.load() == 0;
}

// run fn(rank) on one host thread per GPU; the first failing rank's message becomes the group's
int run_all(orbfe_dist *d, const std::function<int(int)> &fn) {
    const int n = (int) d->ranks.size();
    for (auto &r : d->ranks) { r.rc = ORBFE_OK; r.err.clear(); }
    d->failed.store(0); d->arrived.store(0);
    std::vector<std::thread> th;
    for (int i = 1; i < n; ++i) th.emplace_back([&, i] { cudaSetDevice(d->ranks[(size_t) i].device); fn(i); });
    cudaSetDevice(d->ranks[0].device);
    fn(0);
    for (auto &t : th) t.join();
    for (int i = 0; i < n; ++i)
        if (d->ranks[(size_t) i].rc != ORBFE_OK) return fail(d, d->ranks[(size_t) i].rc, "rank %d (device %d): %s", i, d->ranks[(size_t) i].device, d->ranks[(size_t) i].err.c_str());
    return ORBFE_OK;
}

size_t up(size_t v) { return (v + 255) & ~(size_t) 255; }

}  // namespace

extern "C" {

void orbfe_dist_shard(int n, int rank, int world, int *lo, int *hi) {
    const int base = n / world, rem = n % world;
    const int l = rank * base + (rank < rem ? rank : rem);
    if (lo) *lo = l;
    if (hi) *hi = l + base + (rank < rem ? 1 : 0);
}

int orbfe_dist_init(const orbfe_config *cfg, int n_gpus, const int *devices, orbfe_dist **out) {
    if (!cfg || !out || n_gpus < 1) return fail(nullptr, ORBFE_E_ARG, "orbfe_dist_init: null argument or n_gpus < 1");
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) return fail(nullptr, ORBFE_E_CUDA, "no CUDA device (this library has no CPU fallback)");
    std::vector<int> devs((size_t) n_gpus);
    for (int i = 0; i < n_gpus; ++i) {
        devs[(size_t) i] = devices ? devices[i] : i;
        if (devs[(size_t) i] < 0 || devs[(size_t) i] >= n_dev) return fail(nullptr, ORBFE_E_ARG, "device %d out of range (%d devices)", devs[(size_t) i], n_dev);
        for (int j = 0; j < i; ++j) if (devs[(size_t) j] == devs[(size_t) i]) return fail(nullptr, ORBFE_E_ARG, "device %d listed twice", devs[(size_t) i]);
    }
    orbfe_dist *d = new orbfe_dist();
    d->ranks.resize((size_t) n_gpus);
    std::vector<ncclComm_t> comms((size_t) n_gpus);
    if (n_gpus > 1) {
        ncclResult_t nr = ncclCommInitAll(comms.data(), n_gpus, devs.data());
        if (nr != ncclSuccess) { fail(nullptr, ORBFE_E_CUDA, "ncclCommInitAll: %s", ncclGetErrorString(nr)); delete d; return ORBFE_E_CUDA; }
    }
    for (int i = 0; i < n_gpus; ++i) {
        Rank &r = d->ranks[(size_t) i];
        r.device = devs[(size_t) i]; r.comm = n_gpus > 1 ? comms[(size_t) i] : nullptr;
        orbfe_config c = *cfg; c.device = r.device;
        int rc = orbfe_create(&c, &r.h);
        if (rc == ORBFE_OK && (cudaSetDevice(r.device) != cudaSuccess || cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking) != cudaSuccess)) rc = ORBFE_E_CUDA;
        if (rc != ORBFE_OK) {
            fail(nullptr, rc, "rank %d (device %d): %s", i, r.device, r.h ? "stream creation failed" : orbfe_last_error(nullptr));
            orbfe_dist_destroy(d);
            return rc;
        }
    }
    *out = d;
    return ORBFE_OK;
}

void orbfe_dist_destroy(orbfe_dist *d) {
    if (!d) return;
    for (auto &r : d->ranks) {
        cudaSetDevice(r.device);
        if (r.stream) { cudaStreamSynchronize(r.stream); cudaStreamDestroy(r.stream); }
        cudaFree(r.scratch);
        if (r.comm) ncclCommDestroy(r.comm);
        if (r.h) orbfe_destroy(r.h);
    }
    delete d;
}

int orbfe_dist_size(const orbfe_dist *d) { return d ? (int) d->ranks.size() : 0; }
orbfe_handle *orbfe_dist_handle(orbfe_dist *d, int rank) { return (d && rank >= 0 && rank < (int) d->ranks.size()) ? d->ranks[(size_t) rank].h : nullptr; }
const char *orbfe_dist_last_error(const orbfe_dist *d) { return d ? d->err.c_str() : g_init_error.c_str(); }

int orbfe_extract_batch_sharded(orbfe_dist *d, const uint8_t *frames, int n_frames, int width, int height, size_t row_stride, size_t frame_stride,
                                orbfe_keypoint *kps, uint8_t *desc, int cap, int *n_per_frame) {
    if (!d) return ORBFE_E_ARG;
    if (!frames || !kps || !desc || !n_per_frame || n_frames < 0 || cap < 1) return fail(d, ORBFE_E_ARG, "orbfe_extract_batch_sharded: invalid argument");
    const int world = (int) d->ranks.size();
    return run_all(d, [&](int i) -> int {
        Rank &r = d->ranks[(size_t) i];
        int lo, hi; orbfe_dist_shard(n_frames, i, world, &lo, &hi);
        if (hi == lo) return ORBFE_OK;
        RK_ORBFE(r, orbfe_extract_batch(r.h, frames + (size_t) lo * frame_stride, hi - lo, width, height, row_stride, frame_stride, kps + (size_t) lo * cap,
                                        desc + (size_t) lo * cap * 32, cap, n_per_frame + lo));
        return ORBFE_OK;
    });
}

int orbfe_extract_batch_sharded_device(orbfe_dist *d, const uint8_t *const *d_frames, const int *n_frames, int width, int height, size_t row_stride,
                                       size_t frame_stride, int root, orbfe_keypoint *d_kps_root, uint8_t *d_desc_root, int cap, int *d_n_root) {
    if (!d) return ORBFE_E_ARG;
    const int world = (int) d->ranks.size();
    if (!d_frames || !n_frames || !d_kps_root || !d_desc_root || !d_n_root || cap < 1 || root < 0 || root >= world)
        return fail(d, ORBFE_E_ARG, "orbfe_extract_batch_sharded_device: invalid argument");
    std::vector<int> first((size_t) world + 1, 0);
    for (int i = 0; i < world; ++i) { if (n_frames[i] < 0 || (n_frames[i] && !d_frames[i])) return fail(d, ORBFE_E_ARG, "rank %d: invalid frame block", i); first[(size_t) i + 1] = first[(size_t) i] + n_frames[i]; }
    const size_t kp_row = sizeof(orbfe_keypoint) * (size_t) cap, ds_row = (size_t) cap * 32;
    return run_all(d, [&](int i) -> int {
        Rank &r = d->ranks[(size_t) i];
        const int nb = n_frames[i];
        // the root extracts straight into its block of the result; the others into their own scratch slab
        orbfe_keypoint *kp; uint8_t *ds; int *cn;
        if (i == root) { kp = d_kps_root + (size_t) first[(size_t) i] * cap; ds = d_desc_root + (size_t) first[(size_t) i] * ds_row; cn = d_n_root + first[(size_t) i]; }
        else {
            ensure_scratch(r, up(kp_row * nb) + up(ds_row * nb) + up(sizeof(int) * (size_t) nb) + 256);
            kp = (orbfe_keypoint *) r.scratch; ds = (uint8_t *) r.scratch + up(kp_row * nb); cn = (int *) (ds + up(ds_row * nb));
            if (r.rc != ORBFE_OK) { all_ok(d, r); return r.rc; }
        }
        auto extract = [&]() -> int {
            if (nb) RK_ORBFE(r, orbfe_extract_batch_device(r.h, d_frames[i], nb, width, height, row_stride, frame_stride, kp, ds, cap, cn, r.stream, 1));
            return ORBFE_OK;
        };
        extract();
        if (!all_ok(d, r)) return r.rc;
        if (world > 1) {
            // gather of the fixed-capacity slabs at the root: grouped send / recv (one group per rank, matched pairwise by NCCL)
            RK_NCCL(r, ncclGroupStart());
            if (i == root) {
                for (int p = 0; p < world; ++p) {
                    if (p == root || !n_frames[p]) continue;
                    RK_NCCL(r, ncclRecv(d_kps_root + (size_t) first[(size_t) p] * cap, kp_row * n_frames[p], ncclUint8, p, r.comm, r.stream));
                    RK_NCCL(r, ncclRecv(d_desc_root + (size_t) first[(size_t) p] * ds_row, ds_row * n_frames[p], ncclUint8, p, r.comm, r.stream));
                    RK_NCCL(r, ncclRecv(d_n_root + first[(size_t) p], (size_t) n_frames[p], ncclInt32, p, r.comm, r.stream));
                }
            } else if (nb) {
                RK_NCCL(r, ncclSend(kp, kp_row * nb, ncclUint8, root, r.comm, r.stream));
                RK_NCCL(r, ncclSend(ds, ds_row * nb, ncclUint8, root, r.comm, r.stream));
                RK_NCCL(r, ncclSend(cn, (size_t) nb, ncclInt32, root, r.comm, r.stream));
            }
            RK_NCCL(r, ncclGroupEnd());
        }
        RK_CUDA(r, cudaStreamSynchronize(r.stream));
        return ORBFE_OK;
    });
}

int orbfe_allpairs_sharded(orbfe_dist *d, const uint8_t *q, int nq, const uint8_t *t, int nt, const int32_t *excl, int32_t *best_idx, int32_t *best_dist,
                           int32_t *second_dist) {
    if (!d) return ORBFE_E_ARG;
    if (nq < 0 || nt < 0 || (nq && (!q || !best_idx || !best_dist || !second_dist)) || (nt && !t)) return fail(d, ORBFE_E_ARG, "orbfe_allpairs_sharded: invalid argument");
    if (nq == 0) return ORBFE_OK;
    const int world = (int) d->ranks.size();
    const size_t qb = up((size_t) nq * 32), tb = up((size_t) (nt > 0 ? nt : 1) * 32), eb = up((size_t) nq * 8), rb = up(sizeof(int) * (size_t) nq);
    return run_all(d, [&](int i) -> int {
        Rank &r = d->ranks[(size_t) i];
        ensure_scratch(r, qb + tb + eb + 3 * rb + 256);
        uint8_t *dq = (uint8_t *) r.scratch, *dt = dq + qb; int32_t *dex = (int32_t *) (dt + tb);
        int32_t *bi = (int32_t *) ((uint8_t *) dex + eb), *bd = (int32_t *) ((uint8_t *) bi + rb), *sd = (int32_t *) ((uint8_t *) bd + rb);
        auto upload = [&]() -> int {      // one upload, then the tables travel GPU to GPU
            if (i != 0 || r.rc != ORBFE_OK) return r.rc;
            RK_CUDA(r, cudaMemcpyAsync(dq, q, (size_t) nq * 32, cudaMemcpyHostToDevice, r.stream));
            if (nt) RK_CUDA(r, cudaMemcpyAsync(dt, t, (size_t) nt * 32, cudaMemcpyHostToDevice, r.stream));
            if (excl) RK_CUDA(r, cudaMemcpyAsync(dex, excl, (size_t) nq * 8, cudaMemcpyHostToDevice, r.stream));
            return ORBFE_OK;
        };
        upload();
        if (!all_ok(d, r)) return r.rc;
        if (world > 1) {
            RK_NCCL(r, ncclGroupStart());
            RK_NCCL(r, ncclBroadcast(dq, dq, (size_t) nq * 32, ncclUint8, 0, r.comm, r.stream));
            if (nt) RK_NCCL(r, ncclBroadcast(dt, dt, (size_t) nt * 32, ncclUint8, 0, r.comm, r.stream));
            if (excl) RK_NCCL(r, ncclBroadcast(dex, dex, (size_t) nq * 2, ncclInt32, 0, r.comm, r.stream));
            RK_NCCL(r, ncclGroupEnd());
        }
        int lo, hi; orbfe_dist_shard(nq, i, world, &lo, &hi);
        auto search = [&]() -> int {
            // row blocks start on 16-byte boundaries (32-byte rows, 8-byte exclusion pairs) as the single-GPU entry point requires
            if (hi > lo) RK_ORBFE(r, orbfe_hamming_allpairs_excl_device(r.h, dq + (size_t) lo * 32, hi - lo, dt, nt, excl ? dex + 2 * (size_t) lo : nullptr, bi + lo, bd + lo, sd + lo, r.stream, 0));
            return ORBFE_OK;
        };
        search();
        if (!all_ok(d, r)) return r.rc;
        if (world > 1) {
            RK_NCCL(r, ncclGroupStart());
            if (i == 0) {
                for (int p = 1; p < world; ++p) {
                    int plo, phi; orbfe_dist_shard(nq, p, world, &plo, &phi);
                    if (phi == plo) continue;
                    RK_NCCL(r, ncclRecv(bi + plo, (size_t) (phi - plo), ncclInt32, p, r.comm, r.stream));
                    RK_NCCL(r, ncclRecv(bd + plo, (size_t) (phi - plo), ncclInt32, p, r.comm, r.stream));
                    RK_NCCL(r, ncclRecv(sd + plo, (size_t) (phi - plo), ncclInt32, p, r.comm, r.stream));
                }
            } else if (hi > lo) {
                RK_NCCL(r, ncclSend(bi + lo, (size_t) (hi - lo), ncclInt32, 0, r.comm, r.stream));
                RK_NCCL(r, ncclSend(bd + lo, (size_t) (hi - lo), ncclInt32, 0, r.comm, r.stream));
                RK_NCCL(r, ncclSend(sd + lo, (size_t) (hi - lo), ncclInt32, 0, r.comm, r.stream));
            }
            RK_NCCL(r, ncclGroupEnd());
        }
        if (i == 0) {
            RK_CUDA(r, cudaMemcpyAsync(best_idx, bi, sizeof(int) * (size_t) nq, cudaMemcpyDeviceToHost, r.stream));
            RK_CUDA(r, cudaMemcpyAsync(best_dist, bd, sizeof(int) * (size_t) nq, cudaMemcpyDeviceToHost, r.stream));
            RK_CUDA(r, cudaMemcpyAsync(second_dist, sd, sizeof(int) * (size_t) nq, cudaMemcpyDeviceToHost, r.stream));
        }
        RK_CUDA(r, cudaStreamSynchronize(r.stream));
        return ORBFE_OK;
    });
}

}  // extern "C"
