// dist_test.cpp — the multi-GPU C-ABI (include/orbfe_dist.h) against the single-GPU C-ABI, bitwise.
// usage: dist_test <n_gpus> <frames.raw> <n_frames> <w> <h>
//   1. orbfe_extract_batch_sharded (host in / host out)            == orbfe_extract_batch on one GPU
//   2. orbfe_extract_batch_sharded_device (resident, NCCL gather)   == the same, for root = 0 and root = n_gpus - 1
//   3. orbfe_allpairs_sharded (with and without exclusion ranges)   == orbfe_hamming_allpairs_excl on one GPU
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/orbfe_dist.h"

#define CHECK(x) do { int rc__ = (x); if (rc__ != ORBFE_OK) { fprintf(stderr, "%s failed (%d): %s / %s\n", #x, rc__, orbfe_dist_last_error(dist), orbfe_last_error(single)); return 10; } } while (0)

static bool same_frames(const std::vector<int> &n, const std::vector<orbfe_keypoint> &k, const std::vector<uint8_t> &d, const std::vector<int> &n2,
                        const std::vector<orbfe_keypoint> &k2, const std::vector<uint8_t> &d2, int cap) {
    if (n != n2) return false;
    for (size_t b = 0; b < n.size(); ++b) {
        if (memcmp(&k[b * cap], &k2[b * cap], sizeof(orbfe_keypoint) * (size_t) n[b]) != 0) return false;
        if (memcmp(&d[b * cap * 32], &d2[b * cap * 32], (size_t) n[b] * 32) != 0) return false;
    }
    return true;
}

int main(int argc, char **argv) {
    if (argc != 6) return 2;
    const int world = atoi(argv[1]), B = atoi(argv[3]), w = atoi(argv[4]), h = atoi(argv[5]), cap = 1100;
    const size_t fb = (size_t) w * h;
    std::vector<uint8_t> frames(fb * B);
    FILE *f = fopen(argv[2], "rb");
    if (!f || fread(frames.data(), 1, frames.size(), f) != frames.size()) { fprintf(stderr, "cannot read %s\n", argv[2]); return 2; }
    fclose(f);
    orbfe_config cfg{1000, 1.2f, 8, 20, 7, 0, 16, 0};
    orbfe_handle *single = nullptr; orbfe_dist *dist = nullptr;
    if (orbfe_create(&cfg, &single) != ORBFE_OK) { fprintf(stderr, "orbfe_create: %s\n", orbfe_last_error(nullptr)); return 3; }
    if (orbfe_dist_init(&cfg, world, nullptr, &dist) != ORBFE_OK) { fprintf(stderr, "orbfe_dist_init: %s\n", orbfe_dist_last_error(nullptr)); return 3; }
    if (orbfe_dist_size(dist) != world) return 3;

    std::vector<int> n1(B), n2(B), n3(B);
    std::vector<orbfe_keypoint> k1((size_t) B * cap), k2((size_t) B * cap), k3((size_t) B * cap);
    std::vector<uint8_t> d1((size_t) B * cap * 32), d2((size_t) B * cap * 32), d3((size_t) B * cap * 32);
    CHECK(orbfe_extract_batch(single, frames.data(), B, w, h, (size_t) w, fb, k1.data(), d1.data(), cap, n1.data()));
    CHECK(orbfe_extract_batch_sharded(dist, frames.data(), B, w, h, (size_t) w, fb, k2.data(), d2.data(), cap, n2.data()));
    if (!same_frames(n1, k1, d1, n2, k2, d2, cap)) { fprintf(stderr, "sharded host extraction differs\n"); return 4; }
    long long total = 0; for (int v : n1) total += v;

    for (int root : {0, world - 1}) {
        std::vector<const uint8_t *> dptr((size_t) world, nullptr); std::vector<int> cnt((size_t) world, 0);
        int dev0 = 0; cudaGetDevice(&dev0);
        for (int r = 0; r < world; ++r) {
            int lo, hi; orbfe_dist_shard(B, r, world, &lo, &hi);
            cnt[(size_t) r] = hi - lo;
            if (hi == lo) continue;
            cudaSetDevice(r);
            uint8_t *p = nullptr;
            if (cudaMalloc(&p, fb * (hi - lo)) != cudaSuccess || cudaMemcpy(p, frames.data() + fb * lo, fb * (hi - lo), cudaMemcpyHostToDevice) != cudaSuccess) return 5;
            dptr[(size_t) r] = p;
        }
        cudaSetDevice(root);
        orbfe_keypoint *rk; uint8_t *rd; int *rn;
        if (cudaMalloc(&rk, sizeof(orbfe_keypoint) * (size_t) B * cap) || cudaMalloc(&rd, (size_t) B * cap * 32) || cudaMalloc(&rn, sizeof(int) * B)) return 5;
        cudaMemset(rk, 0, sizeof(orbfe_keypoint) * (size_t) B * cap); cudaMemset(rd, 0, (size_t) B * cap * 32);
        CHECK(orbfe_extract_batch_sharded_device(dist, dptr.data(), cnt.data(), w, h, (size_t) w, fb, root, rk, rd, cap, rn));
        cudaSetDevice(root);
        cudaMemcpy(k3.data(), rk, sizeof(orbfe_keypoint) * (size_t) B * cap, cudaMemcpyDeviceToHost);
        cudaMemcpy(d3.data(), rd, (size_t) B * cap * 32, cudaMemcpyDeviceToHost);
        cudaMemcpy(n3.data(), rn, sizeof(int) * B, cudaMemcpyDeviceToHost);
        if (!same_frames(n1, k1, d1, n3, k3, d3, cap)) { fprintf(stderr, "sharded device extraction (root %d) differs\n", root); return 6; }
        cudaFree(rk); cudaFree(rd); cudaFree(rn);
        for (int r = 0; r < world; ++r) if (dptr[(size_t) r]) { cudaSetDevice(r); cudaFree((void *) dptr[(size_t) r]); }
        cudaSetDevice(dev0);
    }

    // key-frame window: the descriptors of all frames back to back, every row skipping its own frame's block
    std::vector<uint8_t> table; std::vector<int32_t> excl;
    for (int b = 0; b < B; ++b) {
        const int lo = (int) (table.size() / 32);
        table.insert(table.end(), &d1[(size_t) b * cap * 32], &d1[(size_t) b * cap * 32] + (size_t) n1[b] * 32);
        for (int i = 0; i < n1[b]; ++i) { excl.push_back(lo); excl.push_back(lo + n1[b]); }
    }
    const int nq = (int) (table.size() / 32);
    std::vector<int32_t> a1(nq), a2(nq), a3(nq), b1(nq), b2(nq), b3(nq);
    for (int with_excl = 0; with_excl < 2; ++with_excl) {
        const int32_t *ex = with_excl ? excl.data() : nullptr;
        CHECK(orbfe_hamming_allpairs_excl(single, table.data(), nq, table.data(), nq, ex, a1.data(), a2.data(), a3.data()));
        CHECK(orbfe_allpairs_sharded(dist, table.data(), nq, table.data(), nq, ex, b1.data(), b2.data(), b3.data()));
        if (a1 != b1 || a2 != b2 || a3 != b3) { fprintf(stderr, "sharded all-pairs (exclusion %d) differs\n", with_excl); return 7; }
        if (with_excl) for (int i = 0; i < nq; ++i) if (b1[i] >= excl[2 * i] && b1[i] < excl[2 * i + 1]) { fprintf(stderr, "row %d matched inside its own block\n", i); return 8; }
    }
    // an error on one rank (capacity too small) comes back as an error, not as a hang
    if (orbfe_extract_batch_sharded(dist, frames.data(), B, w, h, (size_t) w, fb, k2.data(), d2.data(), 10, n2.data()) != ORBFE_E_CAPACITY) { fprintf(stderr, "capacity error not reported\n"); return 9; }
    printf("dist_test ok: %d GPUs, %d frames, %lld key points, %d x %d all-pairs\n", world, B, total, nq, nq);
    orbfe_dist_destroy(dist); orbfe_destroy(single);
    return 0;
}
