// adapter_test.cpp — drives the C++ adapters (reference class names and signatures) end to end on the GPU.
// usage: adapter_test <w> <h> <frameA.raw> <frameB.raw> <out.bin>
// Writes: nA, kpsA (28 B each), descA, nB, kpsB, descB, nMatches, matches12[nA], vecPreMatched[nA*2] — pytest compares them with the
// Python mirror (which is parity-tested against the oracle).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>
#include "../../monoorbslam3_b200/host/ORBExtractor.h"
#include "../../monoorbslam3_b200/host/ORBMatcher.h"
#include "../../monoorbslam3_b200/host/FramePost.h"

using namespace mono_orb_slam3;

struct Frame {                                  // the members of BasicObject/Frame.h the matcher reads
    std::vector<cv::KeyPoint> key_points; cv::Mat descriptors; int num_kps = 0; cv::Mat img;
};

static cv::Mat load(const char *path, int w, int h) {
    cv::Mat m(h, w, CV_8U);
    FILE *f = fopen(path, "rb");
    if (!f || fread(m.data, 1, (size_t) w * h, f) != (size_t) w * h) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return m;
}

int main(int argc, char **argv) {
    if (argc != 6) return 2;
    const int w = atoi(argv[1]), h = atoi(argv[2]);
    ORBExtractor extractor(1000, 1.2f, 8, 20, 7);
    ORBExtractor initial(2000, extractor);                          // Tracking.cpp:24
    if (ORBExtractor::getNumLevels() != 8 || ORBExtractor::getScaleFactor(1) != 1.2f) return 3;
    auto f1 = std::make_shared<Frame>(), f2 = std::make_shared<Frame>();
    f1->img = load(argv[3], w, h); f2->img = load(argv[4], w, h);
    initial(f1->img, f1->key_points, f1->descriptors); f1->num_kps = (int) f1->key_points.size();
    initial(f2->img, f2->key_points, f2->descriptors); f2->num_kps = (int) f2->key_points.size();
    std::vector<cv::KeyPoint> k3; cv::Mat d3;
    extractor(f1->img, k3, d3);                                     // the 1000-feature extractor on the same image
    std::vector<std::vector<cv::KeyPoint>> bk; std::vector<cv::Mat> bd;
    extractor.extractBatch({f1->img, f2->img, f1->img}, bk, bd);
    if (bk[0].size() != k3.size() || bk[2].size() != k3.size() || memcmp(bk[0].data(), k3.data(), k3.size() * sizeof(cv::KeyPoint)) != 0) return 4;
    {   // the streaming form through the plain C-ABI: two batches in flight on pinned buffers, waited in
        // submission order; both must hold what the blocking adapter call returned
        const int B = 3, cap = 1400;
        const size_t fb = (size_t) w * h;
        uint8_t *fr = nullptr; orbfe_keypoint *kp[2] = {}; uint8_t *ds[2] = {}; int *cnt[2] = {};
        if (orbfe_host_alloc((void **) &fr, B * fb)) return 6;
        for (int s = 0; s < 2; ++s)
            if (orbfe_host_alloc((void **) &kp[s], sizeof(orbfe_keypoint) * B * cap) || orbfe_host_alloc((void **) &ds[s], (size_t) B * cap * 32) ||
                orbfe_host_alloc((void **) &cnt[s], sizeof(int) * B)) return 6;
        const cv::Mat *src[3] = {&f1->img, &f2->img, &f1->img};
        for (int b = 0; b < B; ++b) memcpy(fr + b * fb, src[b]->data, fb);
        long long t[2] = {-1, -1};
        for (int s = 0; s < 2; ++s)
            if (orbfe_extract_batch_submit(extractor.handle(), fr, B, w, h, (size_t) w, fb, kp[s], ds[s], cap, cnt[s], &t[s])) return 7;
        if (t[0] < 0 || t[1] <= t[0]) return 7;
        for (int s = 0; s < 2; ++s) {
            if (orbfe_extract_batch_wait(extractor.handle(), t[s])) return 8;
            for (int b = 0; b < B; ++b) {
                if (cnt[s][b] != (int) bk[b].size() || memcmp(kp[s] + (size_t) b * cap, bk[b].data(), bk[b].size() * sizeof(cv::KeyPoint)) != 0 ||
                    memcmp(ds[s] + (size_t) b * cap * 32, bd[b].data, bk[b].size() * 32) != 0) return 9;
            }
        }
        orbfe_host_free(fr);
        for (int s = 0; s < 2; ++s) { orbfe_host_free(kp[s]); orbfe_host_free(ds[s]); orbfe_host_free(cnt[s]); }
    }
    std::vector<cv::Point2f> pre(f1->key_points.size());
    for (size_t i = 0; i < pre.size(); ++i) pre[i] = f1->key_points[i].pt;      // Tracking.cpp:598-600
    std::vector<int> m12;
    ORBMatcher matcher(0.9f, true);
    const int nm = matcher.SearchForInitialization(f1, f2, pre, m12, 100);       // Tracking.cpp:605-607
    if (ORBMatcher::DescriptorDistance(f1->descriptors.row(0), f1->descriptors.row(0)) != 0) return 5;
    FILE *o = fopen(argv[5], "wb");
    for (auto *f: {f1.get(), f2.get()}) {
        fwrite(&f->num_kps, 4, 1, o);
        fwrite(f->key_points.data(), sizeof(cv::KeyPoint), f->key_points.size(), o);
        fwrite(f->descriptors.data, 32, f->key_points.size(), o);
    }
    fwrite(&nm, 4, 1, o);
    fwrite(m12.data(), 4, m12.size(), o);
    fwrite(pre.data(), 8, pre.size(), o);
    const int n3 = (int) k3.size();
    fwrite(&n3, 4, 1, o);
    fwrite(k3.data(), sizeof(cv::KeyPoint), k3.size(), o);
    // Frame::Frame post-processing (Frame.cpp:22-51) with the euroc camera: undistorted key points + grid cell sizes
    CameraParams cam; cam.fx = 458.654f; cam.fy = 457.296f; cam.cx = 367.215f; cam.cy = 248.375f;
    cam.dist = {-0.28340811f, 0.07395907f, 0.00019359f, 1.76187114e-05f};
    std::vector<cv::KeyPoint> un; std::vector<std::vector<std::vector<size_t>>> grid;
    postprocessFrame(cam, w, h, k3, un, grid);
    fwrite(un.data(), sizeof(cv::KeyPoint), un.size(), o);
    const int gc = (int) grid.size(), gr = (int) grid[0].size();
    fwrite(&gc, 4, 1, o); fwrite(&gr, 4, 1, o);
    for (int cx = 0; cx < gc; ++cx)
        for (int cy = 0; cy < gr; ++cy) {
            const int m = (int) grid[cx][cy].size();
            fwrite(&m, 4, 1, o);
            for (size_t v : grid[cx][cy]) { const int vi = (int) v; fwrite(&vi, 4, 1, o); }
        }
    fclose(o);
    printf("adapter ok: %d / %d key points, %d matches, %d (1000-feature extractor)\n", f1->num_kps, f2->num_kps, nm, n3);
    return 0;
}
