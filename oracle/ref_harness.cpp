// Compiles the reference's modules/ORB/ORBExtractor.cpp VERBATIM (from where it lies, -DREF_SRC=...) against
// oracle/cvshim and exports a tiny C interface for ctypes.  TEST INFRASTRUCTURE ONLY; built by oracle/Makefile
// into oracle/_ref/ (git-ignored).  With -DREF_CANONICAL_SORT the single std::sort call at ORBExtractor.cpp:757
// resolves to the stable, size-only overload declared below (no source edit) — the canonical tie-break of
// SURVEY.md §0.5; without it the reference's heap-address tie-break is kept.
#include <algorithm>
#include <chrono>
#include <list>
#include <thread>
#include <utility>
#include <vector>
#include "opencv2/core/core.hpp"

#ifdef REF_CANONICAL_SORT
namespace mono_orb_slam3 {
    class ExtractorNode;
    typedef std::vector<std::pair<int, ExtractorNode *>>::iterator size_node_iter;
    static inline void sort(size_node_iter a, size_node_iter b) {
        std::stable_sort(a, b, [](const std::pair<int, ExtractorNode *> &l, const std::pair<int, ExtractorNode *> &r) { return l.first < r.first; });
    }
}
#endif

#include REF_SRC

using mono_orb_slam3::ORBExtractor;

extern "C" {
void *ref_extractor_create(int n_features, float scale, int n_levels, int ini_th, int min_th) {
    return new ORBExtractor(n_features, scale, n_levels, ini_th, min_th);
}
void ref_extractor_destroy(void *h) { delete (ORBExtractor *) h; }

static int run_one(ORBExtractor &ex, const uint8_t *img, int w, int h, size_t stride, orc_keypoint *kps, uint8_t *desc, int cap) {
    cv::Mat m; m.create(h, w, CV_8U);
    for (int y = 0; y < h; ++y) std::memcpy(m.ptr(y), img + (size_t) y * stride, (size_t) w);
    std::vector<cv::KeyPoint> out; cv::Mat d;
    ex(m, out, d);
    const int n = (int) out.size();
    if (kps && desc) {
        if (n > cap) return -1;
        static_assert(sizeof(cv::KeyPoint) == sizeof(orc_keypoint), "KeyPoint layout");
        if (n) { std::memcpy(kps, out.data(), sizeof(orc_keypoint) * (size_t) n); for (int i = 0; i < n; ++i) std::memcpy(desc + 32 * (size_t) i, d.ptr(i), 32); }
    }
    return n;
}

int ref_extract(void *h, const uint8_t *img, int w, int hgt, size_t stride, orc_keypoint *kps, uint8_t *desc, int cap) {
    return run_one(*(ORBExtractor *) h, img, w, hgt, stride, kps, desc, cap);
}

// Frame-parallel timing of the reference extractor: `threads` host threads, one private extractor copy each
// (ORBExtractor(int, const ORBExtractor&), ORBExtractor.h:32), frames[B][h][w] contiguous.  Returns seconds.
double ref_extract_mt(void *h, const uint8_t *frames, int B, int w, int hgt, int threads, int *counts) {
    ORBExtractor &proto = *(ORBExtractor *) h;
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    const auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
            ORBExtractor ex(proto);
            for (int b = t; b < B; b += threads) counts[b] = run_one(ex, frames + (size_t) b * w * hgt, w, hgt, (size_t) w, nullptr, nullptr, 0);
        });
    for (auto &th: pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
}
