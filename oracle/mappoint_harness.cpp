// TEST INFRASTRUCTURE ONLY (see oracle/orb_oracle.py's header): C entry point around the reference's own BasicObject/MapPoint.cpp,
// compiled VERBATIM (with its own MapPoint.h) after oracle/mappointshim/prelude.h replaced KeyFrame / Map / ORBMatcher by stand-ins.
// Exercises MapPoint::computeDescriptor (MapPoint.cpp:103-152).  Built by oracle/Makefile into oracle/_ref/libref_mappoint.so.
#include <cstdint>
#include <cstring>
#include "BasicObject/MapPoint.h"

using namespace mono_orb_slam3;

extern "C" {

// One map point per group: its observations are key frames holding one descriptor row each (rows [off[g], off[g+1]) of desc; key
// frames flagged in `bad` are skipped by the reference).  The reference iterates its observations in std::map order, i.e. by
// key-frame address: order[] receives, per group, the row indices in that order (bad ones left out), n_order[g] their count, and
// chosen (n_groups x 32) the descriptor the reference picked.  Empty groups are left untouched.
void ref_compute_descriptors(const uint8_t *desc, const uint8_t *bad, const int *off, int n_groups, int *order, int *n_order, uint8_t *chosen) {
    static ORBExtractor statics(1000, 1.2f, 8, 20, 7);             // fills the static scale tables MapPoint::update reads
    Map map;
    for (int g = 0; g < n_groups; ++g) {
        const int b = off[g], n = off[g + 1] - off[g];
        n_order[g] = 0;
        if (n <= 0) continue;
        std::vector<std::shared_ptr<KeyFrame>> kfs((size_t) n);
        for (int i = 0; i < n; ++i) {
            auto kf = std::make_shared<KeyFrame>();
            kf->id = (unsigned long) i; kf->key_points.resize(1); kf->key_points[0].size = 1.f;
            kf->descriptors = cv::Mat(1, 32, CV_8U);
            std::memcpy(kf->descriptors.data, desc + 32 * (size_t) (b + i), 32);
            kf->bad = bad && bad[b + i];
            kf->center = Eigen::Vector3f((float) i, 0.f, -1.f);
            kfs[(size_t) i] = kf;
        }
        auto mp = std::make_shared<MapPoint>(Eigen::Vector3f(0.f, 0.f, 1.f), kfs[0], kfs[(size_t) (n > 1 ? 1 : 0)], Match(0, 0), &map);
        for (int i = 2; i < n; ++i) mp->addObservation(kfs[(size_t) i], 0);
        mp->computeDescriptor();
        int k = 0;
        for (const auto &obs : mp->getObservations()) {
            if (obs.first->isBad()) continue;
            for (int i = 0; i < n; ++i) if (kfs[(size_t) i] == obs.first) { order[b + k++] = b + i; break; }
        }
        n_order[g] = k;
        const cv::Mat d = mp->getDescriptor();
        std::memcpy(chosen + 32 * (size_t) g, d.ptr(), 32);
    }
}

}  // extern "C"
