// TEST INFRASTRUCTURE ONLY (see oracle/orb_oracle.py's header): C entry points around the reference's own modules/ORB/ORBMatcher.cpp,
// compiled VERBATIM from where it lies against stand-in Frame / KeyFrame / MapPoint / Camera / Eigen headers (oracle/matchshim,
// exactly the members the matcher touches), the cv:: shim and the vendored DBoW2 FeatureVector.  The entry points take the same flat
// arrays as the restatement in oracle/orb_oracle.c (orc_search_*), build the stand-in objects, call the reference's methods and
// flatten what they did.  Map points are placed at (u, v, 1) in front of an identity pose and a unit pinhole, so the reference's own
// projection code yields exactly the (u, v) the restatement is given.  Built by oracle/Makefile into oracle/_ref/libref_matcher.so.
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#include "ORBMatcher.h"
#include "Sensor/Camera.h"

using namespace mono_orb_slam3;

namespace {
void ensure_extractor_statics() {            // ORBExtractor's static tables (scale factors, sigmas) are filled by its constructor
    static ORBExtractor ex(1000, 1.2f, 8, 20, 7);
    (void) ex;
}
cv::Mat rows(const uint8_t *d, int n) {
    cv::Mat m(n > 0 ? n : 1, 32, CV_8U);
    if (n > 0) std::memcpy(m.data, d, (size_t) n * 32);
    return m;
}
cv::Mat row(const uint8_t *d) { cv::Mat m(1, 32, CV_8U); std::memcpy(m.data, d, 32); return m; }
template <class F> void fill_frame(F &f, const orc_keypoint *kps, const uint8_t *desc, int n, int w, int h) {
    f.key_points.resize((size_t) n);
    static_assert(sizeof(cv::KeyPoint) == sizeof(orc_keypoint), "layout");
    if (n) std::memcpy(f.key_points.data(), kps, sizeof(orc_keypoint) * (size_t) n);
    f.descriptors = rows(desc, n);
    f.width = w; f.height = h;
    f.finish();
}
// a frame that only carries per-key-point attributes (no window queries are made on it)
template <class F> void fill_attr_frame(F &f, int n, const float *angle, const int *octave, const float *size, const uint8_t *desc) {
    f.key_points.resize((size_t) n);
    for (int i = 0; i < n; ++i) {
        cv::KeyPoint &k = f.key_points[(size_t) i];
        k.angle = angle ? angle[i] : 0.f; k.octave = octave ? octave[i] : 0; k.size = size ? size[i] : 1.f;
    }
    if (desc) f.descriptors = rows(desc, n);
    f.width = 64; f.height = 64;
    f.finish();
}
void fill_fv(DBoW2::FeatureVector &fv, const int *id, const int *off, const int *idx, int n_nodes) {
    for (int k = 0; k < n_nodes; ++k)
        for (int j = off[k]; j < off[k + 1]; ++j) fv.addFeature((DBoW2::NodeId) id[k], (unsigned) idx[j]);
}
std::shared_ptr<MapPoint> point_at(float u, float v, const uint8_t *desc) {
    auto mp = std::make_shared<MapPoint>();
    mp->pos = Eigen::Vector3f(u, v, 1.f); mp->normal = mp->pos;
    mp->min_distance = 0.f; mp->max_distance = 3.0e38f;
    if (desc) mp->descriptor = row(desc);
    return mp;
}
struct ExposedMatcher : ORBMatcher {
    using ORBMatcher::ORBMatcher;
    static void three_maxima(std::vector<int> *h, int &a, int &b, int &c) { ComputeThreeMaxima(h, a, b, c); }
};
}  // namespace

extern "C" {

int ref_descriptor_distance(const uint8_t *a, const uint8_t *b) { return ORBMatcher::DescriptorDistance(row(a), row(b)); }

void ref_compute_three_maxima(const int *counts, int n_bins, int *i1, int *i2, int *i3) {
    std::vector<std::vector<int>> h((size_t) n_bins);
    for (int b = 0; b < n_bins; ++b) h[(size_t) b].assign((size_t) counts[b], 0);
    int a = -1, b = -1, c = -1;
    ExposedMatcher::three_maxima(h.data(), a, b, c);        // the reference hard-codes HISTO_LENGTH = 30 bins
    *i1 = a; *i2 = b; *i3 = c;
}

int ref_search_for_initialization(const orc_keypoint *kps1, const uint8_t *desc1, int n1, const orc_keypoint *kps2, const uint8_t *desc2, int n2,
                                  int img_w, int img_h, float *prematched_xy, int *matches12, int window, float nn_ratio, int check_orientation) {
    ensure_extractor_statics();
    auto f1 = std::make_shared<Frame>(), f2 = std::make_shared<Frame>();
    fill_frame(*f1, kps1, desc1, n1, img_w, img_h); fill_frame(*f2, kps2, desc2, n2, img_w, img_h);
    std::vector<cv::Point2f> pre((size_t) n1);
    for (int i = 0; i < n1; ++i) pre[(size_t) i] = cv::Point2f(prematched_xy[2 * i], prematched_xy[2 * i + 1]);
    std::vector<int> m12;
    ORBMatcher matcher(nn_ratio, check_orientation != 0);
    const int n = matcher.SearchForInitialization(f1, f2, pre, m12, window);
    for (int i = 0; i < n1; ++i) { matches12[i] = m12[(size_t) i]; prematched_xy[2 * i] = pre[(size_t) i].x; prematched_xy[2 * i + 1] = pre[(size_t) i].y; }
    return n;
}

// SearchByProjection(lastFrame | lastKF, curFrame, th = 1): the radius rides in the last frame's key-point size
int ref_search_by_projection(const float *q_u, const float *q_v, const float *q_radius, const int *q_level, const float *q_angle, const uint8_t *q_desc,
                             const uint8_t *q_valid, int nq, const orc_keypoint *kps2, const uint8_t *desc2, int n2, int img_w, int img_h,
                             const uint8_t *occupied, int *assigned, int check_orientation, int from_keyframe) {
    ensure_extractor_statics();
    Camera::instance()->width = 0;                          // unbounded: validity is the caller's q_valid, as in the flat-array interface
    auto cur = std::make_shared<Frame>();
    fill_frame(*cur, kps2, desc2, n2, img_w, img_h);
    auto blocker = std::make_shared<MapPoint>();
    for (int j = 0; j < n2; ++j) if (occupied && occupied[j]) cur->map_points[(size_t) j] = blocker;
    std::vector<std::shared_ptr<MapPoint>> mps((size_t) nq);
    for (int i = 0; i < nq; ++i) if (q_valid[i]) mps[(size_t) i] = point_at(q_u[i], q_v[i], q_desc + 32 * (size_t) i);
    ORBMatcher matcher(0.6f, check_orientation != 0);
    int n;
    if (from_keyframe) {
        auto last = std::make_shared<KeyFrame>();
        fill_attr_frame(*last, nq, q_angle, q_level, q_radius, nullptr);
        last->map_points = mps;
        n = matcher.SearchByProjection(last, cur, 1.f);
    } else {
        auto last = std::make_shared<Frame>();
        fill_attr_frame(*last, nq, q_angle, q_level, q_radius, nullptr);
        last->map_points = mps;
        n = matcher.SearchByProjection(last, cur, 1.f);
    }
    for (int j = 0; j < n2; ++j) {
        assigned[j] = -1;
        const auto &p = cur->map_points[(size_t) j];
        if (p && p != blocker)
            for (int i = 0; i < nq; ++i) if (mps[(size_t) i] == p) { assigned[j] = i; break; }
    }
    return n;
}

// SearchByProjection(frame, local map points, th): the reference derives the radius from th, track_view_cos and the predicted level
int ref_search_local_points(const float *q_u, const float *q_v, const float *q_view_cos, const int *q_level, const uint8_t *q_desc, const uint8_t *q_valid,
                            int nq, float th, const orc_keypoint *kps2, const uint8_t *desc2, int n2, int img_w, int img_h, const uint8_t *occupied,
                            int *assigned, float nn_ratio) {
    ensure_extractor_statics();
    auto fr = std::make_shared<Frame>();
    fill_frame(*fr, kps2, desc2, n2, img_w, img_h);
    auto blocker = std::make_shared<MapPoint>();
    for (int j = 0; j < n2; ++j) if (occupied && occupied[j]) fr->map_points[(size_t) j] = blocker;
    std::vector<std::shared_ptr<MapPoint>> mps((size_t) nq);
    for (int i = 0; i < nq; ++i) {
        auto mp = point_at(q_u[i], q_v[i], q_desc + 32 * (size_t) i);
        mp->track_in_view = q_valid[i] != 0; mp->track_proj_x = q_u[i]; mp->track_proj_y = q_v[i];
        mp->track_view_cos = q_view_cos[i]; mp->track_scale_level = q_level[i];
        mps[(size_t) i] = mp;
    }
    ORBMatcher matcher(nn_ratio, true);
    const int n = matcher.SearchByProjection(fr, mps, th);
    for (int j = 0; j < n2; ++j) {
        assigned[j] = -1;
        const auto &p = fr->map_points[(size_t) j];
        if (p && p != blocker)
            for (int i = 0; i < nq; ++i) if (mps[(size_t) i] == p) { assigned[j] = i; break; }
    }
    return n;
}

int ref_search_by_bow(const uint8_t *desc1, const float *angle1, const uint8_t *valid1, int n1, const int *id1, const int *off1, const int *idx1, int nn1,
                      const uint8_t *desc2, const float *angle2, const uint8_t *occupied2, int n2, const int *id2, const int *off2, const int *idx2, int nn2,
                      int *assigned, float nn_ratio, int check_orientation) {
    ensure_extractor_statics();
    auto kf = std::make_shared<KeyFrame>(); auto fr = std::make_shared<Frame>();
    fill_attr_frame(*kf, n1, angle1, nullptr, nullptr, desc1); fill_attr_frame(*fr, n2, angle2, nullptr, nullptr, desc2);
    fill_fv(kf->feature_vector, id1, off1, idx1, nn1); fill_fv(fr->feature_vector, id2, off2, idx2, nn2);
    std::vector<std::shared_ptr<MapPoint>> mps((size_t) n1);
    for (int i = 0; i < n1; ++i) if (valid1[i]) mps[(size_t) i] = std::make_shared<MapPoint>();
    kf->map_points = mps;
    auto blocker = std::make_shared<MapPoint>();
    for (int j = 0; j < n2; ++j) if (occupied2 && occupied2[j]) fr->map_points[(size_t) j] = blocker;
    ORBMatcher matcher(nn_ratio, check_orientation != 0);
    const int n = matcher.SearchByBow(kf, fr);
    for (int j = 0; j < n2; ++j) {
        assigned[j] = -1;
        const auto &p = fr->map_points[(size_t) j];
        if (p && p != blocker)
            for (int i = 0; i < n1; ++i) if (mps[(size_t) i] == p) { assigned[j] = i; break; }
    }
    return n;
}

int ref_search_for_triangulation(const uint8_t *desc1, const float *angle1, const uint8_t *has_mp1, int n1, const int *id1, const int *off1, const int *idx1, int nn1,
                                 const uint8_t *desc2, const float *angle2, const uint8_t *has_mp2, int n2, const int *id2, const int *off2, const int *idx2, int nn2,
                                 int *matches12, int check_orientation) {
    ensure_extractor_statics();
    auto k1 = std::make_shared<KeyFrame>(), k2 = std::make_shared<KeyFrame>();
    fill_attr_frame(*k1, n1, angle1, nullptr, nullptr, desc1); fill_attr_frame(*k2, n2, angle2, nullptr, nullptr, desc2);
    fill_fv(k1->feature_vector, id1, off1, idx1, nn1); fill_fv(k2->feature_vector, id2, off2, idx2, nn2);
    auto some = std::make_shared<MapPoint>();
    for (int i = 0; i < n1; ++i) if (has_mp1[i]) k1->map_points[(size_t) i] = some;
    for (int j = 0; j < n2; ++j) if (has_mp2[j]) k2->map_points[(size_t) j] = some;
    std::vector<int> m12;
    ORBMatcher matcher(0.6f, check_orientation != 0);
    const int n = matcher.SearchForTriangulation(k1, k2, m12);
    for (int i = 0; i < n1; ++i) matches12[i] = m12[(size_t) i];
    return n;
}

// fuse SearchByProjection(keyFrame, mapPoints, map, th): best_idx1[i] = key point the reference fused map point i with, or -1.
// (The stand-in key frame never stores the fused point, so every match takes the addObservation branch and is recorded.)
int ref_search_fuse(const float *q_u, const float *q_v, const int *q_level, const uint8_t *q_desc, const uint8_t *q_valid, int nq, float th,
                    const orc_keypoint *kps1, const uint8_t *desc1, int n1, int img_w, int img_h, int *best_idx1) {
    ensure_extractor_statics();
    Camera::instance()->width = 0;
    auto kf = std::make_shared<KeyFrame>();
    fill_frame(*kf, kps1, desc1, n1, img_w, img_h);
    std::vector<std::shared_ptr<MapPoint>> mps((size_t) nq);
    for (int i = 0; i < nq; ++i)
        if (q_valid[i]) { mps[(size_t) i] = point_at(q_u[i], q_v[i], q_desc + 32 * (size_t) i); mps[(size_t) i]->predicted_level = q_level[i]; }
    Map map;
    const int n = ORBMatcher::SearchByProjection(kf, mps, &map, th);
    for (int i = 0; i < nq; ++i) best_idx1[i] = mps[(size_t) i] ? mps[(size_t) i]->fused_idx : -1;
    return n;
}

}  // extern "C"
