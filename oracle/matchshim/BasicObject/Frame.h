// Stand-ins for the reference's Frame / KeyFrame / MapPoint / Pose (BasicObject/*.h need Eigen, g2o and the IMU classes, absent in
// this image) with exactly the members modules/ORB/ORBMatcher.cpp touches, so that ORBMatcher.{h,cpp} compile VERBATIM against
// them.  TEST INFRASTRUCTURE ONLY (oracle/matcher_harness.cpp).  The window query getFeaturesInArea forwards to the oracle's
// restatement of Frame.cpp:97-127 / KeyFrame.cpp:181-211 (orc_features_in_area) — the grid itself is not part of ORBMatcher.cpp.
#pragma once
#include <cassert>
#include <climits>
#include <memory>
#include <vector>
#include <Eigen/Core>
#include <opencv2/core/core.hpp>
#include "DBoW2/FeatureVector.h"          // the reference's vendored header
#include "ORBExtractor.h"                 // the reference's own header (static getters used by the matcher)
#include "orb_oracle.h"

namespace mono_orb_slam3 {
    struct Pose {
        Eigen::Matrix3f R; Eigen::Vector3f t;
        Eigen::Vector3f map(const Eigen::Vector3f &P) const { return R * P + t; }
    };

    class KeyFrame;

    class MapPoint {
    public:
        bool bad = false;
        Eigen::Vector3f pos, normal;
        cv::Mat descriptor;                      // 1 x 32
        int num_obs = 0;
        float min_distance = 0, max_distance = 0;
        int predicted_level = 0;
        const KeyFrame *observer = nullptr;
        // what the fuse search did with this point (ORBMatcher.cpp:573-586)
        int fused_idx = -1; bool replaced_by_other = false, replaced_other = false;
        // Frame::isInFrustum outputs (Frame.cpp:129-166)
        bool track_in_view = false; float track_proj_x = 0, track_proj_y = 0, track_view_cos = 0; int track_scale_level = 0;

        bool isBad() const { return bad; }
        Eigen::Vector3f getPos() const { return pos; }
        cv::Mat getDescriptor() const { return descriptor; }
        int getNumObs() const { return num_obs; }
        float getMaxDistanceInvariance() const { return max_distance; }
        float getMinDistanceInvariance() const { return min_distance; }
        Eigen::Vector3f getAverageDirection() const { return normal; }
        int predictScaleLevel(float) const { return predicted_level; }
        bool isObserveKeyFrame(const std::shared_ptr<KeyFrame> &kf) const { return observer == kf.get(); }
        void addObservation(const std::shared_ptr<KeyFrame> &, int idx) { fused_idx = idx; }
        void replace(const std::shared_ptr<MapPoint> &other) { replaced_by_other = true; other->replaced_other = true; }
    };

    class FrameBase {
    public:
        int num_kps = 0, width = 0, height = 0;
        std::vector<cv::KeyPoint> key_points;
        cv::Mat descriptors;
        std::vector<std::shared_ptr<MapPoint>> map_points;
        DBoW2::FeatureVector feature_vector;
        Pose T_cw;
        orc_grid *grid = nullptr;
        bool strict_radius = false;              // KeyFrame::getFeaturesInArea compares with "<", Frame's with "<="
        std::vector<int> queried;                // key-point indices passed to getMapPoint, in call order

        ~FrameBase() { if (grid) orc_grid_destroy(grid); }
        void finish() {
            num_kps = (int) key_points.size();
            map_points.resize((size_t) num_kps);
            grid = orc_grid_build(reinterpret_cast<const orc_keypoint *>(key_points.data()), num_kps, width, height);
        }
        std::vector<size_t> getFeaturesInArea(const float &x, const float &y, const float &r, int minLevel = -1, int maxLevel = -1) const {
            std::vector<int> tmp((size_t) (num_kps > 0 ? num_kps : 1));
            const int n = orc_features_in_area(grid, reinterpret_cast<const orc_keypoint *>(key_points.data()), x, y, r, minLevel, maxLevel,
                                               strict_radius ? 1 : 0, tmp.data(), (int) tmp.size());
            return std::vector<size_t>(tmp.begin(), tmp.begin() + n);
        }
    };

    class Frame : public FrameBase {};

    class KeyFrame : public FrameBase {
    public:
        KeyFrame() { strict_radius = true; }
        std::vector<std::shared_ptr<MapPoint>> getMapPoints() const { return map_points; }
        std::shared_ptr<MapPoint> getMapPoint(int idx) { queried.push_back(idx); return map_points[(size_t) idx]; }
        bool hasMapPoint(int idx) const { return map_points[(size_t) idx] != nullptr; }
        void addMapPoint(const std::shared_ptr<MapPoint> &, int) {}      // bookkeeping of the fuse (ORBMatcher.cpp:573-586) is not recorded
        Pose getPose() const { return T_cw; }
        Eigen::Vector3f getCameraCenter() const { return Eigen::Vector3f(0, 0, 0); }
    };
}
