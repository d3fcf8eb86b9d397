// Force-included (-include) ahead of the reference's BasicObject/MapPoint.cpp so that MapPoint.{h,cpp} compile VERBATIM in an image
// without Eigen / g2o: the include guards of KeyFrame.h, Map.h and ORBMatcher.h are pre-defined (the files are found but contribute
// nothing) and stand-ins with exactly the members MapPoint.cpp touches are declared here.  TEST INFRASTRUCTURE ONLY
// (oracle/mappoint_harness.cpp).  ORBMatcher::DescriptorDistance forwards to the restatement, itself pinned to the reference's
// ORBMatcher.cpp by tests/test_oracle_matcher_ref.py.
#pragma once
#define MONO_ORB_SLAM3_KEYFRAME_H
#define MONO_ORB_SLAM3_MAP_H
#define MONO_ORB_SLAM3_ORBMATCHER_H
#include <map>
#include <memory>
#include <mutex>
#include <vector>
#include <Eigen/Core>
#include <opencv2/core/core.hpp>
#include "ORB/ORBExtractor.h"             // the reference's own header (Match, static scale getters)
#include "orb_oracle.h"

namespace mono_orb_slam3 {
    class MapPoint;
    class KeyFrame {
    public:
        long unsigned int id = 0; unsigned int frame_id = 0;
        std::vector<cv::KeyPoint> key_points; cv::Mat descriptors;
        bool bad = false; Eigen::Vector3f center;
        Eigen::Vector3f getCameraCenter() const { return center; }
        bool isBad() const { return bad; }
        void eraseMapPoint(size_t) {}
        void addMapPoint(const std::shared_ptr<MapPoint> &, size_t) {}
    };
    class Map {
    public:
        void eraseMapPoint(const std::shared_ptr<MapPoint> &) {}
    };
    class ORBMatcher {
    public:
        static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b) { return orc_descriptor_distance(a.ptr(), b.ptr()); }
    };
}
