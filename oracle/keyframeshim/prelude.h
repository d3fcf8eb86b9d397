// Force-included (-include) ahead of the reference's BasicObject/Frame.cpp and KeyFrame.cpp so that both compile VERBATIM, with their
// own Frame.h / KeyFrame.h, in an image without Eigen / g2o: the include guards of Pose.h, MapPoint.h, Map.h, Sensor/Imu.h,
// Sensor/Camera.h, ORB/ORBExtractor.h and ORB/ORBVocabulary.h are pre-defined (the files are found but contribute nothing) and
// stand-ins with exactly the members the two files touch are declared here; Eigen is the name-level stand-in of
// oracle/twoviewshim.  TEST INFRASTRUCTURE ONLY (oracle/keyframe_harness.cpp).
#pragma once
#define MONO_ORB_SLAM3_POSE_H
#define MONO_ORB_SLAM3_MAPPOINT_H
#define MONO_ORB_SLAM3_MAP_H
#define MONO_ORB_SLAM3_IMU_H
#define MONO_ORB_SLAM3_CAMERA_H
#define MONO_ORB_SLAM3_ORBEXTRACTOR_H
#define MONO_ORB_SLAM3_ORBVOCABULARY_H
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <thread>
#include <unordered_map>
#include <vector>
#include <Eigen/Dense>
#include <opencv2/core/core.hpp>
#include "DBoW2/BowVector.h"
#include "DBoW2/FeatureVector.h"

namespace mono_orb_slam3 {
    class KeyFrame;
    class MapPoint;
    struct Pose {
        Eigen::Matrix3f R = Eigen::Matrix3f::Identity(); Eigen::Vector3f t;
        Eigen::Vector3f map(const Eigen::Vector3f &P) const { return R * P + t; }
        Pose inverse() const { Pose p; p.R = R.transpose(); p.t = -(p.R * t); return p; }
        Pose operator*(const Pose &o) const { Pose p; p.R = R * o.R; p.t = R * o.t + t; return p; }
    };
    struct Bias {};
    struct ImuData { Eigen::Vector3f w, a; double t = 0; };
    struct PreIntegrator {
        double delta_t = 0; Bias updated_bias;
        explicit PreIntegrator(const Bias &) {}
        explicit PreIntegrator(const std::shared_ptr<PreIntegrator> &) {}
        void setNewBias(const Bias &) {}
        void IntegrateNewMeasurement(const Eigen::Vector3f &, const Eigen::Vector3f &, double dt) { delta_t += dt; }
    };
    struct ImuCalib {
        Pose T_cb;
        static const ImuCalib *getImuCalib() { static ImuCalib c; return &c; }
    };
    class MapPoint {
    public:
        Eigen::Vector3f pos, normal; float min_distance = 0, max_distance = 0; int predicted_level = 0; bool bad = false;
        bool track_in_view = false; float track_proj_x = 0, track_proj_y = 0, track_view_cos = 0; int track_scale_level = 0;
        std::map<std::shared_ptr<KeyFrame>, size_t> observations;
        Eigen::Vector3f getPos() const { return pos; }
        float getMaxDistanceInvariance() const { return max_distance; }
        float getMinDistanceInvariance() const { return min_distance; }
        Eigen::Vector3f getAverageDirection() const { return normal; }
        int predictScaleLevel(float) const { return predicted_level; }
        bool isBad() const { return bad; }
        std::map<std::shared_ptr<KeyFrame>, size_t> getObservations() const { return observations; }
        int getNumObs() const { return (int) observations.size(); }
        void eraseObservation(const std::shared_ptr<KeyFrame> &kf) { observations.erase(kf); }
    };
    class Map {
    public:
        void eraseKeyFrame(const std::shared_ptr<KeyFrame> &) {}
    };
    class ORBExtractor {          // hands out the key points and descriptors the harness stored
    public:
        std::vector<cv::KeyPoint> kps; cv::Mat desc;
        void operator()(const cv::Mat &, std::vector<cv::KeyPoint> &key_points, cv::Mat &descriptors) { key_points = kps; descriptors = desc; }
    };
    class Camera {
    public:
        int width = 0, height = 0;
        static Camera *instance() { static Camera c; return &c; }
        static const Camera *getCamera() { return instance(); }
        float uncertainty(const cv::Point2f &) const { return 1.f; }
        void undistortKeyPoints(const std::vector<cv::KeyPoint> &raw, std::vector<cv::KeyPoint> &out) const { out = raw; }
        cv::Point2f project(const Eigen::Vector3f &Pc) const { return cv::Point2f(Pc[0] / Pc[2], Pc[1] / Pc[2]); }
        bool isInImage(const cv::Point2f &p) const { return p.x >= 0 && p.x < (float) width && p.y >= 0 && p.y < (float) height; }
    };
    struct Vocabulary {
        void transform(const std::vector<cv::Mat> &, DBoW2::BowVector &, DBoW2::FeatureVector &, int) const {}
    };
    struct ORBVocabulary {
        static const Vocabulary *getORBVocabulary() { static Vocabulary v; return &v; }
    };
}
