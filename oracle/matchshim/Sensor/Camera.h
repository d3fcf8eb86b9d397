// stand-in for the reference's Sensor/Camera.h with the three members ORBMatcher.cpp uses.  The harness feeds map points whose
// camera-frame position is (u, v, 1), so this unit pinhole returns exactly the (u, v) the flat-array oracle is given.
#pragma once
#include <Eigen/Core>
#include <opencv2/core/core.hpp>
namespace mono_orb_slam3 {
    class Camera {
    public:
        int width = 0, height = 0;
        static Camera *instance() { static Camera c; return &c; }
        static const Camera *getCamera() { return instance(); }
        cv::Point2f project(const Eigen::Vector3f &Pc) const { return cv::Point2f(Pc[0] / Pc[2], Pc[1] / Pc[2]); }
        // width == 0: unbounded (the flat-array interface carries validity in q_valid)
        bool isInImage(const cv::Point2f &p) const { return width == 0 || (p.x >= 0 && p.x < (float) width && p.y >= 0 && p.y < (float) height); }
    };
}
