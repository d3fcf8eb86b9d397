// FramePost.h — the part of Frame::Frame (modules/BasicObject/Frame.cpp:22-51) between the extractor call and the first matcher
// call, on top of include/orbfe.h: kp.size *= camera->uncertainty(kp.pt), camera->undistortKeyPoints(), and the 40-px grid.
// Same member layout as the reference's Frame (raw_key_points, key_points, grid[GRID_COLS][GRID_ROWS]).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "cv_compat.h"
#include "ORBMatcher.h"
#include "../../include/orbfe.h"

namespace mono_orb_slam3 {

    // Camera::create's yaml fields (Sensor/Camera.cpp:27-52); scale_mat is Fisheye::scale_mat (height x width CV_32F) or empty
    struct CameraParams {
        bool equidistant = false;                 // DistortionModel "equidistant" (Fisheye) vs "radtan" (Pinhole)
        float fx = 0, fy = 0, cx = 0, cy = 0;
        std::vector<float> dist;
        const float *scale_mat = nullptr;
    };

    inline void postprocessFrame(const CameraParams &cam, int img_w, int img_h, std::vector<cv::KeyPoint> &raw_key_points,
                                 std::vector<cv::KeyPoint> &key_points, std::vector<std::vector<std::vector<size_t>>> &grid) {
        static_assert(sizeof(cv::KeyPoint) == sizeof(orbfe_keypoint), "cv::KeyPoint and orbfe_keypoint are layout-compatible");
        orbfe_camera c{};
        c.model = cam.equidistant ? ORBFE_CAMERA_FISHEYE : ORBFE_CAMERA_PINHOLE;
        c.fx = cam.fx; c.fy = cam.fy; c.cx = cam.cx; c.cy = cam.cy;
        c.n_dist = (int) (cam.dist.size() < 12 ? cam.dist.size() : 12);
        for (int i = 0; i < c.n_dist; ++i) c.dist[i] = cam.dist[(size_t) i];
        c.uncertainty_map = cam.equidistant ? cam.scale_mat : nullptr; c.uncertainty_w = img_w; c.uncertainty_h = img_h;
        int cols = 0, rows = 0;
        orbfe_grid_size(img_w, img_h, &cols, &rows);                                 // GRID_COLS / GRID_ROWS (Frame.cpp:33-41)
        const int n = (int) raw_key_points.size();
        key_points.resize((size_t) n);
        std::vector<int32_t> off((size_t) cols * rows + 1), idx((size_t) (n > 0 ? n : 1));
        int n_in = 0;
        orbfe_handle *h = ORBMatcher::handle();
        if (orbfe_frame_postprocess(h, &c, reinterpret_cast<orbfe_keypoint *>(raw_key_points.data()), n, img_w, img_h,
                                    reinterpret_cast<orbfe_keypoint *>(key_points.data()), off.data(), idx.data(), &n_in) != ORBFE_OK)
            throw std::runtime_error(std::string("orbfe: ") + orbfe_last_error(h));
        grid.assign((size_t) cols, std::vector<std::vector<size_t>>((size_t) rows));   // Frame.cpp:43
        for (int cx = 0; cx < cols; ++cx)
            for (int cy = 0; cy < rows; ++cy) {
                const int c0 = off[(size_t) cx * rows + cy], c1 = off[(size_t) cx * rows + cy + 1];
                grid[(size_t) cx][(size_t) cy].assign(idx.begin() + c0, idx.begin() + c1);
            }
    }
} // mono_orb_slam3
