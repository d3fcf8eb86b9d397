// TEST INFRASTRUCTURE ONLY (see oracle/orb_oracle.py's header): C entry point around the reference's own BasicObject/Frame.cpp,
// compiled VERBATIM from where it lies (with its own Frame.h) after oracle/frameshim/prelude.h has replaced the headers that need
// Eigen / g2o by stand-ins.  Exercises the constructor's 40-px grid (Frame.cpp:32-51), PosInGrid (:90-95) and getFeaturesInArea
// (:97-127) on caller-supplied key points.  Built by oracle/Makefile into oracle/_ref/libref_frame.so.
#include <cstdint>
#include <cstring>
#include "BasicObject/Frame.h"
#include "orb_oracle.h"

using namespace mono_orb_slam3;

extern "C" {

// For every query (x, y, r, min_level, max_level): the indices Frame::getFeaturesInArea returns, in its order.
// out_off has nq + 1 entries; returns the total, or -1 if out_cap is too small.
int ref_frame_features_in_area(const orc_keypoint *kps, int n, int img_w, int img_h, const float *qx, const float *qy, const float *qr,
                               const int *qmin, const int *qmax, int nq, int *out_idx, int out_cap, int *out_off, int *grid_cols, int *grid_rows) {
    ORBExtractor ex;
    ex.kps.resize((size_t) n);
    static_assert(sizeof(cv::KeyPoint) == sizeof(orc_keypoint), "layout");
    if (n) std::memcpy(ex.kps.data(), kps, sizeof(orc_keypoint) * (size_t) n);
    ex.desc = cv::Mat(n > 0 ? n : 1, 32, CV_8U);
    cv::Mat img(img_h, img_w, CV_8U);
    Frame::grid_size_computed = false;                     // the reference computes the grid size once per process
    Bias bias;
    Frame frame(img, 0.0, &ex, bias);
    *grid_cols = Frame::GRID_COLS; *grid_rows = Frame::GRID_ROWS;
    int total = 0;
    out_off[0] = 0;
    for (int i = 0; i < nq; ++i) {
        const std::vector<size_t> v = frame.getFeaturesInArea(qx[i], qy[i], qr[i], qmin[i], qmax[i]);
        if (total + (int) v.size() > out_cap) return -1;
        for (size_t idx : v) out_idx[total++] = (int) idx;
        out_off[i + 1] = total;
    }
    return total;
}

// The grid Frame::Frame builds (Frame.cpp:32-51) as CSR over cells in grid[cx][cy] order (cell = cx * rows + cy): grid_off has
// cols * rows + 1 entries (the caller sizes it with its own cols / rows and checks them against the returned ones).
int ref_frame_grid(const orc_keypoint *kps, int n, int img_w, int img_h, int *grid_off, int off_cap, int *grid_idx, int *grid_cols, int *grid_rows) {
    ORBExtractor ex;
    ex.kps.resize((size_t) n);
    if (n) std::memcpy(ex.kps.data(), kps, sizeof(orc_keypoint) * (size_t) n);
    ex.desc = cv::Mat(n > 0 ? n : 1, 32, CV_8U);
    cv::Mat img(img_h, img_w, CV_8U);
    Frame::grid_size_computed = false;
    Bias bias;
    Frame frame(img, 0.0, &ex, bias);
    const int cols = Frame::GRID_COLS, rows = Frame::GRID_ROWS;
    *grid_cols = cols; *grid_rows = rows;
    if (cols * rows + 1 > off_cap) return -1;
    int total = 0;
    for (int cx = 0; cx < cols; ++cx)
        for (int cy = 0; cy < rows; ++cy) {
            grid_off[cx * rows + cy] = total;
            for (size_t idx : frame.grid[(size_t) cx][(size_t) cy]) grid_idx[total++] = (int) idx;
        }
    grid_off[cols * rows] = total;
    return total;
}

}  // extern "C"
