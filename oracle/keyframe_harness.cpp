// TEST INFRASTRUCTURE ONLY (see oracle/orb_oracle.py's header): C entry point around the reference's own BasicObject/KeyFrame.cpp and
// Frame.cpp, compiled VERBATIM (with their own headers) after oracle/keyframeshim/prelude.h replaced the Eigen / g2o dependent headers by
// stand-ins.  A Frame is built over caller-supplied key points (as in oracle/frame_harness.cpp), a KeyFrame is made from it the way
// Tracking does, and KeyFrame::getFeaturesInArea (KeyFrame.cpp:181-211, strict "<" radius test) is queried.
// Built by oracle/Makefile into oracle/_ref/libref_keyframe.so.
#include <cstdint>
#include <cstring>
#include "BasicObject/KeyFrame.h"
#include "orb_oracle.h"

using namespace mono_orb_slam3;

extern "C" {

int ref_keyframe_features_in_area(const orc_keypoint *kps, int n, int img_w, int img_h, const float *qx, const float *qy, const float *qr,
                                  const int *qmin, const int *qmax, int nq, int *out_idx, int out_cap, int *out_off) {
    ORBExtractor ex;
    ex.kps.resize((size_t) n);
    static_assert(sizeof(cv::KeyPoint) == sizeof(orc_keypoint), "layout");
    if (n) std::memcpy(ex.kps.data(), kps, sizeof(orc_keypoint) * (size_t) n);
    ex.desc = cv::Mat(n > 0 ? n : 1, 32, CV_8U);
    cv::Mat img(img_h, img_w, CV_8U);
    Frame::grid_size_computed = false;
    Bias bias;
    auto frame = std::make_shared<Frame>(img, 0.0, &ex, bias);
    Map map;
    auto kf = std::make_shared<KeyFrame>(frame, &map);
    int total = 0;
    out_off[0] = 0;
    for (int i = 0; i < nq; ++i) {
        const std::vector<size_t> v = kf->getFeaturesInArea(qx[i], qy[i], qr[i], qmin[i], qmax[i]);
        if (total + (int) v.size() > out_cap) return -1;
        for (size_t idx : v) out_idx[total++] = (int) idx;
        out_off[i + 1] = total;
    }
    return total;
}

}  // extern "C"
