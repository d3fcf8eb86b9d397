// cv::FAST shim → oracle/orb_oracle.c (pinned to cv2 4.13.0).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include "opencv2/core/core.hpp"
namespace cv {
    static inline void FAST(const Mat &img, std::vector<KeyPoint> &kps, int threshold, bool nms) {
        kps.clear();
        std::vector<orc_corner> c((size_t) img.rows * img.cols / 2 + 16);
        const int n = orc_fast9_16(img.data, img.cols, img.rows, img.step, threshold, nms ? 1 : 0, c.data(), (int) c.size());
        for (int i = 0; i < n; ++i) kps.emplace_back((float) c[i].x, (float) c[i].y, 7.f, -1.f, (float) c[i].score);
    }
}
