// cv::resize / cv::GaussianBlur shim → oracle/orb_oracle.c (pinned to cv2 4.13.0).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include "opencv2/core/core.hpp"
namespace cv {
    static inline void resize(const Mat &src, Mat &dst, Size sz, double = 0, double = 0, int = INTER_LINEAR) {
        Mat out; out.create(sz.height, sz.width, CV_8U);
        orc_resize_linear_u8(src.data, src.cols, src.rows, src.step, out.data, sz.width, sz.height, out.step);
        dst = out;
    }
    static inline void GaussianBlur(const Mat &src, Mat &dst, Size, double, double, int) {
        Mat out; out.create(src.rows, src.cols, CV_8U);
        orc_gaussian_blur7_u8(src.data, src.cols, src.rows, src.step, out.data, out.step);
        for (int y = 0; y < src.rows; ++y) std::memcpy(dst.ptr(y), out.ptr(y), (size_t) src.cols);   // in-place call in the reference
    }
}
