// cv_compat.h — the handful of OpenCV types the ORB front-end adapters exchange with their callers.
// With real OpenCV on the include path (the reference's build, CMakeLists.txt:13) the real headers are used and this file adds
// nothing; without it (this image has no OpenCV C++ headers) minimal layout-compatible stand-ins are defined so the adapters and
// their tests build.  Product code (not test infrastructure): it carries types only, no image processing.
#pragma once
#if defined(__has_include)
#if __has_include(<opencv2/core/core.hpp>) && !defined(ORBFE_FORCE_CV_COMPAT)
#include <opencv2/core/core.hpp>
#define ORBFE_HAVE_OPENCV 1
#endif
#endif

#ifndef ORBFE_HAVE_OPENCV
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>

#define CV_8U 0
#define CV_8UC1 0

namespace cv {

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };

struct KeyPoint {                    // 7 x 4 bytes, the layout of cv::KeyPoint
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

// Single-channel 8-bit matrix header with shared ownership: enough for "image in, N x 32 descriptor matrix out".
class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char *data = nullptr;
    size_t step = 0;

    Mat() {}
    Mat(int r, int c, int /*type*/) { create(r, c, CV_8U); }
    Mat(int r, int c, int /*type*/, void *ext, size_t st = 0) : rows(r), cols(c), data((unsigned char *) ext), step(st ? st : (size_t) c) {}
    void create(int r, int c, int /*type*/) {
        if (r == rows && c == cols && data && step == (size_t) c) return;
        buf_.reset(new unsigned char[(size_t) (r > 0 ? r : 0) * (size_t) (c > 0 ? c : 0) + 1]);
        rows = r; cols = c; step = (size_t) c; data = buf_.get();
    }
    bool empty() const { return !data || rows <= 0 || cols <= 0; }
    int type() const { return CV_8UC1; }
    bool isContinuous() const { return step == (size_t) cols; }
    unsigned char *ptr(int r = 0) { return data + (size_t) r * step; }
    const unsigned char *ptr(int r = 0) const { return data + (size_t) r * step; }
    template <class T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + (size_t) r * step); }
    template <class T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + (size_t) r * step); }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat rowRange(int a, int b) const { Mat m; m.rows = b - a; m.cols = cols; m.step = step; m.data = data + (size_t) a * step; m.buf_ = buf_; return m; }
    Mat clone() const {
        Mat m; m.create(rows, cols, CV_8U);
        for (int r = 0; r < rows; ++r) std::memcpy(m.ptr(r), ptr(r), (size_t) cols);
        return m;
    }
    void copyTo(Mat &dst) const { dst = clone(); }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }

private:
    std::shared_ptr<unsigned char[]> buf_;
};

}  // namespace cv
#endif  // !ORBFE_HAVE_OPENCV
