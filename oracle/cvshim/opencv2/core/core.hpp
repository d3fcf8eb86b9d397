// Minimal cv:: surface needed to compile the reference's modules/ORB/ORBExtractor.{h,cpp} VERBATIM in an
// image without OpenCV C++.  TEST INFRASTRUCTURE ONLY.  The four arithmetic primitives
// (resize, GaussianBlur, FAST, fastAtan2) forward to oracle/orb_oracle.c, which is pinned bit-exact to
// cv2 4.13.0 by tests/test_oracle_primitives.py.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>       // the real opencv2/core pulls these in; the vendored DBoW2 headers rely on that
#include <string>
#include <vector>
#include "orb_oracle.h"

typedef unsigned char uchar;
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

static inline int cvRound(double v) { return orc_round(v); }
static inline int cvRound(float v) { return orc_roundf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int) v; return i - (v < i); }
static inline int cvFloor(float v) { int i = (int) v; return i - (v < i); }
static inline int cvFloor(int v) { return v; }
static inline int cvCeil(double v) { int i = (int) v; return i + (v > i); }
static inline int cvCeil(float v) { int i = (int) v; return i + (v > i); }
static inline int cvCeil(int v) { return v; }

namespace cv {
    template<typename T> struct Point_ {
        T x, y;
        Point_() : x(0), y(0) {}
        Point_(T x_, T y_) : x(x_), y(y_) {}
        Point_ &operator*=(float s) { x = (T) (x * s); y = (T) (y * s); return *this; }
    };
    typedef Point_<int> Point;
    typedef Point_<int> Point2i;
    typedef Point_<float> Point2f;

    template<typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {} };
    typedef Point3_<float> Point3f;
    struct RNG {            // name-only (TwoViewReconstruction draws its RANSAC sets with it; the oracle harness supplies its own hypotheses)
        unsigned long long state = 0xffffffffu;
        int uniform(int a, int b) { state = state * 4164903690ULL + (state >> 32); return a + (int) ((unsigned) state % (unsigned) (b - a)); }
    };

    struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };

    struct KeyPoint {     // same 28-byte layout as the real class
        Point2f pt; float size, angle, response; int octave, class_id;
        KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
        KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
    };

    enum { INTER_LINEAR = 1, BORDER_REFLECT_101 = 4 };

    struct MatStep {
        size_t v;
        MatStep(size_t s = 0) : v(s) {}
        operator size_t() const { return v; }
    };

    struct MatZeros { int rows, cols; };   // stand-in for the MatExpr returned by Mat::zeros

    class Mat {
    public:
        int rows, cols; uchar *data; MatStep step;
        Mat() : rows(0), cols(0), data(nullptr), step(0) {}
        Mat(int r, int c, int /*type*/) : Mat() { create(r, c, CV_8U); }
        void create(int r, int c, int type) {
            if (data && r == rows && c == cols && isContinuous() && type == CV_8U) return;
            const size_t es = type == CV_32F ? 4 : 1;           // (DBoW2's FORB::toMat32F; cols then counts elements, step bytes)
            buf_ = std::shared_ptr<uchar>(new uchar[(size_t) r * c * es > 0 ? (size_t) r * c * es : 1], std::default_delete<uchar[]>());
            data = buf_.get(); rows = r; cols = c; step = (size_t) c * es;
        }
        void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
        bool isContinuous() const { return (size_t) step == (size_t) cols || rows == 1; }
        bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
        int type() const { return CV_8UC1; }
        size_t step1() const { return step; }
        template<typename T> T &at(int y, int x) { return *(T *) (data + (size_t) y * step + x * sizeof(T)); }
        template<typename T> const T &at(int y, int x) const { return *(const T *) (data + (size_t) y * step + x * sizeof(T)); }
        uchar *ptr(int y = 0) { return data + (size_t) y * step; }
        const uchar *ptr(int y = 0) const { return data + (size_t) y * step; }
        template<typename T> T *ptr(int y = 0) { return (T *) (data + (size_t) y * step); }
        template<typename T> const T *ptr(int y = 0) const { return (const T *) (data + (size_t) y * step); }
        Mat clone() const {
            Mat m; if (empty()) return m;
            m.create(rows, cols, CV_8U);
            for (int y = 0; y < rows; ++y) std::memcpy(m.ptr(y), ptr(y), (size_t) cols);
            return m;
        }
        Mat row(int y) const { return rowRange(y, y + 1); }
        Mat rowRange(int a, int b) const { Mat m = *this; m.data = data + (size_t) a * step; m.rows = b - a; return m; }
        Mat colRange(int a, int b) const { Mat m = *this; m.data = data + a; m.cols = b - a; return m; }
        static MatZeros zeros(int r, int c, int /*type*/) { return MatZeros{r, c}; }
        // real OpenCV evaluates a MatExpr INTO an existing header of matching size/type (zero-fills in place)
        Mat &operator=(const MatZeros &z) {
            if (!(data && rows == z.rows && cols == z.cols)) { data = nullptr; create(z.rows, z.cols, CV_8U); }
            for (int y = 0; y < rows; ++y) std::memset(ptr(y), 0, (size_t) cols);
            return *this;
        }
    private:
        std::shared_ptr<uchar> buf_;
    };

    static inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

    // Name-only stand-ins for the YAML storage classes that the vendored DBoW2's TemplatedVocabulary.h mentions in its virtual
    // save / load members (instantiated with the class, never called by the oracle harness, which loads text files)
    struct FileNode {
        FileNode operator[](const std::string &) const { return FileNode(); }
        FileNode operator[](const char *) const { return FileNode(); }
        FileNode operator[](int) const { return FileNode(); }
        size_t size() const { return 0; }
        operator int() const { return 0; }
        operator double() const { return 0; }
        operator std::string() const { return std::string(); }
    };
    struct FileStorage {
        enum { READ = 0, WRITE = 1 };
        FileStorage() {}
        FileStorage(const std::string &, int) {}
        bool isOpened() const { return false; }
        FileNode operator[](const std::string &) const { return FileNode(); }
        template <class T> FileStorage &operator<<(const T &) { return *this; }
    };
}
