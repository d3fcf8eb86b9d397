// Adapter with the reference's extractor interface (modules/ORB/ORBExtractor.h:26-98) over the C-ABI of include/orbfe.h: namespace,
// class name, the two constructors, operator() and the static pyramid getters are the reference's, so Frame / ORBMatcher / MapPoint /
// LocalMapping code that calls them compiles unchanged.  Nothing else of the reference's class exists here — no ExtractorNode, no
// pyramid / FAST / quadtree members, no pattern table: that work happens in the CUDA kernels behind the handle.
#pragma once
#include <cassert>
#include <cmath>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "cv_compat.h"
#include "../../include/orbfe.h"

namespace mono_orb_slam3 {
    using Match = std::pair<unsigned int, unsigned int>;

    namespace detail {
        // The reference keeps the pyramid description in class statics shared by every extractor of the process (ORBExtractor.cpp:416-422);
        // the adapter keeps the same process-wide table in one place and lets the static getters read it.
        struct PyramidTable {
            int levels = 1;
            float factor = 1.f, log_factor = 1.f;
            std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
            void fill(orbfe_handle *h, float scaleFactor, int nLevels);
            float of(const std::vector<float> &column, int level) const { assert(level >= 0 && level < levels); return column[(size_t) level]; }
        };
        PyramidTable &pyramid();
    }

    class ORBExtractor {
    public:
        // ---- what the reference's callers use
        // device / maxBatch are additions with defaults (the reference has five arguments): device < 0 = defaultDevice()
        explicit ORBExtractor(int nFeatures = 1000, float scaleFactor = 1.2, int nLevels = 8, int iniThFast = 20, int minThFast = 10, int device = -1,
                              int maxBatch = 64);
        ORBExtractor(int nFeatures, const ORBExtractor &orbExtractor);       // Tracking.cpp:24: the initial extractor, another feature budget
        void operator()(const cv::Mat &image, std::vector<cv::KeyPoint> &keyPoints, cv::Mat &descriptors);

        static int getNumLevels() { return detail::pyramid().levels; }
        static float getLogScaleFactor() { return detail::pyramid().log_factor; }
        static float getMaxScaleFactor() { const auto &p = detail::pyramid(); return p.scale[(size_t) p.levels - 1]; }
        static float getScaleFactor(int level = 0) { const auto &p = detail::pyramid(); return p.of(p.scale, level); }
        static float getInvScaleFactor(int level) { const auto &p = detail::pyramid(); return p.of(p.inv_scale, level); }
        static float getSquareSigma(int level) { const auto &p = detail::pyramid(); return p.of(p.sigma2, level); }
        static float getInvSquareSigma(int level) { const auto &p = detail::pyramid(); return p.of(p.inv_sigma2, level); }
        static std::vector<float> getScaleFactors() { return detail::pyramid().scale; }
        static std::vector<float> getInvScaleFactors() { return detail::pyramid().inv_scale; }
        static std::vector<float> getSquareSigmas() { return detail::pyramid().sigma2; }

        void print() const;

        // public in the reference, read by nobody there: filled only on request (one download per level)
        bool keep_image_pyramid = false;
        std::vector<cv::Mat> image_pyramid;

        // ---- additions
        // equally sized frames in one call (offline workloads); outputs per frame
        void extractBatch(const std::vector<cv::Mat> &frames, std::vector<std::vector<cv::KeyPoint>> &keyPoints, std::vector<cv::Mat> &descriptors);
        orbfe_handle *handle() const { return handle_; }
        // CUDA device of extractors (and of the matcher's per-thread handles) created without an explicit one: setDefaultDevice(), else the
        // environment variable ORBFE_DEVICE, else 0
        static int defaultDevice();
        static void setDefaultDevice(int device);
        const std::vector<int> &featuresPerLevel() const { return quota_; }

        ~ORBExtractor();
        ORBExtractor(const ORBExtractor &) = delete;
        ORBExtractor &operator=(const ORBExtractor &) = delete;

    private:
        void open(float scaleFactor, int nLevels);
        int capacity() const;                   // key points one frame can produce: the quotas plus the quadtree's slack

        int budget_, fast_ini_, fast_min_;      // nFeatures, iniThFast, minThFast
        int device_ = 0, max_batch_ = 64;
        std::vector<int> quota_;                // per-level feature quota, as the library computed it
        orbfe_handle *handle_ = nullptr;
    };
} // mono_orb_slam3
