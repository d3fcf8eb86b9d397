// ORBExtractor.h — drop-in for the reference's modules/ORB/ORBExtractor.h: same namespace, class name, constructor arguments,
// operator() and static getters (ORBExtractor.h:26-98), implemented on the B200 through the C-ABI of include/orbfe.h.
// What is gone: ExtractorNode, ComputePyramid, ComputeKeyPointsOctTree, DistributeOctree, the pattern / u_max members — they
// live in the CUDA kernels now.  `image_pyramid` (public but unread in the reference) can still be filled on request.
#pragma once
#include <cassert>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>
#include "cv_compat.h"
#include "../../include/orbfe.h"

namespace mono_orb_slam3 {
    typedef std::pair<unsigned int, unsigned int> Match;

    class ORBExtractor {
    public:
        explicit ORBExtractor(int nFeatures = 1000, float scaleFactor = 1.2, int nLevels = 8, int iniThFast = 20, int minThFast = 10);

        // the "initial extractor" form (Tracking.cpp:24): another feature budget, same pyramid and thresholds
        ORBExtractor(int nFeatures, const ORBExtractor &orbExtractor);

        ~ORBExtractor();
        ORBExtractor(const ORBExtractor &) = delete;
        ORBExtractor &operator=(const ORBExtractor &) = delete;

        // Compute the pyramid features and descriptors on an image (CV_8UC1); key points are dispersed with the quadtree.
        void operator()(const cv::Mat &image, std::vector<cv::KeyPoint> &keyPoints, cv::Mat &descriptors);

        // Batched form for offline workloads: `frames` are equally sized images; outputs per frame.
        void extractBatch(const std::vector<cv::Mat> &frames, std::vector<std::vector<cv::KeyPoint>> &keyPoints, std::vector<cv::Mat> &descriptors);

        void print() const;

        inline static float getScaleFactor(int level = 0) { assert(level >= 0 && level < n_levels); return scale_factors[level]; }
        inline static float getLogScaleFactor() { return log_sale_factor; }
        inline static float getMaxScaleFactor() { return scale_factors[n_levels - 1]; }
        inline static std::vector<float> getScaleFactors() { return scale_factors; }
        inline static float getInvScaleFactor(int level) { assert(level >= 0 && level < n_levels); return inv_scale_factors[level]; }
        inline static std::vector<float> getInvScaleFactors() { return inv_scale_factors; }
        inline static int getNumLevels() { return n_levels; }
        inline static std::vector<float> getSquareSigmas() { return square_sigmas; }
        inline static float getSquareSigma(int level) { assert(level >= 0 && level < n_levels); return square_sigmas[level]; }
        inline static float getInvSquareSigma(int level) { assert(level >= 0 && level < n_levels); return inv_square_sigmas[level]; }

        // When set, operator() also downloads the pyramid levels into image_pyramid (off by default: nothing in the reference reads it).
        bool keep_image_pyramid = false;
        std::vector<cv::Mat> image_pyramid;

        orbfe_handle *handle() const { return handle_; }
        const std::vector<int> &featuresPerLevel() const { return n_features_per_level; }

    protected:
        void createHandle(float scaleFactor, int nLevels);

        int n_features;       // target num of features
        int ini_th_fast;      // initial threshold of FAST
        int min_th_fast;      // minimum threshold of FAST, used for cells where the initial threshold finds nothing

        // pyramid information: process-global like the reference's statics (ORBExtractor.cpp:416-422)
        static float scale_factor;
        static float log_sale_factor;
        static int n_levels;
        static std::vector<float> scale_factors;
        static std::vector<float> inv_scale_factors;
        static std::vector<float> square_sigmas;
        static std::vector<float> inv_square_sigmas;

        std::vector<int> n_features_per_level;
        orbfe_handle *handle_ = nullptr;
    };
} // mono_orb_slam3
