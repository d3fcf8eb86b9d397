// ORBExtractor.cpp — adapter from the reference's ORBExtractor interface to liborbfe.so (see ORBExtractor.h).
#include "ORBExtractor.h"

#include <iostream>

namespace mono_orb_slam3 {
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbfe_keypoint), "cv::KeyPoint must be 7 x 4 bytes");

    float ORBExtractor::scale_factor = 1.f;
    float ORBExtractor::log_sale_factor = 1.f;
    int ORBExtractor::n_levels = 1;
    std::vector<float> ORBExtractor::scale_factors;
    std::vector<float> ORBExtractor::inv_scale_factors;
    std::vector<float> ORBExtractor::square_sigmas;
    std::vector<float> ORBExtractor::inv_square_sigmas;

    void ORBExtractor::createHandle(float scaleFactor, int nLevels) {
        orbfe_config cfg{};
        cfg.n_features = n_features; cfg.scale_factor = scaleFactor; cfg.n_levels = nLevels;
        cfg.ini_th_fast = ini_th_fast; cfg.min_th_fast = min_th_fast; cfg.device = 0; cfg.max_batch = 64; cfg.flags = 0;
        if (orbfe_create(&cfg, &handle_) != ORBFE_OK) throw std::runtime_error(std::string("orbfe_create: ") + orbfe_last_error(nullptr));
        n_features_per_level.resize(nLevels);
        for (int l = 0; l < nLevels; ++l) n_features_per_level[l] = orbfe_features_per_level(handle_, l);
    }

    ORBExtractor::ORBExtractor(int nFeatures, float scaleFactor, int nLevels, int iniThFast, int minThFast)
            : n_features(nFeatures), ini_th_fast(iniThFast), min_th_fast(minThFast) {
        createHandle(scaleFactor, nLevels);
        // the static pyramid tables are (re)initialised by the primary constructor, like the reference (ORBExtractor.cpp:427-439);
        // the values come from the handle so that host and device agree bit for bit
        scale_factor = scaleFactor;
        log_sale_factor = std::log(scaleFactor);
        n_levels = nLevels;
        scale_factors.resize(n_levels); inv_scale_factors.resize(n_levels);
        square_sigmas.resize(n_levels); inv_square_sigmas.resize(n_levels);
        for (int i = 0; i < n_levels; i++) {
            scale_factors[i] = orbfe_scale_factor(handle_, i);
            inv_scale_factors[i] = 1.f / scale_factors[i];
            square_sigmas[i] = scale_factors[i] * scale_factors[i];
            inv_square_sigmas[i] = 1.f / square_sigmas[i];
        }
        image_pyramid.resize(n_levels);
    }

    ORBExtractor::ORBExtractor(int nFeatures, const ORBExtractor &orbExtractor)
            : n_features(nFeatures), ini_th_fast(orbExtractor.ini_th_fast), min_th_fast(orbExtractor.min_th_fast) {
        createHandle(scale_factor, n_levels);
        image_pyramid.resize(n_levels);
    }

    ORBExtractor::~ORBExtractor() { orbfe_destroy(handle_); }

    void ORBExtractor::operator()(const cv::Mat &image, std::vector<cv::KeyPoint> &_keyPoints, cv::Mat &descriptors) {
        if (image.empty()) return;
        assert(image.type() == CV_8UC1);
        int cap = 0;
        for (int l = 0; l < n_levels; ++l) cap += n_features_per_level[l] + 40;
        std::vector<cv::KeyPoint> kps((size_t) cap);
        cv::Mat desc(cap, 32, CV_8U);
        int n = 0;
        const int rc = orbfe_extract(handle_, image.data, image.cols, image.rows, (size_t) image.step, reinterpret_cast<orbfe_keypoint *>(kps.data()),
                                     desc.data, cap, &n);
        if (rc != ORBFE_OK) throw std::runtime_error(std::string("orbfe_extract: ") + orbfe_last_error(handle_));
        if (keep_image_pyramid) {
            for (int l = 0; l < n_levels; ++l) {
                int w = 0, h = 0;
                orbfe_level_size(handle_, l, &w, &h);
                image_pyramid[l].create(h, w, CV_8U);
                orbfe_get_level_image(handle_, 0, l, image_pyramid[l].data);
            }
        }
        if (n == 0) return;                                   // the reference leaves its outputs untouched (ORBExtractor.cpp:512)
        kps.resize((size_t) n);
        _keyPoints.swap(kps);
        descriptors = desc.rowRange(0, n).clone();
    }

    void ORBExtractor::extractBatch(const std::vector<cv::Mat> &frames, std::vector<std::vector<cv::KeyPoint>> &keyPoints, std::vector<cv::Mat> &descriptors) {
        const int B = (int) frames.size();
        keyPoints.assign((size_t) B, {}); descriptors.assign((size_t) B, cv::Mat());
        if (B == 0) return;
        const int w = frames[0].cols, h = frames[0].rows;
        std::vector<unsigned char> packed((size_t) B * w * h);
        for (int b = 0; b < B; ++b) {
            if (frames[b].cols != w || frames[b].rows != h) throw std::invalid_argument("extractBatch: frames must have equal size");
            for (int r = 0; r < h; ++r) std::memcpy(&packed[((size_t) b * h + r) * w], frames[b].ptr(r), (size_t) w);
        }
        int cap = 0;
        for (int l = 0; l < n_levels; ++l) cap += n_features_per_level[l] + 40;
        std::vector<cv::KeyPoint> kps((size_t) B * cap);
        std::vector<unsigned char> desc((size_t) B * cap * 32);
        std::vector<int> n((size_t) B);
        const int rc = orbfe_extract_batch(handle_, packed.data(), B, w, h, (size_t) w, (size_t) w * h, reinterpret_cast<orbfe_keypoint *>(kps.data()),
                                           desc.data(), cap, n.data());
        if (rc != ORBFE_OK) throw std::runtime_error(std::string("orbfe_extract_batch: ") + orbfe_last_error(handle_));
        for (int b = 0; b < B; ++b) {
            keyPoints[b].assign(kps.begin() + (size_t) b * cap, kps.begin() + (size_t) b * cap + n[b]);
            descriptors[b].create(n[b], 32, CV_8U);
            if (n[b]) std::memcpy(descriptors[b].data, &desc[(size_t) b * cap * 32], (size_t) n[b] * 32);
        }
    }

    void ORBExtractor::print() const {
        std::cout << std::endl << "ORB Pyramid Information: " << std::endl;
        std::cout << " - Features: " << n_features << "(at initial stage)" << std::endl;
        std::cout << " - ScaleFactor: " << scale_factor << std::endl;
        std::cout << " - Levels: " << n_levels << std::endl;
        std::cout << " - IniThFAST: " << ini_th_fast << std::endl;
        std::cout << " - MinThFAST: " << min_th_fast << std::endl;
        std::cout << std::endl;
    }
} // mono_orb_slam3
