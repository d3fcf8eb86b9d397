// ORBExtractor.cpp — adapter from the reference's ORBExtractor interface to liborbfe.so (see ORBExtractor.h).
#include "ORBExtractor.h"

#include <cstdlib>
#include <iostream>

namespace mono_orb_slam3 {
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbfe_keypoint), "cv::KeyPoint must be 7 x 4 bytes");

    namespace detail {
        PyramidTable &pyramid() { static PyramidTable table; return table; }

        // values come from the handle so that host and device agree bit for bit
        void PyramidTable::fill(orbfe_handle *h, float scaleFactor, int nLevels) {
            levels = nLevels; factor = scaleFactor; log_factor = std::log(scaleFactor);
            scale.assign((size_t) nLevels, 1.f); inv_scale = scale; sigma2 = scale; inv_sigma2 = scale;
            for (int l = 0; l < nLevels; ++l) {
                const float s = orbfe_scale_factor(h, l);
                scale[(size_t) l] = s; inv_scale[(size_t) l] = 1.f / s;
                sigma2[(size_t) l] = s * s; inv_sigma2[(size_t) l] = 1.f / (s * s);
            }
        }
    }

    void ORBExtractor::open(float scaleFactor, int nLevels) {
        orbfe_config cfg{};
        cfg.n_features = budget_; cfg.scale_factor = scaleFactor; cfg.n_levels = nLevels;
        cfg.ini_th_fast = fast_ini_; cfg.min_th_fast = fast_min_; cfg.device = device_; cfg.max_batch = max_batch_; cfg.flags = 0;
        if (orbfe_create(&cfg, &handle_) != ORBFE_OK) throw std::runtime_error(std::string("orbfe_create: ") + orbfe_last_error(nullptr));
        quota_.resize((size_t) nLevels);
        for (int l = 0; l < nLevels; ++l) quota_[(size_t) l] = orbfe_features_per_level(handle_, l);
        image_pyramid.resize((size_t) nLevels);
    }

    // the library's own bound (per level max(quota + 4, 4 * roots + 4) once the frame geometry is known, a safe estimate before)
    int ORBExtractor::capacity() const { return orbfe_max_keypoints(handle_); }

    static int g_default_device = -1;
    int ORBExtractor::defaultDevice() {
        if (g_default_device >= 0) return g_default_device;
        const char *e = std::getenv("ORBFE_DEVICE");
        return e ? std::atoi(e) : 0;
    }
    void ORBExtractor::setDefaultDevice(int device) { g_default_device = device; }

    // the primary constructor (re)initialises the process-wide pyramid table, like the reference's (ORBExtractor.cpp:427-439)
    ORBExtractor::ORBExtractor(int nFeatures, float scaleFactor, int nLevels, int iniThFast, int minThFast, int device, int maxBatch)
            : budget_(nFeatures), fast_ini_(iniThFast), fast_min_(minThFast), device_(device >= 0 ? device : defaultDevice()), max_batch_(maxBatch > 0 ? maxBatch : 64) {
        open(scaleFactor, nLevels);
        detail::pyramid().fill(handle_, scaleFactor, nLevels);
    }

    ORBExtractor::ORBExtractor(int nFeatures, const ORBExtractor &orbExtractor)
            : budget_(nFeatures), fast_ini_(orbExtractor.fast_ini_), fast_min_(orbExtractor.fast_min_), device_(orbExtractor.device_),
              max_batch_(orbExtractor.max_batch_) {
        open(detail::pyramid().factor, detail::pyramid().levels);
    }

    ORBExtractor::~ORBExtractor() { orbfe_destroy(handle_); }

    void ORBExtractor::operator()(const cv::Mat &image, std::vector<cv::KeyPoint> &_keyPoints, cv::Mat &descriptors) {
        if (image.empty()) return;
        assert(image.type() == CV_8UC1);
        const int cap = capacity();
        std::vector<cv::KeyPoint> kps((size_t) cap);
        cv::Mat desc(cap, 32, CV_8U);
        int n = 0;
        const int rc = orbfe_extract(handle_, image.data, image.cols, image.rows, (size_t) image.step, reinterpret_cast<orbfe_keypoint *>(kps.data()),
                                     desc.data, cap, &n);
        if (rc != ORBFE_OK) throw std::runtime_error(std::string("orbfe_extract: ") + orbfe_last_error(handle_));
        if (keep_image_pyramid) {
            for (int l = 0; l < detail::pyramid().levels; ++l) {
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
        const int cap = capacity();
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
        const detail::PyramidTable &p = detail::pyramid();
        std::cout << "\nORB extractor (B200): " << budget_ << " features, scale factor " << p.factor << ", " << p.levels << " levels, FAST thresholds "
                  << fast_ini_ << " / " << fast_min_ << "\n" << std::endl;
    }
} // mono_orb_slam3
