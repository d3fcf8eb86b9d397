// ORBMatcher.cpp — the non-template parts of the matcher adapter (see ORBMatcher.h).
#include "ORBMatcher.h"

namespace mono_orb_slam3 {

    // One handle (= one CUDA stream + scratch) per host thread: the reference runs ORBMatcher on the tracking and on the local-mapping
    // thread at the same time (System.cpp:55).  The handle follows the extractor's process-wide pyramid (scale factor, levels) and
    // device, and is re-created when an extractor with other values has been constructed since.
    orbfe_handle *ORBMatcher::handle() {
        struct Holder {
            orbfe_handle *h = nullptr;
            float factor = 0.f; int levels = 0, device = -1;
            void open() {
                const detail::PyramidTable &p = detail::pyramid();
                const float f = p.factor > 1.f ? p.factor : 1.2f;
                const int l = p.scale.empty() ? 8 : p.levels, d = ORBExtractor::defaultDevice();
                if (h && f == factor && l == levels && d == device) return;
                if (h) orbfe_destroy(h);
                h = nullptr;
                orbfe_config cfg{1000, f, l, 20, 7, d, 1, 0};
                if (orbfe_create(&cfg, &h) != ORBFE_OK) throw std::runtime_error(std::string("orbfe_create: ") + orbfe_last_error(nullptr));
                factor = f; levels = l; device = d;
            }
            ~Holder() { if (h) orbfe_destroy(h); }
        };
        static thread_local Holder holder;
        holder.open();
        return holder.h;
    }

    static void flatten(const ORBMatcher::FeatureVector &fv, std::vector<int> &ids, std::vector<int> &off, std::vector<int> &idx) {
        ids.clear(); idx.clear(); off.assign(1, 0);
        for (const auto &node: fv) {                               // std::map iterates in ascending node id, like the reference's merge (:443-512)
            ids.push_back((int) node.first);
            for (unsigned int i: node.second) idx.push_back((int) i);
            off.push_back((int) idx.size());
        }
    }

    int ORBMatcher::SearchForTriangulation(const cv::Mat &desc1, const std::vector<float> &angle1, const std::vector<uint8_t> &hasMapPoint1, const FeatureVector &fv1,
                                           const cv::Mat &desc2, const std::vector<float> &angle2, const std::vector<uint8_t> &hasMapPoint2, const FeatureVector &fv2,
                                           std::vector<int> &matches12) const {
        std::vector<int> id1, off1, idx1, id2, off2, idx2;
        flatten(fv1, id1, off1, idx1); flatten(fv2, id2, off2, idx2);
        matches12.assign((size_t) desc1.rows, -1);
        int n = 0;
        check(orbfe_search_for_triangulation(handle(), desc1.data, angle1.data(), hasMapPoint1.data(), desc1.rows, id1.data(), off1.data(), idx1.data(), (int) id1.size(),
                                             desc2.data, angle2.data(), hasMapPoint2.data(), desc2.rows, id2.data(), off2.data(), idx2.data(), (int) id2.size(),
                                             matches12.data(), be_check_orientation ? 1 : 0, &n));
        return n;
    }

    int ORBMatcher::SearchByBow(const cv::Mat &desc1, const std::vector<float> &angle1, const std::vector<uint8_t> &validMapPoint1, const FeatureVector &fv1,
                                const cv::Mat &desc2, const std::vector<float> &angle2, const std::vector<uint8_t> &occupied2, const FeatureVector &fv2,
                                std::vector<int> &assigned) const {
        std::vector<int> id1, off1, idx1, id2, off2, idx2;
        flatten(fv1, id1, off1, idx1); flatten(fv2, id2, off2, idx2);
        assigned.assign((size_t) desc2.rows, -1);
        int n = 0;
        check(orbfe_search_by_bow(handle(), desc1.data, angle1.data(), validMapPoint1.data(), desc1.rows, id1.data(), off1.data(), idx1.data(), (int) id1.size(),
                                  desc2.data, angle2.data(), occupied2.data(), desc2.rows, id2.data(), off2.data(), idx2.data(), (int) id2.size(),
                                  assigned.data(), nn_ratio, be_check_orientation ? 1 : 0, &n));
        return n;
    }

    int ORBMatcher::HammingAllPairs(const cv::Mat &q, const cv::Mat &t, std::vector<int> &bestIdx, std::vector<int> &bestDist, std::vector<int> &secondDist) {
        bestIdx.assign((size_t) q.rows, -1); bestDist.assign((size_t) q.rows, 257); secondDist.assign((size_t) q.rows, 257);
        check(orbfe_hamming_allpairs(handle(), q.data, q.rows, t.data, t.rows, bestIdx.data(), bestDist.data(), secondDist.data()));
        return q.rows;
    }

    // Host version kept for callers that build their own rotation histograms; the device resolves use the same rule.
    void ORBMatcher::ComputeThreeMaxima(std::vector<int> *histo, int &ind1, int &ind2, int &ind3) {
        int max1 = 0, max2 = -1, max3 = -2;
        for (int i = 0; i < 30; ++i) {
            const int n = (int) histo[i].size();
            if (n > max1) { max3 = max2; max2 = max1; max1 = n; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (n > max2) { max3 = max2; max2 = n; ind3 = ind2; ind2 = i; }
            else if (n > max3) { max3 = n; ind3 = i; }
        }
        if (max2 < max1 / 10) { ind2 = -1; ind3 = -1; }
        else if (max3 < max1 / 10) ind3 = -1;
    }
} // mono_orb_slam3
