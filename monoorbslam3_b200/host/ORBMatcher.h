// ORBMatcher.h — drop-in for the reference's modules/ORB/ORBMatcher.h on top of include/orbfe.h.
// Same class name, constructor and method names.  The reference's methods take Frame / KeyFrame / MapPoint objects; the adapter is a
// template over the frame type so that it compiles against the reference's own BasicObject/Frame.h unchanged and against any
// struct with the same members (key_points, descriptors, num_kps, img).  Projection searches take the flattened queries the
// reference's loop heads compute (projection, radius, level: ORBMatcher.cpp:212-229, 354-369) — see INTEGRATION.md §3.
#pragma once
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "cv_compat.h"
#include "../../include/orbfe.h"

namespace mono_orb_slam3 {

    class ORBMatcher {
    public:
        explicit ORBMatcher(float nnRatio = 0.6, bool checkOrientation = true) : nn_ratio(nnRatio), be_check_orientation(checkOrientation) {}

        // Hamming distance between two 256-bit descriptors (single pair: evaluated on the host, a launch is not worth it)
        static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b) {
            const uint32_t *pa = a.ptr<uint32_t>(), *pb = b.ptr<uint32_t>();
            int dist = 0;
            for (int i = 0; i < 8; ++i) dist += __builtin_popcount(pa[i] ^ pb[i]);
            return dist;
        }

        /// Initialization
        template <class FrameT>
        int SearchForInitialization(const std::shared_ptr<FrameT> &frame1, const std::shared_ptr<FrameT> &frame2,
                                    std::vector<cv::Point2f> &vecPreMatched, std::vector<int> &matches12, int windowSize = 100) const {
            matches12.assign((size_t) frame1->num_kps, -1);
            int n = 0;
            check(orbfe_search_for_initialization(handle(), reinterpret_cast<const orbfe_keypoint *>(frame1->key_points.data()), frame1->descriptors.data,
                                                  frame1->num_kps, reinterpret_cast<const orbfe_keypoint *>(frame2->key_points.data()),
                                                  frame2->descriptors.data, frame2->num_kps, frame2->img.cols, frame2->img.rows,
                                                  reinterpret_cast<float *>(vecPreMatched.data()), matches12.data(), windowSize, nn_ratio,
                                                  be_check_orientation ? 1 : 0, &n));
            return n;
        }

        /// Tracking: the flattened form of "project every map point, search its window"
        struct Queries {
            std::vector<float> u, v, radius, angle;      // projection, th * kp.size (or th * {2.5|4} * scale), last key-point angle
            std::vector<int> level;                      // last octave / predicted level
            std::vector<uint8_t> valid;                  // 0: map point missing, bad, behind the camera or outside the image
            cv::Mat descriptors;                         // N x 32, MapPoint::getDescriptor()
            int size() const { return (int) u.size(); }
        };

        // SearchByProjection(lastFrame | lastKF, curFrame, th): assigned[j] = query written into curFrame->map_points[j], or -1
        template <class FrameT>
        int SearchByProjection(const Queries &q, const std::shared_ptr<FrameT> &curFrame, const std::vector<uint8_t> &occupied, std::vector<int> &assigned) const {
            assigned.assign((size_t) curFrame->num_kps, -1);
            int n = 0;
            check(orbfe_search_by_projection(handle(), q.u.data(), q.v.data(), q.radius.data(), q.level.data(), q.angle.data(), q.descriptors.data,
                                             q.valid.data(), q.size(), reinterpret_cast<const orbfe_keypoint *>(curFrame->key_points.data()),
                                             curFrame->descriptors.data, curFrame->num_kps, curFrame->img.cols, curFrame->img.rows, occupied.data(),
                                             assigned.data(), be_check_orientation ? 1 : 0, &n));
            return n;
        }

        // SearchByProjection(frame, mapPoints, th)
        template <class FrameT>
        int SearchLocalPoints(const Queries &q, const std::shared_ptr<FrameT> &frame, const std::vector<uint8_t> &occupied, std::vector<int> &assigned) const {
            assigned.assign((size_t) frame->num_kps, -1);
            int n = 0;
            check(orbfe_search_local_points(handle(), q.u.data(), q.v.data(), q.radius.data(), q.level.data(), q.descriptors.data, q.valid.data(), q.size(),
                                            reinterpret_cast<const orbfe_keypoint *>(frame->key_points.data()), frame->descriptors.data, frame->num_kps,
                                            frame->img.cols, frame->img.rows, occupied.data(), assigned.data(), nn_ratio, &n));
            return n;
        }

        /// Local Mapping: feature vectors are DBoW2::FeatureVector = std::map<NodeId, std::vector<unsigned>>
        typedef std::map<unsigned int, std::vector<unsigned int>> FeatureVector;
        int SearchForTriangulation(const cv::Mat &desc1, const std::vector<float> &angle1, const std::vector<uint8_t> &hasMapPoint1, const FeatureVector &fv1,
                                   const cv::Mat &desc2, const std::vector<float> &angle2, const std::vector<uint8_t> &hasMapPoint2, const FeatureVector &fv2,
                                   std::vector<int> &matches12) const;

        /// Relocalisation / reference-key-frame tracking: SearchByBow(keyFrame, frame) (ORBMatcher.cpp:118-201).
        /// validMapPoint1[i]: key-frame key point i has a map point that is not bad; occupied2[j]: frame->map_points[j] is set.
        /// assigned[j] = key-frame key-point index whose map point goes into frame->map_points[j], or -1.
        int SearchByBow(const cv::Mat &desc1, const std::vector<float> &angle1, const std::vector<uint8_t> &validMapPoint1, const FeatureVector &fv1,
                        const cv::Mat &desc2, const std::vector<float> &angle2, const std::vector<uint8_t> &occupied2, const FeatureVector &fv2,
                        std::vector<int> &assigned) const;

        /// Offline: brute-force best / second best of every row of `q` against `t`
        static int HammingAllPairs(const cv::Mat &q, const cv::Mat &t, std::vector<int> &bestIdx, std::vector<int> &bestDist, std::vector<int> &secondDist);

        static orbfe_handle *handle();      // one handle (= one CUDA stream) per host thread: tracking and local mapping run concurrently

    protected:
        static void check(int rc) { if (rc != ORBFE_OK) throw std::runtime_error(std::string("orbfe: ") + orbfe_last_error(handle())); }
        static void ComputeThreeMaxima(std::vector<int> *histo, int &ind1, int &ind2, int &ind3);

        float nn_ratio;
        bool be_check_orientation;
    };
} // mono_orb_slam3
