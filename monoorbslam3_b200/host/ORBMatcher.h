// ORBMatcher.h — drop-in for the reference's modules/ORB/ORBMatcher.h on top of include/orbfe.h.
// Same class name, constructor and method names.
//   * Compiled inside the reference tree (define ORBFE_REFERENCE_TYPES; the include path then provides BasicObject/Frame.h,
//     BasicObject/Map.h and Sensor/Camera.h exactly as for the reference's own header) the class has the reference's seven signatures
//     (ORBMatcher.h:14-45) on shared_ptr<Frame> / shared_ptr<KeyFrame> / vector<shared_ptr<MapPoint>>: each method does what the head and
//     the tail of the reference's loop do on the host — validity tests, Pose::map, Camera::project, radius and level, descriptor
//     gathering, and the write-back into map_points / matches12 (ORBMatcher.cpp:212-229, 246, 354-369, 417-447, 524-551, 573-586) —
//     and runs the window / node search and the greedy resolve on the device through the C-ABI.  Tracking.cpp and LocalMapping.cpp
//     compile against it unchanged.
//   * Without ORBFE_REFERENCE_TYPES the header needs no reference type at all: SearchForInitialization is a template over any frame
//     type with key_points / descriptors / num_kps (and img or width / height), and the projection searches take the flattened
//     queries (struct Queries) that the methods above build.
#pragma once
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "cv_compat.h"
#include "ORBExtractor.h"
#include "../../include/orbfe.h"
#ifdef ORBFE_REFERENCE_TYPES
#include "BasicObject/Frame.h"
#include "BasicObject/Map.h"
#include "Sensor/Camera.h"
#endif

namespace mono_orb_slam3 {

    namespace detail {
        // image size of a frame: the reference's Frame keeps the image (Frame.h:52); other frame types may carry width / height
        template <class F> auto frame_size(const F &f, int &w, int &h, int) -> decltype(f.img.cols, void()) { w = f.img.cols; h = f.img.rows; }
        template <class F> auto frame_size(const F &f, int &w, int &h, long) -> decltype(f.width, void()) { w = f.width; h = f.height; }
    }

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
            int n = 0, w = 0, h = 0;
            detail::frame_size(*frame2, w, h, 0);
            check(orbfe_search_for_initialization(handle(), reinterpret_cast<const orbfe_keypoint *>(frame1->key_points.data()), frame1->descriptors.data,
                                                  frame1->num_kps, reinterpret_cast<const orbfe_keypoint *>(frame2->key_points.data()),
                                                  frame2->descriptors.data, frame2->num_kps, w, h,
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
            void resize(int n) {
                u.assign((size_t) n, 0.f); v = u; radius = u; angle = u; level.assign((size_t) n, 0); valid.assign((size_t) n, 0);
                descriptors.create(n > 0 ? n : 1, 32, CV_8U);
            }
            void setDescriptor(int i, const cv::Mat &d) { std::memcpy(descriptors.ptr(i), d.ptr(0), 32); }
        };

        // SearchByProjection(lastFrame | lastKF, curFrame, th): assigned[j] = query written into curFrame->map_points[j], or -1
        template <class FrameT>
        int SearchByProjection(const Queries &q, const std::shared_ptr<FrameT> &curFrame, const std::vector<uint8_t> &occupied, std::vector<int> &assigned) const {
            assigned.assign((size_t) curFrame->num_kps, -1);
            int n = 0, w = 0, h = 0;
            detail::frame_size(*curFrame, w, h, 0);
            check(orbfe_search_by_projection(handle(), q.u.data(), q.v.data(), q.radius.data(), q.level.data(), q.angle.data(), q.descriptors.data,
                                             q.valid.data(), q.size(), reinterpret_cast<const orbfe_keypoint *>(curFrame->key_points.data()),
                                             curFrame->descriptors.data, curFrame->num_kps, w, h, occupied.data(),
                                             assigned.data(), be_check_orientation ? 1 : 0, &n));
            return n;
        }

        // SearchByProjection(frame, mapPoints, th)
        template <class FrameT>
        int SearchLocalPoints(const Queries &q, const std::shared_ptr<FrameT> &frame, const std::vector<uint8_t> &occupied, std::vector<int> &assigned) const {
            assigned.assign((size_t) frame->num_kps, -1);
            int n = 0, w = 0, h = 0;
            detail::frame_size(*frame, w, h, 0);
            check(orbfe_search_local_points(handle(), q.u.data(), q.v.data(), q.radius.data(), q.level.data(), q.descriptors.data, q.valid.data(), q.size(),
                                            reinterpret_cast<const orbfe_keypoint *>(frame->key_points.data()), frame->descriptors.data, frame->num_kps,
                                            w, h, occupied.data(), assigned.data(), nn_ratio, &n));
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


#ifdef ORBFE_REFERENCE_TYPES
        // ------------------------------------------------------------------------------------------------------------------
        // The reference's own signatures (ORBMatcher.h:25-45).  Heads and tails of the reference loops on the host, search on the device.
        // ------------------------------------------------------------------------------------------------------------------

        /// Tracking
        [[nodiscard]] int SearchByBow(const std::shared_ptr<KeyFrame> &keyFrame, const std::shared_ptr<Frame> &frame) const {
            const std::vector<std::shared_ptr<MapPoint>> mapPoints = keyFrame->getMapPoints();                    // ORBMatcher.cpp:119
            std::vector<uint8_t> valid1((size_t) keyFrame->num_kps, 0), occupied2((size_t) frame->num_kps, 0);
            for (int i = 0; i < keyFrame->num_kps; ++i) valid1[(size_t) i] = mapPoints[(size_t) i] != nullptr && !mapPoints[(size_t) i]->isBad();   // :143-144
            for (int j = 0; j < frame->num_kps; ++j) occupied2[(size_t) j] = frame->map_points[(size_t) j] != nullptr;    // :151
            std::vector<int> assigned;
            const int n = SearchByBow(keyFrame->descriptors, angles(keyFrame->key_points), valid1, featureVector(keyFrame->feature_vector), frame->descriptors,
                                      angles(frame->key_points), occupied2, featureVector(frame->feature_vector), assigned);
            for (int j = 0; j < frame->num_kps; ++j)
                if (assigned[(size_t) j] >= 0) frame->map_points[(size_t) j] = mapPoints[(size_t) assigned[(size_t) j]];  // :165 (rotation rejects come back as -1)
            return n;
        }

        [[nodiscard]] int SearchByProjection(const std::shared_ptr<Frame> &lastFrame, const std::shared_ptr<Frame> &curFrame, float th = 5) const {
            return projectLast(lastFrame->map_points, lastFrame->key_points, lastFrame->num_kps, curFrame, th);   // ORBMatcher.cpp:203-274
        }

        [[nodiscard]] int SearchByProjection(const std::shared_ptr<KeyFrame> &lastKF, const std::shared_ptr<Frame> &curFrame, float th = 5) const {
            return projectLast(lastKF->getMapPoints(), lastKF->key_points, lastKF->num_kps, curFrame, th);        // ORBMatcher.cpp:276-348
        }

        [[nodiscard]] int SearchByProjection(const std::shared_ptr<Frame> &frame, const std::vector<std::shared_ptr<MapPoint>> &mapPoints, float th = 3) const {
            const int nq = (int) mapPoints.size();                                                                // ORBMatcher.cpp:350-415
            Queries q; q.resize(nq);
            for (int i = 0; i < nq; ++i) {
                const std::shared_ptr<MapPoint> &mp = mapPoints[(size_t) i];
                if (!mp->track_in_view || mp->isBad()) continue;                                                  // :354-357
                const int predictLevel = mp->track_scale_level;
                float radius = th;                                                                                // :361-364
                if (mp->track_view_cos > 0.998) radius *= 2.5f; else radius *= 4.f;
                radius *= ORBExtractor::getScaleFactor(predictLevel);
                q.u[(size_t) i] = mp->track_proj_x; q.v[(size_t) i] = mp->track_proj_y; q.radius[(size_t) i] = radius; q.level[(size_t) i] = predictLevel;
                q.setDescriptor(i, mp->getDescriptor());
                q.valid[(size_t) i] = 1;
            }
            std::vector<uint8_t> occupied((size_t) frame->num_kps, 0);
            for (int j = 0; j < frame->num_kps; ++j) { const auto &p = frame->map_points[(size_t) j]; occupied[(size_t) j] = p != nullptr && !p->isBad(); }   // :380
            std::vector<int> assigned;
            const int n = SearchLocalPoints(q, frame, occupied, assigned);
            for (int j = 0; j < frame->num_kps; ++j)
                if (assigned[(size_t) j] >= 0) frame->map_points[(size_t) j] = mapPoints[(size_t) assigned[(size_t) j]];  // :406
            return n;
        }

        /// Local Mapping
        int SearchForTriangulation(const std::shared_ptr<KeyFrame> &keyFrame1, const std::shared_ptr<KeyFrame> &keyFrame2, std::vector<int> &matches12) const {
            std::vector<uint8_t> has1((size_t) keyFrame1->num_kps, 0), has2((size_t) keyFrame2->num_kps, 0);      // :452, :466
            for (int i = 0; i < keyFrame1->num_kps; ++i) has1[(size_t) i] = keyFrame1->hasMapPoint(i);
            for (int j = 0; j < keyFrame2->num_kps; ++j) has2[(size_t) j] = keyFrame2->hasMapPoint(j);
            return SearchForTriangulation(keyFrame1->descriptors, angles(keyFrame1->key_points), has1, featureVector(keyFrame1->feature_vector),
                                          keyFrame2->descriptors, angles(keyFrame2->key_points), has2, featureVector(keyFrame2->feature_vector), matches12);
        }

        // fuse (ORBMatcher.cpp:524-592).  Precondition: a map point appears once in mapPoints (the reference's own callers pass lists
        // without duplicates: LocalMapping.cpp:268-301); the searches of all points run on the device first, the observation / replace
        // bookkeeping then walks them in order and re-tests isBad / isObserveKeyFrame at each turn, as the sequential loop does.
        static int SearchByProjection(const std::shared_ptr<KeyFrame> &keyFrame, const std::vector<std::shared_ptr<MapPoint>> &mapPoints, Map *pointMap, float th = 3) {
            (void) pointMap;
            const Camera *camera = Camera::getCamera();
            const Pose Tcw = keyFrame->getPose();
            const Eigen::Vector3f Ow = keyFrame->getCameraCenter();
            const int nq = (int) mapPoints.size();
            Queries q; q.resize(nq);
            for (int i = 0; i < nq; ++i) {
                const std::shared_ptr<MapPoint> &mp = mapPoints[(size_t) i];
                if (mp == nullptr || mp->isBad() || mp->isObserveKeyFrame(keyFrame)) continue;                     // :532
                const Eigen::Vector3f Pw = mp->getPos();
                const Eigen::Vector3f Pc = Tcw.R * Pw + Tcw.t;
                if (Pc[2] < 0) continue;
                const cv::Point2f p = camera->project(Pc);
                if (!camera->isInImage(p)) continue;
                const Eigen::Vector3f OP = Pw - Ow;
                const float distance = OP.norm();
                if (distance < mp->getMinDistanceInvariance() || distance > mp->getMaxDistanceInvariance()) continue;   // :545
                const Eigen::Vector3f Pn = mp->getAverageDirection();
                if (OP.dot(Pn) < 0.5 * distance) continue;                                                        // :548
                const int predictLevel = mp->predictScaleLevel(distance);
                q.u[(size_t) i] = p.x; q.v[(size_t) i] = p.y; q.radius[(size_t) i] = th * ORBExtractor::getScaleFactor(predictLevel);
                q.level[(size_t) i] = predictLevel;
                q.setDescriptor(i, mp->getDescriptor());
                q.valid[(size_t) i] = 1;
            }
            std::vector<int> bestIdx((size_t) (nq > 0 ? nq : 1), -1);
            int n_found = 0, w = 0, h = 0;
            detail::frame_size(*keyFrame, w, h, 0);
            const std::vector<float> sigma2 = ORBExtractor::getSquareSigmas();                                   // the gate of :564
            check(orbfe_search_fuse_sigma(handle(), q.u.data(), q.v.data(), q.radius.data(), q.level.data(), q.descriptors.data, q.valid.data(), nq,
                                          reinterpret_cast<const orbfe_keypoint *>(keyFrame->key_points.data()), keyFrame->descriptors.data, keyFrame->num_kps,
                                          w, h, sigma2.data(), (int) sigma2.size(), bestIdx.data(), nullptr, &n_found));
            int numMatch = 0;
            for (int i = 0; i < nq; ++i) {                                                                        // :573-586, in list order
                if (!q.valid[(size_t) i] || bestIdx[(size_t) i] < 0) continue;
                const std::shared_ptr<MapPoint> &mp = mapPoints[(size_t) i];
                if (mp->isBad() || mp->isObserveKeyFrame(keyFrame)) continue;                                      // state an earlier replace may have changed
                const int bestIdx1 = bestIdx[(size_t) i];
                std::shared_ptr<MapPoint> mp1 = keyFrame->getMapPoint(bestIdx1);
                if (mp1 == nullptr) {
                    mp->addObservation(keyFrame, bestIdx1);
                    keyFrame->addMapPoint(mp, bestIdx1);
                } else if (!mp1->isBad()) {
                    if (mp1->getNumObs() > mp->getNumObs()) mp->replace(mp1);
                    else mp1->replace(mp);
                }
                numMatch++;
            }
            return numMatch;
        }

    private:
        static std::vector<float> angles(const std::vector<cv::KeyPoint> &kps) {
            std::vector<float> a(kps.size());
            for (size_t i = 0; i < kps.size(); ++i) a[i] = kps[i].angle;
            return a;
        }
        template <class FV> static FeatureVector featureVector(const FV &fv) {       // DBoW2::FeatureVector is a std::map<NodeId, std::vector<unsigned int>>
            FeatureVector out;
            for (const auto &node: fv) out.emplace_hint(out.end(), (unsigned int) node.first, node.second);
            return out;
        }
        // shared body of SearchByProjection(lastFrame | lastKF, curFrame, th): ORBMatcher.cpp:212-229 / 286-303 (head), 246 / 320 (tail)
        int projectLast(const std::vector<std::shared_ptr<MapPoint>> &lastPoints, const std::vector<cv::KeyPoint> &lastKps, int numLast,
                        const std::shared_ptr<Frame> &curFrame, float th) const {
            const Camera *camera = Camera::getCamera();
            const Pose &Tcw = curFrame->T_cw;
            Queries q; q.resize(numLast);
            for (int i = 0; i < numLast; ++i) {
                const std::shared_ptr<MapPoint> &mp = lastPoints[(size_t) i];
                if (mp == nullptr || mp->isBad()) continue;
                const Eigen::Vector3f Pw = mp->getPos();
                const Eigen::Vector3f Pc = Tcw.map(Pw);
                if (Pc[2] < 0) continue;
                const cv::Point2f p = camera->project(Pc);
                if (!camera->isInImage(p)) continue;
                const cv::KeyPoint &kp = lastKps[(size_t) i];
                q.u[(size_t) i] = p.x; q.v[(size_t) i] = p.y; q.radius[(size_t) i] = th * kp.size; q.level[(size_t) i] = kp.octave; q.angle[(size_t) i] = kp.angle;
                q.setDescriptor(i, mp->getDescriptor());
                q.valid[(size_t) i] = 1;
            }
            std::vector<uint8_t> occupied((size_t) curFrame->num_kps, 0);
            for (int j = 0; j < curFrame->num_kps; ++j) occupied[(size_t) j] = curFrame->map_points[(size_t) j] != nullptr;   // :235
            std::vector<int> assigned;
            const int n = SearchByProjection(q, curFrame, occupied, assigned);
            for (int j = 0; j < curFrame->num_kps; ++j)
                if (assigned[(size_t) j] >= 0) curFrame->map_points[(size_t) j] = lastPoints[(size_t) assigned[(size_t) j]];  // :246 (rotation rejects come back as -1)
            return n;
        }

    public:
#endif  // ORBFE_REFERENCE_TYPES

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
