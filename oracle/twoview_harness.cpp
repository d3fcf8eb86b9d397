// TEST INFRASTRUCTURE ONLY (see oracle/orb_oracle.py's header): C entry points around the reference's own
// Frontend/TwoViewReconstruction.cpp, compiled VERBATIM (with its own header) against oracle/twoviewshim/Eigen/Dense (name-level
// stand-in: element access and the 3 x 3 inverse are real, the decompositions are not executed) and the cv:: shim.  Exercises
// CheckHomography (TwoViewReconstruction.cpp:226-288) and CheckFundamental (:290-345) on caller-supplied matches and hypotheses.
// Built by oracle/Makefile into oracle/_ref/libref_twoview.so.
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>
#include <Eigen/Dense>
#include <opencv2/core/core.hpp>  // everything the reference's header includes comes first, so that the redefinition below touches only it
#define private public            // the scoring functions and the match arrays are private members of the reference's class
#include "Frontend/TwoViewReconstruction.h"
#undef private

using namespace mono_orb_slam3;

namespace {
void fill(TwoViewReconstruction &tv, const float *pts1, const float *pts2, int n) {
    tv.key_points1.resize((size_t) n); tv.key_points2.resize((size_t) n); tv.match_pairs.resize((size_t) n);
    for (int i = 0; i < n; ++i) {
        tv.key_points1[(size_t) i].pt = cv::Point2f(pts1[2 * i], pts1[2 * i + 1]);
        tv.key_points2[(size_t) i].pt = cv::Point2f(pts2[2 * i], pts2[2 * i + 1]);
        tv.match_pairs[(size_t) i] = std::make_pair(i, i);
    }
    tv.num_matches = n;
}
Eigen::Matrix3f mat(const float *m) { Eigen::Matrix3f M; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M(i, j) = m[3 * i + j]; return M; }
}  // namespace

extern "C" {

// score of CheckHomography(H21); inliers[n]; H12_used = the inverse the function worked with (the stand-in's, handed to the restatement)
float ref_check_homography(const float *H21, const float *pts1, const float *pts2, int n, float sigma, uint8_t *inliers, float *H12_used) {
    TwoViewReconstruction tv(Eigen::Matrix3f::Identity(), sigma, 200);
    fill(tv, pts1, pts2, n);
    const Eigen::Matrix3f H = mat(H21);
    const Eigen::Matrix3f Hi = H.inverse();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) H12_used[3 * i + j] = Hi(i, j);
    std::vector<bool> in;
    const float s = tv.CheckHomography(H, in);
    for (int i = 0; i < n; ++i) inliers[i] = in[(size_t) i];
    return s;
}

float ref_check_fundamental(const float *F21, const float *pts1, const float *pts2, int n, float sigma, uint8_t *inliers) {
    TwoViewReconstruction tv(Eigen::Matrix3f::Identity(), sigma, 200);
    fill(tv, pts1, pts2, n);
    std::vector<bool> in;
    const float s = tv.CheckFundamental(mat(F21), in);
    for (int i = 0; i < n; ++i) inliers[i] = in[(size_t) i];
    return s;
}

}  // extern "C"
