// orbfe_geom.cu — RANSAC hypothesis scoring of TwoViewReconstruction on the device:
//   CheckHomography   Frontend/TwoViewReconstruction.cpp:226-288   symmetric transfer error of every match under H21 / H12
//   CheckFundamental  Frontend/TwoViewReconstruction.cpp:290-345   point-to-epipolar-line distances under F21
// for all hypotheses of FindHomography / FindFundamental (:86-160, 200 iterations) at once.  float32 arithmetic in the reference's
// operation order with explicit round-to-nearest intrinsics (the reference is built without FMA), and the score of a hypothesis is
// summed sequentially over the matches in match order, so scores and inlier flags are bit-identical, not just close.
// One CTA per hypothesis: the threads evaluate the per-match terms into shared memory, thread 0 adds them up in order.
#include "orbfe_internal.cuh"

namespace orbfe {

constexpr int kGeomChunk = 4096;       // matches per shared-memory round

template <bool kFundamental>
__global__ void __launch_bounds__(256) k_check_hypotheses(const float *M21, const float *M12, const float2 *p1, const float2 *p2, int n, float inv_sigma2,
                                                           float *scores, uint8_t *inliers) {
    __shared__ float s_t[2 * kGeomChunk];
    const int hyp = blockIdx.x, tid = threadIdx.x;
    const float *A = M21 + 9 * (size_t) hyp;
    const float m11 = A[0], m12 = A[1], m13 = A[2], m21 = A[3], m22 = A[4], m23 = A[5], m31 = A[6], m32 = A[7], m33 = A[8];
    float i11 = 0, i12 = 0, i13 = 0, i21 = 0, i22 = 0, i23 = 0, i31 = 0, i32 = 0, i33 = 0;
    if (!kFundamental) {
        const float *B = M12 + 9 * (size_t) hyp;
        i11 = B[0]; i12 = B[1]; i13 = B[2]; i21 = B[3]; i22 = B[4]; i23 = B[5]; i31 = B[6]; i32 = B[7]; i33 = B[8];
    }
    const float th = kFundamental ? 3.841f : 5.991f, th_score = 5.991f;
    float score = 0.f;
    for (int base = 0; base < n; base += kGeomChunk) {
        const int m = min(kGeomChunk, n - base);
        for (int j = tid; j < m; j += 256) {
            const float2 a = p1[base + j], b = p2[base + j];
            const float u1 = a.x, v1 = a.y, u2 = b.x, v2 = b.y;
            float chi1, chi2;
            if (!kFundamental) {
                // x1' = H12 * x2 (:259-265)
                const float w2 = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(__fmul_rn(i31, u2), __fmul_rn(i32, v2)), i33));
                const float u2in1 = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(i11, u2), __fmul_rn(i12, v2)), i13), w2);
                const float v2in1 = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(i21, u2), __fmul_rn(i22, v2)), i23), w2);
                const float du1 = __fsub_rn(u1, u2in1), dv1 = __fsub_rn(v1, v2in1);
                chi1 = __fmul_rn(__fadd_rn(__fmul_rn(du1, du1), __fmul_rn(dv1, dv1)), inv_sigma2);
                // x2' = H21 * x1 (:272-278)
                const float w1 = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(__fmul_rn(m31, u1), __fmul_rn(m32, v1)), m33));
                const float u1in2 = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(m11, u1), __fmul_rn(m12, v1)), m13), w1);
                const float v1in2 = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(m21, u1), __fmul_rn(m22, v1)), m23), w1);
                const float du2 = __fsub_rn(u2, u1in2), dv2 = __fsub_rn(v2, v1in2);
                chi2 = __fmul_rn(__fadd_rn(__fmul_rn(du2, du2), __fmul_rn(dv2, dv2)), inv_sigma2);
            } else {
                // l2 = F21 * x1 (:314-321)
                const float a2 = __fadd_rn(__fadd_rn(__fmul_rn(m11, u1), __fmul_rn(m12, v1)), m13);
                const float b2 = __fadd_rn(__fadd_rn(__fmul_rn(m21, u1), __fmul_rn(m22, v1)), m23);
                const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(m31, u1), __fmul_rn(m32, v1)), m33);
                const float num2 = __fadd_rn(__fadd_rn(__fmul_rn(a2, u2), __fmul_rn(b2, v2)), c2);
                chi1 = __fmul_rn(__fdiv_rn(__fmul_rn(num2, num2), __fadd_rn(__fmul_rn(a2, a2), __fmul_rn(b2, b2))), inv_sigma2);
                // l1 = x2^T * F21 (:328-335)
                const float a1 = __fadd_rn(__fadd_rn(__fmul_rn(m11, u2), __fmul_rn(m21, v2)), m31);
                const float b1 = __fadd_rn(__fadd_rn(__fmul_rn(m12, u2), __fmul_rn(m22, v2)), m32);
                const float c1 = __fadd_rn(__fadd_rn(__fmul_rn(m13, u2), __fmul_rn(m23, v2)), m33);
                const float num1 = __fadd_rn(__fadd_rn(__fmul_rn(a1, u1), __fmul_rn(b1, v1)), c1);
                chi2 = __fmul_rn(__fdiv_rn(__fmul_rn(num1, num1), __fadd_rn(__fmul_rn(a1, a1), __fmul_rn(b1, b1))), inv_sigma2);
            }
            const bool in1 = !(chi1 > th), in2 = !(chi2 > th);                 // "if (chi > th) inlier = false; else score += ..." (NaN adds)
            s_t[2 * j] = in1 ? __fsub_rn(th_score, chi1) : 0.f;
            s_t[2 * j + 1] = in2 ? __fsub_rn(th_score, chi2) : 0.f;
            if (inliers) inliers[(size_t) hyp * n + base + j] = (uint8_t) (in1 && in2);
        }
        __syncthreads();
        if (tid == 0)
            for (int j = 0; j < 2 * m; ++j) score = __fadd_rn(score, s_t[j]);   // match order; a rejected term adds +0, which changes nothing
        __syncthreads();
    }
    if (tid == 0) scores[hyp] = score;
}

template <bool kFundamental>
static int run_check(Handle *h, const float *M21, const float *M12, int n_hyp, const float *pts1, const float *pts2, int n, float sigma, float *scores,
                     uint8_t *inliers) {
    if (n_hyp < 0 || n < 0 || (n_hyp && (!M21 || !scores || (!kFundamental && !M12))) || (n && (!pts1 || !pts2)) || !(sigma > 0))
        return set_error(h, ORBFE_E_ARG, "hypothesis scoring: invalid argument");
    if (n_hyp == 0) return ORBFE_OK;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const size_t mb = sizeof(float) * 9 * (size_t) n_hyp, pb = sizeof(float) * 2 * (size_t) std::max(n, 1);
    auto up256 = [](size_t v) { return (v + 255) & ~(size_t) 255; };
    const size_t need = 2 * up256(mb) + 2 * up256(pb) + up256(sizeof(float) * (size_t) n_hyp) + up256((size_t) n_hyp * std::max(n, 1)) + 1024;
    int rc = ensure_match_scratch(h, need);
    if (rc) return rc;
    uint8_t *p = (uint8_t *) h->d_match;
    float *d21 = (float *) p; p += up256(mb);
    float *d12 = (float *) p; p += up256(mb);
    float2 *dp1 = (float2 *) p; p += up256(pb);
    float2 *dp2 = (float2 *) p; p += up256(pb);
    float *ds = (float *) p; p += up256(sizeof(float) * (size_t) n_hyp);
    uint8_t *di = inliers ? p : nullptr;
    ORBFE_CUDA(h, cudaMemcpyAsync(d21, M21, mb, cudaMemcpyHostToDevice, st));
    if (!kFundamental) ORBFE_CUDA(h, cudaMemcpyAsync(d12, M12, mb, cudaMemcpyHostToDevice, st));
    if (n) { ORBFE_CUDA(h, cudaMemcpyAsync(dp1, pts1, pb, cudaMemcpyHostToDevice, st)); ORBFE_CUDA(h, cudaMemcpyAsync(dp2, pts2, pb, cudaMemcpyHostToDevice, st)); }
    const float sigma2 = sigma * sigma;                                          // TwoViewReconstruction.h:20
    k_check_hypotheses<kFundamental><<<n_hyp, 256, 0, st>>>(d21, d12, dp1, dp2, n, 1.f / sigma2, ds, di);
    h->launches++;
    ORBFE_CUDA(h, cudaGetLastError());
    ORBFE_CUDA(h, cudaMemcpyAsync(scores, ds, sizeof(float) * (size_t) n_hyp, cudaMemcpyDeviceToHost, st));
    if (inliers && n) ORBFE_CUDA(h, cudaMemcpyAsync(inliers, di, (size_t) n_hyp * n, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaStreamSynchronize(st));
    return ORBFE_OK;
}

}  // namespace orbfe

using namespace orbfe;

extern "C" {

int orbfe_check_homography(orbfe_handle *h, const float *H21, const float *H12, int n_hyp, const float *pts1, const float *pts2, int n_matches, float sigma,
                           float *scores, uint8_t *inliers) {
    if (!h) return ORBFE_E_ARG;
    return run_check<false>(h, H21, H12, n_hyp, pts1, pts2, n_matches, sigma, scores, inliers);
}

int orbfe_check_fundamental(orbfe_handle *h, const float *F21, int n_hyp, const float *pts1, const float *pts2, int n_matches, float sigma, float *scores,
                            uint8_t *inliers) {
    if (!h) return ORBFE_E_ARG;
    return run_check<true>(h, F21, nullptr, n_hyp, pts1, pts2, n_matches, sigma, scores, inliers);
}

}  // extern "C"
