// orbfe_frame.cu — the Frame constructor's post-processing of the extractor output, on the device:
//
//   kp.size *= camera->uncertainty(kp.pt)                       BasicObject/Frame.cpp:24-26, Sensor/Pinhole.cpp:55-57, Fisheye.cpp:110-112
//   camera->undistortKeyPoints(raw_key_points, key_points)      Frame.cpp:28, Pinhole.cpp:59-84 (cv::undistortPoints, R = I, P = K),
//                                                               Fisheye.cpp:114-117 (copy)
//   grid[x / 40][y / 40].push_back(i) for PosInGrid(kp)         Frame.cpp:31-51, 90-95 (GRID_SIZE = 40, Frame.h:18)
//
// cv::undistortPoints is an un-vendored OpenCV call; its algorithm (cvUndistortPointsInternal: 5 fixed-point iterations of
// the rad-tan inverse in double, criteria MAX_ITER 5, then re-projection with P) is restated here with explicit round-to-nearest
// double intrinsics — no FMA contraction — which reproduces cv2 4.13.0 bit for bit (tests/golden/frame_post.npz).
#include "orbfe_internal.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace orbfe {

constexpr int kGridSize = 40;                // GRID_SIZE, Frame.h:18
constexpr int kGridIdxBits = 14;             // key = cell << 14 | key-point index
constexpr int kGridMaxKp = 1 << kGridIdxBits;

struct CamDev {
    int model;                               // ORBFE_CAMERA_PINHOLE / ORBFE_CAMERA_FISHEYE
    double fx, fy, cx, cy, k[12];            // k: OpenCV order k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4
    int undistort;                           // Pinhole.cpp:62: skipped when dist_coeffs[0] == 0
    const float *unc; int unc_w, unc_h;      // Fisheye scale_mat (row-major height x width) or null
};

// one thread per key-point slot of one frame
__global__ void k_frame_post(const CamDev cam, orbfe_keypoint *raw, orbfe_keypoint *un, const int *n_per_frame, int cap) {
    const int frame = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_per_frame[frame]) return;
    orbfe_keypoint kp = raw[(size_t) frame * cap + i];
    if (cam.unc) {                           // Fisheye::uncertainty: scale_mat.at<float>(p.y, p.x), float -> int conversion truncates
        const int px = min(max((int) kp.x, 0), cam.unc_w - 1), py = min(max((int) kp.y, 0), cam.unc_h - 1);
        kp.size = __fmul_rn(kp.size, __ldg(&cam.unc[(size_t) py * cam.unc_w + px]));
        raw[(size_t) frame * cap + i].size = kp.size;
    }                                        // Pinhole::uncertainty returns 1.f: size * 1.f is the same float
    if (cam.undistort) {
        const double ifx = __ddiv_rn(1.0, cam.fx), ify = __ddiv_rn(1.0, cam.fy);
        double x = __dmul_rn(__dsub_rn((double) kp.x, cam.cx), ifx), y = __dmul_rn(__dsub_rn((double) kp.y, cam.cy), ify);
        const double x0 = x, y0 = y;
        const double *k = cam.k;
#pragma unroll 1
        for (int j = 0; j < 5; ++j) {
            const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
            const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[7], r2), k[6]), r2), k[5]), r2));
            const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[4], r2), k[1]), r2), k[0]), r2));
            const double icdist = __ddiv_rn(num, den);
            if (icdist < 0) { x = x0; y = y0; break; }           // "test: undistortPoints.regression_14583"
            // deltaX = 2*k[2]*x*y + k[3]*(r2 + 2*x*x) + k[8]*r2 + k[9]*r2*r2   (left to right, as the C expression evaluates)
            const double dx = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, k[2]), x), y),
                                                            __dmul_rn(k[3], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x)))),
                                                  __dmul_rn(k[8], r2)),
                                        __dmul_rn(__dmul_rn(k[9], r2), r2));
            const double dy = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(k[2], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))),
                                                            __dmul_rn(__dmul_rn(__dmul_rn(2.0, k[3]), x), y)),
                                                  __dmul_rn(k[10], r2)),
                                        __dmul_rn(__dmul_rn(k[11], r2), r2));
            x = __dmul_rn(__dsub_rn(x0, dx), icdist);
            y = __dmul_rn(__dsub_rn(y0, dy), icdist);
        }
        // xx = RR[0][0]*x + RR[0][1]*y + RR[0][2] with RR = P = K; ww = 1. / (0*x + 0*y + 1)
        const double xx = __dadd_rn(__dadd_rn(__dmul_rn(cam.fx, x), __dmul_rn(0.0, y)), cam.cx);
        const double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(cam.fy, y)), cam.cy);
        const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(0.0, y)), 1.0));
        kp.x = (float) __dmul_rn(xx, ww); kp.y = (float) __dmul_rn(yy, ww);
    }
    un[(size_t) frame * cap + i] = kp;
}

// Frame grid as CSR, one CTA per frame: key = cell << 14 | index for key points inside the image (PosInGrid, Frame.cpp:90-95),
// bitonic sort in shared memory (cell-major, insertion order inside a cell = the reference's push_back order), then
// grid_idx = sorted indices and grid_off[c] = first position whose cell is >= c.  cell = cx * rows + cy (grid[cx][cy]).
__global__ void __launch_bounds__(1024) k_frame_grid(const orbfe_keypoint *un, const int *n_per_frame, int cap, int img_w, int img_h,
                                                      int cols, int rows, int sort_cap, int *grid_off, int *grid_idx, int *n_in_grid) {
    extern __shared__ uint32_t keys[];
    const int frame = blockIdx.x, tid = threadIdx.x;
    const int n = min(n_per_frame[frame], cap);
    const orbfe_keypoint *kp = un + (size_t) frame * cap;
    int cap2 = 32; while (cap2 < n) cap2 <<= 1;
    cap2 = min(cap2, sort_cap);
    for (int i = tid; i < cap2; i += 1024) {
        uint32_t key = 0xffffffffu;
        if (i < n) {
            const int x = (int) floorf(kp[i].x), y = (int) floorf(kp[i].y);                 // cvFloor
            if (x >= 0 && x < img_w && y >= 0 && y < img_h) key = ((uint32_t) ((x / kGridSize) * rows + y / kGridSize) << kGridIdxBits) | (uint32_t) i;
        }
        keys[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= cap2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < cap2; i += 1024) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint32_t a = keys[i], b = keys[ixj];
                    if ((a > b) == ((i & k) == 0)) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    int *off = grid_off + (size_t) frame * (cols * rows + 1), *idx = grid_idx + (size_t) frame * cap;
    for (int i = tid; i < n; i += 1024) if (keys[i] != 0xffffffffu) idx[i] = (int) (keys[i] & (kGridMaxKp - 1));
    for (int c = tid; c <= cols * rows; c += 1024) {                                        // lower bound of c << 14
        const uint32_t want = (uint32_t) c << kGridIdxBits;
        int lo = 0, hi = cap2;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < want) lo = mid + 1; else hi = mid; }
        // invalid keys (0xffffffff) sort last: for c == cols*rows the bound is the number of key points inside the image
        off[c] = lo;
        if (c == cols * rows && n_in_grid) n_in_grid[frame] = lo;
    }
}

static int grid_dims(int img_w, int img_h, int &cols, int &rows) {      // Frame.cpp:33-41
    cols = img_w % kGridSize == 0 ? img_w / kGridSize : img_w / kGridSize + 1;
    rows = img_h % kGridSize == 0 ? img_h / kGridSize : img_h / kGridSize + 1;
    return cols * rows;
}

static int cam_to_dev(Handle *h, const orbfe_camera *cam, int img_w, int img_h, CamDev &c) {
    if (!cam) return set_error(h, ORBFE_E_ARG, "camera is null");
    if (cam->model != ORBFE_CAMERA_PINHOLE && cam->model != ORBFE_CAMERA_FISHEYE) return set_error(h, ORBFE_E_ARG, "unknown camera model %d", cam->model);
    memset(&c, 0, sizeof c);
    c.model = cam->model;
    c.fx = (double) cam->fx; c.fy = (double) cam->fy; c.cx = (double) cam->cx; c.cy = (double) cam->cy;     // mat_K is CV_32F (Camera.cpp:20-21)
    const int nd = std::min(std::max(cam->n_dist, 0), 12);
    for (int i = 0; i < nd; ++i) c.k[i] = (double) cam->dist[i];
    c.undistort = cam->model == ORBFE_CAMERA_PINHOLE && nd > 0 && cam->dist[0] != 0.f;                       // Pinhole.cpp:62
    c.unc = nullptr;
    if (cam->model == ORBFE_CAMERA_FISHEYE && cam->uncertainty_map) {
        // device copy of scale_mat, cached on the handle by (host pointer, size)
        if (cam->uncertainty_w != img_w || cam->uncertainty_h != img_h) return set_error(h, ORBFE_E_ARG, "uncertainty map must be image-sized");
        const size_t bytes = sizeof(float) * (size_t) img_w * img_h;
        if (h->unc_host != cam->uncertainty_map || h->unc_bytes != bytes) {
            ORBFE_CUDA(h, cudaStreamSynchronize(h->stream));
            cudaFree(h->d_unc); h->d_unc = nullptr; h->unc_host = nullptr; h->unc_bytes = 0;
            ORBFE_CUDA(h, cudaMalloc(&h->d_unc, bytes));
            ORBFE_CUDA(h, cudaMemcpy(h->d_unc, cam->uncertainty_map, bytes, cudaMemcpyHostToDevice));
            h->unc_host = cam->uncertainty_map; h->unc_bytes = bytes;
        }
        c.unc = h->d_unc; c.unc_w = img_w; c.unc_h = img_h;
    }
    return ORBFE_OK;
}

// launches on `st`; all pointers are device pointers, slabs of `cap` key points per frame
int frame_post_launch(Handle *h, const CamDev &c, orbfe_keypoint *d_raw, orbfe_keypoint *d_un, const int *d_n, int n_frames, int cap,
                      int img_w, int img_h, int *d_grid_off, int *d_grid_idx, int *d_n_in_grid, cudaStream_t st) {
    if (n_frames <= 0) return ORBFE_OK;
    if (cap > kGridMaxKp) return set_error(h, ORBFE_E_ARG, "frame grid: at most %d key points per frame (got capacity %d)", kGridMaxKp, cap);
    int cols, rows; grid_dims(img_w, img_h, cols, rows);
    if ((long long) cols * rows >= (1ll << (32 - kGridIdxBits)) - 1) return set_error(h, ORBFE_E_ARG, "frame grid: image too large");
    k_frame_post<<<dim3((cap + 255) / 256, n_frames), 256, 0, st>>>(c, d_raw, d_un, d_n, cap);
    int sort_cap = 32; while (sort_cap < cap) sort_cap <<= 1;
    k_frame_grid<<<n_frames, 1024, (size_t) sort_cap * 4, st>>>(d_un, d_n, cap, img_w, img_h, cols, rows, sort_cap, d_grid_off, d_grid_idx, d_n_in_grid);
    h->launches += 2;
    ORBFE_CUDA(h, cudaGetLastError());
    return ORBFE_OK;
}

// grid of one device-resident key-point array (used by the window searches of orbfe_match.cu): d_n holds the count
int frame_device_setup(Handle *h) {        // per-device opt-in of k_frame_grid's sort buffer (called by orbfe_create)
    ORBFE_CUDA(h, cudaFuncSetAttribute(k_frame_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, kGridMaxKp * 4));
    return ORBFE_OK;
}

int frame_grid_launch(Handle *h, const orbfe_keypoint *d_kps, const int *d_n, int cap, int img_w, int img_h, int *d_grid_off, int *d_grid_idx, cudaStream_t st) {
    if (cap > kGridMaxKp) return set_error(h, ORBFE_E_ARG, "frame grid: at most %d key points per frame (got %d)", kGridMaxKp, cap);
    int cols, rows; grid_dims(img_w, img_h, cols, rows);
    int sort_cap = 32; while (sort_cap < cap) sort_cap <<= 1;
    k_frame_grid<<<1, 1024, (size_t) sort_cap * 4, st>>>(d_kps, d_n, cap, img_w, img_h, cols, rows, sort_cap, d_grid_off, d_grid_idx, nullptr);
    h->launches++;
    ORBFE_CUDA(h, cudaGetLastError());
    return ORBFE_OK;
}

}  // namespace orbfe

using namespace orbfe;

extern "C" {

int orbfe_grid_size(int img_w, int img_h, int *cols, int *rows) {
    if (img_w <= 0 || img_h <= 0) return ORBFE_E_ARG;
    int c, r; grid_dims(img_w, img_h, c, r);
    if (cols) *cols = c;
    if (rows) *rows = r;
    return ORBFE_OK;
}

int orbfe_frame_postprocess_device(orbfe_handle *h, const orbfe_camera *cam, orbfe_keypoint *d_kps_raw, orbfe_keypoint *d_kps_un,
                                   const int *d_n_per_frame, int n_frames, int cap, int img_w, int img_h,
                                   int32_t *d_grid_off, int32_t *d_grid_idx, int32_t *d_n_in_grid, void *stream, int sync) {
    if (!h) return ORBFE_E_ARG;
    if (!d_kps_raw || !d_kps_un || !d_n_per_frame || !d_grid_off || !d_grid_idx || n_frames < 0 || cap < 1 || img_w <= 0 || img_h <= 0)
        return set_error(h, ORBFE_E_ARG, "orbfe_frame_postprocess_device: invalid argument");
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    CamDev c; int rc = cam_to_dev(h, cam, img_w, img_h, c);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t) stream : h->stream;
    if ((rc = frame_post_launch(h, c, d_kps_raw, d_kps_un, d_n_per_frame, n_frames, cap, img_w, img_h, d_grid_off, d_grid_idx, d_n_in_grid, st))) return rc;
    if (sync) ORBFE_CUDA(h, cudaStreamSynchronize(st));
    return ORBFE_OK;
}

int orbfe_frame_postprocess(orbfe_handle *h, const orbfe_camera *cam, orbfe_keypoint *kps_raw, int n, int img_w, int img_h,
                            orbfe_keypoint *kps_un, int32_t *grid_off, int32_t *grid_idx, int *n_in_grid) {
    if (!h) return ORBFE_E_ARG;
    if (n < 0 || img_w <= 0 || img_h <= 0 || !grid_off || (n > 0 && (!kps_raw || !kps_un || !grid_idx)))
        return set_error(h, ORBFE_E_ARG, "orbfe_frame_postprocess: invalid argument");
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    int cols, rows; const int nc = grid_dims(img_w, img_h, cols, rows);
    if (n == 0) { for (int c = 0; c <= nc; ++c) grid_off[c] = 0; if (n_in_grid) *n_in_grid = 0; return ORBFE_OK; }
    CamDev c; int rc = cam_to_dev(h, cam, img_w, img_h, c);
    if (rc) return rc;
    cudaStream_t st = h->stream;
    const size_t kb = sizeof(orbfe_keypoint) * (size_t) n;
    const size_t need = 2 * (kb + 256) + sizeof(int) * ((size_t) nc + 1 + n + 2) + 1024;
    if ((rc = ensure_match_scratch(h, need))) return rc;
    uint8_t *p = (uint8_t *) h->d_match;
    orbfe_keypoint *d_raw = (orbfe_keypoint *) p; p += (kb + 255) & ~(size_t) 255;
    orbfe_keypoint *d_un = (orbfe_keypoint *) p; p += (kb + 255) & ~(size_t) 255;
    int *d_n = (int *) p; p += 256;
    int *d_off = (int *) p; p += (sizeof(int) * ((size_t) nc + 1) + 255) & ~(size_t) 255;
    int *d_idx = (int *) p;
    int *d_nin = d_n + 1;
    ORBFE_CUDA(h, cudaMemcpyAsync(d_raw, kps_raw, kb, cudaMemcpyHostToDevice, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(d_n, &n, sizeof(int), cudaMemcpyHostToDevice, st));
    if ((rc = frame_post_launch(h, c, d_raw, d_un, d_n, 1, n, img_w, img_h, d_off, d_idx, d_nin, st))) return rc;
    int nin = 0;
    ORBFE_CUDA(h, cudaMemcpyAsync(kps_raw, d_raw, kb, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(kps_un, d_un, kb, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(grid_off, d_off, sizeof(int) * ((size_t) nc + 1), cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(&nin, d_nin, sizeof(int), cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaStreamSynchronize(st));
    if (nin > 0) {
        ORBFE_CUDA(h, cudaMemcpyAsync(grid_idx, d_idx, sizeof(int) * (size_t) nin, cudaMemcpyDeviceToHost, st));
        ORBFE_CUDA(h, cudaStreamSynchronize(st));
    }
    if (n_in_grid) *n_in_grid = nin;
    return ORBFE_OK;
}

}  // extern "C"
