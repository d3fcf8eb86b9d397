// matcher_ref_test.cpp — drives the seven reference-shaped methods of the C++ matcher adapter (monoorbslam3_b200/host/ORBMatcher.h with
// ORBFE_REFERENCE_TYPES) on Frame / KeyFrame / MapPoint objects, on the GPU.  The object types are the stand-ins of oracle/matchshim —
// the same headers the reference's own ORBMatcher.cpp is compiled against for oracle/_ref/libref_matcher.so — so pytest can hand the
// same flat inputs to the verbatim reference (through oracle/ref_matcher.py) and compare what both did to the objects.
// usage: matcher_ref_test <in.bin> <out.bin>        (formats: see tests/test_cpp_matcher_ref_gpu.py)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>
#include "ORBMatcher.h"

using namespace mono_orb_slam3;

static FILE *fin, *fout;
template <class T> static std::vector<T> take(size_t n) {
    std::vector<T> v(n);
    if (n && fread(v.data(), sizeof(T), n, fin) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return v;
}
static int take_int() { return take<int>(1)[0]; }
static float take_float() { return take<float>(1)[0]; }
template <class T> static void put(const std::vector<T> &v) { if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), fout); }
static void put_int(int v) { fwrite(&v, sizeof v, 1, fout); }

static cv::Mat rows(const std::vector<uint8_t> &d, int n) { cv::Mat m(n > 0 ? n : 1, 32, CV_8U); if (n) memcpy(m.data, d.data(), (size_t) n * 32); return m; }
static cv::Mat row(const uint8_t *d) { cv::Mat m(1, 32, CV_8U); memcpy(m.data, d, 32); return m; }

template <class F> static std::shared_ptr<F> read_frame(int w, int h) {
    auto f = std::make_shared<F>();
    const int n = take_int();
    auto kps = take<cv::KeyPoint>((size_t) n);
    auto desc = take<uint8_t>((size_t) n * 32);
    f->key_points = kps; f->descriptors = rows(desc, n); f->width = w; f->height = h;
    f->finish();
    return f;
}
static void read_fv(DBoW2::FeatureVector &fv) {
    const int nn = take_int();
    auto id = take<int>((size_t) nn), off = take<int>((size_t) nn + 1);
    auto idx = take<int>((size_t) off[(size_t) nn]);
    for (int k = 0; k < nn; ++k)
        for (int j = off[(size_t) k]; j < off[(size_t) k + 1]; ++j) fv.addFeature((DBoW2::NodeId) id[(size_t) k], (unsigned) idx[(size_t) j]);
}
// map points of a projection search: state 0 = none, 1 = good, 2 = bad, 3 = behind the camera
static std::vector<std::shared_ptr<MapPoint>> read_points(int n) {
    auto state = take<uint8_t>((size_t) n);
    auto u = take<float>((size_t) n), v = take<float>((size_t) n);
    auto desc = take<uint8_t>((size_t) n * 32);
    std::vector<std::shared_ptr<MapPoint>> mps((size_t) n);
    for (int i = 0; i < n; ++i) {
        if (!state[(size_t) i]) continue;
        auto mp = std::make_shared<MapPoint>();
        const float z = state[(size_t) i] == 3 ? -1.f : 1.f;
        mp->pos = Eigen::Vector3f(u[(size_t) i] * z, v[(size_t) i] * z, z); mp->normal = Eigen::Vector3f(u[(size_t) i], v[(size_t) i], 1.f);
        mp->min_distance = 0.f; mp->max_distance = 3.0e38f;
        mp->bad = state[(size_t) i] == 2;
        mp->descriptor = row(desc.data() + 32 * (size_t) i);
        mps[(size_t) i] = mp;
    }
    return mps;
}
static void occupy(FrameBase &f, const std::shared_ptr<MapPoint> &blocker) {
    auto occ = take<uint8_t>((size_t) f.num_kps);
    for (int j = 0; j < f.num_kps; ++j) f.map_points[(size_t) j] = occ[(size_t) j] ? blocker : nullptr;
}
// which query's map point ended up in each slot (-1: none / the pre-existing blocker)
static std::vector<int> owners(const FrameBase &f, const std::vector<std::shared_ptr<MapPoint>> &mps, const std::shared_ptr<MapPoint> &blocker) {
    std::vector<int> out((size_t) f.num_kps, -1);
    for (int j = 0; j < f.num_kps; ++j) {
        const auto &p = f.map_points[(size_t) j];
        if (p && p != blocker)
            for (size_t i = 0; i < mps.size(); ++i) if (mps[i] == p) { out[(size_t) j] = (int) i; break; }
    }
    return out;
}

int main(int argc, char **argv) {
    if (argc != 3) return 2;
    fin = fopen(argv[1], "rb"); fout = fopen(argv[2], "wb");
    if (!fin || !fout) return 2;
    const int w = take_int(), h = take_int();
    ORBExtractor extractor(1000, 1.2f, 8, 20, 7);            // fills the process-wide pyramid table the matcher reads (as Tracking's constructor does)
    Camera::instance()->width = w; Camera::instance()->height = h;
    auto blocker = std::make_shared<MapPoint>();

    {   // 1. SearchForInitialization(frame1, frame2, vecPreMatched, matches12, windowSize)
        auto f1 = read_frame<Frame>(w, h), f2 = read_frame<Frame>(w, h);
        const int window = take_int(); const float ratio = take_float(); const int orient = take_int();
        auto pre = take<cv::Point2f>((size_t) f1->num_kps);
        std::vector<int> m12;
        ORBMatcher matcher(ratio, orient != 0);
        put_int(matcher.SearchForInitialization(f1, f2, pre, m12, window));
        put(m12); put(pre);
    }
    for (int from_kf = 0; from_kf < 2; ++from_kf) {   // 2. / 3. SearchByProjection(lastFrame | lastKF, curFrame, th)
        auto cur = read_frame<Frame>(w, h);
        const int nq = take_int(); const float th = take_float(); const int orient = take_int();
        auto kps = take<cv::KeyPoint>((size_t) nq);
        auto mps = read_points(nq);
        occupy(*cur, blocker);
        ORBMatcher matcher(0.6f, orient != 0);
        int n;
        if (from_kf) {
            auto last = std::make_shared<KeyFrame>(); last->key_points = kps; last->width = w; last->height = h; last->finish(); last->map_points = mps;
            n = matcher.SearchByProjection(last, cur, th);
        } else {
            auto last = std::make_shared<Frame>(); last->key_points = kps; last->width = w; last->height = h; last->finish(); last->map_points = mps;
            n = matcher.SearchByProjection(last, cur, th);
        }
        put_int(n); put(owners(*cur, mps, blocker));
    }
    {   // 4. SearchByProjection(frame, mapPoints, th)
        auto fr = read_frame<Frame>(w, h);
        const int nq = take_int(); const float th = take_float(); const float ratio = take_float();
        auto mps = read_points(nq);
        auto in_view = take<uint8_t>((size_t) nq); auto vc = take<float>((size_t) nq); auto lvl = take<int>((size_t) nq);
        auto pu = take<float>((size_t) nq), pv = take<float>((size_t) nq);
        for (int i = 0; i < nq; ++i) {
            auto &mp = mps[(size_t) i];
            mp->track_in_view = in_view[(size_t) i] != 0; mp->track_view_cos = vc[(size_t) i]; mp->track_scale_level = lvl[(size_t) i];
            mp->track_proj_x = pu[(size_t) i]; mp->track_proj_y = pv[(size_t) i];
        }
        occupy(*fr, blocker);
        ORBMatcher matcher(ratio, true);
        put_int(matcher.SearchByProjection(fr, mps, th)); put(owners(*fr, mps, blocker));
    }
    {   // 5. SearchForTriangulation(keyFrame1, keyFrame2, matches12)
        auto k1 = read_frame<KeyFrame>(w, h), k2 = read_frame<KeyFrame>(w, h);
        read_fv(k1->feature_vector); read_fv(k2->feature_vector);
        occupy(*k1, blocker); occupy(*k2, blocker);
        const int orient = take_int();
        std::vector<int> m12;
        ORBMatcher matcher(0.6f, orient != 0);
        put_int(matcher.SearchForTriangulation(k1, k2, m12)); put(m12);
    }
    {   // 6. SearchByBow(keyFrame, frame)
        auto kf = read_frame<KeyFrame>(w, h); auto fr = read_frame<Frame>(w, h);
        read_fv(kf->feature_vector); read_fv(fr->feature_vector);
        auto state = take<uint8_t>((size_t) kf->num_kps);            // 0 none, 1 good, 2 bad
        std::vector<std::shared_ptr<MapPoint>> mps((size_t) kf->num_kps);
        for (int i = 0; i < kf->num_kps; ++i) if (state[(size_t) i]) { mps[(size_t) i] = std::make_shared<MapPoint>(); mps[(size_t) i]->bad = state[(size_t) i] == 2; }
        kf->map_points = mps;
        occupy(*fr, blocker);
        const float ratio = take_float(); const int orient = take_int();
        ORBMatcher matcher(ratio, orient != 0);
        put_int(matcher.SearchByBow(kf, fr)); put(owners(*fr, mps, blocker));
    }
    {   // 7. fuse: static SearchByProjection(keyFrame, mapPoints, map, th)
        auto kf = read_frame<KeyFrame>(w, h);
        const int nq = take_int(); const float th = take_float();
        auto mps = read_points(nq);
        auto lvl = take<int>((size_t) nq); auto nobs = take<int>((size_t) nq);
        for (int i = 0; i < nq; ++i) if (mps[(size_t) i]) { mps[(size_t) i]->predicted_level = lvl[(size_t) i]; mps[(size_t) i]->num_obs = nobs[(size_t) i]; }
        auto slot_obs = take<int>((size_t) kf->num_kps);             // key-frame slots that already hold a point: its observation count, or -1
        for (int j = 0; j < kf->num_kps; ++j)
            if (slot_obs[(size_t) j] >= 0) { kf->map_points[(size_t) j] = std::make_shared<MapPoint>(); kf->map_points[(size_t) j]->num_obs = slot_obs[(size_t) j]; }
        Map map;
        put_int(ORBMatcher::SearchByProjection(kf, mps, &map, th));
        std::vector<int> fused((size_t) nq, -1), replaced((size_t) nq, 0);
        for (int i = 0; i < nq; ++i) if (mps[(size_t) i]) { fused[(size_t) i] = mps[(size_t) i]->fused_idx; replaced[(size_t) i] = (mps[(size_t) i]->replaced_by_other ? 1 : 0) | (mps[(size_t) i]->replaced_other ? 2 : 0); }
        put(fused); put(replaced); put(kf->queried);
        put_int(-12345);                                             // end marker (kf->queried has a data-dependent length)
    }
    fclose(fin); fclose(fout);
    return 0;
}
