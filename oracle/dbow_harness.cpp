// TEST INFRASTRUCTURE ONLY (see oracle/orb_oracle.py's header): C entry points around the reference's own vendored DBoW2
// (thirdParty/DBoW2, compiled VERBATIM from where it lies, with oracle/cvshim for cv::Mat and oracle/boostshim for the
// serialisation hooks nothing here instantiates).  The vocabulary type is the one of modules/ORB/ORBVocabulary.h:12.
// Built by oracle/Makefile into oracle/_ref/libref_dbow.so; used by tests to pin oracle/bow.py (and through it the GPU descent).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <opencv2/core/core.hpp>
#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

namespace {
typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> Vocabulary;
struct OpenVocabulary : Vocabulary {          // the per-feature descent (TemplatedVocabulary.h:1217-1259) is a protected member
    using Vocabulary::transform;
};
std::vector<cv::Mat> rows_of(const uint8_t *desc, int n) {     // Frame::computeBow (Frame.cpp:168-178): one 1x32 header per row
    std::vector<cv::Mat> v((size_t) n);
    for (int i = 0; i < n; ++i) { v[(size_t) i].create(1, 32, CV_8U); std::memcpy(v[(size_t) i].ptr(), desc + 32 * (size_t) i, 32); }
    return v;
}
}  // namespace

extern "C" {

void *ref_voc_load(const char *path) {
    OpenVocabulary *v = new OpenVocabulary();
    if (!v->loadFromTextFile(path) || v->empty()) { delete v; return nullptr; }
    return v;
}
void ref_voc_free(void *p) { delete static_cast<OpenVocabulary *>(p); }
int ref_voc_words(void *p) { return (int) static_cast<OpenVocabulary *>(p)->size(); }
int ref_voc_k(void *p) { return static_cast<OpenVocabulary *>(p)->getBranchingFactor(); }
int ref_voc_depth(void *p) { return static_cast<OpenVocabulary *>(p)->getDepthLevels(); }

// per feature: word id, node id `levelsup` levels above the leaves, word weight
void ref_voc_transform_each(void *p, const uint8_t *desc, int n, int levelsup, int32_t *word_id, int32_t *node_id, double *weight) {
    const OpenVocabulary *v = static_cast<OpenVocabulary *>(p);
    std::vector<cv::Mat> f = rows_of(desc, n);
    for (int i = 0; i < n; ++i) {
        DBoW2::WordId w = 0; DBoW2::NodeId nid = 0; DBoW2::WordValue val = 0;
        v->transform(f[(size_t) i], w, val, &nid, levelsup);
        word_id[i] = (int32_t) w; node_id[i] = (int32_t) nid; weight[i] = val;
    }
}

// vocabulary->transform(vecDescriptor, bow_vector, feature_vector, levelsup): both maps flattened in std::map order.
// Returns 0, or -1 if a capacity is too small (the needed sizes are written either way).
int ref_voc_transform(void *p, const uint8_t *desc, int n, int levelsup, int32_t *bow_id, double *bow_val, int bow_cap, int *n_bow,
                      int32_t *fv_node, int32_t *fv_off, int node_cap, int *n_nodes, int32_t *fv_idx, int idx_cap, int *n_idx) {
    const OpenVocabulary *v = static_cast<OpenVocabulary *>(p);
    std::vector<cv::Mat> f = rows_of(desc, n);
    DBoW2::BowVector bv; DBoW2::FeatureVector fv;
    v->transform(f, bv, fv, levelsup);
    size_t total = 0;
    for (const auto &e : fv) total += e.second.size();
    *n_bow = (int) bv.size(); *n_nodes = (int) fv.size(); *n_idx = (int) total;
    if ((int) bv.size() > bow_cap || (int) fv.size() > node_cap || (int) total > idx_cap) return -1;
    int i = 0;
    for (const auto &e : bv) { bow_id[i] = (int32_t) e.first; bow_val[i] = e.second; ++i; }
    i = 0; int o = 0;
    fv_off[0] = 0;
    for (const auto &e : fv) {
        fv_node[i] = (int32_t) e.first;
        for (unsigned idx : e.second) fv_idx[o++] = (int32_t) idx;
        fv_off[++i] = o;
    }
    return 0;
}

}  // extern "C"
