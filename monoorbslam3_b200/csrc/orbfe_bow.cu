// orbfe_bow.cu — DBoW2 vocabulary descent on the device: what Frame::computeBow / KeyFrame::computeBow
// (BasicObject/Frame.cpp:168-178, KeyFrame.cpp:213-223) ask of thirdParty/DBoW2/DBoW2/TemplatedVocabulary.h:
//   transform(feature, word_id, weight, nid, levelsup)   :1217-1259   descend from the root taking the child with the strictly
//                                                                      smallest FORB::distance (FORB.cpp:81-101), first minimum wins
//   transform(features, BowVector, FeatureVector, levelsup) :1127-1172 the FeatureVector groups the features by the node reached at
//                                                                      level L - levelsup (std::map order), stopped words excluded
// The vocabulary is given the way loadFromTextFile (:1338-1420) reads it: per node parent id, leaf flag, descriptor, weight, in
// file order; children lists and word ids are rebuilt exactly as the loader does.
#include "orbfe_internal.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

struct orbfe_vocab {
    orbfe_handle *h = nullptr;
    int k = 0, L = 0, n_nodes = 0, n_words = 0;
    uint4 *d_desc = nullptr;          // [n_nodes][2]
    int *d_child_off = nullptr;       // [n_nodes + 1]
    int *d_child_idx = nullptr;       // children in list order
    int *d_word_id = nullptr;         // [n_nodes]
    double *d_weight = nullptr;       // [n_nodes]
};

namespace orbfe {

__device__ __forceinline__ int hamming256v(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// one warp per feature: lanes = children of the current node (more than 32 children are walked in rounds)
__global__ void __launch_bounds__(256) k_bow_transform(const uint4 *feat, int n, const uint4 *vdesc, const int *child_off, const int *child_idx,
                                                        const int *word_of, const double *weight_of, int L, int levelsup,
                                                        int *word_id, int *node_id, double *weight) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const uint4 f0 = __ldg(feat + 2 * (size_t) i), f1 = __ldg(feat + 2 * (size_t) i + 1);
    const int nid_level = L - levelsup;
    int nid = 0, final_id = 0, level = 0;                       // :1224, :1226
    while (true) {
        const int cb = __ldg(child_off + final_id), ce = __ldg(child_off + final_id + 1);
        if (cb == ce) break;                                     // isLeaf(): no children
        ++level;
        uint32_t best = 0xffffffffu;                             // dist << 16 | position in the children list: first minimum wins
        for (int c0 = cb; c0 < ce; c0 += 32) {
            const int c = c0 + lane;
            if (c < ce) {
                const int id = __ldg(child_idx + c);
                const int d = hamming256v(f0, f1, __ldg(vdesc + 2 * (size_t) id), __ldg(vdesc + 2 * (size_t) id + 1));
                best = min(best, ((uint32_t) d << 16) | (uint32_t) (c - cb));
            }
        }
        best = __reduce_min_sync(0xffffffffu, best);
        final_id = __ldg(child_idx + cb + (int) (best & 0xffffu));
        if (level == nid_level) nid = final_id;                  // :1247-1248
    }
    if (lane == 0) { word_id[i] = __ldg(word_of + final_id); node_id[i] = nid; weight[i] = __ldg(weight_of + final_id); }
}

}  // namespace orbfe

using namespace orbfe;

extern "C" {

int orbfe_vocab_create(orbfe_handle *h, int k, int L, int n_nodes, const int32_t *parent, const uint8_t *is_leaf, const uint8_t *desc,
                       const double *weight, orbfe_vocab **out) {
    if (!h) return ORBFE_E_ARG;
    if (!out || n_nodes < 2 || !parent || !is_leaf || !desc || !weight || L < 1) return set_error(h, ORBFE_E_ARG, "orbfe_vocab_create: invalid argument");
    *out = nullptr;
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    // children lists in file order, word ids in leaf order (TemplatedVocabulary.h:1386-1414)
    std::vector<int> cnt((size_t) n_nodes + 1, 0), word((size_t) n_nodes, 0);
    int n_words = 0;
    for (int i = 1; i < n_nodes; ++i) {
        if (parent[i] < 0 || parent[i] >= i) return set_error(h, ORBFE_E_ARG, "orbfe_vocab_create: node %d has parent %d (must precede it)", i, parent[i]);
        cnt[(size_t) parent[i] + 1]++;
        if (is_leaf[i]) word[(size_t) i] = n_words++;
    }
    for (int i = 0; i < n_nodes; ++i) cnt[(size_t) i + 1] += cnt[(size_t) i];
    std::vector<int> idx((size_t) std::max(n_nodes - 1, 1)), fill(cnt.begin(), cnt.end() - 1);
    for (int i = 1; i < n_nodes; ++i) idx[(size_t) fill[(size_t) parent[i]]++] = i;
    orbfe_vocab *v = new orbfe_vocab();
    v->h = h; v->k = k; v->L = L; v->n_nodes = n_nodes; v->n_words = n_words;
    cudaError_t e;
    if ((e = cudaMalloc(&v->d_desc, 32 * (size_t) n_nodes)) != cudaSuccess || (e = cudaMalloc(&v->d_child_off, sizeof(int) * ((size_t) n_nodes + 1))) != cudaSuccess ||
        (e = cudaMalloc(&v->d_child_idx, sizeof(int) * idx.size())) != cudaSuccess || (e = cudaMalloc(&v->d_word_id, sizeof(int) * (size_t) n_nodes)) != cudaSuccess ||
        (e = cudaMalloc(&v->d_weight, sizeof(double) * (size_t) n_nodes)) != cudaSuccess ||
        (e = cudaMemcpy(v->d_desc, desc, 32 * (size_t) n_nodes, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(v->d_child_off, cnt.data(), sizeof(int) * ((size_t) n_nodes + 1), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(v->d_child_idx, idx.data(), sizeof(int) * idx.size(), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(v->d_word_id, word.data(), sizeof(int) * (size_t) n_nodes, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(v->d_weight, weight, sizeof(double) * (size_t) n_nodes, cudaMemcpyHostToDevice)) != cudaSuccess) {
        cudaFree(v->d_desc); cudaFree(v->d_child_off); cudaFree(v->d_child_idx); cudaFree(v->d_word_id); cudaFree(v->d_weight);
        delete v;
        return set_error(h, ORBFE_E_CUDA, "orbfe_vocab_create: %s", cudaGetErrorString(e));
    }
    *out = v;
    return ORBFE_OK;
}

void orbfe_vocab_destroy(orbfe_vocab *v) {
    if (!v) return;
    cudaSetDevice(v->h->device);
    cudaFree(v->d_desc); cudaFree(v->d_child_off); cudaFree(v->d_child_idx); cudaFree(v->d_word_id); cudaFree(v->d_weight);
    delete v;
}

int orbfe_vocab_words(const orbfe_vocab *v) { return v ? v->n_words : 0; }

int orbfe_vocab_transform(orbfe_vocab *v, const uint8_t *desc, int n, int levelsup, int32_t *word_id, int32_t *node_id, double *weight,
                          int32_t *fv_node_id, int32_t *fv_off, int32_t *fv_idx, int *fv_n_nodes) {
    if (!v) return ORBFE_E_ARG;
    orbfe_handle *h = v->h;
    if (n < 0 || (n && (!desc || !word_id || !node_id || !weight))) return set_error(h, ORBFE_E_ARG, "orbfe_vocab_transform: invalid argument");
    if (fv_n_nodes) *fv_n_nodes = 0;
    if (n == 0) { if (fv_off) fv_off[0] = 0; return ORBFE_OK; }
    ORBFE_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const size_t need = 32 * (size_t) n + 256 + (sizeof(int) * 2 + sizeof(double)) * (size_t) n + 1024;
    int rc = ensure_match_scratch(h, need);
    if (rc) return rc;
    uint8_t *p = (uint8_t *) h->d_match;
    uint4 *d_f = (uint4 *) p; p += (32 * (size_t) n + 255) & ~(size_t) 255;
    double *d_w = (double *) p; p += (sizeof(double) * (size_t) n + 255) & ~(size_t) 255;
    int *d_wid = (int *) p; p += (sizeof(int) * (size_t) n + 255) & ~(size_t) 255;
    int *d_nid = (int *) p;
    ORBFE_CUDA(h, cudaMemcpyAsync(d_f, desc, 32 * (size_t) n, cudaMemcpyHostToDevice, st));
    k_bow_transform<<<(n + 7) / 8, 256, 0, st>>>(d_f, n, v->d_desc, v->d_child_off, v->d_child_idx, v->d_word_id, v->d_weight, v->L, levelsup, d_wid, d_nid, d_w);
    h->launches++;
    ORBFE_CUDA(h, cudaGetLastError());
    ORBFE_CUDA(h, cudaMemcpyAsync(word_id, d_wid, sizeof(int) * (size_t) n, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(node_id, d_nid, sizeof(int) * (size_t) n, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaMemcpyAsync(weight, d_w, sizeof(double) * (size_t) n, cudaMemcpyDeviceToHost, st));
    ORBFE_CUDA(h, cudaStreamSynchronize(st));
    if (fv_node_id && fv_off && fv_idx && fv_n_nodes) {
        // FeatureVector (TemplatedVocabulary.h:1156-1160, FeatureVector::addFeature): marshalling of the per-feature node ids into
        // std::map order — ascending node id, ascending feature index inside a node; stopped words (weight 0) are left out
        std::vector<int> order; order.reserve((size_t) n);
        for (int i = 0; i < n; ++i) if (weight[i] > 0) order.push_back(i);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return node_id[a] < node_id[b]; });
        int nn = 0;
        for (size_t j = 0; j < order.size(); ++j) {
            if (j == 0 || node_id[order[j]] != node_id[order[j - 1]]) { fv_node_id[nn] = node_id[order[j]]; fv_off[nn] = (int) j; ++nn; }
            fv_idx[j] = order[j];
        }
        fv_off[nn] = (int) order.size();
        *fv_n_nodes = nn;
    }
    return ORBFE_OK;
}

}  // extern "C"
