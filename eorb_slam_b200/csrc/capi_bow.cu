// capi_bow.cu — C ABI of the bag-of-words / undistortion row (include/eorb_b200.h, section "bag of words"):
// DBoW2 TemplatedVocabulary::transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1258) as called by
// Frame::ComputeBoW (src/Frame.cc:796-803), and Frame::UndistortKeyPoints (src/Frame.cc:805-840).
// Host code only; the compute steps are the kernels of bow_kernels.cu.  No CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/eorb_b200.h"
#include "bow_kernels.h"

using namespace eorb;

extern "C" int eorb_internal_fail(int code, const char* msg);
static int bFail(int code, const char* what, const char* detail) {
    char buf[400];
    snprintf(buf, sizeof(buf), "%s%s%s", what, detail ? ": " : "", detail ? detail : "");
    return eorb_internal_fail(code, buf);
}
#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) return bFail(EORB_ERR_CUDA, #call, cudaGetErrorString(e__));      \
    } while (0)

struct eorb_vocab {
    int device = 0, k = 0, L = 0, scoring = 0, weighting = 0, nnodes = 0;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    int* d_childStart = nullptr; int* d_children = nullptr; uint8_t* d_desc = nullptr; double* d_weight = nullptr; uint32_t* d_wordId = nullptr;
    // per-call buffers for EORB_BOW_MAX_FEATURES features
    uint8_t* d_feats = nullptr;
    BowOut o{};
    uint8_t* h_pin = nullptr;   // pinned mirror of the outputs (one allocation)
    long long launches = 0;
};

// kernel attributes (dynamic shared-memory limits) are per device: configured once for every device a handle is created on
static std::mutex g_cfgMu;
static bool g_cfgDone[64] = {false};
static cudaError_t configureDevice(int device) {
    std::lock_guard<std::mutex> lk(g_cfgMu);
    if (device >= 0 && device < 64 && g_cfgDone[device]) return cudaSuccess;
    const cudaError_t e = bow_configure();
    if (e == cudaSuccess && device >= 0 && device < 64) g_cfgDone[device] = true;
    return e;
}

extern "C" int eorb_vocab_create(int device, int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent, const uint8_t* is_leaf,
                                 const uint8_t* desc, const double* weight, eorb_vocab** out) {
    if (!out || nnodes < 1 || !parent || !is_leaf || !desc || !weight) return bFail(EORB_ERR_ARG, "eorb_vocab_create", "null argument");
    if (k < 0 || k > 20 || L < 1 || L > 10 || scoring < 0 || scoring > 5 || weighting < 0 || weighting > 3)   // loadFromTextFile's check (:1353)
        return bFail(EORB_ERR_ARG, "eorb_vocab_create", "k / L / scoring / weighting out of DBoW2's range");
    for (int nid = 1; nid < nnodes; nid++)
        if (parent[nid] < 0 || parent[nid] >= nid) return bFail(EORB_ERR_ARG, "eorb_vocab_create", "parent id must precede the node (file order)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return bFail(EORB_ERR_CUDA, "no CUDA device: eorb_b200 has no CPU fallback", nullptr); }
    if (device < 0 || device >= ndev) return bFail(EORB_ERR_ARG, "eorb_vocab_create", "device out of range");
    CU(cudaSetDevice(device));
    { const cudaError_t ec = configureDevice(device); if (ec != cudaSuccess) return bFail(EORB_ERR_CUDA, "bow_configure", cudaGetErrorString(ec)); }
    // children lists in id order, word ids in order of appearance — exactly what loadFromTextFile builds (:1375-1412)
    std::vector<int> cnt(nnodes, 0), childStart(nnodes + 1, 0), children(std::max(nnodes - 1, 1));
    std::vector<uint32_t> wordId(nnodes, 0);
    uint32_t nwords = 0;
    for (int nid = 1; nid < nnodes; nid++) { cnt[parent[nid]]++; if (is_leaf[nid]) wordId[nid] = nwords++; }
    for (int i = 0; i < nnodes; i++) childStart[i + 1] = childStart[i] + cnt[i];
    std::vector<int> fill(childStart.begin(), childStart.end() - 1);
    for (int nid = 1; nid < nnodes; nid++) children[fill[parent[nid]]++] = nid;
    eorb_vocab* v = new eorb_vocab();
    v->device = device; v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting; v->nnodes = nnodes;
    CU(cudaStreamCreateWithFlags(&v->ownStream, cudaStreamNonBlocking));
    v->stream = v->ownStream;
    CU(cudaMalloc((void**)&v->d_childStart, (size_t)(nnodes + 1) * sizeof(int)));
    CU(cudaMalloc((void**)&v->d_children, children.size() * sizeof(int)));
    CU(cudaMalloc((void**)&v->d_desc, (size_t)nnodes * 32));
    CU(cudaMalloc((void**)&v->d_weight, (size_t)nnodes * sizeof(double)));
    CU(cudaMalloc((void**)&v->d_wordId, (size_t)nnodes * sizeof(uint32_t)));
    CU(cudaMemcpy(v->d_childStart, childStart.data(), (size_t)(nnodes + 1) * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(v->d_children, children.data(), children.size() * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(v->d_desc, desc, (size_t)nnodes * 32, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(v->d_weight, weight, (size_t)nnodes * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(v->d_wordId, wordId.data(), (size_t)nnodes * sizeof(uint32_t), cudaMemcpyHostToDevice));
    const size_t M = EORB_BOW_MAX_FEATS;
    CU(cudaMalloc((void**)&v->d_feats, M * 32));
    CU(cudaMalloc((void**)&v->o.wordId, M * 4)); CU(cudaMalloc((void**)&v->o.weight, M * 8)); CU(cudaMalloc((void**)&v->o.nodeId, M * 4));
    CU(cudaMalloc((void**)&v->o.bowIds, M * 4)); CU(cudaMalloc((void**)&v->o.bowVals, M * 8)); CU(cudaMalloc((void**)&v->o.counts, 2 * sizeof(int)));
    CU(cudaMalloc((void**)&v->o.fvNodes, M * 4)); CU(cudaMalloc((void**)&v->o.fvStart, (M + 1) * 4)); CU(cudaMalloc((void**)&v->o.fvFeats, M * 4));
    CU(cudaMallocHost((void**)&v->h_pin, 16 + M * 32 + 16));
    *out = v;
    return EORB_OK;
}

extern "C" int eorb_vocab_destroy(eorb_vocab* v) {
    if (!v) return EORB_OK;
    cudaSetDevice(v->device);
    cudaStreamSynchronize(v->stream);
    cudaFree(v->d_childStart); cudaFree(v->d_children); cudaFree(v->d_desc); cudaFree(v->d_weight); cudaFree(v->d_wordId); cudaFree(v->d_feats);
    cudaFree(v->o.wordId); cudaFree(v->o.weight); cudaFree(v->o.nodeId); cudaFree(v->o.bowIds); cudaFree(v->o.bowVals); cudaFree(v->o.counts);
    cudaFree(v->o.fvNodes); cudaFree(v->o.fvStart); cudaFree(v->o.fvFeats);
    cudaFreeHost(v->h_pin);
    cudaStreamDestroy(v->ownStream);
    delete v;
    return EORB_OK;
}
extern "C" int eorb_vocab_set_stream(eorb_vocab* v, void* s) {
    if (!v) return bFail(EORB_ERR_ARG, "eorb_vocab_set_stream", "null handle");
    CU(cudaStreamSynchronize(v->stream));
    v->stream = (cudaStream_t)s;
    return EORB_OK;
}
extern "C" int eorb_vocab_reset_stream(eorb_vocab* v) {
    if (!v) return bFail(EORB_ERR_ARG, "eorb_vocab_reset_stream", "null handle");
    CU(cudaStreamSynchronize(v->stream));
    v->stream = v->ownStream;
    return EORB_OK;
}
extern "C" long long eorb_vocab_launch_count(const eorb_vocab* v) { return v ? v->launches : 0; }

static int transformRun(eorb_vocab* v, const uint8_t* d_feats, int n, int levelsup, uint32_t* bow_ids, double* bow_vals, int* nbow,
                        uint32_t* fv_nodes, int32_t* fv_start, uint32_t* fv_feats, int* nfv, uint32_t* word_id, uint32_t* node_id) {
    const int norm = v->scoring == 1 ? 2 : (v->scoring == 5 ? 0 : 1);   // ScoringObject::mustNormalize: L2_NORM -> L2, DOT_PRODUCT -> none, else L1
    const int accumulate = v->weighting == 0 || v->weighting == 1;      // TF_IDF, TF -> addWeight; IDF, BINARY -> addIfNotExist
    VocabDev vd{v->d_childStart, v->d_children, v->d_desc, v->d_weight, v->d_wordId, v->nnodes, v->L};
    CU(launch_bow_transform(vd, d_feats, n, levelsup, accumulate, norm, v->o, v->stream, &v->launches));
    // one stream-ordered batch of copies into the pinned mirror (n entries of everything: the counts are not known yet), one sync
    uint8_t* hp = v->h_pin;
    int* hc = reinterpret_cast<int*>(hp);                                  // [0] nbow, [1] nfv
    double* hVals = reinterpret_cast<double*>(hp + 16);
    uint32_t* hIds = reinterpret_cast<uint32_t*>(hVals + n);
    uint32_t* hNodes = hIds + n;
    uint32_t* hFeats = hNodes + n;
    int32_t* hStart = reinterpret_cast<int32_t*>(hFeats + n);
    uint32_t* hWord = reinterpret_cast<uint32_t*>(hStart + n + 1);
    uint32_t* hNode = hWord + n;
    CU(cudaMemcpyAsync(hc, v->o.counts, 2 * sizeof(int), cudaMemcpyDeviceToHost, v->stream));
    CU(cudaMemcpyAsync(hVals, v->o.bowVals, (size_t)n * 8, cudaMemcpyDeviceToHost, v->stream));
    CU(cudaMemcpyAsync(hIds, v->o.bowIds, (size_t)n * 4, cudaMemcpyDeviceToHost, v->stream));
    CU(cudaMemcpyAsync(hNodes, v->o.fvNodes, (size_t)n * 4, cudaMemcpyDeviceToHost, v->stream));
    CU(cudaMemcpyAsync(hFeats, v->o.fvFeats, (size_t)n * 4, cudaMemcpyDeviceToHost, v->stream));
    CU(cudaMemcpyAsync(hStart, v->o.fvStart, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, v->stream));
    if (word_id) CU(cudaMemcpyAsync(hWord, v->o.wordId, (size_t)n * 4, cudaMemcpyDeviceToHost, v->stream));
    if (node_id) CU(cudaMemcpyAsync(hNode, v->o.nodeId, (size_t)n * 4, cudaMemcpyDeviceToHost, v->stream));
    CU(cudaStreamSynchronize(v->stream));
    const int nb = hc[0], nf = hc[1];
    memcpy(bow_ids, hIds, (size_t)nb * 4); memcpy(bow_vals, hVals, (size_t)nb * 8);
    memcpy(fv_nodes, hNodes, (size_t)nf * 4); memcpy(fv_start, hStart, (size_t)(nf + 1) * 4);
    memcpy(fv_feats, hFeats, (size_t)hStart[nf] * 4);
    if (word_id) memcpy(word_id, hWord, (size_t)n * 4);
    if (node_id) memcpy(node_id, hNode, (size_t)n * 4);
    *nbow = nb; *nfv = nf;
    return EORB_OK;
}

static int transformCheck(eorb_vocab* v, const void* feats, int n, const void* a, const void* b, const void* c, const void* d, const void* e,
                          const void* f, const void* g) {
    if (!v) return bFail(EORB_ERR_ARG, "eorb_vocab_transform", "null handle");
    if (n < 0 || (n > 0 && !feats) || !c || !g || !e) return bFail(EORB_ERR_ARG, "eorb_vocab_transform", "null argument");
    if (n > 0 && (!a || !b || !d || !f)) return bFail(EORB_ERR_ARG, "eorb_vocab_transform", "null output");
    if (n > EORB_BOW_MAX_FEATS) return bFail(EORB_ERR_CAPACITY, "eorb_vocab_transform", "more than EORB_BOW_MAX_FEATURES features");
    return EORB_OK;
}

extern "C" int eorb_vocab_transform(eorb_vocab* v, const uint8_t* feats, int n, int levelsup, uint32_t* bow_ids, double* bow_vals, int* nbow,
                                    uint32_t* fv_nodes, int32_t* fv_start, uint32_t* fv_feats, int* nfv, uint32_t* word_id, uint32_t* node_id) {
    int rc = transformCheck(v, feats, n, bow_ids, bow_vals, nbow, fv_nodes, fv_start, fv_feats, nfv);
    if (rc != EORB_OK) return rc;
    *nbow = 0; *nfv = 0; fv_start[0] = 0;
    if (n == 0 || v->nnodes <= 1) return EORB_OK;   // empty() vocabulary or no features: empty vectors (:1134-1137)
    CU(cudaSetDevice(v->device));
    CU(cudaMemcpyAsync(v->d_feats, feats, (size_t)n * 32, cudaMemcpyHostToDevice, v->stream));
    return transformRun(v, v->d_feats, n, levelsup, bow_ids, bow_vals, nbow, fv_nodes, fv_start, fv_feats, nfv, word_id, node_id);
}

extern "C" int eorb_vocab_transform_device(eorb_vocab* v, const uint8_t* d_feats, int n, int levelsup, uint32_t* bow_ids, double* bow_vals,
                                           int* nbow, uint32_t* fv_nodes, int32_t* fv_start, uint32_t* fv_feats, int* nfv, uint32_t* word_id,
                                           uint32_t* node_id) {
    int rc = transformCheck(v, d_feats, n, bow_ids, bow_vals, nbow, fv_nodes, fv_start, fv_feats, nfv);
    if (rc != EORB_OK) return rc;
    *nbow = 0; *nfv = 0; fv_start[0] = 0;
    if (n == 0 || v->nnodes <= 1) return EORB_OK;
    if ((uintptr_t)d_feats & 15) return bFail(EORB_ERR_ARG, "eorb_vocab_transform_device", "descriptors must be 16-byte aligned");
    CU(cudaSetDevice(v->device));
    return transformRun(v, d_feats, n, levelsup, bow_ids, bow_vals, nbow, fv_nodes, fv_start, fv_feats, nfv, word_id, node_id);
}

// The same with the RESULTS left in HBM as well (the tracking chain: extract -> transform -> SearchByBoW without a host round trip of
// the vectors): the BowVector and the CSR FeatureVector are copied device-to-device into the caller's arrays (n entries each,
// d_fv_start n + 1), only the two counts cross PCIe.
extern "C" int eorb_vocab_transform_resident(eorb_vocab* v, const uint8_t* d_feats, int n, int levelsup, uint32_t* d_bow_ids, double* d_bow_vals,
                                             int* nbow, uint32_t* d_fv_nodes, int32_t* d_fv_start, uint32_t* d_fv_feats, int* nfv) {
    if (!v) return bFail(EORB_ERR_ARG, "eorb_vocab_transform_resident", "null handle");
    if (n < 0 || (n > 0 && !d_feats) || !nbow || !nfv || !d_fv_nodes || !d_fv_start || !d_fv_feats)
        return bFail(EORB_ERR_ARG, "eorb_vocab_transform_resident", "null argument");
    if (n > EORB_BOW_MAX_FEATS) return bFail(EORB_ERR_CAPACITY, "eorb_vocab_transform_resident", "more than EORB_BOW_MAX_FEATURES features");
    *nbow = 0; *nfv = 0;
    CU(cudaSetDevice(v->device));
    if (n == 0 || v->nnodes <= 1) { CU(cudaMemsetAsync(d_fv_start, 0, sizeof(int32_t), v->stream)); CU(cudaStreamSynchronize(v->stream)); return EORB_OK; }
    if ((uintptr_t)d_feats & 15) return bFail(EORB_ERR_ARG, "eorb_vocab_transform_resident", "descriptors must be 16-byte aligned");
    const int norm = v->scoring == 1 ? 2 : (v->scoring == 5 ? 0 : 1);
    const int accumulate = v->weighting == 0 || v->weighting == 1;
    VocabDev vd{v->d_childStart, v->d_children, v->d_desc, v->d_weight, v->d_wordId, v->nnodes, v->L};
    CU(launch_bow_transform(vd, d_feats, n, levelsup, accumulate, norm, v->o, v->stream, &v->launches));
    int* hc = reinterpret_cast<int*>(v->h_pin);
    CU(cudaMemcpyAsync(hc, v->o.counts, 2 * sizeof(int), cudaMemcpyDeviceToHost, v->stream));
    CU(cudaMemcpyAsync(d_fv_nodes, v->o.fvNodes, (size_t)n * 4, cudaMemcpyDeviceToDevice, v->stream));
    CU(cudaMemcpyAsync(d_fv_start, v->o.fvStart, (size_t)(n + 1) * 4, cudaMemcpyDeviceToDevice, v->stream));
    CU(cudaMemcpyAsync(d_fv_feats, v->o.fvFeats, (size_t)n * 4, cudaMemcpyDeviceToDevice, v->stream));
    if (d_bow_ids) CU(cudaMemcpyAsync(d_bow_ids, v->o.bowIds, (size_t)n * 4, cudaMemcpyDeviceToDevice, v->stream));
    if (d_bow_vals) CU(cudaMemcpyAsync(d_bow_vals, v->o.bowVals, (size_t)n * 8, cudaMemcpyDeviceToDevice, v->stream));
    CU(cudaStreamSynchronize(v->stream));
    *nbow = hc[0]; *nfv = hc[1];
    return EORB_OK;
}

// ------------------------------------------------------------------------------------------------ undistortion
extern "C" int eorb_undistort_keypoints_device(const eorb_keypoint* d_in, eorb_keypoint* d_out, int n, const float* K4, const float* dist5,
                                               void* cuda_stream) {
    if (n < 0 || (n > 0 && (!d_in || !d_out)) || !K4 || !dist5) return bFail(EORB_ERR_ARG, "eorb_undistort_keypoints_device", "null argument");
    CU(launch_undistort_keypoints(d_in, d_out, n, K4, dist5, (cudaStream_t)cuda_stream));
    return EORB_OK;
}

extern "C" int eorb_undistort_keypoints(const eorb_keypoint* kps, int n, const float* K4, const float* dist5, eorb_keypoint* out) {
    if (n < 0 || (n > 0 && (!kps || !out)) || !K4 || !dist5) return bFail(EORB_ERR_ARG, "eorb_undistort_keypoints", "null argument");
    if (n == 0) return EORB_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return bFail(EORB_ERR_CUDA, "no CUDA device: eorb_b200 has no CPU fallback", nullptr); }
    if (dist5[0] == 0.0f) {   // mDistCoef.at<float>(0) == 0.0 -> mvKeysUn = mvKeys (Frame.cc:807-811)
        if (out != kps) memcpy(out, kps, (size_t)n * sizeof(eorb_keypoint));
        return EORB_OK;
    }
    eorb_keypoint* d = nullptr;
    CU(cudaMalloc((void**)&d, (size_t)n * sizeof(eorb_keypoint)));
    cudaError_t e = cudaMemcpy(d, kps, (size_t)n * sizeof(eorb_keypoint), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_undistort_keypoints(d, d, n, K4, dist5, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(out, d, (size_t)n * sizeof(eorb_keypoint), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return bFail(EORB_ERR_CUDA, "eorb_undistort_keypoints", cudaGetErrorString(e));
    return EORB_OK;
}
