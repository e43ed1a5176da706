// capi_guided.cu — C ABI of the guided-matching row (include/eorb_b200.h, section "guided matching"):
// Frame::AssignFeaturesToGrid / GetFeaturesInArea (src/Frame.cc:431-460, 709-793) and
// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:714-831).  Host code only; the compute steps are the kernels of
// guided_kernels.cu.  No CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/eorb_b200.h"
#include "guided_kernels.h"

using namespace eorb;

extern "C" int eorb_internal_fail(int code, const char* msg);
static int gFail(int code, const char* what, const char* detail) {
    char buf[400];
    snprintf(buf, sizeof(buf), "%s%s%s", what, detail ? ": " : "", detail ? detail : "");
    return eorb_internal_fail(code, buf);
}
#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) return gFail(EORB_ERR_CUDA, #call, cudaGetErrorString(e__));      \
    } while (0)

struct eorb_guided {
    int device = 0;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    // staging for the host entry points (frames 1 and 2) + work buffers, grown on demand
    eorb_keypoint* d_kps[2] = {nullptr, nullptr}; uint8_t* d_desc[2] = {nullptr, nullptr}; int kpCap[2] = {0, 0};
    float* d_prev = nullptr; int32_t* d_m12 = nullptr; int q1Cap = 0;
    // SearchByProjection staging: camera-frame points, validity, observations (frame 1), match table (frame 2)
    float* d_x3 = nullptr; uint8_t* d_valid = nullptr; int32_t* d_obs = nullptr; int pCap1 = 0;
    int32_t* d_mc = nullptr; int pCap2 = 0;
    uint8_t* d_held = nullptr; int heldCap = 0;
    // rectified-stereo / relocalisation extras: per query (n1) the predicted right column and the predicted level, per slot (n2) mvuRight
    float* d_qUr = nullptr; float* d_xr = nullptr; int32_t* d_lvl = nullptr; int exCap1 = 0;
    float* d_ur2 = nullptr; int exCap2 = 0;
    // SearchByBoW host call: every input packed into ONE pinned blob -> one H2D copy; [nmatches | match table] -> one D2H copy
    unsigned char* d_blob = nullptr; unsigned char* h_blob = nullptr; size_t blobCap = 0;
    unsigned char* d_outb = nullptr; unsigned char* h_outb = nullptr; size_t outbCap = 0;
    int* d_bowWork = nullptr;                           // rotation histogram, match count, block ticket
         // SearchByProjection (map points): slots of F held on entry
    GuidedWork w{};
    int workN1 = 0, workN2 = 0;
    int* d_nm = nullptr; int* h_nm = nullptr;          // [nmatches, total candidates] device + pinned mirror
    eorb_area_query* d_q = nullptr; int* d_cnt = nullptr; int* d_out = nullptr; size_t qCap = 0, outCap = 0;
    long long launches = 0;
};

// kernel attributes (dynamic shared-memory limits) are per device: configured once for every device a handle is created on
static std::mutex g_cfgMu;
static bool g_cfgDone[64] = {false};
static cudaError_t configureDevice(int device) {
    std::lock_guard<std::mutex> lk(g_cfgMu);
    if (device >= 0 && device < 64 && g_cfgDone[device]) return cudaSuccess;
    const cudaError_t e = guided_configure();
    if (e == cudaSuccess && device >= 0 && device < 64) g_cfgDone[device] = true;
    return e;
}

static GuidedGrid gridGeom(const float* b) {
    GuidedGrid g;
    g.minX = b[0]; g.minY = b[1];
    g.wInv = (float)EORB_GRID_COLS / (b[2] - b[0]);   // Frame.cc:165-166, float arithmetic
    g.hInv = (float)EORB_GRID_ROWS / (b[3] - b[1]);
    return g;
}

extern "C" int eorb_guided_create(int device, eorb_guided** out) {
    if (!out) return gFail(EORB_ERR_ARG, "eorb_guided_create", "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return gFail(EORB_ERR_CUDA, "no CUDA device: eorb_b200 has no CPU fallback", nullptr); }
    if (device < 0 || device >= ndev) return gFail(EORB_ERR_ARG, "eorb_guided_create", "device out of range");
    CU(cudaSetDevice(device));
    { const cudaError_t ec = configureDevice(device); if (ec != cudaSuccess) return gFail(EORB_ERR_CUDA, "guided_configure", cudaGetErrorString(ec)); }
    eorb_guided* g = new eorb_guided();
    g->device = device;
    CU(cudaStreamCreateWithFlags(&g->ownStream, cudaStreamNonBlocking));
    g->stream = g->ownStream;
    CU(cudaMalloc((void**)&g->w.cellStart, (EORB_GRID_CELLS + 1) * sizeof(int)));
    CU(cudaMalloc((void**)&g->w.assigned, sizeof(int)));
    CU(cudaMalloc((void**)&g->d_nm, 2 * sizeof(int)));
    g->w.total = g->d_nm + 1;
    CU(cudaMallocHost((void**)&g->h_nm, 2 * sizeof(int)));
    *out = g;
    return EORB_OK;
}

extern "C" int eorb_guided_destroy(eorb_guided* g) {
    if (!g) return EORB_OK;
    cudaSetDevice(g->device);
    cudaStreamSynchronize(g->stream);
    for (int k = 0; k < 2; k++) { cudaFree(g->d_kps[k]); cudaFree(g->d_desc[k]); }
    cudaFree(g->d_prev); cudaFree(g->d_m12); cudaFree(g->d_x3); cudaFree(g->d_valid); cudaFree(g->d_obs); cudaFree(g->d_mc); cudaFree(g->d_held); cudaFree(g->d_qUr); cudaFree(g->d_xr); cudaFree(g->d_lvl); cudaFree(g->d_ur2); cudaFree(g->d_blob); cudaFreeHost(g->h_blob); cudaFree(g->d_outb); cudaFreeHost(g->h_outb); cudaFree(g->d_bowWork); cudaFree(g->w.q);
    cudaFree(g->w.cellStart); cudaFree(g->w.cellIdx); cudaFree(g->w.assigned); cudaFree(g->w.candOff); cudaFree(g->w.candCnt);
    cudaFree(g->w.top); cudaFree(g->w.bin); cudaFree(g->w.cand);
    cudaFree(g->d_nm); cudaFreeHost(g->h_nm); cudaFree(g->d_q); cudaFree(g->d_cnt); cudaFree(g->d_out);
    cudaStreamDestroy(g->ownStream);
    delete g;
    return EORB_OK;
}

extern "C" int eorb_guided_set_stream(eorb_guided* g, void* s) {
    if (!g) return gFail(EORB_ERR_ARG, "eorb_guided_set_stream", "null handle");
    CU(cudaStreamSynchronize(g->stream));
    g->stream = (cudaStream_t)s;
    return EORB_OK;
}
extern "C" int eorb_guided_reset_stream(eorb_guided* g) {
    if (!g) return gFail(EORB_ERR_ARG, "eorb_guided_reset_stream", "null handle");
    CU(cudaStreamSynchronize(g->stream));
    g->stream = g->ownStream;
    return EORB_OK;
}
extern "C" long long eorb_guided_launch_count(const eorb_guided* g) { return g ? g->launches : 0; }

static int stageFrame(eorb_guided* g, int k, const eorb_keypoint* kps, const uint8_t* desc, int n) {
    if (n > g->kpCap[k]) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_kps[k]); cudaFree(g->d_desc[k]);
        g->d_kps[k] = nullptr; g->d_desc[k] = nullptr; g->kpCap[k] = 0;
        const int cap = std::max(1024, n);
        CU(cudaMalloc((void**)&g->d_kps[k], (size_t)cap * sizeof(eorb_keypoint)));
        CU(cudaMalloc((void**)&g->d_desc[k], (size_t)cap * 32));
        g->kpCap[k] = cap;
    }
    if (n > 0) {
        CU(cudaMemcpyAsync(g->d_kps[k], kps, (size_t)n * sizeof(eorb_keypoint), cudaMemcpyHostToDevice, g->stream));
        if (desc) CU(cudaMemcpyAsync(g->d_desc[k], desc, (size_t)n * 32, cudaMemcpyHostToDevice, g->stream));
    }
    return EORB_OK;
}

static int reserveWork(eorb_guided* g, int n1, int n2) {
    if (n2 > g->workN2) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->w.cellIdx); g->w.cellIdx = nullptr; g->workN2 = 0;
        const int cap = std::max(1024, n2);
        CU(cudaMalloc((void**)&g->w.cellIdx, (size_t)cap * sizeof(int)));
        g->workN2 = cap;
    }
    if (n1 > g->workN1) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->w.candOff); cudaFree(g->w.candCnt); cudaFree(g->w.top); cudaFree(g->w.bin); cudaFree(g->w.q);
        g->w.candOff = g->w.candCnt = nullptr; g->w.top = nullptr; g->w.bin = nullptr; g->w.q = nullptr; g->workN1 = 0;
        const int cap = std::max(1024, n1);
        CU(cudaMalloc((void**)&g->w.candOff, (size_t)cap * sizeof(int)));
        CU(cudaMalloc((void**)&g->w.candCnt, (size_t)cap * sizeof(int)));
        CU(cudaMalloc((void**)&g->w.top, (size_t)cap * EORB_GUIDED_TOP * sizeof(unsigned long long)));
        CU(cudaMalloc((void**)&g->w.bin, (size_t)cap));
        CU(cudaMalloc((void**)&g->w.q, (size_t)cap * sizeof(eorb_area_query)));
        g->workN1 = cap;
    }
    if (!g->w.cand) {
        g->w.candCap = 1 << 20;   // 4 MB; grown to the exact need when a search overflows it
        CU(cudaMalloc((void**)&g->w.cand, (size_t)g->w.candCap * sizeof(uint32_t)));
    }
    return EORB_OK;
}

static int checkFrames(const void* k1, int n1, const void* k2, int n2, const float* b) {
    if (n1 < 0 || n2 < 0 || (n1 > 0 && !k1) || (n2 > 0 && !k2) || !b) return gFail(EORB_ERR_ARG, "guided matching", "null argument");
    if (n1 > EORB_GUIDED_MAX_KEYPOINTS || n2 > EORB_GUIDED_MAX_KEYPOINTS) return gFail(EORB_ERR_CAPACITY, "guided matching", "more than EORB_GUIDED_MAX_KEYPOINTS keypoints");
    if (!(b[2] > b[0]) || !(b[3] > b[1])) return gFail(EORB_ERR_ARG, "guided matching", "empty image bounds");
    return EORB_OK;
}

extern "C" int eorb_guided_frame_grid(eorb_guided* g, const eorb_keypoint* kps, int n, const float* bounds4, int* cell_start, int* cell_idx,
                                      int* assigned) {
    if (!g || !cell_start || (n > 0 && !cell_idx)) return gFail(EORB_ERR_ARG, "eorb_guided_frame_grid", "null argument");
    int rc = checkFrames(nullptr, 0, kps, n, bounds4);
    if (rc != EORB_OK) return rc;
    CU(cudaSetDevice(g->device));
    if ((rc = stageFrame(g, 1, kps, nullptr, n)) != EORB_OK) return rc;
    if ((rc = reserveWork(g, 0, n)) != EORB_OK) return rc;
    CU(launch_frame_grid(g->d_kps[1], n, gridGeom(bounds4), g->w.cellStart, g->w.cellIdx, g->w.assigned, g->stream));
    g->launches++;
    CU(cudaMemcpyAsync(cell_start, g->w.cellStart, (EORB_GRID_CELLS + 1) * sizeof(int), cudaMemcpyDeviceToHost, g->stream));
    CU(cudaMemcpyAsync(g->h_nm, g->w.assigned, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    const int na = g->h_nm[0];
    if (na > 0) CU(cudaMemcpy(cell_idx, g->w.cellIdx, (size_t)na * sizeof(int), cudaMemcpyDeviceToHost));
    if (assigned) *assigned = na;
    return EORB_OK;
}

extern "C" int eorb_guided_features_in_area(eorb_guided* g, const eorb_keypoint* kps, int n, const float* bounds4, const eorb_area_query* queries,
                                            int nq, int* counts, int* idx_out, int cap_per_query) {
    if (!g || nq < 0 || (nq > 0 && (!queries || !counts)) || cap_per_query < 0 || (cap_per_query > 0 && nq > 0 && !idx_out))
        return gFail(EORB_ERR_ARG, "eorb_guided_features_in_area", "bad argument");
    int rc = checkFrames(nullptr, 0, kps, n, bounds4);
    if (rc != EORB_OK) return rc;
    if (nq == 0) return EORB_OK;
    CU(cudaSetDevice(g->device));
    if ((rc = stageFrame(g, 1, kps, nullptr, n)) != EORB_OK) return rc;
    if ((rc = reserveWork(g, 0, n)) != EORB_OK) return rc;
    const size_t outNeed = std::max<size_t>((size_t)nq * cap_per_query, 1);
    if ((size_t)nq > g->qCap || outNeed > g->outCap) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_q); cudaFree(g->d_cnt); cudaFree(g->d_out);
        g->d_q = nullptr; g->d_cnt = nullptr; g->d_out = nullptr;
        g->qCap = std::max<size_t>(g->qCap, (size_t)nq); g->outCap = std::max(g->outCap, outNeed);
        CU(cudaMalloc((void**)&g->d_q, g->qCap * sizeof(eorb_area_query)));
        CU(cudaMalloc((void**)&g->d_cnt, g->qCap * sizeof(int)));
        CU(cudaMalloc((void**)&g->d_out, g->outCap * sizeof(int)));
    }
    const GuidedGrid gg = gridGeom(bounds4);
    CU(cudaMemcpyAsync(g->d_q, queries, (size_t)nq * sizeof(eorb_area_query), cudaMemcpyHostToDevice, g->stream));
    CU(launch_frame_grid(g->d_kps[1], n, gg, g->w.cellStart, g->w.cellIdx, g->w.assigned, g->stream));
    CU(launch_features_in_area(g->d_kps[1], gg, g->w.cellStart, g->w.cellIdx, g->d_q, nq, g->d_cnt, g->d_out, cap_per_query, g->stream));
    g->launches += 2;
    CU(cudaMemcpyAsync(counts, g->d_cnt, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, g->stream));
    if (cap_per_query > 0) CU(cudaMemcpyAsync(idx_out, g->d_out, (size_t)nq * cap_per_query * sizeof(int), cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    return EORB_OK;
}

// runs the three kernels; when the candidate buffer was too small it is grown to the exact need and the search repeated
static int searchInitRun(eorb_guided* g, const eorb_keypoint* d_k1, const uint8_t* d_d1, int n1, const eorb_keypoint* d_k2, const uint8_t* d_d2,
                         int n2, const float* bounds4, float* d_prev, int window, float nnratio, int checkOri, int32_t* d_m12, int* nmatches) {
    int rc = reserveWork(g, n1, n2);
    if (rc != EORB_OK) return rc;
    GuidedFrame f1{d_k1, d_d1, n1}, f2{d_k2, d_d2, n2};
    const GuidedGrid gg = gridGeom(bounds4);
    for (int attempt = 0; attempt < 2; attempt++) {
        CU(launch_search_init(f1, f2, gg, d_prev, window, nnratio, checkOri, g->w, d_m12, g->d_nm, g->stream, &g->launches));
        CU(cudaMemcpyAsync(g->h_nm, g->d_nm, 2 * sizeof(int), cudaMemcpyDeviceToHost, g->stream));
        CU(cudaStreamSynchronize(g->stream));
        if (g->h_nm[1] <= g->w.candCap) { if (nmatches) *nmatches = g->h_nm[0]; return EORB_OK; }
        cudaFree(g->w.cand); g->w.cand = nullptr;
        g->w.candCap = g->h_nm[1] + g->h_nm[1] / 4;
        CU(cudaMalloc((void**)&g->w.cand, (size_t)g->w.candCap * sizeof(uint32_t)));
    }
    return gFail(EORB_ERR_STATE, "eorb_guided_search_for_initialization", "candidate buffer overflow after growth");
}

extern "C" int eorb_guided_search_for_initialization_device(eorb_guided* g, const eorb_keypoint* d_kps1, const uint8_t* d_desc1, int n1,
                                                            const eorb_keypoint* d_kps2, const uint8_t* d_desc2, int n2, const float* bounds4,
                                                            float* d_prev_xy, int window_size, float nnratio, int check_ori,
                                                            int32_t* d_matches12, int* nmatches) {
    if (!g) return gFail(EORB_ERR_ARG, "eorb_guided_search_for_initialization_device", "null handle");
    int rc = checkFrames(d_kps1, n1, d_kps2, n2, bounds4);
    if (rc != EORB_OK) return rc;
    if (nmatches) *nmatches = 0;
    if (n1 == 0) return EORB_OK;
    if (!d_prev_xy || !d_matches12 || !d_desc1 || (n2 > 0 && !d_desc2)) return gFail(EORB_ERR_ARG, "eorb_guided_search_for_initialization_device", "null argument");
    if (((uintptr_t)d_desc1 | (uintptr_t)d_desc2) & 15) return gFail(EORB_ERR_ARG, "eorb_guided_search_for_initialization_device", "descriptors must be 16-byte aligned");
    CU(cudaSetDevice(g->device));
    return searchInitRun(g, d_kps1, d_desc1, n1, d_kps2, d_desc2, n2, bounds4, d_prev_xy, window_size, nnratio, check_ori, d_matches12, nmatches);
}

extern "C" int eorb_guided_search_for_initialization(eorb_guided* g, const eorb_keypoint* kps1, const uint8_t* desc1, int n1,
                                                     const eorb_keypoint* kps2, const uint8_t* desc2, int n2, const float* bounds4, float* prev_xy,
                                                     int window_size, float nnratio, int check_ori, int32_t* matches12, int* nmatches) {
    if (!g) return gFail(EORB_ERR_ARG, "eorb_guided_search_for_initialization", "null handle");
    int rc = checkFrames(kps1, n1, kps2, n2, bounds4);
    if (rc != EORB_OK) return rc;
    if (nmatches) *nmatches = 0;
    if (n1 == 0) return EORB_OK;   // vnMatches12 is empty, nothing to do (ORBmatcher.cc:718)
    if (!prev_xy || !matches12 || !desc1 || (n2 > 0 && !desc2)) return gFail(EORB_ERR_ARG, "eorb_guided_search_for_initialization", "null argument");
    CU(cudaSetDevice(g->device));
    if ((rc = stageFrame(g, 0, kps1, desc1, n1)) != EORB_OK) return rc;
    if ((rc = stageFrame(g, 1, kps2, desc2, n2)) != EORB_OK) return rc;
    if (n1 > g->q1Cap) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_prev); cudaFree(g->d_m12);
        g->d_prev = nullptr; g->d_m12 = nullptr; g->q1Cap = 0;
        const int cap = std::max(1024, n1);
        CU(cudaMalloc((void**)&g->d_prev, (size_t)cap * 2 * sizeof(float)));
        CU(cudaMalloc((void**)&g->d_m12, (size_t)cap * sizeof(int32_t)));
        g->q1Cap = cap;
    }
    CU(cudaMemcpyAsync(g->d_prev, prev_xy, (size_t)n1 * 2 * sizeof(float), cudaMemcpyHostToDevice, g->stream));
    rc = searchInitRun(g, g->d_kps[0], g->d_desc[0], n1, g->d_kps[1], g->d_desc[1], n2, bounds4, g->d_prev, window_size, nnratio, check_ori,
                       g->d_m12, nmatches);
    if (rc != EORB_OK) return rc;
    CU(cudaMemcpyAsync(matches12, g->d_m12, (size_t)n1 * sizeof(int32_t), cudaMemcpyDeviceToHost, g->stream));
    CU(cudaMemcpyAsync(prev_xy, g->d_prev, (size_t)n1 * 2 * sizeof(float), cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    return EORB_OK;
}

// ------------------------------------------------------------------------------------------------ SearchByProjection
static int projConst(const float* bounds4, const float* K4, const float* scale_factors, int nlevels, float th, GuidedProj& pr) {
    if (!K4 || !scale_factors || nlevels < 1 || nlevels > 32) return gFail(EORB_ERR_ARG, "eorb_guided_search_by_projection", "bad intrinsics / scale table (1..32 levels)");
    pr.fx = K4[0]; pr.fy = K4[1]; pr.cx = K4[2]; pr.cy = K4[3];
    pr.minX = bounds4[0]; pr.minY = bounds4[1]; pr.maxX = bounds4[2]; pr.maxY = bounds4[3];
    pr.th = th; pr.nlevels = nlevels;
    for (int i = 0; i < 32; i++) pr.scale[i] = i < nlevels ? scale_factors[i] : 1.0f;
    return EORB_OK;
}

// per-query / per-slot scratch of the stereo and relocalisation variants
static int reserveExtras(eorb_guided* g, int n1, int n2) {
    if (n1 > g->exCap1) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_qUr); cudaFree(g->d_xr); cudaFree(g->d_lvl);
        g->d_qUr = nullptr; g->d_xr = nullptr; g->d_lvl = nullptr; g->exCap1 = 0;
        const int cap = std::max(1024, n1);
        CU(cudaMalloc((void**)&g->d_qUr, (size_t)cap * sizeof(float)));
        CU(cudaMalloc((void**)&g->d_xr, (size_t)cap * sizeof(float)));
        CU(cudaMalloc((void**)&g->d_lvl, (size_t)cap * sizeof(int32_t)));
        g->exCap1 = cap;
    }
    if (n2 > g->exCap2) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_ur2); g->d_ur2 = nullptr; g->exCap2 = 0;
        const int cap = std::max(1024, n2);
        CU(cudaMalloc((void**)&g->d_ur2, (size_t)cap * sizeof(float)));
        g->exCap2 = cap;
    }
    return EORB_OK;
}
static int reserveHeld(eorb_guided* g, int n2) {
    if (n2 > g->heldCap) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_held); g->d_held = nullptr; g->heldCap = 0;
        const int cap = std::max(1024, n2);
        CU(cudaMalloc((void**)&g->d_held, (size_t)cap));
        g->heldCap = cap;
    }
    return EORB_OK;
}

// what the variants add to the monocular frame-to-frame search (all pointers are device pointers, any may be null)
struct ProjExtras { int levelMode = 0; int reloc = 0; float mbf = 0.f; const int32_t* d_level1 = nullptr; const float* d_uRight2 = nullptr;
                    const uint8_t* d_held2 = nullptr; int thHigh = 100; };

static int searchProjRun(eorb_guided* g, const float* d_x3, const uint8_t* d_valid, const int32_t* d_obs, const eorb_keypoint* d_k1,
                         const uint8_t* d_dmp, int n1, const eorb_keypoint* d_k2, const uint8_t* d_d2, int n2, const float* bounds4,
                         const GuidedProj& pr, int checkOri, int32_t* d_mc, int* nmatches, const ProjExtras& ex = ProjExtras()) {
    int rc = reserveWork(g, n1, n2);
    if (rc != EORB_OK) return rc;
    GuidedProjMode md{ex.levelMode, ex.reloc, ex.mbf, ex.d_level1, nullptr};
    if (ex.d_uRight2) {
        if ((rc = reserveExtras(g, n1, 0)) != EORB_OK) return rc;
        md.qUr = g->d_qUr;
    }
    if (n1 > g->q1Cap) {   // claim[n1] lives in the matches12 staging buffer
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_prev); cudaFree(g->d_m12);
        g->d_prev = nullptr; g->d_m12 = nullptr; g->q1Cap = 0;
        const int cap = std::max(1024, n1);
        CU(cudaMalloc((void**)&g->d_prev, (size_t)cap * 2 * sizeof(float)));
        CU(cudaMalloc((void**)&g->d_m12, (size_t)cap * sizeof(int32_t)));
        g->q1Cap = cap;
    }
    GuidedFrame f2{d_k2, d_d2, n2};
    const GuidedGrid gg = gridGeom(bounds4);
    for (int attempt = 0; attempt < 2; attempt++) {
        CU(launch_search_proj(d_x3, d_valid, d_obs, d_k1, d_dmp, n1, f2, gg, pr, checkOri, g->w, g->d_m12, d_mc, g->d_nm, g->stream, &g->launches, md,
                              ex.d_uRight2, ex.d_held2, ex.thHigh));
        CU(cudaMemcpyAsync(g->h_nm, g->d_nm, 2 * sizeof(int), cudaMemcpyDeviceToHost, g->stream));
        CU(cudaStreamSynchronize(g->stream));
        if (g->h_nm[1] <= g->w.candCap) { if (nmatches) *nmatches = g->h_nm[0]; return EORB_OK; }
        cudaFree(g->w.cand); g->w.cand = nullptr;
        g->w.candCap = g->h_nm[1] + g->h_nm[1] / 4;
        CU(cudaMalloc((void**)&g->w.cand, (size_t)g->w.candCap * sizeof(uint32_t)));
    }
    return gFail(EORB_ERR_STATE, "eorb_guided_search_by_projection", "candidate buffer overflow after growth");
}

extern "C" int eorb_guided_search_by_projection_device(eorb_guided* g, const float* d_x3Dc, const uint8_t* d_valid1, const int32_t* d_obs1,
                                                       const eorb_keypoint* d_kps1, const uint8_t* d_descMP, int n1, const eorb_keypoint* d_kps2,
                                                       const uint8_t* d_desc2, int n2, const float* bounds4, const float* K4,
                                                       const float* scale_factors, int nlevels, float th, int check_ori,
                                                       int32_t* d_match_cur, int* nmatches) {
    if (!g) return gFail(EORB_ERR_ARG, "eorb_guided_search_by_projection_device", "null handle");
    int rc = checkFrames(d_kps1, n1, d_kps2, n2, bounds4);
    if (rc != EORB_OK) return rc;
    if (nmatches) *nmatches = 0;
    if (n2 == 0) return EORB_OK;
    if (!d_match_cur || !d_desc2 || (n1 > 0 && (!d_x3Dc || !d_valid1 || !d_obs1 || !d_descMP)))
        return gFail(EORB_ERR_ARG, "eorb_guided_search_by_projection_device", "null argument");
    if (((uintptr_t)d_descMP | (uintptr_t)d_desc2) & 15) return gFail(EORB_ERR_ARG, "eorb_guided_search_by_projection_device", "descriptors must be 16-byte aligned");
    GuidedProj pr;
    if ((rc = projConst(bounds4, K4, scale_factors, nlevels, th, pr)) != EORB_OK) return rc;
    CU(cudaSetDevice(g->device));
    return searchProjRun(g, d_x3Dc, d_valid1, d_obs1, d_kps1, d_descMP, n1, d_kps2, d_desc2, n2, bounds4, pr, check_ori, d_match_cur, nmatches);
}

// host-pointer runner shared by the three frame-to-frame entry points: obs1 == null and level1 != null for the relocalisation search
static int searchProjHost(eorb_guided* g, const char* what, const float* x3Dc, const uint8_t* valid1, const int32_t* obs1, const int32_t* level1,
                          const eorb_keypoint* kps1, const uint8_t* descMP, int n1, const eorb_keypoint* kps2, const uint8_t* desc2,
                          const uint8_t* held2, const float* u_right2, int n2, const float* bounds4, const float* K4, const float* scale_factors,
                          int nlevels, float th, int check_ori, int level_mode, float mbf, int reloc, int th_high, int32_t* match_cur,
                          int* nmatches) {
    if (!g) return gFail(EORB_ERR_ARG, what, "null handle");
    int rc = checkFrames(kps1, n1, kps2, n2, bounds4);
    if (rc != EORB_OK) return rc;
    if (nmatches) *nmatches = 0;
    if (n2 == 0) return EORB_OK;
    if (!match_cur || !desc2 || (n1 > 0 && (!x3Dc || !valid1 || (!obs1 && !reloc) || (reloc && !level1) || !descMP))) return gFail(EORB_ERR_ARG, what, "null argument");
    if (level_mode < 0 || level_mode > 2) return gFail(EORB_ERR_ARG, what, "level_mode must be 0, 1 (forward) or 2 (backward)");
    GuidedProj pr;
    if ((rc = projConst(bounds4, K4, scale_factors, nlevels, th, pr)) != EORB_OK) return rc;
    CU(cudaSetDevice(g->device));
    if ((rc = stageFrame(g, 0, kps1, descMP, n1)) != EORB_OK) return rc;
    if ((rc = stageFrame(g, 1, kps2, desc2, n2)) != EORB_OK) return rc;
    if (n1 > g->pCap1) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_x3); cudaFree(g->d_valid); cudaFree(g->d_obs);
        g->d_x3 = nullptr; g->d_valid = nullptr; g->d_obs = nullptr; g->pCap1 = 0;
        const int cap = std::max(1024, n1);
        CU(cudaMalloc((void**)&g->d_x3, (size_t)cap * 3 * sizeof(float)));
        CU(cudaMalloc((void**)&g->d_valid, (size_t)cap));
        CU(cudaMalloc((void**)&g->d_obs, (size_t)cap * sizeof(int32_t)));
        g->pCap1 = cap;
    }
    if (n2 > g->pCap2) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_mc); g->d_mc = nullptr; g->pCap2 = 0;
        const int cap = std::max(1024, n2);
        CU(cudaMalloc((void**)&g->d_mc, (size_t)cap * sizeof(int32_t)));
        g->pCap2 = cap;
    }
    ProjExtras ex;
    ex.levelMode = level_mode; ex.reloc = reloc; ex.mbf = mbf; ex.thHigh = th_high;
    if (level1 || u_right2) {
        if ((rc = reserveExtras(g, n1, n2)) != EORB_OK) return rc;
    }
    if (held2 && (rc = reserveHeld(g, n2)) != EORB_OK) return rc;
    if (n1 > 0) {
        CU(cudaMemcpyAsync(g->d_x3, x3Dc, (size_t)n1 * 3 * sizeof(float), cudaMemcpyHostToDevice, g->stream));
        CU(cudaMemcpyAsync(g->d_valid, valid1, (size_t)n1, cudaMemcpyHostToDevice, g->stream));
        if (obs1) CU(cudaMemcpyAsync(g->d_obs, obs1, (size_t)n1 * sizeof(int32_t), cudaMemcpyHostToDevice, g->stream));
        if (level1) { CU(cudaMemcpyAsync(g->d_lvl, level1, (size_t)n1 * sizeof(int32_t), cudaMemcpyHostToDevice, g->stream)); ex.d_level1 = g->d_lvl; }
    }
    if (u_right2) { CU(cudaMemcpyAsync(g->d_ur2, u_right2, (size_t)n2 * sizeof(float), cudaMemcpyHostToDevice, g->stream)); ex.d_uRight2 = g->d_ur2; }
    if (held2) { CU(cudaMemcpyAsync(g->d_held, held2, (size_t)n2, cudaMemcpyHostToDevice, g->stream)); ex.d_held2 = g->d_held; }
    rc = searchProjRun(g, g->d_x3, g->d_valid, obs1 ? g->d_obs : nullptr, g->d_kps[0], g->d_desc[0], n1, g->d_kps[1], g->d_desc[1], n2, bounds4, pr,
                       check_ori, g->d_mc, nmatches, ex);
    if (rc != EORB_OK) return rc;
    CU(cudaMemcpyAsync(match_cur, g->d_mc, (size_t)n2 * sizeof(int32_t), cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    return EORB_OK;
}

extern "C" int eorb_guided_search_by_projection(eorb_guided* g, const float* x3Dc, const uint8_t* valid1, const int32_t* obs1,
                                                const eorb_keypoint* kps1, const uint8_t* descMP, int n1, const eorb_keypoint* kps2,
                                                const uint8_t* desc2, int n2, const float* bounds4, const float* K4, const float* scale_factors,
                                                int nlevels, float th, int check_ori, int32_t* match_cur, int* nmatches) {
    return searchProjHost(g, "eorb_guided_search_by_projection", x3Dc, valid1, obs1, nullptr, kps1, descMP, n1, kps2, desc2, nullptr, nullptr, n2,
                          bounds4, K4, scale_factors, nlevels, th, check_ori, 0, 0.f, 0, 100, match_cur, nmatches);
}

extern "C" int eorb_guided_search_by_projection_stereo(eorb_guided* g, const float* x3Dc, const uint8_t* valid1, const int32_t* obs1,
                                                       const eorb_keypoint* kps1, const uint8_t* descMP, int n1, const eorb_keypoint* kps2,
                                                       const uint8_t* desc2, const float* u_right2, int n2, const float* bounds4, const float* K4,
                                                       const float* scale_factors, int nlevels, float th, int check_ori, int level_mode, float mbf,
                                                       int32_t* match_cur, int* nmatches) {
    return searchProjHost(g, "eorb_guided_search_by_projection_stereo", x3Dc, valid1, obs1, nullptr, kps1, descMP, n1, kps2, desc2, nullptr, u_right2,
                          n2, bounds4, K4, scale_factors, nlevels, th, check_ori, level_mode, mbf, 0, 100, match_cur, nmatches);
}

extern "C" int eorb_guided_search_by_projection_reloc(eorb_guided* g, const float* x3Dc, const uint8_t* valid1, const int32_t* level1,
                                                      const eorb_keypoint* kps1, const uint8_t* descMP, int n1, const eorb_keypoint* kps2,
                                                      const uint8_t* desc2, const uint8_t* held2, int n2, const float* bounds4, const float* K4,
                                                      const float* scale_factors, int nlevels, float th, int orb_dist, int check_ori,
                                                      int32_t* match_cur, int* nmatches) {
    return searchProjHost(g, "eorb_guided_search_by_projection_reloc", x3Dc, valid1, nullptr, level1, kps1, descMP, n1, kps2, desc2, held2, nullptr,
                          n2, bounds4, K4, scale_factors, nlevels, th, check_ori, 0, 0.f, 1, orb_dist, match_cur, nmatches);
}

static int searchProjDeviceCommon(eorb_guided* g, const char* what, const float* d_x3Dc, const uint8_t* d_valid1, const int32_t* d_obs1,
                                  const eorb_keypoint* d_kps1, const uint8_t* d_descMP, int n1, const eorb_keypoint* d_kps2, const uint8_t* d_desc2,
                                  int n2, const float* bounds4, const float* K4, const float* scale_factors, int nlevels, float th, int check_ori,
                                  const ProjExtras& ex, int32_t* d_match_cur, int* nmatches) {
    if (!g) return gFail(EORB_ERR_ARG, what, "null handle");
    int rc = checkFrames(d_kps1, n1, d_kps2, n2, bounds4);
    if (rc != EORB_OK) return rc;
    if (nmatches) *nmatches = 0;
    if (n2 == 0) return EORB_OK;
    if (!d_match_cur || !d_desc2 || (n1 > 0 && (!d_x3Dc || !d_valid1 || (!d_obs1 && !ex.reloc) || (ex.reloc && !ex.d_level1) || !d_descMP)))
        return gFail(EORB_ERR_ARG, what, "null argument");
    if (ex.levelMode < 0 || ex.levelMode > 2) return gFail(EORB_ERR_ARG, what, "level_mode must be 0, 1 (forward) or 2 (backward)");
    if (((uintptr_t)d_descMP | (uintptr_t)d_desc2) & 15) return gFail(EORB_ERR_ARG, what, "descriptors must be 16-byte aligned");
    GuidedProj pr;
    if ((rc = projConst(bounds4, K4, scale_factors, nlevels, th, pr)) != EORB_OK) return rc;
    CU(cudaSetDevice(g->device));
    return searchProjRun(g, d_x3Dc, d_valid1, d_obs1, d_kps1, d_descMP, n1, d_kps2, d_desc2, n2, bounds4, pr, check_ori, d_match_cur, nmatches, ex);
}

extern "C" int eorb_guided_search_by_projection_stereo_device(eorb_guided* g, const float* d_x3Dc, const uint8_t* d_valid1, const int32_t* d_obs1,
                                                              const eorb_keypoint* d_kps1, const uint8_t* d_descMP, int n1,
                                                              const eorb_keypoint* d_kps2, const uint8_t* d_desc2, const float* d_u_right2, int n2,
                                                              const float* bounds4, const float* K4, const float* scale_factors, int nlevels,
                                                              float th, int check_ori, int level_mode, float mbf, int32_t* d_match_cur,
                                                              int* nmatches) {
    ProjExtras ex;
    ex.levelMode = level_mode; ex.mbf = mbf; ex.d_uRight2 = d_u_right2;
    return searchProjDeviceCommon(g, "eorb_guided_search_by_projection_stereo_device", d_x3Dc, d_valid1, d_obs1, d_kps1, d_descMP, n1, d_kps2, d_desc2,
                                  n2, bounds4, K4, scale_factors, nlevels, th, check_ori, ex, d_match_cur, nmatches);
}

extern "C" int eorb_guided_search_by_projection_reloc_device(eorb_guided* g, const float* d_x3Dc, const uint8_t* d_valid1, const int32_t* d_level1,
                                                             const eorb_keypoint* d_kps1, const uint8_t* d_descMP, int n1,
                                                             const eorb_keypoint* d_kps2, const uint8_t* d_desc2, const uint8_t* d_held2, int n2,
                                                             const float* bounds4, const float* K4, const float* scale_factors, int nlevels,
                                                             float th, int orb_dist, int check_ori, int32_t* d_match_cur, int* nmatches) {
    ProjExtras ex;
    ex.reloc = 1; ex.d_level1 = d_level1; ex.d_held2 = d_held2; ex.thHigh = orb_dist;
    return searchProjDeviceCommon(g, "eorb_guided_search_by_projection_reloc_device", d_x3Dc, d_valid1, nullptr, d_kps1, d_descMP, n1, d_kps2, d_desc2,
                                  n2, bounds4, K4, scale_factors, nlevels, th, check_ori, ex, d_match_cur, nmatches);
}

// ------------------------------------------------------------------------------------------------ keyframe-side window searches
static int searchWindowsRun(eorb_guided* g, const char* what, const eorb_area_query* d_q, const float* d_ur, const uint8_t* d_dmp, int n1,
                            const eorb_keypoint* d_k2, const uint8_t* d_d2, const uint8_t* d_held, const float* d_ur2, int n2, const float* bounds4,
                            const float* query_min_xy, const GuidedCandExtra& cx, int blocking, int thHigh, int32_t* d_bestIdx, int32_t* d_bestDist, int32_t* d_match2,
                            int* nmatches) {
    int rc = reserveWork(g, n1, n2);
    if (rc != EORB_OK) return rc;
    if (n1 > g->q1Cap) {   // claim[n1] of the blocking form lives in the matches12 staging buffer
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_prev); cudaFree(g->d_m12);
        g->d_prev = nullptr; g->d_m12 = nullptr; g->q1Cap = 0;
        const int cap = std::max(1024, n1);
        CU(cudaMalloc((void**)&g->d_prev, (size_t)cap * 2 * sizeof(float)));
        CU(cudaMalloc((void**)&g->d_m12, (size_t)cap * sizeof(int32_t)));
        g->q1Cap = cap;
    }
    if (blocking && !d_match2) {      // the ordered resolve always writes its slot table
        if (n2 > g->pCap2) {
            CU(cudaStreamSynchronize(g->stream));
            cudaFree(g->d_mc); g->d_mc = nullptr; g->pCap2 = 0;
            const int cap = std::max(1024, n2);
            CU(cudaMalloc((void**)&g->d_mc, (size_t)cap * sizeof(int32_t)));
            g->pCap2 = cap;
        }
        d_match2 = g->d_mc;
    }
    GuidedFrame f2{d_k2, d_d2, n2};
    const GuidedGrid gg = gridGeom(bounds4);
    GuidedGrid gq = gg;                       // KeyFrame::GetFeaturesInArea subtracts the keyframe's INT bounds (include/KeyFrame.h:529)
    if (query_min_xy) { gq.minX = query_min_xy[0]; gq.minY = query_min_xy[1]; }
    for (int attempt = 0; attempt < 2; attempt++) {
        CU(launch_search_windows(d_q, d_ur, d_dmp, n1, f2, d_held, d_ur2, gg, gq, cx, blocking, thHigh, g->w, g->d_m12, d_bestIdx, d_bestDist, d_match2,
                                 g->d_nm, g->stream, &g->launches));
        CU(cudaMemcpyAsync(g->h_nm, g->d_nm, 2 * sizeof(int), cudaMemcpyDeviceToHost, g->stream));
        CU(cudaStreamSynchronize(g->stream));
        if (g->h_nm[1] <= g->w.candCap) { if (nmatches) *nmatches = g->h_nm[0]; return EORB_OK; }
        cudaFree(g->w.cand); g->w.cand = nullptr;
        g->w.candCap = g->h_nm[1] + g->h_nm[1] / 4;
        CU(cudaMalloc((void**)&g->w.cand, (size_t)g->w.candCap * sizeof(uint32_t)));
    }
    return gFail(EORB_ERR_STATE, what, "candidate buffer overflow after growth");
}

static int windowsArgs(eorb_guided* g, const char* what, const void* queries, const void* descMP, int n1, const void* kps2, const void* desc2, int n2,
                       const float* bounds4, const float* inv_level_sigma2, int nlevels, int th_high, const void* best_idx, int* nmatches,
                       GuidedCandExtra& cx) {
    if (!g) return gFail(EORB_ERR_ARG, what, "null handle");
    int rc = checkFrames(queries, n1, kps2, n2, bounds4);
    if (rc != EORB_OK) return rc;
    if (nmatches) *nmatches = 0;
    if ((n1 > 0 && (!descMP || !best_idx)) || (n2 > 0 && !desc2)) return gFail(EORB_ERR_ARG, what, "null argument");
    if (th_high < 0 || th_high > 255) return gFail(EORB_ERR_ARG, what, "th_high must lie in [0, 255]");
    cx = GuidedCandExtra{};
    if (inv_level_sigma2) {
        if (nlevels < 1 || nlevels > 32) return gFail(EORB_ERR_ARG, what, "bad level table (1..32 levels)");
        cx.chi2 = 1;
        for (int i = 0; i < 32; i++) cx.invSigma2[i] = inv_level_sigma2[i < nlevels ? i : nlevels - 1];
    }
    return EORB_OK;
}

extern "C" int eorb_guided_search_windows_device(eorb_guided* g, const eorb_area_query* d_queries, const float* d_ur, const uint8_t* d_descMP, int n1,
                                                 const eorb_keypoint* d_kps2, const uint8_t* d_desc2, const uint8_t* d_held2,
                                                 const float* d_u_right2, int n2, const float* bounds4, const float* query_min_xy, const float* inv_level_sigma2,
                                                 int nlevels, int blocking, int th_high, int32_t* d_best_idx, int32_t* d_best_dist, int32_t* d_match2,
                                                 int* nmatches) {
    const char* what = "eorb_guided_search_windows_device";
    GuidedCandExtra cx;
    int rc = windowsArgs(g, what, d_queries, d_descMP, n1, d_kps2, d_desc2, n2, bounds4, inv_level_sigma2, nlevels, th_high, d_best_idx, nmatches, cx);
    if (rc != EORB_OK) return rc;
    if (n1 == 0) return EORB_OK;
    if (((uintptr_t)d_descMP | (uintptr_t)d_desc2) & 15) return gFail(EORB_ERR_ARG, what, "descriptors must be 16-byte aligned");
    CU(cudaSetDevice(g->device));
    if (n2 == 0) {
        CU(cudaMemsetAsync(d_best_idx, 0xff, (size_t)n1 * sizeof(int32_t), g->stream));
        if (d_best_dist) {
            std::vector<int32_t> none((size_t)n1, 256);
            CU(cudaMemcpyAsync(d_best_dist, none.data(), (size_t)n1 * sizeof(int32_t), cudaMemcpyHostToDevice, g->stream));
        }
        CU(cudaStreamSynchronize(g->stream));
        return EORB_OK;
    }
    return searchWindowsRun(g, what, d_queries, d_ur, d_descMP, n1, d_kps2, d_desc2, d_held2, d_u_right2, n2, bounds4, query_min_xy, cx, blocking != 0,
                            th_high, d_best_idx, d_best_dist, d_match2, nmatches);
}

extern "C" int eorb_guided_search_windows(eorb_guided* g, const eorb_area_query* queries, const float* ur, const uint8_t* descMP, int n1,
                                          const eorb_keypoint* kps2, const uint8_t* desc2, const uint8_t* held2, const float* u_right2, int n2,
                                          const float* bounds4, const float* query_min_xy, const float* inv_level_sigma2, int nlevels, int blocking,
                                          int th_high, int32_t* best_idx, int32_t* best_dist, int32_t* match2, int* nmatches) {
    const char* what = "eorb_guided_search_windows";
    GuidedCandExtra cx;
    int rc = windowsArgs(g, what, queries, descMP, n1, kps2, desc2, n2, bounds4, inv_level_sigma2, nlevels, th_high, best_idx, nmatches, cx);
    if (rc != EORB_OK) return rc;
    for (int i = 0; i < n2 && match2; i++) match2[i] = -1;
    for (int i = 0; i < n1; i++) { best_idx[i] = -1; if (best_dist) best_dist[i] = 256; }
    if (n1 == 0 || n2 == 0) return EORB_OK;
    CU(cudaSetDevice(g->device));
    // staging: the map points' descriptors in frame slot 0 (its keypoint array receives the per-point results), the windows in w.q
    if (n1 > g->kpCap[0]) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_kps[0]); cudaFree(g->d_desc[0]);
        g->d_kps[0] = nullptr; g->d_desc[0] = nullptr; g->kpCap[0] = 0;
        const int cap = std::max(1024, n1);
        CU(cudaMalloc((void**)&g->d_kps[0], (size_t)cap * sizeof(eorb_keypoint)));
        CU(cudaMalloc((void**)&g->d_desc[0], (size_t)cap * 32));
        g->kpCap[0] = cap;
    }
    if ((rc = stageFrame(g, 1, kps2, desc2, n2)) != EORB_OK) return rc;
    if ((rc = reserveWork(g, n1, n2)) != EORB_OK) return rc;
    if ((rc = reserveExtras(g, n1, n2)) != EORB_OK) return rc;
    if (held2 && (rc = reserveHeld(g, n2)) != EORB_OK) return rc;
    if (n2 > g->pCap2) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_mc); g->d_mc = nullptr; g->pCap2 = 0;
        const int cap = std::max(1024, n2);
        CU(cudaMalloc((void**)&g->d_mc, (size_t)cap * sizeof(int32_t)));
        g->pCap2 = cap;
    }
    static_assert(2 * sizeof(int32_t) <= sizeof(eorb_keypoint), "best_idx | best_dist are staged in the frame-1 keypoint buffer");
    CU(cudaMemcpyAsync(g->w.q, queries, (size_t)n1 * sizeof(eorb_area_query), cudaMemcpyHostToDevice, g->stream));
    CU(cudaMemcpyAsync(g->d_desc[0], descMP, (size_t)n1 * 32, cudaMemcpyHostToDevice, g->stream));
    if (ur) CU(cudaMemcpyAsync(g->d_qUr, ur, (size_t)n1 * sizeof(float), cudaMemcpyHostToDevice, g->stream));
    if (u_right2) CU(cudaMemcpyAsync(g->d_ur2, u_right2, (size_t)n2 * sizeof(float), cudaMemcpyHostToDevice, g->stream));
    if (held2) CU(cudaMemcpyAsync(g->d_held, held2, (size_t)n2, cudaMemcpyHostToDevice, g->stream));
    int32_t* d_bi = reinterpret_cast<int32_t*>(g->d_kps[0]);           // best_idx[n1] | best_dist[n1] in the (unused) frame-1 keypoint buffer
    int32_t* d_bd = d_bi + n1;
    rc = searchWindowsRun(g, what, g->w.q, ur ? g->d_qUr : nullptr, g->d_desc[0], n1, g->d_kps[1], g->d_desc[1], held2 ? g->d_held : nullptr,
                          u_right2 ? g->d_ur2 : nullptr, n2, bounds4, query_min_xy, cx, blocking != 0, th_high, d_bi, d_bd, g->d_mc, nmatches);
    if (rc != EORB_OK) return rc;
    CU(cudaMemcpyAsync(best_idx, d_bi, (size_t)n1 * sizeof(int32_t), cudaMemcpyDeviceToHost, g->stream));
    if (best_dist) CU(cudaMemcpyAsync(best_dist, d_bd, (size_t)n1 * sizeof(int32_t), cudaMemcpyDeviceToHost, g->stream));
    if (match2 && blocking) CU(cudaMemcpyAsync(match2, g->d_mc, (size_t)n2 * sizeof(int32_t), cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    return EORB_OK;
}

// ------------------------------------------------------------------------------------------------ SearchByProjection (map points)
static_assert(sizeof(eorb_track_point) == sizeof(eorb_keypoint), "the host entry point stages the track points in the frame-1 keypoint buffer");

static int searchMapRun(eorb_guided* g, const eorb_track_point* d_pts, const uint8_t* d_dmp, int n1, const eorb_keypoint* d_k2, const uint8_t* d_d2,
                        const uint8_t* d_held, int n2, const float* bounds4, const GuidedProj& pr, int farPoints, float thFar, float nnratio,
                        int32_t* d_mc, int* nmatches, const float* d_projXR = nullptr, const float* d_uRight2 = nullptr) {
    int rc = reserveWork(g, n1, n2);
    if (rc != EORB_OK) return rc;
    GuidedFrame f2{d_k2, d_d2, n2};
    const GuidedGrid gg = gridGeom(bounds4);
    for (int attempt = 0; attempt < 2; attempt++) {
        CU(launch_search_map_points(d_pts, d_dmp, n1, f2, d_held, gg, pr, farPoints, thFar, nnratio, g->w, d_mc, g->d_nm, g->stream, &g->launches,
                                    d_projXR, d_uRight2));
        CU(cudaMemcpyAsync(g->h_nm, g->d_nm, 2 * sizeof(int), cudaMemcpyDeviceToHost, g->stream));
        CU(cudaStreamSynchronize(g->stream));
        if (g->h_nm[1] <= g->w.candCap) { if (nmatches) *nmatches = g->h_nm[0]; return EORB_OK; }
        cudaFree(g->w.cand); g->w.cand = nullptr;
        g->w.candCap = g->h_nm[1] + g->h_nm[1] / 4;
        CU(cudaMalloc((void**)&g->w.cand, (size_t)g->w.candCap * sizeof(uint32_t)));
    }
    return gFail(EORB_ERR_STATE, "eorb_guided_search_by_projection_map_points", "candidate buffer overflow after growth");
}

static int mapConst(const float* scale_factors, int nlevels, float th, GuidedProj& pr) {
    if (!scale_factors || nlevels < 1 || nlevels > 32) return gFail(EORB_ERR_ARG, "eorb_guided_search_by_projection_map_points", "bad scale table (1..32 levels)");
    if (!(th > 0.0f)) return gFail(EORB_ERR_ARG, "eorb_guided_search_by_projection_map_points", "th must be positive");
    pr = GuidedProj{};
    pr.th = th; pr.nlevels = nlevels;
    for (int i = 0; i < 32; i++) pr.scale[i] = i < nlevels ? scale_factors[i] : 1.0f;
    return EORB_OK;
}

static int searchMapDeviceCommon(eorb_guided* g, const char* what, const eorb_track_point* d_pts, const float* d_proj_xr, const uint8_t* d_descMP,
                                 int n1, const eorb_keypoint* d_kps2, const uint8_t* d_desc2, const uint8_t* d_held2, const float* d_u_right2, int n2,
                                 const float* bounds4, const float* scale_factors, int nlevels, float th, int far_points, float th_far,
                                 float nnratio, int32_t* d_match_cur, int* nmatches) {
    if (!g) return gFail(EORB_ERR_ARG, what, "null handle");
    int rc = checkFrames(d_pts, n1, d_kps2, n2, bounds4);
    if (rc != EORB_OK) return rc;
    if (nmatches) *nmatches = 0;
    if (n2 == 0) return EORB_OK;
    if (!d_match_cur || !d_desc2 || (n1 > 0 && !d_descMP)) return gFail(EORB_ERR_ARG, what, "null argument");
    if (((uintptr_t)d_descMP | (uintptr_t)d_desc2) & 15) return gFail(EORB_ERR_ARG, what, "descriptors must be 16-byte aligned");
    GuidedProj pr;
    if ((rc = mapConst(scale_factors, nlevels, th, pr)) != EORB_OK) return rc;
    CU(cudaSetDevice(g->device));
    const bool stereo = d_proj_xr && d_u_right2;
    return searchMapRun(g, d_pts, d_descMP, n1, d_kps2, d_desc2, d_held2, n2, bounds4, pr, far_points, th_far, nnratio, d_match_cur, nmatches,
                        stereo ? d_proj_xr : nullptr, stereo ? d_u_right2 : nullptr);
}

extern "C" int eorb_guided_search_by_projection_map_points_device(eorb_guided* g, const eorb_track_point* d_pts, const uint8_t* d_descMP, int n1,
                                                                  const eorb_keypoint* d_kps2, const uint8_t* d_desc2, const uint8_t* d_held2,
                                                                  int n2, const float* bounds4, const float* scale_factors, int nlevels, float th,
                                                                  int far_points, float th_far, float nnratio, int32_t* d_match_cur,
                                                                  int* nmatches) {
    return searchMapDeviceCommon(g, "eorb_guided_search_by_projection_map_points_device", d_pts, nullptr, d_descMP, n1, d_kps2, d_desc2, d_held2, nullptr,
                                 n2, bounds4, scale_factors, nlevels, th, far_points, th_far, nnratio, d_match_cur, nmatches);
}

extern "C" int eorb_guided_search_by_projection_map_points_stereo_device(eorb_guided* g, const eorb_track_point* d_pts, const float* d_proj_xr,
                                                                         const uint8_t* d_descMP, int n1, const eorb_keypoint* d_kps2,
                                                                         const uint8_t* d_desc2, const uint8_t* d_held2, const float* d_u_right2,
                                                                         int n2, const float* bounds4, const float* scale_factors, int nlevels,
                                                                         float th, int far_points, float th_far, float nnratio,
                                                                         int32_t* d_match_cur, int* nmatches) {
    return searchMapDeviceCommon(g, "eorb_guided_search_by_projection_map_points_stereo_device", d_pts, d_proj_xr, d_descMP, n1, d_kps2, d_desc2, d_held2,
                                 d_u_right2, n2, bounds4, scale_factors, nlevels, th, far_points, th_far, nnratio, d_match_cur, nmatches);
}

static int searchMapHost(eorb_guided* g, const char* what, const eorb_track_point* pts, const float* proj_xr, const uint8_t* descMP, int n1,
                         const eorb_keypoint* kps2, const uint8_t* desc2, const uint8_t* held2, const float* u_right2, int n2, const float* bounds4,
                         const float* scale_factors, int nlevels, float th, int far_points, float th_far, float nnratio, int32_t* match_cur,
                         int* nmatches) {
    if (!g) return gFail(EORB_ERR_ARG, what, "null handle");
    int rc = checkFrames(pts, n1, kps2, n2, bounds4);
    if (rc != EORB_OK) return rc;
    if (nmatches) *nmatches = 0;
    if (n2 == 0) return EORB_OK;
    if (!match_cur || !desc2 || (n1 > 0 && !descMP)) return gFail(EORB_ERR_ARG, what, "null argument");
    GuidedProj pr;
    if ((rc = mapConst(scale_factors, nlevels, th, pr)) != EORB_OK) return rc;
    CU(cudaSetDevice(g->device));
    if ((rc = stageFrame(g, 0, reinterpret_cast<const eorb_keypoint*>(pts), descMP, n1)) != EORB_OK) return rc;
    if ((rc = stageFrame(g, 1, kps2, desc2, n2)) != EORB_OK) return rc;
    if (n2 > g->pCap2) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_mc); g->d_mc = nullptr; g->pCap2 = 0;
        const int cap = std::max(1024, n2);
        CU(cudaMalloc((void**)&g->d_mc, (size_t)cap * sizeof(int32_t)));
        g->pCap2 = cap;
    }
    if (held2 && (rc = reserveHeld(g, n2)) != EORB_OK) return rc;
    if (held2) CU(cudaMemcpyAsync(g->d_held, held2, (size_t)n2, cudaMemcpyHostToDevice, g->stream));
    const bool stereo = proj_xr && u_right2 && n1 > 0;     // the rectified-stereo column test needs both sides (ORBmatcher.cc:91-96)
    if (stereo) {
        if ((rc = reserveExtras(g, n1, n2)) != EORB_OK) return rc;
        CU(cudaMemcpyAsync(g->d_xr, proj_xr, (size_t)n1 * sizeof(float), cudaMemcpyHostToDevice, g->stream));
        CU(cudaMemcpyAsync(g->d_ur2, u_right2, (size_t)n2 * sizeof(float), cudaMemcpyHostToDevice, g->stream));
    }
    rc = searchMapRun(g, reinterpret_cast<const eorb_track_point*>(g->d_kps[0]), g->d_desc[0], n1, g->d_kps[1], g->d_desc[1],
                      held2 ? g->d_held : nullptr, n2, bounds4, pr, far_points, th_far, nnratio, g->d_mc, nmatches, stereo ? g->d_xr : nullptr,
                      stereo ? g->d_ur2 : nullptr);
    if (rc != EORB_OK) return rc;
    CU(cudaMemcpyAsync(match_cur, g->d_mc, (size_t)n2 * sizeof(int32_t), cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    return EORB_OK;
}

extern "C" int eorb_guided_search_by_projection_map_points(eorb_guided* g, const eorb_track_point* pts, const uint8_t* descMP, int n1,
                                                           const eorb_keypoint* kps2, const uint8_t* desc2, const uint8_t* held2, int n2,
                                                           const float* bounds4, const float* scale_factors, int nlevels, float th, int far_points,
                                                           float th_far, float nnratio, int32_t* match_cur, int* nmatches) {
    return searchMapHost(g, "eorb_guided_search_by_projection_map_points", pts, nullptr, descMP, n1, kps2, desc2, held2, nullptr, n2, bounds4,
                         scale_factors, nlevels, th, far_points, th_far, nnratio, match_cur, nmatches);
}

extern "C" int eorb_guided_search_by_projection_map_points_stereo(eorb_guided* g, const eorb_track_point* pts, const float* proj_xr,
                                                                  const uint8_t* descMP, int n1, const eorb_keypoint* kps2, const uint8_t* desc2,
                                                                  const uint8_t* held2, const float* u_right2, int n2, const float* bounds4,
                                                                  const float* scale_factors, int nlevels, float th, int far_points, float th_far,
                                                                  float nnratio, int32_t* match_cur, int* nmatches) {
    return searchMapHost(g, "eorb_guided_search_by_projection_map_points_stereo", pts, proj_xr, descMP, n1, kps2, desc2, held2, u_right2, n2, bounds4,
                         scale_factors, nlevels, th, far_points, th_far, nnratio, match_cur, nmatches);
}

// ------------------------------------------------------------------------------------------------ SearchByBoW
static int checkFeatureVector(const char* who, const uint32_t* nodes, const int32_t* start, const uint32_t* feats, int nn, int n) {
    if (nn < 0 || (nn > 0 && (!nodes || !start || !feats))) return gFail(EORB_ERR_ARG, who, "null FeatureVector");
    if (nn == 0) return EORB_OK;
    if (start[0] != 0 || start[nn] < 0 || start[nn] > n) return gFail(EORB_ERR_ARG, who, "FeatureVector offsets out of range");
    for (int q = 0; q < nn; q++) {
        if (start[q + 1] < start[q]) return gFail(EORB_ERR_ARG, who, "FeatureVector offsets not ascending");
        if (q > 0 && nodes[q] <= nodes[q - 1]) return gFail(EORB_ERR_ARG, who, "FeatureVector node ids not ascending");
    }
    for (int p = 0; p < start[nn]; p++)
        if (feats[p] >= (uint32_t)n) return gFail(EORB_ERR_ARG, who, "FeatureVector feature index out of range");
    return EORB_OK;
}

// device-resident SearchByBoW, both forms: d_valid_f / d_match12 given = ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) (ORBmatcher.cc:833-990)
static int guidedBowDevice(eorb_guided* g, const char* who, const eorb_keypoint* d_kps_kf, const uint8_t* d_desc_kf, const uint8_t* d_valid_kf, int n1,
                           const uint32_t* d_kf_nodes, const int32_t* d_kf_start, const uint32_t* d_kf_feats, int nkf, const eorb_keypoint* d_kps_f,
                           const uint8_t* d_desc_f, const uint8_t* d_valid_f, int n2, const uint32_t* d_f_nodes, const int32_t* d_f_start,
                           const uint32_t* d_f_feats, int nf, float nnratio, int check_ori, int32_t* d_match_f, int32_t* d_match12, int* nmatches) {
    if (!g) return gFail(EORB_ERR_ARG, who, "null handle");
    if (n1 < 0 || n2 < 0 || nkf < 0 || nf < 0) return gFail(EORB_ERR_ARG, who, "negative size");
    if (n1 > EORB_GUIDED_MAX_KEYPOINTS || n2 > EORB_GUIDED_MAX_KEYPOINTS) return gFail(EORB_ERR_CAPACITY, who, "more than EORB_GUIDED_MAX_KEYPOINTS keypoints");
    if (nmatches) *nmatches = 0;
    CU(cudaSetDevice(g->device));
    if (d_match12 && n1 > 0) CU(cudaMemsetAsync(d_match12, 0xff, (size_t)n1 * sizeof(int32_t), g->stream));
    if (n2 == 0) return EORB_OK;
    if (!d_match_f || !d_kps_f || !d_desc_f || (n1 > 0 && (!d_kps_kf || !d_desc_kf || !d_valid_kf)) ||
        (nkf > 0 && (!d_kf_nodes || !d_kf_start || !d_kf_feats)) || (nf > 0 && (!d_f_nodes || !d_f_start || !d_f_feats)))
        return gFail(EORB_ERR_ARG, who, "null argument");
    if (((uintptr_t)d_desc_kf | (uintptr_t)d_desc_f) & 15) return gFail(EORB_ERR_ARG, who, "descriptors must be 16-byte aligned");
    GuidedBowSide a{d_kps_kf, d_desc_kf, d_kf_nodes, d_kf_start, d_kf_feats, nkf, n1}, b{d_kps_f, d_desc_f, d_f_nodes, d_f_start, d_f_feats, nf, n2};
    if (!g->d_bowWork) CU(cudaMalloc((void**)&g->d_bowWork, 64 * sizeof(int)));
    CU(launch_search_by_bow(a, d_valid_kf, b, nnratio, check_ori, d_match_f, g->d_bowWork, g->d_nm, g->stream, &g->launches, d_valid_f, d_match12));
    CU(cudaMemcpyAsync(g->h_nm, g->d_nm, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    if (nmatches) *nmatches = g->h_nm[0];
    return EORB_OK;
}

extern "C" int eorb_guided_search_by_bow_device(eorb_guided* g, const eorb_keypoint* d_kps_kf, const uint8_t* d_desc_kf, const uint8_t* d_valid_kf,
                                                int n1, const uint32_t* d_kf_nodes, const int32_t* d_kf_start, const uint32_t* d_kf_feats, int nkf,
                                                const eorb_keypoint* d_kps_f, const uint8_t* d_desc_f, int n2, const uint32_t* d_f_nodes,
                                                const int32_t* d_f_start, const uint32_t* d_f_feats, int nf, float nnratio, int check_ori,
                                                int32_t* d_match_f, int* nmatches) {
    return guidedBowDevice(g, "eorb_guided_search_by_bow_device", d_kps_kf, d_desc_kf, d_valid_kf, n1, d_kf_nodes, d_kf_start, d_kf_feats, nkf, d_kps_f,
                           d_desc_f, nullptr, n2, d_f_nodes, d_f_start, d_f_feats, nf, nnratio, check_ori, d_match_f, nullptr, nmatches);
}

extern "C" int eorb_guided_search_by_bow_kf_device(eorb_guided* g, const eorb_keypoint* d_kps1, const uint8_t* d_desc1, const uint8_t* d_valid1, int n1,
                                                   const uint32_t* d_nodes1, const int32_t* d_start1, const uint32_t* d_feats1, int nn1,
                                                   const eorb_keypoint* d_kps2, const uint8_t* d_desc2, const uint8_t* d_valid2, int n2,
                                                   const uint32_t* d_nodes2, const int32_t* d_start2, const uint32_t* d_feats2, int nn2, float nnratio,
                                                   int check_ori, int32_t* d_match12, int32_t* d_scratch_n2, int* nmatches) {
    const char* who = "eorb_guided_search_by_bow_kf_device";
    if (!d_match12 || (n2 > 0 && (!d_scratch_n2 || !d_valid2))) return gFail(EORB_ERR_ARG, who, "null argument");
    return guidedBowDevice(g, who, d_kps1, d_desc1, d_valid1, n1, d_nodes1, d_start1, d_feats1, nn1, d_kps2, d_desc2, d_valid2, n2, d_nodes2, d_start2,
                           d_feats2, nn2, nnratio, check_ori, d_scratch_n2, d_match12, nmatches);
}

// host-buffer SearchByBoW, both forms (valid_f / match12 given = the keyframe-keyframe form): one H2D blob, one launch, one D2H blob
static int guidedBowHost(eorb_guided* g, const char* who, const eorb_keypoint* kps_kf, const uint8_t* desc_kf, const uint8_t* valid_kf, int n1,
                         const uint32_t* kf_nodes, const int32_t* kf_start, const uint32_t* kf_feats, int nkf, const eorb_keypoint* kps_f,
                         const uint8_t* desc_f, const uint8_t* valid_f, int n2, const uint32_t* f_nodes, const int32_t* f_start, const uint32_t* f_feats,
                         int nf, float nnratio, int check_ori, int32_t* match_f, int32_t* match12, int* nmatches) {
    const bool kfForm = match12 != nullptr;
    if (!g) return gFail(EORB_ERR_ARG, who, "null handle");
    if (n1 < 0 || n2 < 0) return gFail(EORB_ERR_ARG, who, "negative size");
    if (n1 > EORB_GUIDED_MAX_KEYPOINTS || n2 > EORB_GUIDED_MAX_KEYPOINTS) return gFail(EORB_ERR_CAPACITY, who, "more than EORB_GUIDED_MAX_KEYPOINTS keypoints");
    if (nmatches) *nmatches = 0;
    if (kfForm) for (int i = 0; i < n1; i++) match12[i] = -1;
    if (n2 == 0) return EORB_OK;
    if ((!kfForm && !match_f) || !kps_f || !desc_f || (n1 > 0 && (!kps_kf || !desc_kf || !valid_kf)) || (kfForm && !valid_f))
        return gFail(EORB_ERR_ARG, who, "null argument");
    int rc;
    if ((rc = checkFeatureVector(who, kf_nodes, kf_start, kf_feats, nkf, n1)) != EORB_OK) return rc;
    if ((rc = checkFeatureVector(who, f_nodes, f_start, f_feats, nf, n2)) != EORB_OK) return rc;
    CU(cudaSetDevice(g->device));
    // one blob, every part 16-byte aligned: [kf kps | kf desc | f kps | f desc | kf nodes | kf start | kf feats | f nodes | f start | f feats | valid kf | valid f]
    const int nfk = nkf > 0 ? kf_start[nkf] : 0, nff = nf > 0 ? f_start[nf] : 0;
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const int NP = 12;
    const int32_t zero2[2] = {0, 0};
    const size_t sz[NP] = {(size_t)n1 * sizeof(eorb_keypoint), (size_t)n1 * 32, (size_t)n2 * sizeof(eorb_keypoint), (size_t)n2 * 32, (size_t)nkf * 4,
                           (size_t)(nkf + 1) * 4, (size_t)nfk * 4, (size_t)nf * 4, (size_t)(nf + 1) * 4, (size_t)nff * 4, (size_t)n1,
                           kfForm ? (size_t)n2 : 0};
    const void* src[NP] = {kps_kf, desc_kf, kps_f, desc_f, kf_nodes, nkf > 0 ? (const void*)kf_start : (const void*)zero2, kf_feats,
                           f_nodes, nf > 0 ? (const void*)f_start : (const void*)zero2, f_feats, valid_kf, valid_f};
    size_t off[NP + 1]; off[0] = 0;
    for (int i = 0; i < NP; i++) off[i + 1] = off[i] + al(sz[i]);
    if (off[NP] > g->blobCap) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_blob); cudaFreeHost(g->h_blob); g->d_blob = nullptr; g->h_blob = nullptr; g->blobCap = 0;
        const size_t cap = std::max<size_t>(off[NP] + off[NP] / 4, 1 << 18);
        CU(cudaMalloc((void**)&g->d_blob, cap));
        CU(cudaMallocHost((void**)&g->h_blob, cap));
        g->blobCap = cap;
    }
    // output blob: [nmatches (16 bytes) | match_f[n2] | match12[n1] (keyframe-keyframe form)]
    const size_t o12 = 16 + al((size_t)n2 * sizeof(int32_t));
    const size_t outBytes = o12 + (kfForm ? (size_t)n1 * sizeof(int32_t) : 0);
    if (outBytes > g->outbCap) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_outb); cudaFreeHost(g->h_outb); g->d_outb = nullptr; g->h_outb = nullptr; g->outbCap = 0;
        const size_t cap = std::max<size_t>(outBytes + outBytes / 4, 1 << 14);
        CU(cudaMalloc((void**)&g->d_outb, cap));
        CU(cudaMallocHost((void**)&g->h_outb, cap));
        g->outbCap = cap;
    }
    for (int i = 0; i < NP; i++)
        if (sz[i] > 0) std::memcpy(g->h_blob + off[i], src[i], sz[i]);
    CU(cudaMemcpyAsync(g->d_blob, g->h_blob, off[NP], cudaMemcpyHostToDevice, g->stream));
    unsigned char* B = g->d_blob;
    GuidedBowSide a{(const eorb_keypoint*)(B + off[0]), B + off[1], (const uint32_t*)(B + off[4]), (const int32_t*)(B + off[5]), (const uint32_t*)(B + off[6]), nkf, n1};
    GuidedBowSide b{(const eorb_keypoint*)(B + off[2]), B + off[3], (const uint32_t*)(B + off[7]), (const int32_t*)(B + off[8]), (const uint32_t*)(B + off[9]), nf, n2};
    if (!g->d_bowWork) CU(cudaMalloc((void**)&g->d_bowWork, 64 * sizeof(int)));
    CU(launch_search_by_bow(a, B + off[10], b, nnratio, check_ori, (int32_t*)(g->d_outb + 16), g->d_bowWork, (int*)g->d_outb, g->stream, &g->launches,
                            kfForm ? B + off[11] : nullptr, kfForm && n1 > 0 ? (int32_t*)(g->d_outb + o12) : nullptr));
    CU(cudaMemcpyAsync(g->h_outb, g->d_outb, outBytes, cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    if (match_f) std::memcpy(match_f, g->h_outb + 16, (size_t)n2 * sizeof(int32_t));
    if (kfForm && n1 > 0) std::memcpy(match12, g->h_outb + o12, (size_t)n1 * sizeof(int32_t));
    if (nmatches) *nmatches = *(const int*)g->h_outb;
    return EORB_OK;
}

extern "C" int eorb_guided_search_by_bow(eorb_guided* g, const eorb_keypoint* kps_kf, const uint8_t* desc_kf, const uint8_t* valid_kf, int n1,
                                         const uint32_t* kf_nodes, const int32_t* kf_start, const uint32_t* kf_feats, int nkf,
                                         const eorb_keypoint* kps_f, const uint8_t* desc_f, int n2, const uint32_t* f_nodes, const int32_t* f_start,
                                         const uint32_t* f_feats, int nf, float nnratio, int check_ori, int32_t* match_f, int* nmatches) {
    return guidedBowHost(g, "eorb_guided_search_by_bow", kps_kf, desc_kf, valid_kf, n1, kf_nodes, kf_start, kf_feats, nkf, kps_f, desc_f, nullptr, n2,
                         f_nodes, f_start, f_feats, nf, nnratio, check_ori, match_f, nullptr, nmatches);
}

extern "C" int eorb_guided_search_by_bow_kf(eorb_guided* g, const eorb_keypoint* kps1, const uint8_t* desc1, const uint8_t* valid1, int n1,
                                            const uint32_t* nodes1, const int32_t* start1, const uint32_t* feats1, int nn1, const eorb_keypoint* kps2,
                                            const uint8_t* desc2, const uint8_t* valid2, int n2, const uint32_t* nodes2, const int32_t* start2,
                                            const uint32_t* feats2, int nn2, float nnratio, int check_ori, int32_t* match12, int* nmatches) {
    if (!match12 && n1 > 0) return gFail(EORB_ERR_ARG, "eorb_guided_search_by_bow_kf", "null output");
    return guidedBowHost(g, "eorb_guided_search_by_bow_kf", kps1, desc1, valid1, n1, nodes1, start1, feats1, nn1, kps2, desc2, valid2, n2, nodes2, start2,
                         feats2, nn2, nnratio, check_ori, nullptr, match12, nmatches);
}

// ------------------------------------------------------------------------------------------------ SearchForTriangulation
extern "C" int eorb_guided_search_for_triangulation(eorb_guided* g, const eorb_keypoint* kps1, const uint8_t* desc1, const uint8_t* flags1, int n1,
                                                    const uint32_t* nodes1, const int32_t* start1, const uint32_t* feats1, int nn1,
                                                    const eorb_keypoint* kps2, const uint8_t* desc2, const uint8_t* flags2, int n2,
                                                    const uint32_t* nodes2, const int32_t* start2, const uint32_t* feats2, int nn2, const float* F12,
                                                    const float* epipole2, const float* scale_factors2, const float* level_sigma2_2, int nlevels,
                                                    int coarse, int check_ori, int32_t* match12, int* nmatches) {
    const char* who = "eorb_guided_search_for_triangulation";
    if (!g) return gFail(EORB_ERR_ARG, who, "null handle");
    if (n1 < 0 || n2 < 0) return gFail(EORB_ERR_ARG, who, "negative size");
    if (n1 > EORB_GUIDED_MAX_KEYPOINTS || n2 > EORB_GUIDED_MAX_KEYPOINTS) return gFail(EORB_ERR_CAPACITY, who, "more than EORB_GUIDED_MAX_KEYPOINTS keypoints");
    if (nmatches) *nmatches = 0;
    if (n1 > 0 && !match12) return gFail(EORB_ERR_ARG, who, "null output");
    for (int i = 0; i < n1; i++) match12[i] = -1;
    if (n1 == 0 || n2 == 0) return EORB_OK;
    if (!kps1 || !desc1 || !flags1 || !kps2 || !desc2 || !flags2 || !F12 || !epipole2 || !scale_factors2 || !level_sigma2_2) return gFail(EORB_ERR_ARG, who, "null argument");
    if (nlevels < 1 || nlevels > 32) return gFail(EORB_ERR_ARG, who, "bad level tables (1..32 levels)");
    int rc;
    if ((rc = checkFeatureVector(who, nodes1, start1, feats1, nn1, n1)) != EORB_OK) return rc;
    if ((rc = checkFeatureVector(who, nodes2, start2, feats2, nn2, n2)) != EORB_OK) return rc;
    CU(cudaSetDevice(g->device));
    GuidedTriGeom tg;
    for (int i = 0; i < 9; i++) tg.F[i] = F12[i];
    tg.ep[0] = epipole2[0]; tg.ep[1] = epipole2[1]; tg.coarse = coarse ? 1 : 0;
    for (int i = 0; i < 32; i++) { tg.scale2[i] = scale_factors2[i < nlevels ? i : nlevels - 1]; tg.sigma2[i] = level_sigma2_2[i < nlevels ? i : nlevels - 1]; }
    // one blob, every part 16-byte aligned: [kps1 | desc1 | kps2 | desc2 | nodes1 | start1 | feats1 | nodes2 | start2 | feats2 | flags1 | flags2]
    const int nf1 = nn1 > 0 ? start1[nn1] : 0, nf2 = nn2 > 0 ? start2[nn2] : 0;
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const int NP = 12;
    const int32_t zero2[2] = {0, 0};
    const size_t sz[NP] = {(size_t)n1 * sizeof(eorb_keypoint), (size_t)n1 * 32, (size_t)n2 * sizeof(eorb_keypoint), (size_t)n2 * 32, (size_t)nn1 * 4,
                           (size_t)(nn1 + 1) * 4, (size_t)nf1 * 4, (size_t)nn2 * 4, (size_t)(nn2 + 1) * 4, (size_t)nf2 * 4, (size_t)n1, (size_t)n2};
    const void* src[NP] = {kps1, desc1, kps2, desc2, nodes1, nn1 > 0 ? (const void*)start1 : (const void*)zero2, feats1,
                           nodes2, nn2 > 0 ? (const void*)start2 : (const void*)zero2, feats2, flags1, flags2};
    size_t off[NP + 1]; off[0] = 0;
    for (int i = 0; i < NP; i++) off[i + 1] = off[i] + al(sz[i]);
    if (off[NP] > g->blobCap) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_blob); cudaFreeHost(g->h_blob); g->d_blob = nullptr; g->h_blob = nullptr; g->blobCap = 0;
        const size_t cap = std::max<size_t>(off[NP] + off[NP] / 4, 1 << 18);
        CU(cudaMalloc((void**)&g->d_blob, cap));
        CU(cudaMallocHost((void**)&g->h_blob, cap));
        g->blobCap = cap;
    }
    // output blob: [nmatches (16 bytes) | match12[n1] | rotation bin per feature of keyframe 1 (device scratch)]
    const size_t oBin = 16 + al((size_t)n1 * sizeof(int32_t));
    const size_t outBytes = oBin + al((size_t)n1);
    if (outBytes > g->outbCap) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_outb); cudaFreeHost(g->h_outb); g->d_outb = nullptr; g->h_outb = nullptr; g->outbCap = 0;
        const size_t cap = std::max<size_t>(outBytes + outBytes / 4, 1 << 14);
        CU(cudaMalloc((void**)&g->d_outb, cap));
        CU(cudaMallocHost((void**)&g->h_outb, cap));
        g->outbCap = cap;
    }
    for (int i = 0; i < NP; i++)
        if (sz[i] > 0) std::memcpy(g->h_blob + off[i], src[i], sz[i]);
    CU(cudaMemcpyAsync(g->d_blob, g->h_blob, off[NP], cudaMemcpyHostToDevice, g->stream));
    unsigned char* B = g->d_blob;
    GuidedBowSide a{(const eorb_keypoint*)(B + off[0]), B + off[1], (const uint32_t*)(B + off[4]), (const int32_t*)(B + off[5]), (const uint32_t*)(B + off[6]), nn1, n1};
    GuidedBowSide b{(const eorb_keypoint*)(B + off[2]), B + off[3], (const uint32_t*)(B + off[7]), (const int32_t*)(B + off[8]), (const uint32_t*)(B + off[9]), nn2, n2};
    if (!g->d_bowWork) CU(cudaMalloc((void**)&g->d_bowWork, 64 * sizeof(int)));
    CU(launch_search_triangulation(a, B + off[10], b, B + off[11], tg, check_ori, nf1, (int32_t*)(g->d_outb + 16), (signed char*)(g->d_outb + oBin), g->d_bowWork,
                                   (int*)g->d_outb, g->stream, &g->launches));
    CU(cudaMemcpyAsync(g->h_outb, g->d_outb, oBin, cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    std::memcpy(match12, g->h_outb + 16, (size_t)n1 * sizeof(int32_t));
    if (nmatches) *nmatches = *(const int*)g->h_outb;
    return EORB_OK;
}

/* the same with every array resident in HBM (FeatureVectors in the CSR form eorb_vocab_transform_resident leaves there); nentries1 = start1[nn1],
   d_match12 (n1 entries) is written on the device, *nmatches after a stream synchronisation */
extern "C" int eorb_guided_search_for_triangulation_device(eorb_guided* g, const eorb_keypoint* d_kps1, const uint8_t* d_desc1, const uint8_t* d_flags1,
                                                           int n1, const uint32_t* d_nodes1, const int32_t* d_start1, const uint32_t* d_feats1, int nn1,
                                                           int nentries1, const eorb_keypoint* d_kps2, const uint8_t* d_desc2, const uint8_t* d_flags2,
                                                           int n2, const uint32_t* d_nodes2, const int32_t* d_start2, const uint32_t* d_feats2, int nn2,
                                                           const float* F12, const float* epipole2, const float* scale_factors2,
                                                           const float* level_sigma2_2, int nlevels, int coarse, int check_ori, int32_t* d_match12,
                                                           int* nmatches) {
    const char* who = "eorb_guided_search_for_triangulation_device";
    if (!g) return gFail(EORB_ERR_ARG, who, "null handle");
    if (n1 < 0 || n2 < 0 || nn1 < 0 || nn2 < 0 || nentries1 < 0 || nentries1 > n1) return gFail(EORB_ERR_ARG, who, "bad size");
    if (n1 > EORB_GUIDED_MAX_KEYPOINTS || n2 > EORB_GUIDED_MAX_KEYPOINTS) return gFail(EORB_ERR_CAPACITY, who, "more than EORB_GUIDED_MAX_KEYPOINTS keypoints");
    if (nmatches) *nmatches = 0;
    if (n1 == 0) return EORB_OK;
    if (!d_match12 || !F12 || !epipole2 || !scale_factors2 || !level_sigma2_2) return gFail(EORB_ERR_ARG, who, "null argument");
    if (n2 > 0 && nn1 > 0 && nn2 > 0 && (!d_kps1 || !d_desc1 || !d_flags1 || !d_nodes1 || !d_start1 || !d_feats1 || !d_kps2 || !d_desc2 || !d_flags2 || !d_nodes2 ||
                                        !d_start2 || !d_feats2))
        return gFail(EORB_ERR_ARG, who, "null argument");
    if (nlevels < 1 || nlevels > 32) return gFail(EORB_ERR_ARG, who, "bad level tables (1..32 levels)");
    if (((uintptr_t)d_desc1 | (uintptr_t)d_desc2) & 15) return gFail(EORB_ERR_ARG, who, "descriptors must be 16-byte aligned");
    CU(cudaSetDevice(g->device));
    GuidedTriGeom tg;
    for (int i = 0; i < 9; i++) tg.F[i] = F12[i];
    tg.ep[0] = epipole2[0]; tg.ep[1] = epipole2[1]; tg.coarse = coarse ? 1 : 0;
    for (int i = 0; i < 32; i++) { tg.scale2[i] = scale_factors2[i < nlevels ? i : nlevels - 1]; tg.sigma2[i] = level_sigma2_2[i < nlevels ? i : nlevels - 1]; }
    const size_t outBytes = 16 + (((size_t)n1 + 15) & ~(size_t)15);            // [nmatches | rotation bin per feature]
    if (outBytes > g->outbCap) {
        CU(cudaStreamSynchronize(g->stream));
        cudaFree(g->d_outb); cudaFreeHost(g->h_outb); g->d_outb = nullptr; g->h_outb = nullptr; g->outbCap = 0;
        const size_t cap = std::max<size_t>(outBytes + outBytes / 4, 1 << 14);
        CU(cudaMalloc((void**)&g->d_outb, cap));
        CU(cudaMallocHost((void**)&g->h_outb, cap));
        g->outbCap = cap;
    }
    if (!g->d_bowWork) CU(cudaMalloc((void**)&g->d_bowWork, 64 * sizeof(int)));
    GuidedBowSide a{d_kps1, d_desc1, d_nodes1, d_start1, d_feats1, (n2 > 0 ? nn1 : 0), n1};
    GuidedBowSide b{d_kps2, d_desc2, d_nodes2, d_start2, d_feats2, nn2, n2};
    CU(launch_search_triangulation(a, d_flags1, b, d_flags2, tg, check_ori, nentries1, d_match12, (signed char*)(g->d_outb + 16), g->d_bowWork,
                                   (int*)g->d_outb, g->stream, &g->launches));
    CU(cudaMemcpyAsync(g->h_outb, g->d_outb, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    if (nmatches) *nmatches = *(const int*)g->h_outb;
    return EORB_OK;
}

