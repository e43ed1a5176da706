// capi_lk.cu — C ABI of the pyramidal Lucas-Kanade tracker (include/eorb_b200.h, section "LK tracker").
// Mirrors EORB_SLAM::ELK_Tracker (include/Event/KLT_Tracker.h, src/Event/KLT_Tracker.cpp:14-91): setRefImage keeps the
// reference frame and points, trackCurrImage runs cv::calcOpticalFlowPyrLK against every new frame.  Host code only;
// the compute steps are the kernels of lk_kernels.cu.  No CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/eorb_b200.h"
#include "lk_kernels.h"

using namespace eorb;

extern "C" int eorb_internal_fail(int code, const char* msg);
static int lkFail(int code, const char* what, const char* detail) {
    char buf[400];
    snprintf(buf, sizeof(buf), "%s%s%s", what, detail ? ": " : "", detail ? detail : "");
    return eorb_internal_fail(code, buf);
}
#define CU(call)                                                                         \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) return lkFail(EORB_ERR_CUDA, #call, cudaGetErrorString(e__)); \
    } while (0)

struct eorb_lk {
    int device = 0, maxW = 0, maxH = 0, maxPts = 0;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    // pyramid slabs: level l of image k lives at base[k] + off[l]; derivatives of the reference in d_deriv + doff[l]
    uint8_t* d_ref = nullptr; uint8_t* d_cur = nullptr; short2* d_deriv = nullptr;
    size_t off[EORB_LK_MAX_LEVELS] = {0}, doff[EORB_LK_MAX_LEVELS] = {0}, slabBytes = 0, derivCount = 0;
    int lw[EORB_LK_MAX_LEVELS] = {0}, lh[EORB_LK_MAX_LEVELS] = {0}, lpitch[EORB_LK_MAX_LEVELS] = {0};
    int w = 0, h = 0, win = 0, maxLevel = -1, nref = 0;
    // outputs share one device buffer [next float2[maxPts] | err float[maxPts] | status u8[maxPts]] and one pinned mirror: one D2H
    float2* d_prev = nullptr; float2* d_next = nullptr; uint8_t* d_status = nullptr; float* d_err = nullptr;
    uint8_t* d_out = nullptr; uint8_t* h_out = nullptr; uint8_t* h_img = nullptr; float* h_init = nullptr;
    // ELK_Tracker state on the device (track_and_match): mRefKPoints, and mLastTrackedPts = d_next itself while lastValid
    eorb_keypoint* d_refKps = nullptr; uint8_t* d_match = nullptr; uint8_t* h_match = nullptr;   // d_match: [kps | pxDisp | counts | matched]
    bool haveRefKps = false, lastValid = false;
    long long launches = 0;
};

static inline int roundUp(int v, int a) { return (v + a - 1) / a * a; }

extern "C" int eorb_lk_create(int device, int max_width, int max_height, int max_points, eorb_lk** out) {
    if (!out || max_width < 1 || max_height < 1 || max_points < 1) return lkFail(EORB_ERR_ARG, "eorb_lk_create", "bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return lkFail(EORB_ERR_CUDA, "no CUDA device: eorb_b200 has no CPU fallback", nullptr); }
    if (device < 0 || device >= ndev) return lkFail(EORB_ERR_ARG, "eorb_lk_create", "device out of range");
    CU(cudaSetDevice(device));
    eorb_lk* h = new eorb_lk();
    h->device = device; h->maxW = max_width; h->maxH = max_height; h->maxPts = max_points;
    CU(cudaStreamCreateWithFlags(&h->ownStream, cudaStreamNonBlocking));
    h->stream = h->ownStream;
    size_t bytes = 0, dcount = 0;
    int w = max_width, hh = max_height;
    for (int l = 0; l < EORB_LK_MAX_LEVELS; l++) {
        bytes += (size_t)roundUp(w, 16) * hh; dcount += (size_t)w * hh;
        w = (w + 1) / 2; hh = (hh + 1) / 2;
    }
    h->slabBytes = bytes; h->derivCount = dcount;
    CU(cudaMalloc((void**)&h->d_ref, bytes)); CU(cudaMalloc((void**)&h->d_cur, bytes));
    CU(cudaMalloc((void**)&h->d_deriv, dcount * sizeof(short2)));
    CU(cudaMalloc((void**)&h->d_prev, (size_t)max_points * sizeof(float2)));
    const size_t outBytes = (size_t)max_points * (sizeof(float2) + sizeof(float) + 1);
    CU(cudaMalloc((void**)&h->d_out, outBytes));
    h->d_next = (float2*)h->d_out; h->d_err = (float*)(h->d_out + (size_t)max_points * sizeof(float2));
    h->d_status = h->d_out + (size_t)max_points * (sizeof(float2) + sizeof(float));
    CU(cudaMallocHost((void**)&h->h_out, outBytes));
    CU(cudaMallocHost((void**)&h->h_img, (size_t)max_width * max_height));
    CU(cudaMallocHost((void**)&h->h_init, (size_t)max_points * sizeof(float2)));
    CU(cudaMalloc((void**)&h->d_refKps, (size_t)max_points * sizeof(eorb_keypoint)));
    CU(cudaMalloc((void**)&h->d_match, (size_t)max_points * (sizeof(eorb_keypoint) + sizeof(float) + 1) + 16));
    CU(cudaMallocHost((void**)&h->h_match, (size_t)max_points * (sizeof(eorb_keypoint) + sizeof(float) + 1) + 16));
    *out = h;
    return EORB_OK;
}

extern "C" int eorb_lk_destroy(eorb_lk* h) {
    if (!h) return EORB_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaFree(h->d_ref); cudaFree(h->d_cur); cudaFree(h->d_deriv); cudaFree(h->d_prev); cudaFree(h->d_out);
    cudaFree(h->d_refKps); cudaFree(h->d_match); cudaFreeHost(h->h_match);
    cudaFreeHost(h->h_out); cudaFreeHost(h->h_img); cudaFreeHost(h->h_init);
    cudaStreamDestroy(h->ownStream);
    delete h;
    return EORB_OK;
}

extern "C" int eorb_lk_set_stream(eorb_lk* h, void* s) {
    if (!h) return lkFail(EORB_ERR_ARG, "null handle", nullptr);
    CU(cudaStreamSynchronize(h->stream));
    h->stream = (cudaStream_t)s;
    return EORB_OK;
}
extern "C" int eorb_lk_reset_stream(eorb_lk* h) {
    if (!h) return lkFail(EORB_ERR_ARG, "null handle", nullptr);
    CU(cudaStreamSynchronize(h->stream));
    h->stream = h->ownStream;
    return EORB_OK;
}
extern "C" long long eorb_lk_launch_count(const eorb_lk* h) { return h ? h->launches : 0; }

// level geometry as buildOpticalFlowPyramid derives it: halve (round up) until a side would be <= the window
static void lkGeometry(eorb_lk* h, int w, int hgt, int win, int maxLevel) {
    size_t o = 0, d = 0;
    int lw = w, lh = hgt, l = 0;
    for (;; l++) {
        h->lw[l] = lw; h->lh[l] = lh; h->lpitch[l] = roundUp(lw, 16); h->off[l] = o; h->doff[l] = d;
        o += (size_t)h->lpitch[l] * lh; d += (size_t)lw * lh;
        const int nw = (lw + 1) / 2, nh = (lh + 1) / 2;
        if (l + 1 > maxLevel || l + 1 >= EORB_LK_MAX_LEVELS || nw <= win || nh <= win) break;
        lw = nw; lh = nh;
    }
    h->maxLevel = l; h->w = w; h->h = hgt; h->win = win;
}

static int lkBuildPyramid(eorb_lk* h, uint8_t* slab, const uint8_t* img, size_t stride, bool deviceSrc) {
    if (deviceSrc) {
        CU(cudaMemcpy2DAsync(slab, h->lpitch[0], img, stride, h->w, h->h, cudaMemcpyDeviceToDevice, h->stream));
    } else {   // through the pinned staging image: a truly asynchronous H2D instead of the driver's pageable path
        for (int y = 0; y < h->h; y++) memcpy(h->h_img + (size_t)y * h->w, img + (size_t)y * stride, (size_t)h->w);
        CU(cudaMemcpy2DAsync(slab, h->lpitch[0], h->h_img, h->w, h->w, h->h, cudaMemcpyHostToDevice, h->stream));
    }
    for (int l = 1; l <= h->maxLevel; l++) {
        CU(launch_lk_pyrdown(slab + h->off[l - 1], h->lw[l - 1], h->lh[l - 1], h->lpitch[l - 1], slab + h->off[l], h->lw[l], h->lh[l], h->lpitch[l], h->stream));
        h->launches++;
    }
    return EORB_OK;
}

static int lkSetRef(eorb_lk* h, const uint8_t* img, int w, int hgt, size_t stride, bool deviceImg, const float* pts_xy, int n, int win, int max_level) {
    if (!h) return lkFail(EORB_ERR_ARG, "null handle", nullptr);
    if (!img || w <= 0 || hgt <= 0 || n <= 0 || !pts_xy) return EORB_EMPTY;   // assert(!image.empty() && !refPts.empty()) :22
    if (w > h->maxW || hgt > h->maxH || (size_t)roundUp(w, 16) * hgt > (size_t)roundUp(h->maxW, 16) * h->maxH) return lkFail(EORB_ERR_CAPACITY, "image exceeds the tracker's capacity", nullptr);
    if (n > h->maxPts) return lkFail(EORB_ERR_CAPACITY, "more points than max_points", nullptr);
    if (win < 3 || win > EORB_LK_MAX_WIN || max_level < 0) return lkFail(EORB_ERR_ARG, "window must be 3..33, maxLevel >= 0", nullptr);
    CU(cudaSetDevice(h->device));
    lkGeometry(h, w, hgt, win, max_level);
    int rc = lkBuildPyramid(h, h->d_ref, img, stride, deviceImg);
    if (rc != EORB_OK) return rc;
    for (int l = 0; l <= h->maxLevel; l++) {
        CU(launch_lk_scharr(h->d_ref + h->off[l], h->lw[l], h->lh[l], h->lpitch[l], h->d_deriv + h->doff[l], h->stream));
        h->launches++;
    }
    CU(cudaMemcpyAsync(h->d_prev, pts_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));   // pts_xy may be pageable caller memory
    h->nref = n; h->haveRefKps = false; h->lastValid = false;
    return EORB_OK;
}

extern "C" int eorb_lk_set_ref(eorb_lk* h, const uint8_t* img, int w, int hgt, size_t stride, const float* pts_xy, int n, int win, int max_level) {
    return lkSetRef(h, img, w, hgt, stride, false, pts_xy, n, win, max_level);
}
extern "C" int eorb_lk_set_ref_device(eorb_lk* h, const uint8_t* d_img, int w, int hgt, size_t stride, const float* pts_xy, int n, int win,
                                      int max_level) {
    return lkSetRef(h, d_img, w, hgt, stride, true, pts_xy, n, win, max_level);
}

// builds the current frame's pyramid and launches the tracker.  initMode 0: start from the reference points; 1: init_xy (host) is the
// initial flow; 2: the points left in d_next by the previous call are (ELK_Tracker::mLastTrackedPts, resident)
static int lkLaunchTrack(eorb_lk* h, const uint8_t* img, size_t stride, bool deviceImg, int initMode, const float* init_xy, int max_iter, double eps,
                         float min_eig, bool wantErr) {
    const int n = h->nref;
    int rc = lkBuildPyramid(h, h->d_cur, img, stride, deviceImg);
    if (rc != EORB_OK) return rc;
    if (initMode == 1) {
        memcpy(h->h_init, init_xy, (size_t)n * sizeof(float2));
        CU(cudaMemcpyAsync(h->d_next, h->h_init, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
    }
    LkLevels L{};
    L.maxLevel = h->maxLevel;
    for (int l = 0; l <= h->maxLevel; l++) {
        L.lv[l].I = h->d_ref + h->off[l]; L.lv[l].J = h->d_cur + h->off[l]; L.lv[l].dI = h->d_deriv + h->doff[l];
        L.lv[l].w = h->lw[l]; L.lv[l].h = h->lh[l]; L.lv[l].pitch = h->lpitch[l];
    }
    LkParams p{};
    p.win = h->win; p.maxIter = std::min(std::max(max_iter, 0), 100); p.useInitialFlow = initMode != 0 ? 1 : 0;
    const double e = std::min(std::max(eps, 0.), 10.);
    p.epsilon2 = e * e; p.minEigThreshold = min_eig;
    CU(launch_lk_track(L, p, h->d_prev, h->d_next, n, h->d_status, wantErr ? h->d_err : nullptr, h->stream));
    h->launches++;
    h->lastValid = true;   // d_next now holds this call's tracked points
    return EORB_OK;
}

static int lkTrack(eorb_lk* h, const uint8_t* img, size_t stride, bool deviceImg, const float* init_xy, int max_iter, double eps, float min_eig,
                   float* out_xy, uint8_t* status, float* err) {
    if (!h) return lkFail(EORB_ERR_ARG, "null handle", nullptr);
    if (h->nref <= 0) return lkFail(EORB_ERR_STATE, "set the reference image and points first", nullptr);   // KLT_Tracker.cpp:52-56
    if (!img) return EORB_EMPTY;
    if (!out_xy || !status) return lkFail(EORB_ERR_ARG, "null output", nullptr);
    CU(cudaSetDevice(h->device));
    const int n = h->nref;
    int rc = lkLaunchTrack(h, img, stride, deviceImg, init_xy ? 1 : 0, init_xy, max_iter, eps, min_eig, err != nullptr);
    if (rc != EORB_OK) return rc;
    const size_t outBytes = (size_t)h->maxPts * (sizeof(float2) + sizeof(float) + 1);
    CU(cudaMemcpyAsync(h->h_out, h->d_out, outBytes, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    memcpy(out_xy, h->h_out, (size_t)n * sizeof(float2));
    if (err) memcpy(err, h->h_out + (size_t)h->maxPts * sizeof(float2), (size_t)n * sizeof(float));
    memcpy(status, h->h_out + (size_t)h->maxPts * (sizeof(float2) + sizeof(float)), (size_t)n);
    return h->maxLevel;
}

// ---- ELK_Tracker with its state on the device ------------------------------------------------------------------------------------
// setRefImage(image, vector<KeyPoint>) (KLT_Tracker.cpp:22-46): mRefKPoints = refPts, mRefPoints = mLastTrackedPts = their pt
extern "C" int eorb_lk_set_ref_keypoints(eorb_lk* h, const uint8_t* img, int w, int hgt, size_t stride, int img_on_device,
                                         const eorb_keypoint* ref_kps, int n, int win, int max_level) {
    if (!h) return lkFail(EORB_ERR_ARG, "null handle", nullptr);
    if (!img || w <= 0 || hgt <= 0 || n <= 0 || !ref_kps) return EORB_EMPTY;
    if (n > h->maxPts) return lkFail(EORB_ERR_CAPACITY, "more points than max_points", nullptr);
    std::vector<float> pts((size_t)n * 2);
    for (int i = 0; i < n; i++) { pts[2 * i] = ref_kps[i].x; pts[2 * i + 1] = ref_kps[i].y; }
    int rc = lkSetRef(h, img, w, hgt, stride, img_on_device != 0, pts.data(), n, win, max_level);
    if (rc != EORB_OK) return rc;
    CU(cudaMemcpyAsync(h->d_refKps, ref_kps, (size_t)n * sizeof(eorb_keypoint), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_next, h->d_prev, (size_t)n * sizeof(float2), cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->haveRefKps = true; h->lastValid = true;
    return EORB_OK;
}

// setLastTrackedPts (KLT_Tracker.cpp:252-262).  A list whose size differs from the reference's makes the next trackCurrImage start from
// the reference points without OPTFLOW_USE_INITIAL_FLOW (:63-70), which is what lastValid = false does.
extern "C" int eorb_lk_set_last_tracked(eorb_lk* h, const eorb_keypoint* kps, int n) {
    if (!h) return lkFail(EORB_ERR_ARG, "null handle", nullptr);
    if (h->nref <= 0) return lkFail(EORB_ERR_STATE, "set the reference image and points first", nullptr);
    if (n < 0 || (n > 0 && !kps)) return lkFail(EORB_ERR_ARG, "eorb_lk_set_last_tracked", "bad argument");
    if (n != h->nref) { h->lastValid = false; return EORB_OK; }
    CU(cudaSetDevice(h->device));
    for (int i = 0; i < n; i++) { h->h_init[2 * i] = kps[i].x; h->h_init[2 * i + 1] = kps[i].y; }
    CU(cudaMemcpyAsync(h->d_next, h->h_init, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->lastValid = true;
    return EORB_OK;
}

extern "C" int eorb_lk_get_last_tracked(eorb_lk* h, float* pts_xy) {
    if (!h || !pts_xy) return lkFail(EORB_ERR_ARG, "eorb_lk_get_last_tracked", "null argument");
    if (h->nref <= 0 || !h->lastValid) return lkFail(EORB_ERR_STATE, "no tracked points on the device", nullptr);
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h->h_init, h->d_next, (size_t)h->nref * sizeof(float2), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    memcpy(pts_xy, h->h_init, (size_t)h->nref * sizeof(float2));
    return h->nref;
}

// trackAndMatchCurrImage (:215-234) / trackAndMatchCurrImageInit (:236-242): LK from the resident last tracked points, refineTrackedPts
// (+ refineFirstOctaveLevel) on the device; the tracked points stay in HBM as the next call's initial flow.
static int lkTrackAndMatch(eorb_lk* h, const uint8_t* img, size_t stride, bool deviceImg, int max_iter, double eps, float min_eig,
                           int firstOctaveOnly, eorb_keypoint* tracked, uint8_t* matched, float* pxDisp, int* counts, bool deviceOut) {
    if (!h) return lkFail(EORB_ERR_ARG, "null handle", nullptr);
    if (h->nref <= 0 || !h->haveRefKps) return EORB_EMPTY;   // "No reference info., did you forget to init. tracker??" -> returns 0 matches (:218-221)
    if (!img) return EORB_EMPTY;
    if (!tracked || !matched || !pxDisp || !counts) return lkFail(EORB_ERR_ARG, "null output", nullptr);
    CU(cudaSetDevice(h->device));
    const int n = h->nref;
    int rc = lkLaunchTrack(h, img, stride, deviceImg, h->lastValid ? 2 : 0, nullptr, max_iter, eps, min_eig, false);
    if (rc != EORB_OK) return rc;
    eorb_keypoint* dk = deviceOut ? tracked : (eorb_keypoint*)h->d_match;
    float* dd = deviceOut ? pxDisp : (float*)(h->d_match + (size_t)n * sizeof(eorb_keypoint));
    int* dc = deviceOut ? counts : (int*)(h->d_match + (size_t)n * (sizeof(eorb_keypoint) + sizeof(float)));
    uint8_t* dm = deviceOut ? matched : h->d_match + (size_t)n * (sizeof(eorb_keypoint) + sizeof(float)) + 8;
    CU(launch_lk_refine(h->d_next, h->d_status, h->d_refKps, n, h->w, h->h, firstOctaveOnly, dk, dm, dd, dc, h->stream));
    h->launches++;
    if (deviceOut) return h->maxLevel;
    const size_t bytes = (size_t)n * (sizeof(eorb_keypoint) + sizeof(float) + 1) + 8;
    CU(cudaMemcpyAsync(h->h_match, h->d_match, bytes, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    memcpy(tracked, h->h_match, (size_t)n * sizeof(eorb_keypoint));
    memcpy(counts, h->h_match + (size_t)n * (sizeof(eorb_keypoint) + sizeof(float)), 8);
    memcpy(pxDisp, h->h_match + (size_t)n * sizeof(eorb_keypoint), (size_t)counts[1] * sizeof(float));
    memcpy(matched, h->h_match + (size_t)n * (sizeof(eorb_keypoint) + sizeof(float)) + 8, (size_t)n);
    return h->maxLevel;
}

extern "C" int eorb_lk_track_and_match(eorb_lk* h, const uint8_t* img, size_t stride, int img_on_device, int max_iter, double eps, float min_eig,
                                       int first_octave_only, eorb_keypoint* tracked, uint8_t* matched, float* px_disp, int* counts2) {
    return lkTrackAndMatch(h, img, stride, img_on_device != 0, max_iter, eps, min_eig, first_octave_only, tracked, matched, px_disp, counts2, false);
}
extern "C" int eorb_lk_track_and_match_device(eorb_lk* h, const uint8_t* d_img, size_t stride, int max_iter, double eps, float min_eig,
                                              int first_octave_only, eorb_keypoint* d_tracked, uint8_t* d_matched, float* d_px_disp,
                                              int* d_counts2) {
    return lkTrackAndMatch(h, d_img, stride, true, max_iter, eps, min_eig, first_octave_only, d_tracked, d_matched, d_px_disp, d_counts2, true);
}

extern "C" int eorb_lk_track(eorb_lk* h, const uint8_t* img, size_t stride, const float* init_xy, int max_iter, double eps, float min_eig,
                             float* out_xy, uint8_t* status, float* err) {
    return lkTrack(h, img, stride, false, init_xy, max_iter, eps, min_eig, out_xy, status, err);
}
extern "C" int eorb_lk_track_device(eorb_lk* h, const uint8_t* d_img, size_t stride, const float* init_xy, int max_iter, double eps,
                                    float min_eig, float* out_xy, uint8_t* status, float* err) {
    return lkTrack(h, d_img, stride, true, init_xy, max_iter, eps, min_eig, out_xy, status, err);
}
