// guided_kernels.h — launch interface of the guided-matching kernels: the Frame grid
// (Frame::AssignFeaturesToGrid / GetFeaturesInArea, src/Frame.cc:431-460, 709-793) and
// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:714-831).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eorb_b200.h"

namespace eorb {

#define EORB_GRID_COLS 64          // FRAME_GRID_COLS (include/Frame.h:46)
#define EORB_GRID_ROWS 48          // FRAME_GRID_ROWS (include/Frame.h:45)
#define EORB_GRID_CELLS (EORB_GRID_COLS * EORB_GRID_ROWS)
#define EORB_GUIDED_MAX_KPS 16384  // keypoints per frame (the reference extracts <= 5 * nFeatures = 5000)
#define EORB_GUIDED_TOP 32         // sorted head of every query's candidate list kept for the sequential pass

// grid geometry, computed on the host in float exactly as Frame.cc:165-166
struct GuidedGrid { float minX, minY, wInv, hInv; };

struct GuidedFrame {
    const eorb_keypoint* kps; const uint8_t* desc; int n;
};

struct GuidedWork {
    // frame-2 grid (CSR): cellStart[EORB_GRID_CELLS + 1], cellIdx[n2]
    int* cellStart; int* cellIdx; int* assigned;
    // per query (frame-1 keypoint): candidate range, sorted head (dist<<32 | pos<<16 | i2), histogram bin
    int* candOff; int* candCnt; unsigned long long* top; signed char* bin;
    eorb_area_query* q;            // per query window of SearchByProjection (guided_project_kernel)
    uint32_t* cand; int candCap;   // all candidates in the reference's visiting order: dist<<16 | i2
    int* total;                    // candidates produced (may exceed candCap: the host grows the buffer and retries)
};

cudaError_t launch_frame_grid(const eorb_keypoint* d_kps, int n, GuidedGrid g, int* d_cellStart, int* d_cellIdx, int* d_assigned,
                              cudaStream_t st);
// batch of GetFeaturesInArea queries: q = (x, y, r, minLevel, maxLevel as floats/ints packed in eorb_area_query)
cudaError_t launch_features_in_area(const eorb_keypoint* d_kps, GuidedGrid g, const int* d_cellStart, const int* d_cellIdx,
                                    const eorb_area_query* d_q, int nq, int* d_count, int* d_out, int capPerQuery, cudaStream_t st);
cudaError_t launch_search_init(const GuidedFrame& f1, const GuidedFrame& f2, GuidedGrid g, float* d_prevXY, int window, float nnratio,
                               int checkOri, const GuidedWork& w, int32_t* d_matches12, int* d_nmatches, cudaStream_t st,
                               long long* launches);
// SearchByProjection(CurrentFrame, LastFrame, th, bMono = true): projection constants (Pinhole intrinsics, image bounds, scale table)
struct GuidedProj { float fx, fy, cx, cy, minX, minY, maxX, maxY, th; int nlevels; float scale[32]; };
// variants of the projection window (guided_project_kernel): levelMode 0 = [oct-1, oct+1], 1 = bForward [oct, inf), 2 = bBackward
// [0, oct] (ORBmatcher.cc:2024-2029); level1 != nullptr: the predicted level per map point instead of the keypoint octave and
// reloc = 1: no depth-sign test (SearchByProjection(CurrentFrame, pKF, sAlreadyFound, ...), :2189-2312); qUr != nullptr receives
// ur = u - mbf / zc per query for the rectified-stereo column test (:2049-2055)
struct GuidedProjMode { int levelMode; int reloc; float mbf; const int32_t* level1; float* qUr; };
// d_obs1 == nullptr: every set slot blocks; d_held2 (may be null): slots taken on entry; thHigh: acceptance threshold
cudaError_t launch_search_proj(const float* d_x3Dc, const uint8_t* d_valid1, const int32_t* d_obs1, const eorb_keypoint* d_kps1,
                               const uint8_t* d_descMP, int n1, const GuidedFrame& f2, GuidedGrid g, const GuidedProj& pr, int checkOri,
                               const GuidedWork& w, int32_t* d_claim, int32_t* d_matchCur, int* d_nmatches, cudaStream_t st,
                               long long* launches, const GuidedProjMode& md, const float* d_uRight2, const uint8_t* d_held2, int thHigh);
// SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints), monocular (ORBmatcher.cc:44-148); pr carries th, nlevels, scale
cudaError_t launch_search_map_points(const eorb_track_point* d_pts, const uint8_t* d_descMP, int n1, const GuidedFrame& f2,
                                     const uint8_t* d_held2, GuidedGrid g, const GuidedProj& pr, int farPoints, float thFar, float nnratio,
                                     const GuidedWork& w, int32_t* d_matchCur, int* d_nmatches, cudaStream_t st, long long* launches,
                                     const float* d_projXR /* mTrackProjXR per point or null */, const float* d_uRight2 /* F.mvuRight or null */);
// SearchByBoW(pKF, F, vpMapPointMatches), monocular (ORBmatcher.cc:276-478): one side = keypoints, descriptors and the
// FeatureVector in CSR form (nodes ascending, features of node q = feats[start[q] .. start[q+1]))
struct GuidedBowSide { const eorb_keypoint* kps; const uint8_t* desc; const uint32_t* nodes; const int32_t* start; const uint32_t* feats; int nnodes, n; };
cudaError_t launch_search_by_bow(const GuidedBowSide& kf, const uint8_t* d_validKF, const GuidedBowSide& f, float nnratio, int checkOri,
                                 int32_t* d_matchF, int* d_work /* 64 ints */, int* d_nmatches, cudaStream_t st, long long* launches,
                                 const uint8_t* d_validF = nullptr /* keyframe-keyframe form: good map points of side 2 */,
                                 int32_t* d_match12 = nullptr /* keyframe-keyframe form: result indexed by side 1 (kf.n entries) */);
// Keyframe-side window searches (eorb_guided_search_windows): the caller's windows (x, y, r, level range) per map point; static
// candidate filters: slots taken on entry (non-blocking searches) and the reprojection gate of Fuse (chi2, invSigma2 = mvInvLevelSigma2).
// g = geometry the keypoints were binned with (the Frame's float bounds), gq = geometry of the window lookup (the KeyFrame's int bounds)
struct GuidedCandExtra { const uint8_t* held2 = nullptr; int chi2 = 0; float invSigma2[32] = {0}; };
cudaError_t launch_search_windows(const eorb_area_query* d_q, const float* d_qUr, const uint8_t* d_descMP, int n1, const GuidedFrame& f2,
                                  const uint8_t* d_held2, const float* d_uRight2, GuidedGrid g, GuidedGrid gq, const GuidedCandExtra& cx, int blocking,
                                  int thHigh, const GuidedWork& w, int32_t* d_claim, int32_t* d_bestIdx, int32_t* d_bestDist, int32_t* d_match2, int* d_nmatches,
                                  cudaStream_t st, long long* launches);
// SearchForTriangulation (ORBmatcher.cc:975-1214), pinhole keyframes: F = F12 row-major as Pinhole::epipolarConstrain forms it, ep = the epipole
// in the second image, scale2 / sigma2 = mvScaleFactors / mvLevelSigma2 of the second keyframe, coarse = bCoarse.  flags bit 0: the feature takes
// part (no map point; stereo when bOnlyStereo), bit 1: it has a right-image column (bStereo)
struct GuidedTriGeom { float F[9]; float ep[2]; float scale2[32]; float sigma2[32]; int coarse; };
cudaError_t launch_search_triangulation(const GuidedBowSide& k1, const uint8_t* d_flags1, const GuidedBowSide& k2, const uint8_t* d_flags2,
                                        const GuidedTriGeom& tg, int checkOri, int nEntries1 /* start1[nnodes1] */, int32_t* d_match12, signed char* d_binOf,
                                        int* d_work /* 64 ints */, int* d_nmatches, cudaStream_t st, long long* launches);
cudaError_t guided_configure();

}  // namespace eorb
