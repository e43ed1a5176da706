// lk_kernels.h — launch interface of the pyramidal Lucas-Kanade kernels
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eorb_b200.h"

namespace eorb {

#define EORB_LK_MAX_LEVELS 8
#define EORB_LK_MAX_WIN 33

struct LkLevelDev {
    const uint8_t* I;      // reference image level
    const uint8_t* J;      // current image level
    const short2* dI;      // (Ix, Iy) of the reference level, w*h entries
    int w, h, pitch;       // pitch in bytes of I and J
};

struct LkLevels {
    LkLevelDev lv[EORB_LK_MAX_LEVELS];
    int maxLevel;          // last valid level index
};

struct LkParams {
    int win, maxIter, useInitialFlow;
    double epsilon2;       // squared, clamped
    float minEigThreshold;
};

cudaError_t launch_lk_pyrdown(const uint8_t* src, int w, int h, int pitch, uint8_t* dst, int dw, int dh, int dpitch, cudaStream_t st);
cudaError_t launch_lk_scharr(const uint8_t* src, int w, int h, int pitch, short2* dst, cudaStream_t st);
cudaError_t launch_lk_track(const LkLevels& L, const LkParams& p, const float2* prevPts, float2* nextPts, int n, uint8_t* status, float* err,
                            cudaStream_t st);
// ELK_Tracker::refineTrackedPts / refineFirstOctaveLevel on one LK result (one block; see lk_kernels.cu)
cudaError_t launch_lk_refine(const float2* curr, const uint8_t* status, const eorb_keypoint* ref, int n, int W, int H, int firstOctaveOnly,
                             eorb_keypoint* tracked, uint8_t* matched, float* pxDisp, int* counts, cudaStream_t st);

}  // namespace eorb
