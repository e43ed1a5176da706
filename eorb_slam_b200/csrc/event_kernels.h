// event_kernels.h — launch interface of the event-frame kernels
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eorb_b200.h"

namespace eorb {

// per-window constants prepared on the host (angle-axis of the window pose is extracted once in double,
// like Eigen::AngleAxisd(R) at EventConversion.cc:303)
struct EvWindow {
    long long begin, end;      // event range
    double angle, axis[3];     // SE3: whole-window rotation
    double t[3];               // SE3: whole-window translation
};

struct EvConst {
    int mode, width, height, pol;
    float sigma, sig2, norm;   // norm = 2*pi*sig2
    int half;                  // ceil(3*sigma)
    float depth;
    float fx, fy, cx, cy;
    float se2[4];
    int se2_n;
    int cam;                   // 0 Pinhole, 1 KannalaBrandt8
    float kb[4];               // k1..k4
};

cudaError_t launch_ev_splat(const eorb_event* d_evs, const EvWindow* d_wins, int nwin, long long maxEventsPerWindow,
                            const EvConst& c, float* d_img, cudaStream_t st, long long* launches);
cudaError_t launch_ev_frames(const eorb_event* d_evs, const EvWindow* d_wins, int nwin, long long maxEventsPerWindow, const EvConst& c,
                             int normMode, float* d_img, float* d_minmax, uint8_t* d_u8, cudaStream_t st, long long* launches,
                             float2* d_xyScratch = nullptr);   // [>= last window end] floats pairs: enables the one-pass warp for multi-band frames
#define EORB_EV_FOCUS_MAX_CELLS 1024
cudaError_t launch_ev_focus(const float* d_img, int nwin, int W, int H, int patch, int what, int avg, float* d_out, cudaStream_t st,
                            long long* launches);
cudaError_t launch_ev_jac(const eorb_event* d_evs, const EvWindow* d_win, long long nev, const EvConst& c, int globalMean, float* d_frames7,
                          float* d_out6, cudaStream_t st, long long* launches);
cudaError_t launch_ev_normalize(const float* d_img, int nwin, int npix, int normMode, float* d_minmax, uint8_t* d_u8,
                                cudaStream_t st, long long* launches);

}  // namespace eorb
