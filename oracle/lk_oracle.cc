// oracle/lk_oracle.cc — CPU ORACLE (test infrastructure only) for the first "next" row of SURVEY.md §8f:
// pyramidal Lucas-Kanade tracking of event-frame keypoints.
//
// Reference call: ELK_Tracker::trackCurrImage, src/Event/KLT_Tracker.cpp:49-91 —
//   cv::calcOpticalFlowPyrLK(mRefFrame, currImage, mRefPoints, kpts, status, err, Size(win, win), maxLevel,
//                            TermCriteria(COUNT+EPS, maxItr, eps) [, OPTFLOW_USE_INITIAL_FLOW])
// with win = 23, maxLevel = 1, maxItr = 10, eps = 0.03 (Examples/Event/EvETHZ.yaml:205-208).
//
// The arithmetic lives in OpenCV (video/src/lkpyramid.cpp, imgproc pyrDown), which is not vendored in the
// reference; its published algorithm is restated here:
//   * pyramid: level l+1 = pyrDown(level l): separable [1 4 6 4 1] on u8 in integers, (sum + 128) >> 8, size
//     ((w+1)/2, (h+1)/2), BORDER_REFLECT_101; levels stop when a side is <= the window (buildOpticalFlowPyramid);
//   * derivatives: calcSharrDeriv — Ix = [3 10 3]^T smoothing x [-1 0 1], Iy = [-1 0 1]^T x [3 10 3], int16,
//     REFLECT_101 inside the level; the derivative image is read with a ZERO border, the images with a
//     REFLECT_101 border (both of width = window);
//   * per point and level (LKTrackerInvoker): bilinear patch in 14-bit fixed point
//     (I: DESCALE(.., 9) -> int16 = 32 x intensity; Ix, Iy: DESCALE(.., 14)), 2x2 gradient matrix, min-eigenvalue
//     test (1e-4), Newton iterations with the COUNT / EPS / oscillation stops, L1 patch error / (32 * win^2).
// Where OpenCV is not a function of its inputs — its SSE/AVX paths accumulate the float sums A11..b2 in four or
// eight lanes, the scalar path sequentially — this oracle accumulates the INTEGER products exactly (int64) and
// converts once; that is within one float rounding of any summation order and makes the CUDA path bit-comparable.
// Pin: tests/test_oracle_lk.py checks pyrDown bit-exactly and the tracked points against cv2.calcOpticalFlowPyrLK
// (cv2 4.13.0) within 0.02 px on >= 99 % of the points (status equal), the residual being OpenCV's own
// summation-order freedom.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "oracle.h"

extern "C" {

static inline int lk_reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

// cv::pyrDown on CV_8UC1 (PyrDownVec / PyrDown_<FixPtCast<uchar, 8>, ...>), BORDER_REFLECT_101
void orc_pyrdown_u8(const uint8_t* src, int w, int h, size_t stride, uint8_t* dst, int dw, int dh, size_t dstride) {
    std::vector<int> rows((size_t)5 * dw);
    for (int y = 0; y < dh; y++) {
        for (int k = 0; k < 5; k++) {
            const int sy = lk_reflect101(2 * y + k - 2, h);
            const uint8_t* s = src + (size_t)sy * stride;
            int* r = rows.data() + (size_t)k * dw;
            for (int x = 0; x < dw; x++) {
                const int x0 = lk_reflect101(2 * x - 2, w), x1 = lk_reflect101(2 * x - 1, w), x2 = lk_reflect101(2 * x, w);
                const int x3 = lk_reflect101(2 * x + 1, w), x4 = lk_reflect101(2 * x + 2, w);
                r[x] = s[x2] * 6 + (s[x1] + s[x3]) * 4 + s[x0] + s[x4];
            }
        }
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < dw; x++) {
            const int v = rows[2 * dw + x] * 6 + (rows[dw + x] + rows[3 * dw + x]) * 4 + rows[x] + rows[4 * dw + x];
            d[x] = (uint8_t)((v + 128) >> 8);
        }
    }
}

// calcSharrDeriv (lkpyramid.cpp): dst[2*(y*w+x)] = Ix, [..+1] = Iy, int16
void orc_scharr_deriv(const uint8_t* src, int w, int h, size_t stride, int16_t* dst) {
    std::vector<int> t0(w + 2), t1(w + 2);
    for (int y = 0; y < h; y++) {
        const uint8_t* r0 = src + (size_t)(y > 0 ? y - 1 : (h > 1 ? 1 : 0)) * stride;
        const uint8_t* r1 = src + (size_t)y * stride;
        const uint8_t* r2 = src + (size_t)(y < h - 1 ? y + 1 : (h > 1 ? h - 2 : 0)) * stride;
        for (int x = 0; x < w; x++) {
            t0[x + 1] = (r0[x] + r2[x]) * 3 + r1[x] * 10;
            t1[x + 1] = r2[x] - r0[x];
        }
        const int x0 = w > 1 ? 1 : 0, x1 = w > 1 ? w - 2 : 0;
        t0[0] = t0[x0 + 1]; t0[w + 1] = t0[x1 + 1];
        t1[0] = t1[x0 + 1]; t1[w + 1] = t1[x1 + 1];
        int16_t* d = dst + (size_t)y * w * 2;
        for (int x = 0; x < w; x++) {
            d[2 * x] = (int16_t)(t0[x + 2] - t0[x]);
            d[2 * x + 1] = (int16_t)((t1[x + 2] + t1[x]) * 3 + t1[x + 1] * 10);
        }
    }
}

struct LkLevel {
    int w, h;
    std::vector<uint8_t> I, J;
    std::vector<int16_t> dI;
};

static inline int lk_descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
static inline int img_at(const std::vector<uint8_t>& im, int w, int h, int x, int y) {   // REFLECT_101 border
    return im[(size_t)lk_reflect101(y, h) * w + lk_reflect101(x, w)];
}
static inline int der_at(const std::vector<int16_t>& d, int w, int h, int x, int y, int c) {   // zero border
    if (x < 0 || x >= w || y < 0 || y >= h) return 0;
    return d[((size_t)y * w + x) * 2 + c];
}

// cv::calcOpticalFlowPyrLK.  pts are (x, y) float pairs; nextPts is in/out (read when useInitialFlow != 0).
// Returns the number of pyramid levels - 1 actually used.
int orc_lk_track(const uint8_t* prevImg, const uint8_t* nextImg, int w, int h, size_t stride, const float* prevPts, float* nextPts, int n,
                 int win, int maxLevel, int maxIter, double eps, int useInitialFlow, float minEigThreshold, uint8_t* status, float* err) {
    // ---- pyramids (buildOpticalFlowPyramid)
    std::vector<LkLevel> pyr(1);
    pyr[0].w = w; pyr[0].h = h;
    pyr[0].I.resize((size_t)w * h); pyr[0].J.resize((size_t)w * h);
    for (int y = 0; y < h; y++) {
        memcpy(&pyr[0].I[(size_t)y * w], prevImg + (size_t)y * stride, w);
        memcpy(&pyr[0].J[(size_t)y * w], nextImg + (size_t)y * stride, w);
    }
    for (int l = 1; l <= maxLevel; l++) {
        const int pw = pyr[l - 1].w, ph = pyr[l - 1].h;
        const int dw = (pw + 1) / 2, dh = (ph + 1) / 2;
        if (dw <= win || dh <= win) break;
        pyr.emplace_back();
        LkLevel& L = pyr.back();
        L.w = dw; L.h = dh; L.I.resize((size_t)dw * dh); L.J.resize((size_t)dw * dh);
        orc_pyrdown_u8(pyr[l - 1].I.data(), pw, ph, pw, L.I.data(), dw, dh, dw);
        orc_pyrdown_u8(pyr[l - 1].J.data(), pw, ph, pw, L.J.data(), dw, dh, dw);
    }
    maxLevel = (int)pyr.size() - 1;
    for (auto& L : pyr) { L.dI.resize((size_t)L.w * L.h * 2); orc_scharr_deriv(L.I.data(), L.w, L.h, L.w, L.dI.data()); }

    maxIter = std::min(std::max(maxIter, 0), 100);
    double epsilon = std::min(std::max(eps, 0.), 10.);
    epsilon *= epsilon;
    for (int i = 0; i < n; i++) { status[i] = 1; if (err) err[i] = 0.f; }

    const float halfWin = (float)(win - 1) * 0.5f;
    const int W_BITS = 14;
    const float FLT_SCALE = 1.f / (1 << 20);
    std::vector<int> Iw((size_t)win * win), Ixw((size_t)win * win), Iyw((size_t)win * win);

    for (int level = maxLevel; level >= 0; level--) {
        const LkLevel& L = pyr[level];
        for (int p = 0; p < n; p++) {
            const float sc = (float)(1. / (1 << level));
            float px = prevPts[2 * p] * sc, py = prevPts[2 * p + 1] * sc;
            float nx, ny;
            if (level == maxLevel) {
                if (useInitialFlow) { nx = nextPts[2 * p] * sc; ny = nextPts[2 * p + 1] * sc; }
                else { nx = px; ny = py; }
            } else {
                nx = nextPts[2 * p] * 2.f; ny = nextPts[2 * p + 1] * 2.f;
            }
            nextPts[2 * p] = nx; nextPts[2 * p + 1] = ny;
            px -= halfWin; py -= halfWin;
            const int ipx = (int)floorf(px), ipy = (int)floorf(py);
            if (ipx < -win || ipx >= L.w || ipy < -win || ipy >= L.h) {
                if (level == 0) { status[p] = 0; if (err) err[p] = 0.f; }
                continue;
            }
            float a = px - (float)ipx, b = py - (float)ipy;
            int iw00 = (int)lrintf((1.f - a) * (1.f - b) * (1 << W_BITS));
            int iw01 = (int)lrintf(a * (1.f - b) * (1 << W_BITS));
            int iw10 = (int)lrintf((1.f - a) * b * (1 << W_BITS));
            int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
            int64_t sA11 = 0, sA12 = 0, sA22 = 0;
            for (int y = 0; y < win; y++)
                for (int x = 0; x < win; x++) {
                    const int X = ipx + x, Y = ipy + y;
                    const int ival = lk_descale(img_at(L.I, L.w, L.h, X, Y) * iw00 + img_at(L.I, L.w, L.h, X + 1, Y) * iw01 +
                                                img_at(L.I, L.w, L.h, X, Y + 1) * iw10 + img_at(L.I, L.w, L.h, X + 1, Y + 1) * iw11, W_BITS - 5);
                    const int ixval = lk_descale(der_at(L.dI, L.w, L.h, X, Y, 0) * iw00 + der_at(L.dI, L.w, L.h, X + 1, Y, 0) * iw01 +
                                                 der_at(L.dI, L.w, L.h, X, Y + 1, 0) * iw10 + der_at(L.dI, L.w, L.h, X + 1, Y + 1, 0) * iw11, W_BITS);
                    const int iyval = lk_descale(der_at(L.dI, L.w, L.h, X, Y, 1) * iw00 + der_at(L.dI, L.w, L.h, X + 1, Y, 1) * iw01 +
                                                 der_at(L.dI, L.w, L.h, X, Y + 1, 1) * iw10 + der_at(L.dI, L.w, L.h, X + 1, Y + 1, 1) * iw11, W_BITS);
                    Iw[(size_t)y * win + x] = (int16_t)ival; Ixw[(size_t)y * win + x] = (int16_t)ixval; Iyw[(size_t)y * win + x] = (int16_t)iyval;
                    sA11 += (int64_t)ixval * ixval; sA12 += (int64_t)ixval * iyval; sA22 += (int64_t)iyval * iyval;
                }
            const float A11 = (float)sA11 * FLT_SCALE, A12 = (float)sA12 * FLT_SCALE, A22 = (float)sA22 * FLT_SCALE;
            float D = A11 * A22 - A12 * A12;
            const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
            if (minEig < minEigThreshold || D < 1.1920929e-07f) {
                if (level == 0) status[p] = 0;
                continue;
            }
            D = 1.f / D;
            nx -= halfWin; ny -= halfWin;
            float pdx = 0.f, pdy = 0.f;
            for (int j = 0; j < maxIter; j++) {
                const int inx = (int)floorf(nx), iny = (int)floorf(ny);
                if (inx < -win || inx >= L.w || iny < -win || iny >= L.h) {
                    if (level == 0) status[p] = 0;
                    break;
                }
                a = nx - (float)inx; b = ny - (float)iny;
                iw00 = (int)lrintf((1.f - a) * (1.f - b) * (1 << W_BITS));
                iw01 = (int)lrintf(a * (1.f - b) * (1 << W_BITS));
                iw10 = (int)lrintf((1.f - a) * b * (1 << W_BITS));
                iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
                int64_t sb1 = 0, sb2 = 0;
                for (int y = 0; y < win; y++)
                    for (int x = 0; x < win; x++) {
                        const int X = inx + x, Y = iny + y;
                        const int diff = lk_descale(img_at(L.J, L.w, L.h, X, Y) * iw00 + img_at(L.J, L.w, L.h, X + 1, Y) * iw01 +
                                                    img_at(L.J, L.w, L.h, X, Y + 1) * iw10 + img_at(L.J, L.w, L.h, X + 1, Y + 1) * iw11, W_BITS - 5) -
                                         Iw[(size_t)y * win + x];
                        sb1 += (int64_t)diff * Ixw[(size_t)y * win + x]; sb2 += (int64_t)diff * Iyw[(size_t)y * win + x];
                    }
                const float b1 = (float)sb1 * FLT_SCALE, b2 = (float)sb2 * FLT_SCALE;
                const float dx = (A12 * b2 - A22 * b1) * D, dy = (A12 * b1 - A11 * b2) * D;
                nx += dx; ny += dy;
                nextPts[2 * p] = nx + halfWin; nextPts[2 * p + 1] = ny + halfWin;
                if ((double)dx * dx + (double)dy * dy <= epsilon) break;
                if (j > 0 && (double)fabsf(dx + pdx) < 0.01 && (double)fabsf(dy + pdy) < 0.01) {   // 0.01 is a double literal in OpenCV
                    nextPts[2 * p] -= dx * 0.5f; nextPts[2 * p + 1] -= dy * 0.5f;
                    break;
                }
                pdx = dx; pdy = dy;
            }
            if (status[p] && err && level == 0) {
                const float ex = nextPts[2 * p] - halfWin, ey = nextPts[2 * p + 1] - halfWin;
                const int iex = (int)floorf(ex), iey = (int)floorf(ey);
                if (iex < -win || iex >= L.w || iey < -win || iey >= L.h) { status[p] = 0; continue; }
                const float aa = ex - (float)iex, bb = ey - (float)iey;
                iw00 = (int)lrintf((1.f - aa) * (1.f - bb) * (1 << W_BITS));
                iw01 = (int)lrintf(aa * (1.f - bb) * (1 << W_BITS));
                iw10 = (int)lrintf((1.f - aa) * bb * (1 << W_BITS));
                iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
                int64_t serr = 0;
                for (int y = 0; y < win; y++)
                    for (int x = 0; x < win; x++) {
                        const int X = iex + x, Y = iey + y;
                        const int diff = lk_descale(img_at(L.J, L.w, L.h, X, Y) * iw00 + img_at(L.J, L.w, L.h, X + 1, Y) * iw01 +
                                                    img_at(L.J, L.w, L.h, X, Y + 1) * iw10 + img_at(L.J, L.w, L.h, X + 1, Y + 1) * iw11, W_BITS - 5) -
                                         Iw[(size_t)y * win + x];
                        serr += diff < 0 ? -diff : diff;
                    }
                err[p] = (float)serr * 1.f / (float)(32 * win * win);
            }
        }
    }
    return maxLevel;
}

/* ELK_Tracker::refineTrackedPts (src/Event/KLT_Tracker.cpp:105-155) followed, when first_octave_only is set, by
 * refineFirstOctaveLevel (:157-183) as trackAndMatchCurrImageInit chains them (:236-242).  The reference's vectors are passed flat:
 * matches12 / cnt_matches are in-out (resize(n, -1) / resize(n, 1) keep what the caller left in them, :113-133), px_disp receives the
 * push_back sequence (:147), counts2 = {nMatches returned, px_disp entries}. */
void orc_lk_refine(const float* curr_xy, const uint8_t* status, const orc_keypoint* ref_kps, int n, int img_w, int img_h, int first_octave_only,
                   orc_keypoint* tracked, int32_t* matches12, int32_t* cnt_matches, float* px_disp, int32_t* counts2) {
    unsigned nMatches = 0;
    int nd = 0;
    for (int i = 0; i < n; i++) {
        const float cx = curr_xy[2 * i], cy = curr_xy[2 * i + 1];
        const orc_keypoint pre = ref_kps[i];
        orc_keypoint k = pre;                       /* KeyPoint(currPt, prePt.size, prePt.angle, prePt.response, prePt.octave, prePt.class_id) :138 */
        k.x = cx; k.y = cy;
        tracked[i] = k;
        const bool inImage = cx >= 0 && cx < (float)img_w && cy >= 0 && cy < (float)img_h;   /* isInImage :99-102 */
        if (status[i] == 1 && inImage) {            /* :140 */
            cnt_matches[i]++;
            matches12[i] = i;
            nMatches++;
            const float dx = cx - pre.x, dy = cy - pre.y;
            px_disp[nd++] = sqrtf(dx * dx + dy * dy);   /* sqrtf(powf(dx, 2) + powf(dy, 2)) :147; powf(x, 2) is x * x (exact square, one rounding) */
        }
    }
    if (first_octave_only) {                        /* :157-183, nTrackedPts == n */
        for (int i = 0; i < n; i++) {
            const int cur = matches12[i];
            if (cur >= 0 && cur < n && ref_kps[i].octave > 0) { matches12[i] = -1; cnt_matches[i]--; nMatches--; }
        }
    }
    counts2[0] = (int32_t)nMatches; counts2[1] = nd;
}

}  // extern "C"
