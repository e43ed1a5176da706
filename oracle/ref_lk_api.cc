// oracle/ref_lk_api.cc — C entry point of oracle/_ref/libref.so for ELK_Tracker's post-LK bookkeeping.  TEST INFRASTRUCTURE ONLY.
//
// The function BODIES are the reference's own text, cut out of src/Event/KLT_Tracker.cpp at build time (oracle/Makefile ->
// _ref/gen_lk.inc):  isInImage :99-102, ELK_Tracker::refineTrackedPts :105-155, ELK_Tracker::refineFirstOctaveLevel (member) :157-183.
// They are compiled against the stand-in class below, which declares exactly the members those bodies touch (the real header pulls
// in opencv2/video/tracking.hpp).  This file only fills the members from flat arrays and reads the vectors back.
#include <cassert>
#include <cmath>
#include <cstring>
#include <iostream>
#include <vector>

#include <opencv2/core/core.hpp>   // ref_mock
#include "oracle.h"

namespace EORB_SLAM {
class ELK_Tracker {   // include/Event/KLT_Tracker.h:27-99, the members used by the cut bodies
public:
    unsigned refineTrackedPts(const std::vector<cv::Point2f>& currPts, const std::vector<uchar>& status, const std::vector<float>& err,
                              std::vector<cv::KeyPoint>& p1, std::vector<int>& vMatches12, std::vector<int>& vCntMatches,
                              std::vector<float>& vPxDisp);
    unsigned refineFirstOctaveLevel(std::vector<cv::KeyPoint>& trackedKPts, std::vector<int>& vMatches12, unsigned& nMatches,
                                    std::vector<int>& vCntMatches);
    cv::Mat mRefFrame;
    std::vector<cv::KeyPoint> mRefKPoints;
    std::vector<cv::Point2f> mRefPoints;
};
using namespace cv;    // KLT_Tracker.cpp:8-9
using namespace std;
#include "gen_lk.inc"
}  // namespace EORB_SLAM

extern "C" {

/* same contract as orc_lk_refine (oracle/lk_oracle.cc); matches12 / cnt_matches in-out of size n, or have_prev = 0 for empty vectors */
int ref_lk_refine(const float* curr_xy, const uint8_t* status, const orc_keypoint* ref_kps, int n, int img_w, int img_h, int first_octave_only,
                  int have_prev, orc_keypoint* tracked, int32_t* matches12, int32_t* cnt_matches, float* px_disp, int32_t* counts2) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(orc_keypoint), "cv::KeyPoint layout");
    EORB_SLAM::ELK_Tracker T;
    T.mRefFrame = cv::Mat(img_h, img_w, CV_8UC1);
    T.mRefKPoints.resize((size_t)n); T.mRefPoints.resize((size_t)n);
    std::memcpy(T.mRefKPoints.data(), ref_kps, (size_t)n * sizeof(orc_keypoint));
    for (int i = 0; i < n; i++) T.mRefPoints[(size_t)i] = T.mRefKPoints[(size_t)i].pt;
    std::vector<cv::Point2f> curr((size_t)n);
    for (int i = 0; i < n; i++) curr[(size_t)i] = cv::Point2f(curr_xy[2 * i], curr_xy[2 * i + 1]);
    std::vector<uchar> st(status, status + n);
    std::vector<float> err((size_t)n, 0.f), disp;
    std::vector<cv::KeyPoint> p1;
    std::vector<int> m12, cnt;
    if (have_prev) { m12.assign(matches12, matches12 + n); cnt.assign(cnt_matches, cnt_matches + n); }
    unsigned nm = T.refineTrackedPts(curr, st, err, p1, m12, cnt, disp);
    if (first_octave_only) nm = T.refineFirstOctaveLevel(p1, m12, nm, cnt);
    if ((int)p1.size() != n || (int)m12.size() != n || (int)cnt.size() != n) return -1;
    std::memcpy(tracked, p1.data(), (size_t)n * sizeof(orc_keypoint));
    for (int i = 0; i < n; i++) { matches12[i] = m12[(size_t)i]; cnt_matches[i] = cnt[(size_t)i]; }
    for (size_t i = 0; i < disp.size(); i++) px_disp[i] = disp[i];
    counts2[0] = (int32_t)nm; counts2[1] = (int32_t)disp.size();
    return 0;
}

}  // extern "C"
