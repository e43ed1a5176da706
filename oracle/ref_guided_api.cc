// oracle/ref_guided_api.cc — C entry points of oracle/_ref/libref.so for the tracking-thread matchers.  TEST INFRASTRUCTURE ONLY.
//
// The function BODIES are the reference's own text, cut out of the sources at build time (oracle/Makefile -> _ref/gen_matcher.inc):
//     src/ORBmatcher.cc   :35-37   TH_HIGH / TH_LOW / HISTO_LENGTH
//                         :44-218  SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)
//                         :221-227 RadiusByViewingCos          :276-478  SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)
//                         :714-831 SearchForInitialization     :1969-2187 SearchByProjection(CurrentFrame, LastFrame, th, bMono)
//                         :2189-2312 SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist)   :2314-2355 ComputeThreeMaxima
//     src/Frame.cc        :431-460 AssignFeaturesToGrid   :709-777 GetFeaturesInArea   :783-793 PosInGrid
//     src/CameraModels/Pinhole.cpp :36-40 project(const cv::Mat&)   (in _ref/gen_event_deps.inc with the other camera functions)
// compiled against the stand-in Frame / KeyFrame / MapPoint of ref_mock/Frame.h.  This file only builds those objects from flat
// arrays (the layout of the oracle's orc_search_* functions) and reads the results back.  Poses are the identity, so that
// x3Dc = Rcw * x3Dw + tcw reproduces the caller's camera-frame point exactly; the LastFrame pose carries only the z offset that
// selects the forward / backward window of the stereo case (ORBmatcher.cc:1989-1990).
#include <cstring>
#include <set>
#include <vector>

#include "ORBmatcher.h"   // ref_mock
#include "Pinhole.h"      // ref_mock
#include "oracle.h"

using namespace std;      // the cut text is written for `using namespace std` (ORBmatcher.cc:30, Frame.cc)

namespace ORB_SLAM3 {
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
#include "gen_matcher.inc"
}  // namespace ORB_SLAM3

namespace {
using namespace ORB_SLAM3;

typedef Pinhole PinholeMat;   // Pinhole::project(const cv::Mat&) (Pinhole.cpp:36-40) is compiled in ref_event_api.cc with the other camera cuts

void setBounds(const float* b) {
    Frame::mnMinX = b[0]; Frame::mnMinY = b[1]; Frame::mnMaxX = b[2]; Frame::mnMaxY = b[3];
    // Frame.cc:165-166: static_cast<float>(FRAME_GRID_COLS) / (mnMaxX - mnMinX)
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);
}
std::vector<cv::KeyPoint> toKps(const orc_keypoint* k, int n) {
    std::vector<cv::KeyPoint> v((size_t)n);
    if (n) std::memcpy(v.data(), k, (size_t)n * sizeof(orc_keypoint));
    return v;
}
cv::Mat descMat(const uint8_t* d, int n) {
    cv::Mat m(std::max(n, 1), 32, CV_8UC1);
    if (n && d) std::memcpy(m.data, d, (size_t)n * 32);
    return m;
}
cv::Mat eye4() { cv::Mat m = cv::Mat::zeros(4, 4, CV_32FC1); for (int i = 0; i < 4; i++) m.at<float>(i, i) = 1.f; return m; }

// a frame of n keypoints (monocular layout: Nleft == -1), grid assigned
void fillFrame(Frame& F, const orc_keypoint* kps, const uint8_t* desc, int n, const float* scale, int nlevels, const float* u_right) {
    F.N = n; F.Nleft = -1;
    F.mvKeysUn = toKps(kps, n); F.mvKeys = F.mvKeysUn;
    F.mDescriptors = descMat(desc, n);
    F.mvpMapPoints.assign((size_t)n, nullptr);
    F.mvbOutlier.assign((size_t)n, false);
    F.mvuRight.assign((size_t)n, -1.f);
    if (u_right) for (int i = 0; i < n; i++) F.mvuRight[i] = u_right[i];
    F.mvScaleFactors.assign(scale, scale + nlevels);
    F.mTcw = eye4();
    F.AssignFeaturesToGrid();
}
void fillFeatVec(DBoW2::FeatureVector& fv, const uint32_t* nodes, const int32_t* start, const uint32_t* feats, int nn) {
    for (int q = 0; q < nn; q++) fv[nodes[q]] = std::vector<unsigned int>(feats + start[q], feats + start[q + 1]);
}
}  // namespace

extern "C" {

/* Frame::AssignFeaturesToGrid + GetFeaturesInArea (same contract as orc_frame_grid / orc_features_in_area) */
int ref_features_in_area(const orc_keypoint* kps, int n, const float* bounds4, float x, float y, float r, int min_level, int max_level,
                         int* out, int cap) {
    setBounds(bounds4);
    Frame F;
    const float one = 1.f;
    fillFrame(F, kps, nullptr, n, &one, 1, nullptr);
    const std::vector<size_t> v = F.GetFeaturesInArea(x, y, r, min_level, max_level);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int)v[i];
    return (int)v.size();
}

/* ORBmatcher::SearchForInitialization (ORBmatcher.cc:714-831) */
int ref_search_for_initialization(const orc_keypoint* kps1, const uint8_t* desc1, int n1, const orc_keypoint* kps2, const uint8_t* desc2, int n2,
                                  const float* bounds4, float* prev_xy, int window_size, float nnratio, int check_ori, int32_t* matches12) {
    setBounds(bounds4);
    const float one = 1.f;
    Frame F1, F2;
    fillFrame(F1, kps1, desc1, n1, &one, 1, nullptr);
    fillFrame(F2, kps2, desc2, n2, &one, 1, nullptr);
    std::vector<cv::Point2f> prev((size_t)n1);
    for (int i = 0; i < n1; i++) prev[i] = cv::Point2f(prev_xy[2 * i], prev_xy[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int nm = matcher.SearchForInitialization(F1, F2, prev, m12, window_size);
    for (int i = 0; i < n1; i++) { matches12[i] = m12[i]; prev_xy[2 * i] = prev[i].x; prev_xy[2 * i + 1] = prev[i].y; }
    return nm;
}

/* ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1969-2187), rectified-stereo aware:
   level_mode 0 = neither forward nor backward (also the monocular call), 1 = bForward, 2 = bBackward (:1989-1990, :2024-2029);
   u_right2 (n2 floats or NULL) = CurrentFrame.mvuRight, mbf = CurrentFrame.mbf (:2049-2055) */
int ref_search_by_projection(const float* x3Dc, const uint8_t* valid1, const int32_t* obs1, const orc_keypoint* kps1, const uint8_t* descMP, int n1,
                             const orc_keypoint* kps2, const uint8_t* desc2, int n2, const float* bounds4, const float* K4,
                             const float* scale_factors, int nlevels, float th, int check_ori, int level_mode, float mbf, const float* u_right2,
                             int32_t* match_cur) {
    setBounds(bounds4);
    PinholeMat cam(std::vector<float>(K4, K4 + 4));
    Frame cur, last;
    fillFrame(cur, kps2, desc2, n2, scale_factors, nlevels, u_right2);
    cur.mpCamera = &cam; cur.mbf = mbf; cur.mb = 1.0f;
    fillFrame(last, kps1, nullptr, n1, scale_factors, nlevels, nullptr);
    last.mTcw.at<float>(2, 3) = level_mode == 1 ? 2.0f : (level_mode == 2 ? -2.0f : 0.0f);   // tlc = tlw when the current pose is the identity
    std::vector<MapPoint> mps((size_t)n1);
    for (int i = 0; i < n1; i++) {
        mps[i].id = i; mps[i].nObs = obs1[i];
        mps[i].pos = cv::Mat(3, 1, CV_32FC1);
        for (int k = 0; k < 3; k++) mps[i].pos.at<float>(k, 0) = x3Dc[3 * i + k];
        mps[i].desc = cv::Mat(1, 32, CV_8UC1); std::memcpy(mps[i].desc.data, descMP + (size_t)i * 32, 32);
        if (valid1[i]) last.mvpMapPoints[i] = &mps[i];
    }
    ORBmatcher matcher(0.9f, check_ori != 0);
    const bool bMono = level_mode == 0 && !u_right2;
    const int nm = matcher.SearchByProjection(cur, last, th, bMono);
    for (int i = 0; i < n2; i++) match_cur[i] = cur.mvpMapPoints[i] ? cur.mvpMapPoints[i]->id : -1;
    return nm;
}

/* ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints) (ORBmatcher.cc:44-218), rectified-stereo aware:
   proj_xr (n1 floats or NULL) = mTrackProjXR, u_right2 (n2 floats or NULL) = F.mvuRight (:91-96) */
int ref_search_by_projection_map_points(const orc_track_point* pts, const float* proj_xr, const uint8_t* descMP, int n1, const orc_keypoint* kps2,
                                        const uint8_t* desc2, const uint8_t* held2, const float* u_right2, int n2, const float* bounds4,
                                        const float* scale_factors, int nlevels, float th, int far_points, float th_far, float nnratio,
                                        int32_t* match_cur) {
    setBounds(bounds4);
    Frame F;
    fillFrame(F, kps2, desc2, n2, scale_factors, nlevels, u_right2);
    MapPoint heldPoint;                 // a point with observations sitting in the frame on entry (:87-89)
    heldPoint.id = -1; heldPoint.nObs = 1;
    for (int i = 0; i < n2; i++) if (held2 && held2[i]) F.mvpMapPoints[i] = &heldPoint;
    std::vector<MapPoint> mps((size_t)n1);
    std::vector<MapPoint*> vp((size_t)n1);
    for (int i = 0; i < n1; i++) {
        MapPoint& m = mps[i];
        m.id = i; m.nObs = pts[i].observations; m.bad = pts[i].bad != 0;
        m.mTrackProjX = pts[i].proj_x; m.mTrackProjY = pts[i].proj_y; m.mTrackViewCos = pts[i].view_cos; m.mTrackDepth = pts[i].depth;
        m.mnTrackScaleLevel = pts[i].scale_level; m.mbTrackInView = pts[i].in_view != 0;
        m.mTrackProjXR = proj_xr ? proj_xr[i] : 0.f;
        m.desc = cv::Mat(1, 32, CV_8UC1); std::memcpy(m.desc.data, descMP + (size_t)i * 32, 32);
        vp[i] = &m;
    }
    ORBmatcher matcher(nnratio, true);
    const int nm = matcher.SearchByProjection(F, vp, th, far_points != 0, th_far);
    for (int i = 0; i < n2; i++) match_cur[i] = (F.mvpMapPoints[i] && F.mvpMapPoints[i]->id >= 0) ? F.mvpMapPoints[i]->id : -1;
    return nm;
}

/* ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:2189-2312; Tracking::Relocalization).
   valid1[i] = the keyframe has a map point at i that is not bad and not in sAlreadyFound; level1[i] = PredictScale(dist3D, &CurrentFrame)
   of that point (formed by the caller with the distance-invariance gate :2231-2232); held2[i2] != 0: the frame already holds a point */
int ref_search_by_projection_reloc(const float* x3Dc, const uint8_t* valid1, const int32_t* level1, const orc_keypoint* kps1, const uint8_t* descMP,
                                   int n1, const orc_keypoint* kps2, const uint8_t* desc2, const uint8_t* held2, int n2, const float* bounds4,
                                   const float* K4, const float* scale_factors, int nlevels, float th, int orb_dist, int check_ori,
                                   int32_t* match_cur) {
    setBounds(bounds4);
    PinholeMat cam(std::vector<float>(K4, K4 + 4));
    Frame cur;
    fillFrame(cur, kps2, desc2, n2, scale_factors, nlevels, nullptr);
    cur.mpCamera = &cam;
    MapPoint heldPoint;
    heldPoint.id = -1;
    for (int i = 0; i < n2; i++) if (held2 && held2[i]) cur.mvpMapPoints[i] = &heldPoint;
    KeyFrame kf;
    kf.mvKeysUn = toKps(kps1, n1);
    kf.mvpMapPoints.assign((size_t)n1, nullptr);
    std::vector<MapPoint> mps((size_t)n1);
    for (int i = 0; i < n1; i++) {
        mps[i].id = i; mps[i].predictedLevel = level1[i];
        mps[i].pos = cv::Mat(3, 1, CV_32FC1);
        for (int k = 0; k < 3; k++) mps[i].pos.at<float>(k, 0) = x3Dc[3 * i + k];
        mps[i].desc = cv::Mat(1, 32, CV_8UC1); std::memcpy(mps[i].desc.data, descMP + (size_t)i * 32, 32);
        if (valid1[i]) kf.mvpMapPoints[i] = &mps[i];
    }
    ORBmatcher matcher(0.9f, check_ori != 0);
    const std::set<MapPoint*> none;
    const int nm = matcher.SearchByProjection(cur, &kf, none, th, orb_dist);
    for (int i = 0; i < n2; i++) match_cur[i] = (cur.mvpMapPoints[i] && cur.mvpMapPoints[i]->id >= 0) ? cur.mvpMapPoints[i]->id : -1;
    return nm;
}

/* ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) (ORBmatcher.cc:276-478), monocular; same contract as orc_search_by_bow */
int ref_search_by_bow(const orc_keypoint* kps_kf, const uint8_t* desc_kf, const uint8_t* valid_kf, const uint32_t* kf_nodes, const int32_t* kf_start,
                      const uint32_t* kf_feats, int nkf, int n1, const orc_keypoint* kps_f, const uint8_t* desc_f, int n2, const uint32_t* f_nodes,
                      const int32_t* f_start, const uint32_t* f_feats, int nf, float nnratio, int check_ori, int32_t* match_f) {
    const float b[4] = {0.f, 0.f, 1.f, 1.f}, one = 1.f;
    setBounds(b);
    Frame F;
    fillFrame(F, kps_f, desc_f, n2, &one, 1, nullptr);
    fillFeatVec(F.mFeatVec, f_nodes, f_start, f_feats, nf);
    KeyFrame kf;
    kf.mvKeysUn = toKps(kps_kf, n1);
    kf.mDescriptors = descMat(desc_kf, n1);
    fillFeatVec(kf.mFeatVec, kf_nodes, kf_start, kf_feats, nkf);
    std::vector<MapPoint> mps((size_t)n1);
    kf.mvpMapPoints.assign((size_t)n1, nullptr);
    for (int i = 0; i < n1; i++) { mps[i].id = i; if (valid_kf[i]) kf.mvpMapPoints[i] = &mps[i]; }
    std::vector<MapPoint*> out;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int nm = matcher.SearchByBoW(&kf, F, out);
    for (int i = 0; i < n2; i++) match_f[i] = out[i] ? out[i]->id : -1;
    return nm;
}

/* ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) (ORBmatcher.cc:833-990); same contract as orc_search_by_bow_kf */
int ref_search_by_bow_kf(const orc_keypoint* kps1, const uint8_t* desc1, const uint8_t* valid1, const uint32_t* nodes1, const int32_t* start1,
                         const uint32_t* feats1, int nn1, int n1, const orc_keypoint* kps2, const uint8_t* desc2, const uint8_t* valid2,
                         const uint32_t* nodes2, const int32_t* start2, const uint32_t* feats2, int nn2, int n2, float nnratio, int check_ori,
                         int32_t* match12) {
    KeyFrame k1, k2;
    k1.mvKeysUn = toKps(kps1, n1); k1.mDescriptors = descMat(desc1, n1); fillFeatVec(k1.mFeatVec, nodes1, start1, feats1, nn1);
    k2.mvKeysUn = toKps(kps2, n2); k2.mDescriptors = descMat(desc2, n2); fillFeatVec(k2.mFeatVec, nodes2, start2, feats2, nn2);
    std::vector<MapPoint> m1((size_t)n1), m2((size_t)n2);
    k1.mvpMapPoints.assign((size_t)n1, nullptr); k2.mvpMapPoints.assign((size_t)n2, nullptr);
    for (int i = 0; i < n1; i++) { m1[i].id = i; if (valid1[i]) k1.mvpMapPoints[i] = &m1[i]; }
    for (int i = 0; i < n2; i++) { m2[i].id = i; if (valid2[i]) k2.mvpMapPoints[i] = &m2[i]; }
    std::vector<MapPoint*> out;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int nm = matcher.SearchByBoW(&k1, &k2, out);
    for (int i = 0; i < n1; i++) match12[i] = out[i] ? out[i]->id : -1;
    return nm;
}

}  // extern "C"