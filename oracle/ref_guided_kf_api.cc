// oracle/ref_guided_kf_api.cc — C entry points of oracle/_ref/libref.so for the KEYFRAME-side searches of local mapping and loop
// closing.  TEST INFRASTRUCTURE ONLY.
//
// The function BODIES are the reference's own text, cut out of the sources at build time (oracle/Makefile -> _ref/gen_matcher_kf.inc):
//     src/ORBmatcher.cc  :480-593   SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th, ratioHamming)
//                        :595-712   SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpPointsKFs, vpMatched, vpMatchedKF, th, ratioHamming)
//                        :1407-1617 Fuse(KeyFrame*, vpMapPoints, th, bRight)     :1619-1741 Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint)
//                        :1743-1967 SearchBySim3
//     src/KeyFrame.cc    :873-917   GetFeaturesInArea     :919-922 IsInImage
// compiled against the stand-in KeyFrame / MapPoint of ref_mock/Frame.h.  This file builds those objects from flat arrays and reads
// the results back.  Poses are the identity (Scw = I, s12 = 1, R12 = I, t12 = 0), so that the camera-frame point equals the world point
// the caller passes and the small cv::Mat algebra of the functions is exact whatever formula OpenCV uses; what the pin covers is the
// gating, the window lookup, the candidate loop and the acceptance rules.  The keyframe's grid is the Frame's (KeyFrame.cc:73-110 copies
// F.mGrid), assigned with the Frame's float bounds, while the window lookup uses the keyframe's own int bounds (include/KeyFrame.h:529).
#include <climits>
#include <cstring>
#include <set>
#include <vector>

#include "ORBmatcher.h"   // ref_mock
#include "Pinhole.h"      // ref_mock
#include "oracle.h"

using namespace std;

namespace ORB_SLAM3 {
#include "gen_matcher_kf.inc"
}  // namespace ORB_SLAM3

namespace {
using namespace ORB_SLAM3;

cv::Mat eye(int n) { cv::Mat m = cv::Mat::zeros(n, n, CV_32FC1); for (int i = 0; i < n; i++) m.at<float>(i, i) = 1.f; return m; }
cv::Mat vec3(const float* p) { cv::Mat m(3, 1, CV_32FC1); for (int k = 0; k < 3; k++) m.at<float>(k, 0) = p[k]; return m; }

// a keyframe of n keypoints made from a Frame (grid assigned with the frame's float bounds, KeyFrame.cc:73-110)
void fillKeyFrame(KeyFrame& kf, Pinhole* cam, const orc_keypoint* kps, const uint8_t* desc, int n, const float* bounds4, const float* K4,
                  const float* scale, const float* invSigma2, int nlevels, const float* u_right) {
    Frame::mnMinX = bounds4[0]; Frame::mnMinY = bounds4[1]; Frame::mnMaxX = bounds4[2]; Frame::mnMaxY = bounds4[3];
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);
    Frame F;
    F.N = n; F.Nleft = -1;
    F.mvKeysUn.resize((size_t)n);
    if (n) std::memcpy(F.mvKeysUn.data(), kps, (size_t)n * sizeof(orc_keypoint));
    F.AssignFeaturesToGrid();
    kf.N = n; kf.NLeft = -1;
    kf.mvKeysUn = F.mvKeysUn;
    kf.mDescriptors = cv::Mat(std::max(n, 1), 32, CV_8UC1);
    if (n) std::memcpy(kf.mDescriptors.data, desc, (size_t)n * 32);
    kf.mvpMapPoints.assign((size_t)n, nullptr);
    kf.mvuRight.assign((size_t)n, -1.f);
    if (u_right) for (int i = 0; i < n; i++) kf.mvuRight[i] = u_right[i];
    kf.mvScaleFactors.assign(scale, scale + nlevels);
    if (invSigma2) kf.mvInvLevelSigma2.assign(invSigma2, invSigma2 + nlevels); else kf.mvInvLevelSigma2.assign((size_t)nlevels, 1.f);
    kf.fx = K4[0]; kf.fy = K4[1]; kf.cx = K4[2]; kf.cy = K4[3];
    kf.mpCamera = cam;
    kf.mnMinX = (int)Frame::mnMinX; kf.mnMinY = (int)Frame::mnMinY; kf.mnMaxX = (int)Frame::mnMaxX; kf.mnMaxY = (int)Frame::mnMaxY;   // KeyFrame.cc:82
    kf.mfGridElementWidthInv = Frame::mfGridElementWidthInv; kf.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
    kf.mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<std::size_t> >(FRAME_GRID_ROWS));
    for (int i = 0; i < FRAME_GRID_COLS; i++) for (int j = 0; j < FRAME_GRID_ROWS; j++) kf.mGrid[i][j] = F.mGrid[i][j];
    kf.Rcw = eye(3); kf.tcw = cv::Mat::zeros(3, 1, CV_32FC1); kf.Ow = cv::Mat::zeros(3, 1, CV_32FC1);
}

// map points from flat arrays: pt[i] = {pos[3], normal[3], minDist, maxDist}, level, flags bit 0 = bad, bit 1 = IsInKeyFrame
void fillPoints(std::vector<MapPoint>& mps, const float* pt8, const int32_t* level, const int32_t* obs, const uint8_t* flags, const uint8_t* desc, int n) {
    mps.resize((size_t)n);
    for (int i = 0; i < n; i++) {
        MapPoint& m = mps[i];
        m.id = i;
        m.pos = vec3(pt8 + 8 * i); m.normal = vec3(pt8 + 8 * i + 3);
        m.minDist = pt8[8 * i + 6]; m.maxDist = pt8[8 * i + 7];
        m.predictedLevel = level[i];
        m.nObs = obs ? obs[i] : 1;
        m.bad = flags && (flags[i] & 1); m.inKF = flags && (flags[i] & 2);
        m.desc = cv::Mat(1, 32, CV_8UC1); std::memcpy(m.desc.data, desc + (size_t)i * 32, 32);
    }
}
}  // namespace

extern "C" {

/* ORBmatcher::SearchByProjection(KeyFrame*, Scw = I, vpPoints, [vpPointsKFs,] vpMatched, [vpMatchedKF,] th, ratioHamming).
   overload 0 = :480-593 (camera projection), 1 = :595-712 (fx * (x * invz) + cx).  held2[i2]: -1 = vpMatched[i2] empty on entry,
   -2 = holds a point that is not in vpPoints, k >= 0 = holds vpPoints[k] (which is then "already found", :496-497, :509).
   match2[i2] = index of the point vpMatched[i2] holds on return when this call put it there, else -1. */
int ref_search_by_projection_kf(int overload, const float* pt8, const int32_t* level1, const uint8_t* flags1, const uint8_t* descMP, int n1,
                                const orc_keypoint* kps2, const uint8_t* desc2, const int32_t* held2, int n2, const float* bounds4, const float* K4,
                                const float* scale_factors, int nlevels, int th, float ratio_hamming, int32_t* match2) {
    Pinhole cam(std::vector<float>(K4, K4 + 4));
    KeyFrame kf;
    fillKeyFrame(kf, &cam, kps2, desc2, n2, bounds4, K4, scale_factors, nullptr, nlevels, nullptr);
    std::vector<MapPoint> mps;
    fillPoints(mps, pt8, level1, nullptr, flags1, descMP, n1);
    std::vector<MapPoint*> vp((size_t)n1);
    for (int i = 0; i < n1; i++) vp[i] = &mps[i];
    MapPoint foreign; foreign.id = -2;
    std::vector<MapPoint*> vpMatched((size_t)n2, nullptr);
    for (int i = 0; i < n2; i++) vpMatched[i] = held2[i] == -1 ? nullptr : (held2[i] == -2 ? &foreign : &mps[held2[i]]);
    const std::vector<MapPoint*> onEntry = vpMatched;
    ORBmatcher matcher(0.75f, true);
    int nm;
    if (overload == 0) nm = matcher.SearchByProjection(&kf, eye(4), vp, vpMatched, th, ratio_hamming);
    else {
        KeyFrame other;
        std::vector<KeyFrame*> vpKFs((size_t)n1, &other), vpMatchedKF((size_t)n2, nullptr);
        nm = matcher.SearchByProjection(&kf, eye(4), vp, vpKFs, vpMatched, vpMatchedKF, th, ratio_hamming);
        for (int i = 0; i < n2; i++) if ((vpMatchedKF[i] != nullptr) != (vpMatched[i] != onEntry[i])) return -1000;   // the two tables move together
    }
    for (int i = 0; i < n2; i++) match2[i] = (vpMatched[i] != onEntry[i] && vpMatched[i]) ? vpMatched[i]->id : -1;
    return nm;
}

/* ORBmatcher::Fuse.  overload 0 = Fuse(pKF, vpMapPoints, th, bRight = false) (:1407-1617), 1 = Fuse(pKF, Scw = I, vpPoints, th,
   vpReplacePoint) (:1619-1741).  present1[i] = 0: vpMapPoints[i] is NULL (overload 0).  occupied2[i2] != 0: the keyframe holds a map
   point at i2 on entry (observations occ_obs2[i2], bad when occupied2 == 2; overload 1: these are also GetMapPoints(), none of them in
   vpPoints).  Events, in the order the function produces them: {0, i, idx} = AddObservation / AddMapPoint of point i at idx;
   {1, a, b} = a.Replace(b) (overload 0; ids: point i -> i, entry occupant of slot s -> -(s + 2)); {2, i, b} = vpReplacePoint[i] = b
   (overload 1).  Returns nFused; *n_events = number of events written (at most cap). */
int ref_fuse(int overload, const float* pt8, const int32_t* level1, const int32_t* obs1, const uint8_t* flags1, const uint8_t* present1,
             const uint8_t* descMP, int n1, const orc_keypoint* kps2, const uint8_t* desc2, const uint8_t* occupied2, const int32_t* occ_obs2,
             const float* u_right2, int n2, const float* bounds4, const float* K4, float mbf, const float* scale_factors,
             const float* inv_level_sigma2, int nlevels, float th, int32_t* events3, int cap, int* n_events) {
    Pinhole cam(std::vector<float>(K4, K4 + 4));
    KeyFrame kf;
    fillKeyFrame(kf, &cam, kps2, desc2, n2, bounds4, K4, scale_factors, inv_level_sigma2, nlevels, u_right2);
    kf.mbf = mbf;
    std::vector<MapPoint> mps, occ((size_t)n2);
    fillPoints(mps, pt8, level1, obs1, flags1, descMP, n1);
    for (int i = 0; i < n2; i++) {
        occ[i].id = -(i + 2); occ[i].nObs = occ_obs2 ? occ_obs2[i] : 1; occ[i].bad = occupied2 && occupied2[i] == 2;
        if (occupied2 && occupied2[i]) kf.mvpMapPoints[i] = &occ[i];
    }
    std::vector<MapPoint*> vp((size_t)n1);
    for (int i = 0; i < n1; i++) vp[i] = (!present1 || present1[i]) ? &mps[i] : nullptr;
    MapPoint::replaceLog().clear();
    ORBmatcher matcher(0.75f, true);
    std::vector<MapPoint*> vpReplace((size_t)n1, nullptr);
    // the map-graph calls are logged in call order by running point after point would change nothing: the stand-ins record them
    const int nFused = overload == 0 ? matcher.Fuse(&kf, vp, th, false) : matcher.Fuse(&kf, eye(4), vp, th, vpReplace);
    // events in point order: a point produces at most one of {add, replace}; replaces are logged in call order = point order
    int ne = 0;
    auto put = [&](int t, int a, int b) { if (ne < cap) { events3[3 * ne] = t; events3[3 * ne + 1] = a; events3[3 * ne + 2] = b; } ne++; };
    size_t rl = 0;
    const std::vector<std::pair<MapPoint*, MapPoint*> >& log = MapPoint::replaceLog();
    for (int i = 0; i < n1; i++) {
        if (mps[i].addedObsIdx >= 0) put(0, i, mps[i].addedObsIdx);
        if (overload == 1 && vpReplace[i]) put(2, i, vpReplace[i]->id);
        while (overload == 0 && rl < log.size() && (log[rl].first == &mps[i] || log[rl].second == &mps[i]) &&
               !(log[rl].first->id > i || log[rl].second->id > i)) {
            put(1, log[rl].first->id, log[rl].second->id);
            rl++;
        }
    }
    *n_events = ne;
    return nFused;
}

/* ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12 = 1, R12 = I, t12 = 0, th) (:1743-1967), both keyframes at the identity pose.
   present{1,2}[i] = the keyframe holds a map point at i (flags: bit 0 bad); matched12_in[i] = index in KF2 of the point vpMatches12[i]
   holds on entry (-1: none, -2: a point that is not in KF2).  match12[i] = index in KF2 of the point vpMatches12[i] holds on return
   (entries of the input are kept), returns nFound. */
int ref_search_by_sim3(const float* pt8_1, const int32_t* level_1, const uint8_t* flags_1, const uint8_t* present_1, const orc_keypoint* kps1,
                       const uint8_t* desc1, int n1, const float* pt8_2, const int32_t* level_2, const uint8_t* flags_2, const uint8_t* present_2,
                       const orc_keypoint* kps2, const uint8_t* desc2, int n2, const int32_t* matched12_in, const float* bounds4, const float* K4,
                       const float* scale_factors, int nlevels, float th, int32_t* match12) {
    Pinhole cam(std::vector<float>(K4, K4 + 4));
    KeyFrame k1, k2;
    fillKeyFrame(k1, &cam, kps1, desc1, n1, bounds4, K4, scale_factors, nullptr, nlevels, nullptr);
    fillKeyFrame(k2, &cam, kps2, desc2, n2, bounds4, K4, scale_factors, nullptr, nlevels, nullptr);
    std::vector<MapPoint> m1, m2;
    // a map point's descriptor is its keyframe's descriptor here (GetDescriptor() is the point's representative descriptor)
    fillPoints(m1, pt8_1, level_1, nullptr, flags_1, desc1, n1);
    fillPoints(m2, pt8_2, level_2, nullptr, flags_2, desc2, n2);
    for (int i = 0; i < n1; i++) if (present_1[i]) k1.mvpMapPoints[i] = &m1[i];
    for (int i = 0; i < n2; i++) { m2[i].idxInOtherKF = i; if (present_2[i]) k2.mvpMapPoints[i] = &m2[i]; }
    MapPoint foreign; foreign.id = -2; foreign.idxInOtherKF = -1;
    std::vector<MapPoint*> v12((size_t)n1, nullptr);
    for (int i = 0; i < n1; i++) v12[i] = matched12_in[i] == -1 ? nullptr : (matched12_in[i] == -2 ? &foreign : &m2[matched12_in[i]]);
    ORBmatcher matcher(0.75f, true);
    const int nFound = matcher.SearchBySim3(&k1, &k2, v12, 1.0f, eye(3), cv::Mat::zeros(3, 1, CV_32FC1), th);
    for (int i = 0; i < n1; i++) match12[i] = v12[i] ? v12[i]->id : -1;
    return nFound;
}

/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo, bCoarse) (:975-1214) for two keyframes with the same
   pinhole camera, rotations the identity, translations t1w / t2w.  has_mp{1,2}[i] != 0: the keyframe holds a map point at i;
   u_right{1,2} (may be NULL) = mvuRight.  Also returns what the function forms inside and the device path receives from its caller:
   F12_out[9] = K1.t().inv() * SkewSymmetricMatrix(t12) * R12 * K2.inv() (Pinhole.cpp:137-140, the same expression on the same stand-in
   algebra) and ep_out[2] = the epipole (:982-988).  match12[i1] = second index of the pair (i1, .) in vMatchedPairs or -1. */
int ref_search_for_triangulation(const orc_keypoint* kps1, const uint8_t* desc1, const uint8_t* has_mp1, const float* u_right1, int n1,
                                 const uint32_t* nodes1, const int32_t* start1, const uint32_t* feats1, int nn1, const orc_keypoint* kps2,
                                 const uint8_t* desc2, const uint8_t* has_mp2, const float* u_right2, int n2, const uint32_t* nodes2,
                                 const int32_t* start2, const uint32_t* feats2, int nn2, const float* K4, const float* t1w, const float* t2w,
                                 const float* scale_factors, const float* level_sigma2, int nlevels, int only_stereo, int coarse, int check_ori,
                                 int32_t* match12, float* F12_out, float* ep_out) {
    Pinhole cam(std::vector<float>(K4, K4 + 4));
    const float b4[4] = {0.f, 0.f, 1.f, 1.f};
    KeyFrame k1, k2;
    fillKeyFrame(k1, &cam, kps1, desc1, n1, b4, K4, scale_factors, nullptr, nlevels, u_right1);
    fillKeyFrame(k2, &cam, kps2, desc2, n2, b4, K4, scale_factors, nullptr, nlevels, u_right2);
    k1.mvLevelSigma2.assign(level_sigma2, level_sigma2 + nlevels); k2.mvLevelSigma2 = k1.mvLevelSigma2;
    k1.tcw = vec3(t1w); k2.tcw = vec3(t2w);
    k1.Ow = -k1.Rcw.t() * k1.tcw; k2.Ow = -k2.Rcw.t() * k2.tcw;           // KeyFrame::SetPose: Ow = -Rwc * tcw
    for (int q = 0; q < nn1; q++) k1.mFeatVec[nodes1[q]] = std::vector<unsigned int>(feats1 + start1[q], feats1 + start1[q + 1]);
    for (int q = 0; q < nn2; q++) k2.mFeatVec[nodes2[q]] = std::vector<unsigned int>(feats2 + start2[q], feats2 + start2[q + 1]);
    MapPoint held;
    for (int i = 0; i < n1; i++) if (has_mp1[i]) k1.mvpMapPoints[i] = &held;
    for (int i = 0; i < n2; i++) if (has_mp2[i]) k2.mvpMapPoints[i] = &held;
    // what the function computes for itself (:982-1000, Pinhole.cpp:137-140), for the caller of the device path
    {
        cv::Mat R1w = k1.GetRotation(), t1 = k1.GetTranslation(), R2w = k2.GetRotation(), t2 = k2.GetTranslation();
        cv::Mat R12 = R1w * R2w.t();
        cv::Mat t12 = -R1w * R2w.t() * t2 + t1;
        cv::Mat F = cam.toK().t().inv() * cam.SkewSymmetricMatrix(t12) * R12 * cam.toK().inv();
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) F12_out[3 * r + c] = F.at<float>(r, c);
        cv::Mat C2 = R2w * k1.GetCameraCenter() + t2;
        const cv::Point2f ep = cam.project(C2);
        ep_out[0] = ep.x; ep_out[1] = ep.y;
    }
    std::vector<std::pair<size_t, size_t> > pairs;
    ORBmatcher matcher(0.6f, check_ori != 0);
    const int nm = matcher.SearchForTriangulation(&k1, &k2, cv::Mat(), pairs, only_stereo != 0, coarse != 0);
    for (int i = 0; i < n1; i++) match12[i] = -1;
    size_t prev = 0;
    for (size_t k = 0; k < pairs.size(); k++) {
        if (k > 0 && pairs[k].first <= prev) return -1000;               // ascending first index (:1203-1208)
        prev = pairs[k].first;
        match12[pairs[k].first] = (int32_t)pairs[k].second;
    }
    if ((int)pairs.size() != nm) return -1001;
    return nm;
}

}  // extern "C"
