// B200 replacements for the KEYFRAME-side searches of local mapping and loop closing, same signatures as the reference:
//   ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th, ratioHamming)                     src/ORBmatcher.cc:480-593
//   ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpPointsKFs, vpMatched, vpMatchedKF, th, ratio)  :595-712
//   ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th, bRight)                                                              :1407-1617
//   ORBmatcher::Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint)                                            :1619-1741
//   ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th)                                              :1743-1967
// Drop the reference's definitions of these members and link this file.  The split follows the data: what needs MapPoint / KeyFrame
// accessors (bad / already-found tests, pose algebra, projection through the keyframe's camera, IsInImage, distance and viewing-angle
// gates, PredictScale) and what changes the map (Replace, AddObservation, AddMapPoint, vpMatched) stays here on the host, in the order
// of the points, exactly where the reference has it; the window lookup, level filter, reprojection gate and descriptor search of all
// points run as ONE device call (eorb_guided_search_windows).  The keyframe's grid is rebuilt on the device from its undistorted
// keypoints and the image bounds the Frame binned them with (Frame::mnMinX ..., static floats), the windows are looked up with the
// keyframe's own int bounds like KeyFrame::GetFeaturesInArea does (src/KeyFrame.cc:873-917, include/KeyFrame.h:529).
// The small matrix algebra is written out on floats with double accumulation, as cv::gemm does for CV_32F.
// A fisheye rig (NLeft != -1: right-camera grids, Fuse with bRight) is not taken over: those calls return 0 with a message.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <set>
#include <tuple>
#include <utility>
#include <vector>

#include "eorb_b200.h"
#ifdef EORB_SHIM_MOCK
#include "ref_mock.h"
#else
#include "ORBmatcher.h"
#include "Frame.h"
#include "MapPoint.h"
#include "KeyFrame.h"
#endif

namespace ORB_SLAM3
{
namespace {
eorb_guided* kfHandle()
{
    thread_local struct Holder { eorb_guided* h = nullptr; ~Holder() { eorb_guided_destroy(h); } } holder;
    if (!holder.h && eorb_guided_create(0, &holder.h) != EORB_OK) {
        std::fprintf(stderr, "ORBmatcher(b200) keyframe-side search: %s\n", eorb_last_error());
        holder.h = nullptr;
    }
    return holder.h;
}

struct Pose { float R[9]; float t[3]; float Ow[3]; };

// y = (float)(sum_k R[r][k] * x[k]) + t[r]: cv::gemm accumulates CV_32F products in double and rounds once, operator+ adds in float
inline void rigid(const float* R, const float* t, const float* x, float* y)
{
    for (int r = 0; r < 3; r++)
        y[r] = (float)((double)R[3 * r] * x[0] + (double)R[3 * r + 1] * x[1] + (double)R[3 * r + 2] * x[2]) + (t ? t[r] : 0.f);
}
inline void matTo(const cv::Mat& M, float* R9) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R9[3 * r + c] = M.at<float>(r, c); }
inline void vecTo(const cv::Mat& v, float* x3) { for (int r = 0; r < 3; r++) x3[r] = v.at<float>(r, 0); }
inline float norm3(const float* v) { return (float)std::sqrt((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]); }   // cv::norm

// Scw -> Rcw = sRcw / scw, tcw = Scw(0:3, 3) / scw, Ow = -Rcw^T * tcw (:491-495, :606-610, :1628-1632)
Pose decomposeSim3(const cv::Mat& Scw)
{
    Pose P;
    const double d = (double)Scw.at<float>(0, 0) * Scw.at<float>(0, 0) + (double)Scw.at<float>(0, 1) * Scw.at<float>(0, 1) + (double)Scw.at<float>(0, 2) * Scw.at<float>(0, 2);
    const float scw = (float)std::sqrt(d);
    const double inv = 1.0 / (double)scw;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) P.R[3 * r + c] = (float)((double)Scw.at<float>(r, c) * inv);
        P.t[r] = (float)((double)Scw.at<float>(r, 3) * inv);
    }
    for (int r = 0; r < 3; r++) P.Ow[r] = -(float)((double)P.R[r] * P.t[0] + (double)P.R[3 + r] * P.t[1] + (double)P.R[6 + r] * P.t[2]);
    return P;
}

// the keyframe as the device sees it
struct KfView {
    std::vector<eorb_keypoint> kps; std::vector<unsigned char> desc; std::vector<float> uRight, invSigma2;
    float bounds[4]; float qmin[2]; int n = 0;
};
bool packKeyFrame(KeyFrame* pKF, KfView& v, bool wantGate)
{
    if (pKF->numAllKPtsLeft() != -1) {
        std::fprintf(stderr, "ORBmatcher(b200) keyframe-side search: fisheye-rig keyframes (NLeft != -1) are not taken over\n");
        return false;
    }
    const std::vector<cv::KeyPoint> k = pKF->getAllUndistKPtsMono();
    const cv::Mat D = pKF->getAllORBDescriptors();
    v.n = (int)k.size();
    v.kps.resize(v.n); v.desc.resize((size_t)v.n * 32);
    for (int i = 0; i < v.n; i++) {
        eorb_keypoint& o = v.kps[i];
        o.x = k[i].pt.x; o.y = k[i].pt.y; o.size = k[i].size; o.angle = k[i].angle; o.response = k[i].response; o.octave = k[i].octave; o.class_id = k[i].class_id;
        std::memcpy(&v.desc[(size_t)i * 32], D.ptr<unsigned char>(i), 32);
    }
    v.bounds[0] = Frame::mnMinX; v.bounds[1] = Frame::mnMinY; v.bounds[2] = Frame::mnMaxX; v.bounds[3] = Frame::mnMaxY;
    v.qmin[0] = (float)pKF->mnMinX; v.qmin[1] = (float)pKF->mnMinY;
    if (wantGate) {
        v.uRight.assign(pKF->mvuRight.begin(), pKF->mvuRight.end());
        v.uRight.resize(v.n, -1.f);
        const int nl = pKF->getORBNLevels();
        v.invSigma2.resize(nl);
        for (int l = 0; l < nl; l++) v.invSigma2[l] = pKF->getORBInvLevelSigma2(l);
    }
    return true;
}

// the gates every one of these functions applies to a point before it searches (:504-552 and its four copies), and the window.
// projForm 0: pCamera->project(cv::Point3f) (:526, :1472, :1670); 1: invz = 1 / z, fx * (x * invz) + cx (:636-641, :1805-1810, :1885-1890)
struct Gate { const float* pc; const float* po; bool viewGate; int projForm; float th; };
bool windowOf(MapPoint* pMP, KeyFrame* pKF, GeometricCamera* cam, const Gate& gt, eorb_area_query& q, float* invzOut, float* uOut)
{
    const float x = gt.pc[0], y = gt.pc[1], z = gt.pc[2];
    if (z < 0.0f) return false;                                    // depth must be positive
    const float invz = 1 / z;
    float u, v;
    if (gt.projForm == 0) { const cv::Point2f uv = cam->project(cv::Point3f(x, y, z)); u = uv.x; v = uv.y; }
    else { const float xn = x * invz, yn = y * invz; u = pKF->fx * xn + pKF->cx; v = pKF->fy * yn + pKF->cy; }
    if (!pKF->IsInImage(u, v)) return false;
    const float dist = norm3(gt.po);
    if (dist < pMP->GetMinDistanceInvariance() || dist > pMP->GetMaxDistanceInvariance()) return false;
    if (gt.viewGate) {                                             // viewing angle must be less than 60 deg
        const cv::Mat Pn = pMP->GetNormal();
        const double dot = (double)gt.po[0] * Pn.at<float>(0, 0) + (double)gt.po[1] * Pn.at<float>(1, 0) + (double)gt.po[2] * Pn.at<float>(2, 0);
        if (dot < 0.5 * dist) return false;
    }
    const int nPredictedLevel = pMP->PredictScale(dist, pKF);
    q.x = u; q.y = v; q.r = gt.th * pKF->getORBScaleFactor(nPredictedLevel);
    q.min_level = nPredictedLevel - 1; q.max_level = nPredictedLevel;
    if (invzOut) *invzOut = invz;
    if (uOut) *uOut = u;
    return true;
}
void noWindow(eorb_area_query& q) { q.x = 0.f; q.y = 0.f; q.r = -1.f; q.min_level = 0; q.max_level = -1; }
void descOf(MapPoint* pMP, unsigned char* dst) { const cv::Mat d = pMP->GetDescriptor(); std::memcpy(dst, d.ptr<unsigned char>(), 32); }

int projectionSearch(KeyFrame* pKF, const cv::Mat& Scw, const std::vector<MapPoint*>& vpPoints, const std::vector<KeyFrame*>* vpPointsKFs,
                     std::vector<MapPoint*>& vpMatched, std::vector<KeyFrame*>* vpMatchedKF, int th, float ratioHamming, int projForm, float thLow)
{
    eorb_guided* g = kfHandle();
    KfView kv;
    const int n1 = (int)vpPoints.size();
    if (!g || n1 == 0 || !packKeyFrame(pKF, kv, false) || kv.n == 0) return 0;
    const Pose P = decomposeSim3(Scw);
    std::set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint*>(NULL));
    std::vector<eorb_area_query> q(n1);
    std::vector<unsigned char> dmp((size_t)n1 * 32, 0), held(kv.n, 0);
    for (int i = 0; i < n1; i++) {
        noWindow(q[i]);
        MapPoint* pMP = vpPoints[i];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        float pw[3], pc[3], po[3];
        vecTo(pMP->GetWorldPos(), pw);
        rigid(P.R, P.t, pw, pc);
        for (int r = 0; r < 3; r++) po[r] = pw[r] - P.Ow[r];
        const Gate gt{pc, po, true, projForm, (float)th};
        if (windowOf(pMP, pKF, pKF->mpCamera, gt, q[i], nullptr, nullptr)) descOf(pMP, &dmp[(size_t)i * 32]);
    }
    for (int i2 = 0; i2 < kv.n && i2 < (int)vpMatched.size(); i2++) held[i2] = vpMatched[i2] ? 1 : 0;
    std::vector<int> bi(n1, -1), m2(kv.n, -1);
    int nmatches = 0;
    const int thHigh = (int)std::floor(thLow * ratioHamming);      // bestDist <= TH_LOW * ratioHamming, int against float (:585, :700)
    const int rc = eorb_guided_search_windows(g, q.data(), nullptr, dmp.data(), n1, kv.kps.data(), kv.desc.data(), held.data(), nullptr, kv.n, kv.bounds,
                                              kv.qmin, nullptr, 0, 1, thHigh < 0 ? 0 : (thHigh > 255 ? 255 : thHigh), bi.data(), nullptr, m2.data(), &nmatches);
    if (rc != EORB_OK) { std::fprintf(stderr, "ORBmatcher(b200)::SearchByProjection(KeyFrame): %s\n", eorb_last_error()); return 0; }
    for (int i2 = 0; i2 < kv.n; i2++)
        if (m2[i2] >= 0) {
            vpMatched[i2] = vpPoints[m2[i2]];
            if (vpMatchedKF) (*vpMatchedKF)[i2] = (*vpPointsKFs)[m2[i2]];
        }
    return nmatches;
}
} // namespace

int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*> &vpPoints, std::vector<MapPoint*> &vpMatched, int th, float ratioHamming)
{
    return projectionSearch(pKF, Scw, vpPoints, nullptr, vpMatched, nullptr, th, ratioHamming, 0, (float)TH_LOW);
}

int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*> &vpPoints, const std::vector<KeyFrame*> &vpPointsKFs,
                                   std::vector<MapPoint*> &vpMatched, std::vector<KeyFrame*> &vpMatchedKF, int th, float ratioHamming)
{
    return projectionSearch(pKF, Scw, vpPoints, &vpPointsKFs, vpMatched, &vpMatchedKF, th, ratioHamming, 1, (float)TH_LOW);
}

int ORBmatcher::Fuse(KeyFrame *pKF, const std::vector<MapPoint *> &vpMapPoints, const float th, const bool bRight)
{
    if (bRight) { std::fprintf(stderr, "ORBmatcher(b200)::Fuse: the right camera of a fisheye rig is not taken over\n"); return 0; }
    eorb_guided* g = kfHandle();
    KfView kv;
    const int n1 = (int)vpMapPoints.size();
    if (!g || n1 == 0 || !packKeyFrame(pKF, kv, true) || kv.n == 0) return 0;
    Pose P;
    matTo(pKF->GetRotation(), P.R); vecTo(pKF->GetTranslation(), P.t); vecTo(pKF->GetCameraCenter(), P.Ow);
    const float bf = pKF->mbf;
    std::vector<eorb_area_query> q(n1);
    std::vector<float> ur(n1, 0.f);
    std::vector<unsigned char> dmp((size_t)n1 * 32, 0);
    // The geometry of a point does not depend on the fusions before it, so every non-NULL point gets its window; isBad() /
    // IsInKeyFrame() CAN change while the loop runs (Replace marks a point bad, AddObservation puts it into the keyframe), so they are
    // asked when the point's turn comes, below, as in the reference (:1450-1459).
    for (int i = 0; i < n1; i++) {
        noWindow(q[i]);
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP) continue;
        float pw[3], pc[3], po[3], invz = 0.f, u = 0.f;
        vecTo(pMP->GetWorldPos(), pw);
        rigid(P.R, P.t, pw, pc);
        for (int r = 0; r < 3; r++) po[r] = pw[r] - P.Ow[r];
        const Gate gt{pc, po, true, 0, th};
        if (windowOf(pMP, pKF, pKF->mpCamera, gt, q[i], &invz, &u)) { descOf(pMP, &dmp[(size_t)i * 32]); ur[i] = u - bf * invz; }
    }
    std::vector<int> bi(n1, -1);
    const int rc = eorb_guided_search_windows(g, q.data(), ur.data(), dmp.data(), n1, kv.kps.data(), kv.desc.data(), nullptr, kv.uRight.data(), kv.n,
                                              kv.bounds, kv.qmin, kv.invSigma2.data(), (int)kv.invSigma2.size(), 0, TH_LOW, bi.data(), nullptr, nullptr, nullptr);
    if (rc != EORB_OK) { std::fprintf(stderr, "ORBmatcher(b200)::Fuse: %s\n", eorb_last_error()); return 0; }
    int nFused = 0;
    for (int i = 0; i < n1; i++) {
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP || bi[i] < 0) continue;
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        // If there is already a MapPoint replace otherwise add new measurement (:1583-1602)
        const int bestIdx = bi[i];
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, bestIdx);
            pKF->AddMapPoint(pMP, bestIdx);
        }
        nFused++;
    }
    return nFused;
}

int ORBmatcher::Fuse(KeyFrame *pKF, cv::Mat Scw, const std::vector<MapPoint *> &vpPoints, float th, std::vector<MapPoint *> &vpReplacePoint)
{
    eorb_guided* g = kfHandle();
    KfView kv;
    const int n1 = (int)vpPoints.size();
    if (!g || n1 == 0 || !packKeyFrame(pKF, kv, false) || kv.n == 0) return 0;
    const Pose P = decomposeSim3(Scw);
    const std::set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();
    std::vector<eorb_area_query> q(n1);
    std::vector<unsigned char> dmp((size_t)n1 * 32, 0);
    for (int i = 0; i < n1; i++) {
        noWindow(q[i]);
        MapPoint* pMP = vpPoints[i];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;     // nothing in this loop changes either (no Replace here, the set is a copy)
        float pw[3], pc[3], po[3];
        vecTo(pMP->GetWorldPos(), pw);
        rigid(P.R, P.t, pw, pc);
        for (int r = 0; r < 3; r++) po[r] = pw[r] - P.Ow[r];
        const Gate gt{pc, po, true, 0, th};
        if (windowOf(pMP, pKF, pKF->mpCamera, gt, q[i], nullptr, nullptr)) descOf(pMP, &dmp[(size_t)i * 32]);
    }
    std::vector<int> bi(n1, -1);
    const int rc = eorb_guided_search_windows(g, q.data(), nullptr, dmp.data(), n1, kv.kps.data(), kv.desc.data(), nullptr, nullptr, kv.n, kv.bounds, kv.qmin,
                                              nullptr, 0, 0, TH_LOW, bi.data(), nullptr, nullptr, nullptr);
    if (rc != EORB_OK) { std::fprintf(stderr, "ORBmatcher(b200)::Fuse(Scw): %s\n", eorb_last_error()); return 0; }
    int nFused = 0;
    for (int i = 0; i < n1; i++) {
        if (bi[i] < 0) continue;
        MapPoint* pMP = vpPoints[i];
        MapPoint* pMPinKF = pKF->GetMapPoint(bi[i]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) vpReplacePoint[i] = pMPinKF;
        } else {
            pMP->AddObservation(pKF, bi[i]);
            pKF->AddMapPoint(pMP, bi[i]);
        }
        nFused++;
    }
    return nFused;
}

int ORBmatcher::SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint*> &vpMatches12, const float &s12, const cv::Mat &R12, const cv::Mat &t12, const float th)
{
    eorb_guided* g = kfHandle();
    KfView v1, v2;
    if (!g || !packKeyFrame(pKF1, v1, false) || !packKeyFrame(pKF2, v2, false) || v1.n == 0 || v2.n == 0) return 0;
    float R1w[9], t1w[3], R2w[9], t2w[3], R12f[9], t12f[3], sR12[9], sR21[9], t21[3];
    matTo(pKF1->GetRotation(), R1w); vecTo(pKF1->GetTranslation(), t1w);
    matTo(pKF2->GetRotation(), R2w); vecTo(pKF2->GetTranslation(), t2w);
    matTo(R12, R12f); vecTo(t12, t12f);
    // sR12 = s12 * R12, sR21 = (1.0 / s12) * R12.t(), t21 = -sR21 * t12 (:1757-1759)
    const double is12 = 1.0 / s12;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { sR12[3 * r + c] = (float)((double)s12 * R12f[3 * r + c]); sR21[3 * r + c] = (float)(is12 * R12f[3 * c + r]); }
    rigid(sR21, nullptr, t12f, t21);
    for (int r = 0; r < 3; r++) t21[r] = -t21[r];
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const std::vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
    std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
    for (int i = 0; i < N1; i++) {
        MapPoint* pMP = vpMatches12[i];
        if (pMP) {
            vbAlreadyMatched1[i] = true;
            const int idx2 = std::get<0>(pMP->GetIndexInKeyFrame(pKF2));
            if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
        }
    }
    // one direction: the points of keyframe A (pose RAw, tAw) taken into B by (sRBA, tBA), searched in B
    auto direction = [&](const std::vector<MapPoint*>& vpA, const std::vector<bool>& doneA, const float* RAw, const float* tAw, const float* sRBA,
                         const float* tBA, KeyFrame* pKFB, const KfView& vB, std::vector<int>& vnMatch) -> bool {
        const int nA = (int)vpA.size();
        vnMatch.assign(nA, -1);
        std::vector<eorb_area_query> q(nA);
        std::vector<unsigned char> dmp((size_t)nA * 32, 0);
        for (int i = 0; i < nA; i++) {
            noWindow(q[i]);
            MapPoint* pMP = vpA[i];
            if (!pMP || doneA[i] || pMP->isBad()) continue;
            float pw[3], pa[3], pb[3];
            vecTo(pMP->GetWorldPos(), pw);
            rigid(RAw, tAw, pw, pa);
            rigid(sRBA, tBA, pa, pb);
            const Gate gt{pb, pb, false, 1, th};                       // dist3D = cv::norm(p3Dc2) (:1820), no viewing-angle gate
            if (windowOf(pMP, pKFB, nullptr, gt, q[i], nullptr, nullptr)) descOf(pMP, &dmp[(size_t)i * 32]);
        }
        const int rc = eorb_guided_search_windows(g, q.data(), nullptr, dmp.data(), nA, vB.kps.data(), vB.desc.data(), nullptr, nullptr, vB.n, vB.bounds, vB.qmin,
                                                  nullptr, 0, 0, TH_HIGH, vnMatch.data(), nullptr, nullptr, nullptr);
        if (rc != EORB_OK) { std::fprintf(stderr, "ORBmatcher(b200)::SearchBySim3: %s\n", eorb_last_error()); return false; }
        return true;
    };
    std::vector<int> vnMatch1, vnMatch2;
    if (!direction(vpMapPoints1, vbAlreadyMatched1, R1w, t1w, sR21, t21, pKF2, v2, vnMatch1)) return 0;
    if (!direction(vpMapPoints2, vbAlreadyMatched2, R2w, t2w, sR12, t12f, pKF1, v1, vnMatch2)) return 0;
    // Check agreement (:1944-1958)
    int nFound = 0;
    for (int i1 = 0; i1 < N1; i1++) {
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0 && idx2 < N2 && vnMatch2[idx2] == i1) { vpMatches12[i1] = vpMapPoints2[idx2]; nFound++; }
    }
    return nFound;
}
// ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo, bCoarse) (:975-1214; LocalMapping::CreateNewMapPoints,
// src/LocalMapping.cc:507; Tracking.cc:3212) for keyframes with one pinhole camera each.  The host forms what the reference forms once per
// call -- the epipole (:982-988), R12 / t12 (:998-999) and the fundamental matrix Pinhole::epipolarConstrain rebuilds for every candidate pair
// (src/CameraModels/Pinhole.cpp:137-140; the argument F12 is not read by the reference function either) -- the node walk, descriptor search,
// both gates and the rotation filter run on the device.
namespace {
void flattenFV(const DBoW2::FeatureVector& fv, std::vector<unsigned>& nodes, std::vector<int>& start, std::vector<unsigned>& feats)
{
    nodes.clear(); feats.clear(); start.assign(1, 0);
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
        nodes.push_back(it->first);
        feats.insert(feats.end(), it->second.begin(), it->second.end());
        start.push_back((int)feats.size());
    }
}
} // namespace

int ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<pair<size_t, size_t> > &vMatchedPairs, bool bOnlyStereo, bool bCoarse)
{
    (void)F12;
    vMatchedPairs.clear();
    GeometricCamera* pCamera1 = pKF1->mpCamera; GeometricCamera* pCamera2 = pKF2->mpCamera;
    if (pKF1->mpCamera2 || pKF2->mpCamera2 || pCamera1->GetType() != pCamera1->CAM_PINHOLE || pCamera2->GetType() != pCamera2->CAM_PINHOLE) {
        std::fprintf(stderr, "ORBmatcher(b200)::SearchForTriangulation: only pinhole keyframes with one camera are taken over\n");
        return 0;
    }
    eorb_guided* g = kfHandle();
    KfView v1, v2;
    if (!g || !packKeyFrame(pKF1, v1, true) || !packKeyFrame(pKF2, v2, true) || v1.n == 0 || v2.n == 0) return 0;
    float Cw[3], R1w[9], t1w[3], R2w[9], t2w[3], C2[3];
    vecTo(pKF1->GetCameraCenter(), Cw);
    matTo(pKF1->GetRotation(), R1w); vecTo(pKF1->GetTranslation(), t1w);
    matTo(pKF2->GetRotation(), R2w); vecTo(pKF2->GetTranslation(), t2w);
    rigid(R2w, t2w, Cw, C2);                                                   // C2 = R2w * Cw + t2w
    const cv::Point2f ep = pCamera2->project(cv::Point3f(C2[0], C2[1], C2[2]));
    float Fm[9];
#ifndef EORB_SHIM_MOCK
    {   // the reference's own expressions on cv::Mat (ORBmatcher.cc:998-999, Pinhole.cpp:137-140, :175-180)
        const cv::Mat R1 = pKF1->GetRotation(), R2 = pKF2->GetRotation();
        const cv::Mat R12 = R1 * R2.t();
        const cv::Mat t12 = -R1 * R2.t() * pKF2->GetTranslation() + pKF1->GetTranslation();
        const cv::Mat t12x = (cv::Mat_<float>(3, 3) << 0, -t12.at<float>(2), t12.at<float>(1), t12.at<float>(2), 0, -t12.at<float>(0), -t12.at<float>(1), t12.at<float>(0), 0);
        const cv::Mat K1 = pCamera1->toK(), K2 = pCamera2->toK();
        const cv::Mat F = K1.t().inv() * t12x * R12 * K2.inv();
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Fm[3 * r + c] = F.at<float>(r, c);
    }
#else
    {   // the same with the stand-in cv::Mat (no matrix algebra): written out on floats, products accumulated in double
        float R12[9], t12[3], tmp[3];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++)
            R12[3 * r + c] = (float)((double)R1w[3 * r] * R2w[3 * c] + (double)R1w[3 * r + 1] * R2w[3 * c + 1] + (double)R1w[3 * r + 2] * R2w[3 * c + 2]);
        rigid(R12, nullptr, t2w, tmp);
        for (int r = 0; r < 3; r++) t12[r] = -tmp[r] + t1w[r];
        const float tx[9] = {0.f, -t12[2], t12[1], t12[2], 0.f, -t12[0], -t12[1], t12[0], 0.f};
        const double fx1 = pCamera1->getParameter(0), fy1 = pCamera1->getParameter(1), cx1 = pCamera1->getParameter(2), cy1 = pCamera1->getParameter(3);
        const double fx2 = pCamera2->getParameter(0), fy2 = pCamera2->getParameter(1), cx2 = pCamera2->getParameter(2), cy2 = pCamera2->getParameter(3);
        const float K1tInv[9] = {(float)(1.0 / fx1), 0.f, 0.f, 0.f, (float)(1.0 / fy1), 0.f, (float)(-cx1 / fx1), (float)(-cy1 / fy1), 1.f};
        const float K2Inv[9] = {(float)(1.0 / fx2), 0.f, (float)(-cx2 / fx2), 0.f, (float)(1.0 / fy2), (float)(-cy2 / fy2), 0.f, 0.f, 1.f};
        auto mul = [](const float* A, const float* B, float* Cm) {
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++)
                Cm[3 * r + c] = (float)((double)A[3 * r] * B[c] + (double)A[3 * r + 1] * B[3 + c] + (double)A[3 * r + 2] * B[6 + c]);
        };
        float A[9], B[9];
        mul(K1tInv, tx, A); mul(A, R12, B); mul(B, K2Inv, Fm);
    }
#endif
    std::vector<unsigned char> f1(v1.n, 0), f2(v2.n, 0);
    for (int i = 0; i < v1.n; i++) {
        const bool st = v1.uRight[i] >= 0;
        f1[i] = (unsigned char)(((!pKF1->GetMapPoint(i) && (!bOnlyStereo || st)) ? 1 : 0) | (st ? 2 : 0));
    }
    for (int i = 0; i < v2.n; i++) {
        const bool st = v2.uRight[i] >= 0;
        f2[i] = (unsigned char)(((!pKF2->GetMapPoint(i) && (!bOnlyStereo || st)) ? 1 : 0) | (st ? 2 : 0));
    }
    std::vector<unsigned> nodes1, feats1, nodes2, feats2;
    std::vector<int> start1, start2;
    flattenFV(pKF1->mFeatVec, nodes1, start1, feats1); flattenFV(pKF2->mFeatVec, nodes2, start2, feats2);
    const int nl = pKF2->getORBNLevels();
    std::vector<float> sc(nl), sg(nl);
    for (int l = 0; l < nl; l++) { sc[l] = pKF2->getORBScaleFactor(l); sg[l] = pKF2->getORBLevelSigma2(l); }
    const float epf[2] = {ep.x, ep.y};
    std::vector<int> m12(v1.n, -1);
    int nmatches = 0;
    const int rc = eorb_guided_search_for_triangulation(g, v1.kps.data(), v1.desc.data(), f1.data(), v1.n, nodes1.data(), start1.data(), feats1.data(),
                                                        (int)nodes1.size(), v2.kps.data(), v2.desc.data(), f2.data(), v2.n, nodes2.data(), start2.data(),
                                                        feats2.data(), (int)nodes2.size(), Fm, epf, sc.data(), sg.data(), nl, bCoarse ? 1 : 0,
                                                        mbCheckOrientation ? 1 : 0, m12.data(), &nmatches);
    if (rc != EORB_OK) { std::fprintf(stderr, "ORBmatcher(b200)::SearchForTriangulation: %s\n", eorb_last_error()); return 0; }
    vMatchedPairs.reserve(nmatches);
    for (int i = 0; i < v1.n; i++)
        if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)m12[i]));
    return nmatches;
}
} // namespace ORB_SLAM3
