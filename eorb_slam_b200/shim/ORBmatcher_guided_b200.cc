// B200 replacements for ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:714-831) and the monocular path of
// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (:1969-2150, Tracking::TrackWithMotionModel), same signatures: drop the
// reference's definition of this one member and link this file (Tracking.cc's MonocularInitialization keeps calling
// matcher.SearchForInitialization(mInitialFrame, mCurrentFrame, mvbPrevMatched, mvIniMatches, 100) unchanged).
// The Frame grid (AssignFeaturesToGrid / GetFeaturesInArea, src/Frame.cc:431-460, 709-793) is rebuilt on the device from
// F2's undistorted keypoints and the static image bounds, so Frame::mGrid is not read.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
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
#ifdef EORB_SHIM_MOCK
float Frame::mnMinX = 0.f, Frame::mnMaxX = 0.f, Frame::mnMinY = 0.f, Frame::mnMaxY = 0.f;
#endif

namespace {
// one handle per calling thread: matchers are stack objects in the reference (Tracking.cc), handles hold device buffers
eorb_guided* threadHandle()
{
    thread_local struct Holder { eorb_guided* h = nullptr; ~Holder() { eorb_guided_destroy(h); } } holder;
    if (!holder.h && eorb_guided_create(0, &holder.h) != EORB_OK) {
        std::fprintf(stderr, "ORBmatcher(b200)::SearchForInitialization: %s\n", eorb_last_error());
        holder.h = nullptr;
    }
    return holder.h;
}

void packFrame(Frame& F, std::vector<eorb_keypoint>& kps, std::vector<unsigned char>& desc)
{
    const int n = F.numAllKPts();
    kps.resize(n); desc.resize((size_t)n * 32);
    cv::Mat& D = F.getAllORBDescMono();
    for (int i = 0; i < n; i++) {
        const cv::KeyPoint kp = F.getUndistKPtMono(i);
        eorb_keypoint& k = kps[i];
        k.x = kp.pt.x; k.y = kp.pt.y; k.size = kp.size; k.angle = kp.angle; k.response = kp.response;
        k.octave = F.getKPtLevelMono(i); k.class_id = kp.class_id;
        std::memcpy(&desc[(size_t)i * 32], D.ptr<unsigned char>(i), 32);
    }
}
} // namespace

int ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize)
{
    const int n1 = F1.numAllKPts();
    vnMatches12 = std::vector<int>(n1, -1);
    eorb_guided* g = threadHandle();
    if (!g || n1 == 0 || (int)vbPrevMatched.size() < n1) return 0;   // no CPU fallback: no device -> no matches, error on stderr
    std::vector<eorb_keypoint> k1, k2;
    std::vector<unsigned char> d1, d2;
    packFrame(F1, k1, d1); packFrame(F2, k2, d2);
    const float bounds[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY};
    static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "cv::Point2f is two floats");
    int nmatches = 0;
    const int rc = eorb_guided_search_for_initialization(g, k1.data(), d1.data(), n1, k2.data(), d2.data(), (int)k2.size(), bounds,
                                                         reinterpret_cast<float*>(vbPrevMatched.data()), windowSize, mfNNratio,
                                                         mbCheckOrientation ? 1 : 0, vnMatches12.data(), &nmatches);
    if (rc != EORB_OK) {
        std::fprintf(stderr, "ORBmatcher(b200)::SearchForInitialization: %s\n", eorb_last_error());
        vnMatches12.assign(n1, -1);
        return 0;
    }
    return nmatches;
}
int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
{
    // Monocular, rectified-stereo and RGB-D frames (Nleft == -1) run on the device, including the forward / backward level windows and
    // the right-column test (:1989-1990, :2024-2029, :2049-2055).  A fisheye rig (Nleft != -1) has a second pass over the right
    // camera's own grid (:2093-2160): rename the reference body and call it for those frames.
    const int n1 = LastFrame.numAllKPts(), n2 = CurrentFrame.numAllKPts();
    int levelMode = 0;
    if (!bMono && !CurrentFrame.mTcw.empty() && !LastFrame.mTcw.empty()) {
        // tlc = Rlw * twc + tlw with twc = -Rcw^T * tcw (:1980-1987); only its z component is read
        const cv::Mat& Tc = CurrentFrame.mTcw; const cv::Mat& Tl = LastFrame.mTcw;
        float twc[3];
        for (int r = 0; r < 3; r++)
            twc[r] = -(Tc.at<float>(0, r) * Tc.at<float>(0, 3) + Tc.at<float>(1, r) * Tc.at<float>(1, 3) + Tc.at<float>(2, r) * Tc.at<float>(2, 3));
        const float tlcz = Tl.at<float>(2, 0) * twc[0] + Tl.at<float>(2, 1) * twc[1] + Tl.at<float>(2, 2) * twc[2] + Tl.at<float>(2, 3);
        if (tlcz > CurrentFrame.mb) levelMode = 1;            // bForward
        else if (-tlcz > CurrentFrame.mb) levelMode = 2;      // bBackward
    }
    const bool haveRight = CurrentFrame.numKPtsLeft() == -1 && (int)CurrentFrame.mvuRight.size() == n2;
    eorb_guided* g = threadHandle();
    if (!g || n1 == 0 || n2 == 0) return 0;
    std::vector<eorb_keypoint> k1(n1), k2;
    std::vector<unsigned char> dmp((size_t)n1 * 32, 0), d2, valid(n1, 0);
    std::vector<float> x3(3 * (size_t)n1, 0.f);
    std::vector<int> obs(n1, 0);
    const cv::Mat& T = CurrentFrame.mTcw;
    for (int i = 0; i < n1; i++) {
        const cv::KeyPoint kp = LastFrame.getUndistKPtMono(i);
        k1[i].x = kp.pt.x; k1[i].y = kp.pt.y; k1[i].size = kp.size; k1[i].angle = kp.angle; k1[i].response = kp.response;
        k1[i].octave = LastFrame.getKPtLevelMono(i); k1[i].class_id = kp.class_id;
        MapPoint* pMP = LastFrame.getMapPoint(i);
        if (!pMP || LastFrame.getMPOutlier(i)) continue;
        valid[i] = 1;
        obs[i] = pMP->Observations();
        const cv::Mat x3Dw = pMP->GetWorldPos();
#ifdef EORB_SHIM_MOCK
        for (int r = 0; r < 3; r++)   // the cv mock has no matrix product; a real build uses the expression of the reference below
            x3[3 * i + r] = T.at<float>(r, 0) * x3Dw.at<float>(0, 0) + T.at<float>(r, 1) * x3Dw.at<float>(1, 0) + T.at<float>(r, 2) * x3Dw.at<float>(2, 0) + T.at<float>(r, 3);
#else
        const cv::Mat x3Dc = T.rowRange(0, 3).colRange(0, 3) * x3Dw + T.rowRange(0, 3).col(3);   // :2000-2001, OpenCV's own arithmetic
        for (int r = 0; r < 3; r++) x3[3 * i + r] = x3Dc.at<float>(r);
#endif
        const cv::Mat dMP = pMP->GetDescriptor();
        std::memcpy(&dmp[(size_t)i * 32], dMP.ptr<unsigned char>(), 32);
    }
    packFrame(CurrentFrame, k2, d2);
    const float bounds[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY};
    const float K4[4] = {CurrentFrame.mpCamera->getParameter(0), CurrentFrame.mpCamera->getParameter(1), CurrentFrame.mpCamera->getParameter(2),
                         CurrentFrame.mpCamera->getParameter(3)};
    const std::vector<float> sf = CurrentFrame.getAllORBScaleFactors();
    std::vector<int> mc(n2, -1);
    int nmatches = 0;
    const int rc = eorb_guided_search_by_projection_stereo(g, x3.data(), valid.data(), obs.data(), k1.data(), dmp.data(), n1, k2.data(), d2.data(),
                                                           haveRight ? CurrentFrame.mvuRight.data() : nullptr, n2, bounds, K4, sf.data(),
                                                           (int)sf.size(), th, mbCheckOrientation ? 1 : 0, levelMode, CurrentFrame.mbf, mc.data(),
                                                           &nmatches);
    if (rc != EORB_OK) {
        std::fprintf(stderr, "ORBmatcher(b200)::SearchByProjection: %s\n", eorb_last_error());
        return 0;
    }
    for (int i2 = 0; i2 < n2; i2++)
        if (mc[i2] >= 0) CurrentFrame.setMapPoint(i2, LastFrame.getMapPoint(mc[i2]));
    return nmatches;
}

// ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints) (:44-218; Tracking::SearchLocalPoints): monocular,
// rectified-stereo and RGB-D frames (Nleft == -1; the right-column test :91-96 runs on the device).  A fisheye rig (Nleft != -1) keeps
// the reference's implementation for its right-camera block (:149-216).
int ORBmatcher::SearchByProjection(Frame &F, const std::vector<MapPoint*> &vpMapPoints, const float th, const bool bFarPoints, const float thFarPoints)
{
    const int n1 = (int)vpMapPoints.size(), n2 = F.numAllKPts();
    eorb_guided* g = threadHandle();
    if (!g || n1 == 0 || n2 == 0) return 0;
    std::vector<eorb_track_point> pts(n1);
    std::vector<unsigned char> dmp((size_t)n1 * 32, 0), d2, held(n2, 0);
    std::vector<eorb_keypoint> k2;
    std::vector<float> projXR(n1, 0.f);
    const bool haveRight = F.numKPtsLeft() == -1 && (int)F.mvuRight.size() == n2;
    for (int i = 0; i < n1; i++) {
        MapPoint* pMP = vpMapPoints[i];
        eorb_track_point& p = pts[i];
        std::memset(&p, 0, sizeof(p));
        if (!pMP || !pMP->mbTrackInView) continue;
        projXR[i] = pMP->mTrackProjXR;
        p.proj_x = pMP->mTrackProjX; p.proj_y = pMP->mTrackProjY; p.view_cos = pMP->mTrackViewCos; p.depth = pMP->mTrackDepth;
        p.scale_level = pMP->mnTrackScaleLevel; p.observations = pMP->Observations();
        p.in_view = 1; p.bad = pMP->isBad() ? 1 : 0;
        const cv::Mat dMP = pMP->GetDescriptor();
        std::memcpy(&dmp[(size_t)i * 32], dMP.ptr<unsigned char>(), 32);
    }
    packFrame(F, k2, d2);
    for (int i2 = 0; i2 < n2; i2++) {
        MapPoint* q = F.getMapPoint(i2);
        held[i2] = (q && q->Observations() > 0) ? 1 : 0;
    }
    const float bounds[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY};
    const std::vector<float> sf = F.getAllORBScaleFactors();
    std::vector<int> mc(n2, -1);
    int nmatches = 0;
    const int rc = eorb_guided_search_by_projection_map_points_stereo(g, pts.data(), haveRight ? projXR.data() : nullptr, dmp.data(), n1, k2.data(),
                                                                      d2.data(), held.data(), haveRight ? F.mvuRight.data() : nullptr, n2, bounds,
                                                                      sf.data(), (int)sf.size(), th, bFarPoints ? 1 : 0, thFarPoints, mfNNratio,
                                                                      mc.data(), &nmatches);
    if (rc != EORB_OK) {
        std::fprintf(stderr, "ORBmatcher(b200)::SearchByProjection(map points): %s\n", eorb_last_error());
        return 0;
    }
    for (int i2 = 0; i2 < n2; i2++)
        if (mc[i2] >= 0) F.setMapPoint(i2, vpMapPoints[mc[i2]]);
    return nmatches;
}

// ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (:2189-2312; Tracking::Relocalization after the PnP
// refinement).  The per-point host arithmetic of the reference stays here with its own expressions: camera centre Ow = -Rcw^T tcw,
// dist3D = norm(x3Dw - Ow), the distance-invariance gate and MapPoint::PredictScale (:2196, :2226-2237).
int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const std::set<MapPoint*> &sAlreadyFound, const float th, const int ORBdist)
{
    const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
    const int n1 = (int)vpMPs.size(), n2 = CurrentFrame.numAllKPts();
    eorb_guided* g = threadHandle();
    if (!g || n1 == 0 || n2 == 0) return 0;
    const cv::Mat& T = CurrentFrame.mTcw;
    float Ow[3];
    for (int r = 0; r < 3; r++)
        Ow[r] = -(T.at<float>(0, r) * T.at<float>(0, 3) + T.at<float>(1, r) * T.at<float>(1, 3) + T.at<float>(2, r) * T.at<float>(2, 3));
    std::vector<eorb_keypoint> k1(n1), k2;
    std::vector<unsigned char> dmp((size_t)n1 * 32, 0), d2, valid(n1, 0), held(n2, 0);
    std::vector<float> x3(3 * (size_t)n1, 0.f);
    std::vector<int> level(n1, 0);
    for (int i = 0; i < n1; i++) {
        const cv::KeyPoint kp = pKF->getUndistKPtMono(i);
        k1[i].x = kp.pt.x; k1[i].y = kp.pt.y; k1[i].size = kp.size; k1[i].angle = kp.angle; k1[i].response = kp.response;
        k1[i].octave = kp.octave; k1[i].class_id = kp.class_id;
        MapPoint* pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
        const cv::Mat x3Dw = pMP->GetWorldPos();
        float po[3];
        for (int r = 0; r < 3; r++) {
            x3[3 * i + r] = T.at<float>(r, 0) * x3Dw.at<float>(0, 0) + T.at<float>(r, 1) * x3Dw.at<float>(1, 0) + T.at<float>(r, 2) * x3Dw.at<float>(2, 0) + T.at<float>(r, 3);
            po[r] = x3Dw.at<float>(r, 0) - Ow[r];
        }
        const float dist3D = (float)std::sqrt((double)po[0] * po[0] + (double)po[1] * po[1] + (double)po[2] * po[2]);   // cv::norm accumulates in double
        if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
        level[i] = pMP->PredictScale(dist3D, &CurrentFrame);
        valid[i] = 1;
        const cv::Mat dMP = pMP->GetDescriptor();
        std::memcpy(&dmp[(size_t)i * 32], dMP.ptr<unsigned char>(), 32);
    }
    packFrame(CurrentFrame, k2, d2);
    for (int i2 = 0; i2 < n2; i2++) held[i2] = CurrentFrame.getMapPoint(i2) ? 1 : 0;
    const float bounds[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY};
    const float K4[4] = {CurrentFrame.mpCamera->getParameter(0), CurrentFrame.mpCamera->getParameter(1), CurrentFrame.mpCamera->getParameter(2),
                         CurrentFrame.mpCamera->getParameter(3)};
    const std::vector<float> sf = CurrentFrame.getAllORBScaleFactors();
    std::vector<int> mc(n2, -1);
    int nmatches = 0;
    const int rc = eorb_guided_search_by_projection_reloc(g, x3.data(), valid.data(), level.data(), k1.data(), dmp.data(), n1, k2.data(), d2.data(),
                                                          held.data(), n2, bounds, K4, sf.data(), (int)sf.size(), th, ORBdist,
                                                          mbCheckOrientation ? 1 : 0, mc.data(), &nmatches);
    if (rc != EORB_OK) {
        std::fprintf(stderr, "ORBmatcher(b200)::SearchByProjection(relocalisation): %s\n", eorb_last_error());
        return 0;
    }
    for (int i2 = 0; i2 < n2; i2++)
        if (mc[i2] >= 0) CurrentFrame.setMapPoint(i2, vpMPs[mc[i2]]);
    return nmatches;
}

// ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) (:276-478; Tracking::TrackReferenceKeyFrame, Relocalization), monocular frame.
// The FeatureVectors (std::map<NodeId, vector<unsigned>>) are flattened in map order = ascending node id.
namespace {
void flattenFeatureVector(const DBoW2::FeatureVector& fv, std::vector<unsigned>& nodes, std::vector<int>& start, std::vector<unsigned>& feats)
{
    nodes.clear(); feats.clear(); start.assign(1, 0);
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
        nodes.push_back(it->first);
        feats.insert(feats.end(), it->second.begin(), it->second.end());
        start.push_back((int)feats.size());
    }
}
} // namespace

int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame &F, std::vector<MapPoint*> &vpMapPointMatches)
{
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    const int n1 = pKF->numAllKPts(), n2 = F.numAllKPts();
    vpMapPointMatches = std::vector<MapPoint*>(n2, static_cast<MapPoint*>(NULL));
    eorb_guided* g = threadHandle();
    if (!g || n1 == 0 || n2 == 0) return 0;
    std::vector<eorb_keypoint> k1(n1), k2;
    std::vector<unsigned char> d1((size_t)n1 * 32), d2, valid(n1, 0);
    for (int i = 0; i < n1; i++) {
        const cv::KeyPoint kp = pKF->getUndistKPtMono(i);
        k1[i].x = kp.pt.x; k1[i].y = kp.pt.y; k1[i].size = kp.size; k1[i].angle = kp.angle; k1[i].response = kp.response;
        k1[i].octave = kp.octave; k1[i].class_id = kp.class_id;
        const cv::Mat d = pKF->getORBDescriptor(i);
        std::memcpy(&d1[(size_t)i * 32], d.ptr<unsigned char>(), 32);
        MapPoint* pMP = i < (int)vpMapPointsKF.size() ? vpMapPointsKF[i] : NULL;
        valid[i] = (pMP && !pMP->isBad()) ? 1 : 0;
    }
    packFrame(F, k2, d2);
    std::vector<unsigned> an, af, bn, bf;
    std::vector<int> as, bs;
    flattenFeatureVector(pKF->mFeatVec, an, as, af);
    flattenFeatureVector(F.mFeatVec, bn, bs, bf);
    std::vector<int> mf(n2, -1);
    int nmatches = 0;
    const int rc = eorb_guided_search_by_bow(g, k1.data(), d1.data(), valid.data(), n1, an.data(), as.data(), af.data(), (int)an.size(), k2.data(), d2.data(),
                                             n2, bn.data(), bs.data(), bf.data(), (int)bn.size(), mfNNratio, mbCheckOrientation ? 1 : 0, mf.data(), &nmatches);
    if (rc != EORB_OK) {
        std::fprintf(stderr, "ORBmatcher(b200)::SearchByBoW: %s\n", eorb_last_error());
        return 0;
    }
    for (int i2 = 0; i2 < n2; i2++)
        if (mf[i2] >= 0) vpMapPointMatches[i2] = vpMapPointsKF[mf[i2]];
    return nmatches;
}

// ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) (:833-990; loop closing / place recognition), monocular keyframes
int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*> &vpMatches12)
{
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const std::vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    const int n1 = pKF1->numAllKPts(), n2 = pKF2->numAllKPts();
    vpMatches12 = std::vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));
    eorb_guided* g = threadHandle();
    if (!g || n1 == 0 || n2 == 0) return 0;
    struct Side { std::vector<eorb_keypoint> k; std::vector<unsigned char> d, valid; std::vector<unsigned> nodes, feats; std::vector<int> start; };
    Side s[2];
    KeyFrame* kf[2] = {pKF1, pKF2};
    const std::vector<MapPoint*>* mps[2] = {&vpMapPoints1, &vpMapPoints2};
    for (int side = 0; side < 2; side++) {
        const int n = side ? n2 : n1;
        s[side].k.resize(n); s[side].d.resize((size_t)n * 32); s[side].valid.assign(n, 0);
        for (int i = 0; i < n; i++) {
            const cv::KeyPoint kp = kf[side]->getUndistKPtMono(i);
            eorb_keypoint& o = s[side].k[i];
            o.x = kp.pt.x; o.y = kp.pt.y; o.size = kp.size; o.angle = kp.angle; o.response = kp.response; o.octave = kp.octave; o.class_id = kp.class_id;
            const cv::Mat d = kf[side]->getORBDescriptor(i);
            std::memcpy(&s[side].d[(size_t)i * 32], d.ptr<unsigned char>(), 32);
            MapPoint* pMP = i < (int)mps[side]->size() ? (*mps[side])[i] : NULL;
            s[side].valid[i] = (pMP && !pMP->isBad()) ? 1 : 0;
        }
        flattenFeatureVector(kf[side]->mFeatVec, s[side].nodes, s[side].start, s[side].feats);
    }
    std::vector<int> m12(n1, -1);
    int nmatches = 0;
    const int rc = eorb_guided_search_by_bow_kf(g, s[0].k.data(), s[0].d.data(), s[0].valid.data(), n1, s[0].nodes.data(), s[0].start.data(),
                                                s[0].feats.data(), (int)s[0].nodes.size(), s[1].k.data(), s[1].d.data(), s[1].valid.data(), n2,
                                                s[1].nodes.data(), s[1].start.data(), s[1].feats.data(), (int)s[1].nodes.size(), mfNNratio,
                                                mbCheckOrientation ? 1 : 0, m12.data(), &nmatches);
    if (rc != EORB_OK) {
        std::fprintf(stderr, "ORBmatcher(b200)::SearchByBoW(KF, KF): %s\n", eorb_last_error());
        return 0;
    }
    for (int i1 = 0; i1 < n1 && i1 < (int)vpMatches12.size(); i1++)
        if (m12[i1] >= 0) vpMatches12[i1] = vpMapPoints2[m12[i1]];
    return nmatches;
}
} // namespace ORB_SLAM3
