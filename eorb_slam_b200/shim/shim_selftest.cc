// Compiles and links the three shims against the cv mock and libeorb_b200.so and drives them once through the
// reference call shapes.  With a CUDA device it prints counts that tests/test_shim.py cross-checks; without one
// it verifies the loud-failure path (no CPU fallback): -1 / empty results and an error message.
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "ORBextractor.h"
#include "ORBmatcher_b200.h"
#include "EventConversion_b200.h"
#include "KLT_b200.h"
#include "ORBVocabulary_b200.h"
#include "eorb_b200.h"

int main(int argc, char** argv)
{
    const int W = 752, H = 480;
    cv::Mat im(H, W, CV_8UC1);
    unsigned s = 12345;
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) { s = s * 1664525u + 1013904223u; im.at<unsigned char>(y, x) = (unsigned char)(((x / 24 + y / 24) & 1) * 90 + 60 + (s >> 28)); }
    ORB_SLAM3::ORBxParams par(1000, 1.2f, 8, 20, 7, 19, cv::Size(W, H));
    ORB_SLAM3::ORBextractor ex(par);
    ex.mbDownloadPyramid = true;      // as the rectified-stereo Frame constructor does (ComputeStereoMatches reads mvImagePyramid)
    std::vector<cv::KeyPoint> kps, kps2;
    cv::Mat desc, mask;
    std::vector<int> lap = {0, 1000};
    int ret = ex(im, mask, kps, desc, lap);
    int ret2 = ex(im, mask, kps2, lap);
    std::printf("devices=%d ret=%d n=%zu desc_rows=%d ret2=%d n2=%zu levels=%d pyr0=%dx%d\n", eorb_device_count(), ret, kps.size(), desc.rows,
                ret2, kps2.size(), ex.GetLevels(), ex.mvImagePyramid[0].cols, ex.mvImagePyramid[0].rows);
    cv::Mat empty;
    std::printf("empty_ret=%d\n", ex(empty, mask, kps2, lap));
    {   // two threads drive ONE extractor object at once: each gets its own device handle, results are those of the serial call
        std::vector<cv::KeyPoint> ka, kb;
        cv::Mat da, db;
        int ra = -2, rb = -2;
        ex.mbDownloadPyramid = false;   // mvImagePyramid is ONE public member per extractor: concurrent callers must not ask for it
        std::thread ta([&] { std::vector<int> l = {0, 1000}; cv::Mat m; for (int i = 0; i < 3; i++) ra = ex(im, m, ka, da, l); });
        std::thread tb([&] { std::vector<int> l = {0, 1000}; cv::Mat m; for (int i = 0; i < 3; i++) rb = ex(im, m, kb, db, l); });
        ta.join(); tb.join();
        bool same = ra == ret && rb == ret && ka.size() == kps.size() && kb.size() == kps.size() && da.rows == desc.rows && db.rows == desc.rows;
        for (size_t i = 0; same && i < kps.size(); i++)
            same = ka[i].pt.x == kps[i].pt.x && ka[i].pt.y == kps[i].pt.y && kb[i].pt.x == kps[i].pt.x && kb[i].angle == kps[i].angle &&
                   std::memcmp(da.ptr<unsigned char>((int)i), desc.ptr<unsigned char>((int)i), 32) == 0 &&
                   std::memcmp(db.ptr<unsigned char>((int)i), desc.ptr<unsigned char>((int)i), 32) == 0;
        std::printf("threads_equal=%d\n", same ? 1 : 0);
    }
    if (desc.rows >= 2) {
        std::printf("dist01=%d\n", ORB_SLAM3::ORBmatcher::DescriptorDistance(desc.row(0), desc.row(1)));
        ORB_SLAM3::BruteForceBest2 bf(0.9f, true);
        bf.SetTrainDescriptors(desc);
        std::vector<int> m12;
        int nm = bf.Match(desc, kps, kps, m12);
        int self = 0;
        for (size_t i = 0; i < m12.size(); i++) self += (m12[i] == (int)i);
        std::printf("selfmatch=%d of %d (nm=%d)\n", self, (int)m12.size(), nm);
    }
    std::vector<EORB_SLAM::EventData> evs;
    for (int i = 0; i < 2000; i++) { s = s * 1664525u + 1013904223u; float x = (s >> 8) % 24000 / 100.f; s = s * 1664525u + 1013904223u; float y = (s >> 8) % 18000 / 100.f; evs.emplace_back(1e-6 * i, x, y, (s >> 5) & 1); }
    cv::Mat f = EORB_SLAM::EvImConverter::ev2im_gauss(evs, 240, 180, 1.0f, false, false);
    cv::Mat u = EORB_SLAM::EvImConverter::ev2im_gauss(evs, 240, 180, 1.0f);
    double sum = 0; int mx = 0;
    for (int y = 0; y < 180; y++) for (int x = 0; x < 240; x++) { sum += f.at<float>(y, x); if (u.at<unsigned char>(y, x) > mx) mx = u.at<unsigned char>(y, x); }
    ORB_SLAM3::GeometricCamera cam({199.09f, 198.83f, 132.19f, 110.71f});
    cv::Mat T = cv::Mat::zeros(4, 4, CV_32F);
    for (int i = 0; i < 4; i++) T.at<float>(i, i) = 1.f;
    cv::Mat g = EORB_SLAM::EvImConverter::ev2mci_gg_f(evs, &cam, T, 1.0f, 240, 180, 1.0f, false, false);
    double sum2 = 0;
    for (int y = 0; y < 180; y++) for (int x = 0; x < 240; x++) sum2 += g.at<float>(y, x);
    std::printf("ev_sum=%.3f ev_u8_max=%d types=%d,%d mci_sum=%.3f\n", sum, mx, f.type(), u.type(), sum2);
    std::printf("focus=%.6f focus_med=%.6f focus_glob=%.6f mean_loc=%.6f\n", EORB_SLAM::EvImConverter::measureImageFocus(f),
                EORB_SLAM::EvImConverter::measureImageFocusLocal(f, false), EORB_SLAM::EvImConverter::measureImageFocusGlobal(f),
                EORB_SLAM::EvImConverter::imageMeanLocal(f));
    {
        const double Rt[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0.02, -0.01, 0.03};
        double j[6];
        const bool okj = EORB_SLAM::EvImConverter::ev2mci_gg_f_jac(evs, &cam, Rt, 1.0f, 240, 180, 1.0f, false, true, j);
        std::printf("jac_ok=%d jac=%.6g,%.6g,%.6g,%.6g,%.6g,%.6g\n", (int)okj, j[0], j[1], j[2], j[3], j[4], j[5]);
    }
    // ELK_Tracker's call: track the extractor's keypoints from the image into a copy shifted by (2, 1) pixels
    {
        cv::Mat im2(H, W, CV_8UC1);
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) im2.at<unsigned char>(y, x) = im.at<unsigned char>(y >= 1 ? y - 1 : 0, x >= 2 ? x - 2 : 0);
        std::vector<cv::Point2f> p0, p1;
        for (size_t i = 0; i < kps.size(); i++) p0.push_back(kps[i].pt);
        std::vector<unsigned char> st; std::vector<float> er;
        const bool okk = EORB_SLAM::b200::calcOpticalFlowPyrLK(im, im2, p0, p1, st, er, 23, 1, 10, 0.03, false);
        int good = 0, tracked = 0;
        for (size_t i = 0; i < st.size(); i++) if (st[i]) { tracked++; const float dx = p1[i].x - p0[i].x - 2.f, dy = p1[i].y - p0[i].y - 1.f; good += (dx * dx + dy * dy < 0.25f); }
        std::printf("lk_ok=%d lk_n=%zu lk_tracked=%d lk_good=%d\n", (int)okk, p1.size(), tracked, good);
        // the tracker object with its state on the device: two frames, the second from the first one's tracked points
        EORB_SLAM::b200::ELK_Tracker trk(23, 1, 10, 0.03);
        trk.setRefImage(im, kps);
        std::vector<cv::KeyPoint> tk; std::vector<int> m12, cntm; std::vector<float> disp;
        const unsigned nm1 = trk.trackAndMatchCurrImage(im2, tk, m12, cntm, disp);
        int good1 = 0;
        for (size_t i = 0; i < tk.size(); i++) if (m12[i] == (int)i) { const float dx = tk[i].pt.x - kps[i].pt.x - 2.f, dy = tk[i].pt.y - kps[i].pt.y - 1.f; good1 += (dx * dx + dy * dy < 0.25f); }
        const unsigned nm2 = trk.trackAndMatchCurrImageInit(im2, tk, m12, cntm, disp);
        int lvl0 = 0, cnt3 = 0;
        for (size_t i = 0; i < tk.size(); i++) { lvl0 += (m12[i] == (int)i && kps[i].octave == 0); cnt3 += (cntm[i] == 3); }
        std::printf("elk_n=%zu elk_nm1=%u elk_good1=%d elk_disp=%zu elk_nm2=%u elk_lvl0=%d elk_cnt3=%d elk_last=%zu\n", tk.size(), nm1, good1, disp.size(),
                    nm2, lvl0, cnt3, trk.getLastTrackedPts().size());
    }
    // ORBmatcher::SearchForInitialization, the reference signature: a frame against a copy whose keypoints moved by (3, -2)
    {
        ORB_SLAM3::Frame F1, F2;
        ORB_SLAM3::Frame::mnMinX = 0.f; ORB_SLAM3::Frame::mnMinY = 0.f; ORB_SLAM3::Frame::mnMaxX = (float)W; ORB_SLAM3::Frame::mnMaxY = (float)H;
        F1.mvKeysUn = kps; F1.mDescriptors = desc;
        F2.mvKeysUn = kps; F2.mDescriptors = desc;
        for (auto& k : F2.mvKeysUn) { k.pt.x += 3.f; k.pt.y -= 2.f; }
        std::vector<cv::Point2f> prev;
        for (auto& k : F1.mvKeysUn) prev.push_back(k.pt);
        std::vector<int> m12;
        ORB_SLAM3::ORBmatcher matcher(0.9f, true);
        const int nm = matcher.SearchForInitialization(F1, F2, prev, m12, 100);
        int self = 0, lvl0 = 0, moved = 0;
        for (size_t i = 0; i < m12.size(); i++) {
            lvl0 += (kps[i].octave == 0);
            self += (m12[i] == (int)i);
            if (m12[i] >= 0) moved += (prev[i].x == F2.mvKeysUn[m12[i]].pt.x && prev[i].y == F2.mvKeysUn[m12[i]].pt.y);
        }
        std::printf("sfi_nm=%d sfi_self=%d sfi_lvl0=%d sfi_prev=%d\n", nm, self, lvl0, moved);
        // ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono): every last-frame keypoint carries a map point 4 m in
        // front of the camera that projects onto the keypoint's position in the current (shifted) frame; identity pose
        ORB_SLAM3::GeometricCamera cam2({458.654f, 457.296f, 367.215f, 248.375f});
        std::vector<ORB_SLAM3::MapPoint*> mps;
        F1.mvpMapPoints.assign(kps.size(), nullptr); F1.mvbOutlier.assign(kps.size(), false);
        for (size_t i = 0; i < kps.size(); i++) {
            cv::Mat pos(3, 1, CV_32F);
            pos.at<float>(2, 0) = 4.f;
            pos.at<float>(0, 0) = (F2.mvKeysUn[i].pt.x - 367.215f) / 458.654f * 4.f;
            pos.at<float>(1, 0) = (F2.mvKeysUn[i].pt.y - 248.375f) / 457.296f * 4.f;
            mps.push_back(new ORB_SLAM3::MapPoint(pos, desc.row((int)i), 2));
            F1.mvpMapPoints[i] = mps.back();
        }
        F2.mvpMapPoints.assign(kps.size(), nullptr); F2.mvbOutlier.assign(kps.size(), false);
        F2.mTcw = cv::Mat::zeros(4, 4, CV_32F);
        for (int i = 0; i < 4; i++) F2.mTcw.at<float>(i, i) = 1.f;
        F2.mpCamera = &cam2;
        F2.mvScaleFactors.assign(8, 1.f);
        for (int i = 1; i < 8; i++) F2.mvScaleFactors[i] = F2.mvScaleFactors[i - 1] * 1.2f;
        const int np = matcher.SearchByProjection(F2, F1, 15.f, true);
        int same = 0, set = 0;
        for (size_t i = 0; i < kps.size(); i++) { set += (F2.mvpMapPoints[i] != nullptr); same += (F2.mvpMapPoints[i] == F1.mvpMapPoints[i]); }
        std::printf("sbp_nm=%d sbp_set=%d sbp_same=%d\n", np, set, same);
        // ORBmatcher::SearchByProjection(F, vpMapPoints, th, ...): the same map points as a local map, tracking fields set as
        // Frame::isInFrustum would (projection = the keypoint, predicted level = its octave), into an empty frame
        F2.mvpMapPoints.assign(kps.size(), nullptr);
        for (size_t i = 0; i < kps.size(); i++) {
            mps[i]->mbTrackInView = true; mps[i]->mTrackProjX = F2.mvKeysUn[i].pt.x; mps[i]->mTrackProjY = F2.mvKeysUn[i].pt.y;
            mps[i]->mnTrackScaleLevel = F2.mvKeysUn[i].octave; mps[i]->mTrackViewCos = 0.9995f; mps[i]->mTrackDepth = 4.f;
        }
        ORB_SLAM3::ORBmatcher matcher08(0.8f);
        const int nl = matcher08.SearchByProjection(F2, mps, 3.f, false, 50.f);
        int lset = 0, lsame = 0;
        for (size_t i = 0; i < kps.size(); i++) { lset += (F2.mvpMapPoints[i] != nullptr); lsame += (F2.mvpMapPoints[i] == mps[i]); }
        std::printf("slp_nm=%d slp_set=%d slp_same=%d\n", nl, lset, lsame);
        // rectified stereo: every keypoint has a right column consistent with its depth (ur = u - mbf / z), so the right-column test
        // passes everywhere and the result equals the monocular one; level window = forward (the last frame sits 1 m behind)
        F2.mvpMapPoints.assign(kps.size(), nullptr);
        F2.mbf = 40.f; F2.mb = 0.1f;
        F2.mvuRight.assign(kps.size(), -1.f);
        for (size_t i = 0; i < kps.size(); i++) F2.mvuRight[i] = F2.mvKeysUn[i].pt.x - 40.f / 4.f;
        F1.mTcw = F2.mTcw.clone();
        const int ns = matcher.SearchByProjection(F2, F1, 15.f, false);
        int sset = 0;
        for (size_t i = 0; i < kps.size(); i++) sset += (F2.mvpMapPoints[i] != nullptr);
        std::printf("sbs_nm=%d sbs_set=%d\n", ns, sset);
        // ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist): the same points seen from a keyframe
        {
            ORB_SLAM3::KeyFrame KF;
            KF.mvKeysUn = kps; KF.mDescriptors = desc; KF.mvpMapPoints = mps;
            for (size_t i = 0; i < kps.size(); i++) mps[i]->mnPredictedLevel = F2.mvKeysUn[i].octave;
            F2.mvpMapPoints.assign(kps.size(), nullptr);
            std::set<ORB_SLAM3::MapPoint*> found;
            if (!mps.empty()) found.insert(mps[0]);
            const int nr = matcher.SearchByProjection(F2, &KF, found, 10.f, 100);
            int rset = 0, rsame = 0;
            for (size_t i = 0; i < kps.size(); i++) { rset += (F2.mvpMapPoints[i] != nullptr); rsame += (F2.mvpMapPoints[i] == mps[i]); }
            std::printf("sbr_nm=%d sbr_set=%d sbr_same=%d sbr_skipped_found=%d\n", nr, rset, rsame, (int)(mps.empty() || F2.mvpMapPoints[0] != mps[0]));
        }
        // the keyframe-side searches of local mapping / loop closing (ORBmatcher_kf_b200.cc): the same map points against a keyframe at the
        // identity pose whose keypoints are F2's; every point projects onto "its" keypoint and carries that keypoint's descriptor
        {
            ORB_SLAM3::KeyFrame KA, KB;
            for (ORB_SLAM3::KeyFrame* K : {&KA, &KB}) {
                K->mvKeysUn = F2.mvKeysUn; K->mDescriptors = desc; K->mvpMapPoints.assign(kps.size(), nullptr);
                K->fx = 458.654f; K->fy = 457.296f; K->cx = 367.215f; K->cy = 248.375f; K->mpCamera = &cam2;
                K->mnMinX = 0; K->mnMinY = 0; K->mnMaxX = W; K->mnMaxY = H;
                K->mvScaleFactors = F2.mvScaleFactors; K->mvInvLevelSigma2.clear();
                for (float sf : F2.mvScaleFactors) K->mvInvLevelSigma2.push_back(1.f / (sf * sf));
                K->mvuRight.assign(kps.size(), -1.f);
                K->mRcw = cv::Mat::zeros(3, 3, CV_32F); for (int i = 0; i < 3; i++) K->mRcw.at<float>(i, i) = 1.f;
                K->mtcw = cv::Mat::zeros(3, 1, CV_32F); K->mOw = cv::Mat::zeros(3, 1, CV_32F);
            }
            for (size_t i = 0; i < kps.size(); i++) {
                cv::Mat pos = mps[i]->GetWorldPos(), nrm(3, 1, CV_32F);
                const float l = std::sqrt(pos.at<float>(0, 0) * pos.at<float>(0, 0) + pos.at<float>(1, 0) * pos.at<float>(1, 0) + pos.at<float>(2, 0) * pos.at<float>(2, 0));
                for (int r = 0; r < 3; r++) nrm.at<float>(r, 0) = pos.at<float>(r, 0) / l;
                mps[i]->mNormal = nrm; mps[i]->mnPredictedLevel = F2.mvKeysUn[i].octave;
            }
            cv::Mat I4 = cv::Mat::zeros(4, 4, CV_32F); for (int i = 0; i < 4; i++) I4.at<float>(i, i) = 1.f;
            std::vector<ORB_SLAM3::MapPoint*> vpMatched(kps.size(), nullptr);
            const int k1n = matcher.SearchByProjection(&KA, I4, mps, vpMatched, 5, 1.0f);
            int k1same = 0; for (size_t i = 0; i < kps.size(); i++) k1same += (vpMatched[i] != nullptr && vpMatched[i] == mps[i]);
            std::vector<ORB_SLAM3::MapPoint*> vpMatched2(kps.size(), nullptr);
            std::vector<ORB_SLAM3::KeyFrame*> vpKFs(kps.size(), &KB), vpMatchedKF(kps.size(), nullptr);
            const int k2n = matcher.SearchByProjection(&KA, I4, mps, vpKFs, vpMatched2, vpMatchedKF, 5, 1.0f);
            int k2kf = 0; for (size_t i = 0; i < kps.size(); i++) k2kf += (vpMatched2[i] != nullptr && vpMatchedKF[i] == &KB);
            std::printf("kfp_nm=%d kfp_same=%d kfq_nm=%d kfq_kf=%d\n", k1n, k1same, k2n, k2kf);
            const int f1n = matcher.Fuse(&KA, mps, 3.0f, false);               // empty keyframe: every fusion adds an observation
            int f1add = 0; for (size_t i = 0; i < kps.size(); i++) f1add += (KA.mvpMapPoints[i] != nullptr && mps[i]->mnAddedObs >= 0);
            ORB_SLAM3::MapPoint other(mps.empty() ? cv::Mat() : mps[0]->GetWorldPos(), mps.empty() ? cv::Mat() : mps[0]->GetDescriptor(), 9);
            KB.mvpMapPoints.assign(kps.size(), &other);                        // occupied keyframe: every fusion proposes a replacement
            std::vector<ORB_SLAM3::MapPoint*> vpReplace(kps.size(), nullptr);
            const int f2n = matcher.Fuse(&KB, I4, mps, 3.0f, vpReplace);
            int f2rep = 0; for (size_t i = 0; i < kps.size(); i++) f2rep += (vpReplace[i] == &other);
            std::printf("fuse_n=%d fuse_add=%d fuse2_n=%d fuse2_rep=%d\n", f1n, f1add, f2n, f2rep);
            KA.mvpMapPoints = mps; KB.mvpMapPoints = mps;                      // two keyframes seeing the same points, identity relative pose
            cv::Mat R12 = cv::Mat::zeros(3, 3, CV_32F), t12 = cv::Mat::zeros(3, 1, CV_32F); for (int i = 0; i < 3; i++) R12.at<float>(i, i) = 1.f;
            std::vector<ORB_SLAM3::MapPoint*> vp12(kps.size(), nullptr);
            const float s12 = 1.f;
            const int s3n = matcher.SearchBySim3(&KA, &KB, vp12, s12, R12, t12, 7.5f);
            int s3same = 0; for (size_t i = 0; i < kps.size(); i++) s3same += (vp12[i] != nullptr && vp12[i] == mps[i]);
            std::printf("sim3_n=%d sim3_same=%d\n", s3n, s3same);
            // SearchForTriangulation: the first keyframe sees the unshifted keypoints, the second one is moved sideways along the (3, -2) image
            // flow, so every feature lies on its epipolar line; all features under one vocabulary node, no map points yet
            KA.mvKeysUn = kps;
            KA.mvpMapPoints.assign(kps.size(), nullptr); KB.mvpMapPoints.assign(kps.size(), nullptr);
            KB.mtcw.at<float>(0, 0) = -0.03f; KB.mtcw.at<float>(1, 0) = 0.02f;
            KB.mOw.at<float>(0, 0) = 0.03f; KB.mOw.at<float>(1, 0) = -0.02f;
            std::vector<unsigned int> all;
            for (size_t i = 0; i < kps.size(); i++) all.push_back((unsigned int)i);
            KA.mFeatVec.clear(); KB.mFeatVec.clear();
            if (!all.empty()) { KA.mFeatVec[7] = all; KB.mFeatVec[7] = all; }
            std::vector<std::pair<size_t, size_t> > pairs;
            const int tn = matcher.SearchForTriangulation(&KA, &KB, cv::Mat(), pairs, false, false);
            int tsame = 0; for (auto& pr : pairs) tsame += (pr.first == pr.second);
            std::printf("tri_nm=%d tri_pairs=%zu tri_same=%d\n", tn, pairs.size(), tsame);
        }
        for (auto* p : mps) delete p;
    }
    // ORBVocabulary::transform through a text file in ORB-SLAM's vocabulary format (k = 3, L = 2: 3 inner nodes, 9 words whose
    // descriptors are rows of `desc`), then Frame::UndistortKeyPoints with the EuRoC cam0 calibration
    {
        const char* path = argc > 1 ? argv[1] : "/tmp/_eorb_shim_voc.txt";
        int nw = 0;
        if (desc.rows >= 12) {
            FILE* fp = std::fopen(path, "w");
            std::fprintf(fp, "3 2 0 0\n");
            for (int c = 0; c < 3; c++) {
                std::fprintf(fp, "0 0"); for (int b = 0; b < 32; b++) std::fprintf(fp, " %d", desc.at<unsigned char>(c * 4, b)); std::fprintf(fp, " 0\n");
            }
            for (int c = 0; c < 3; c++) for (int w = 0; w < 3; w++) {
                std::fprintf(fp, "%d 1", c + 1); for (int b = 0; b < 32; b++) std::fprintf(fp, " %d", desc.at<unsigned char>(c * 4 + w + 1, b)); std::fprintf(fp, " %.6f\n", 1.0 + 0.25 * nw++);
            }
            std::fclose(fp);
        }
        ORB_SLAM3::ORBVocabularyB200 voc;
        const bool okv = desc.rows >= 12 && voc.loadFromTextFile(path);
        std::vector<cv::Mat> feats;
        for (int i = 0; i < desc.rows; i++) feats.push_back(desc.row(i));
        DBoW2::BowVector bv; DBoW2::FeatureVector fv;
        voc.transform(feats, bv, fv, 1);
        double sum = 0; size_t nf = 0;
        for (auto& e : bv) sum += e.second;
        for (auto& e : fv) nf += e.second.size();
        std::printf("voc_ok=%d voc_words=%u bow=%zu bow_sum=%.9f fv_nodes=%zu fv_feats=%zu\n", (int)okv, voc.size(), bv.size(), sum, fv.size(), nf);
        // ORBmatcher::SearchByBoW(pKF, F, matches): the frame matched against a keyframe made of the same features (every one with
        // a map point) through the FeatureVector computed above: every feature that passes the ratio test finds itself
        {
            ORB_SLAM3::KeyFrame KF; ORB_SLAM3::Frame Fb;
            KF.mvKeysUn = kps; KF.mDescriptors = desc; KF.mFeatVec = fv;
            Fb.mvKeysUn = kps; Fb.mDescriptors = desc; Fb.mFeatVec = fv;
            std::vector<ORB_SLAM3::MapPoint*> own;
            cv::Mat pos = cv::Mat::zeros(3, 1, CV_32F);
            for (int i = 0; i < desc.rows; i++) { own.push_back(new ORB_SLAM3::MapPoint(pos, desc.row(i), 1)); }
            KF.mvpMapPoints = own;
            std::vector<ORB_SLAM3::MapPoint*> mm;
            ORB_SLAM3::ORBmatcher m07(0.7f, true);
            const int nb = okv ? m07.SearchByBoW(&KF, Fb, mm) : 0;
            int bset = 0, bself = 0;
            for (size_t i = 0; i < mm.size(); i++) { bset += (mm[i] != nullptr); bself += (mm[i] != nullptr && mm[i] == own[i]); }
            std::printf("sbb_nm=%d sbb_set=%d sbb_self=%d\n", nb, bset, bself);
            // the keyframe-keyframe form on the same data: every feature matches itself, the result is indexed by the first keyframe
            ORB_SLAM3::KeyFrame KF2; KF2.mvKeysUn = kps; KF2.mDescriptors = desc; KF2.mFeatVec = fv; KF2.mvpMapPoints = own;
            std::vector<ORB_SLAM3::MapPoint*> m12;
            const int nk = okv ? m07.SearchByBoW(&KF, &KF2, m12) : 0;
            int kset = 0, kself = 0;
            for (size_t i = 0; i < m12.size(); i++) { kset += (m12[i] != nullptr); kself += (m12[i] != nullptr && m12[i] == own[i]); }
            std::printf("sbk_nm=%d sbk_set=%d sbk_self=%d\n", nk, kset, kself);
            for (auto* p : own) delete p;
        }
        cv::Mat Kc = cv::Mat::zeros(3, 3, CV_32F), Dc = cv::Mat::zeros(4, 1, CV_32F);
        Kc.at<float>(0, 0) = 458.654f; Kc.at<float>(1, 1) = 457.296f; Kc.at<float>(0, 2) = 367.215f; Kc.at<float>(1, 2) = 248.375f; Kc.at<float>(2, 2) = 1.f;
        Dc.at<float>(0, 0) = -0.28340811f; Dc.at<float>(1, 0) = 0.07395907f; Dc.at<float>(2, 0) = 0.00019359f; Dc.at<float>(3, 0) = 1.76187114e-05f;
        std::vector<cv::KeyPoint> un;
        const bool oku = ORB_SLAM3::b200::UndistortKeyPoints(kps, Kc, Dc, un);
        double shift = 0;
        for (size_t i = 0; i < un.size() && i < kps.size(); i++) shift += std::fabs(un[i].pt.x - kps[i].pt.x) + std::fabs(un[i].pt.y - kps[i].pt.y);
        std::printf("undist_ok=%d undist_n=%zu undist_shift=%.3f\n", (int)(oku && !kps.empty()), un.size(), shift);
    }
    return 0;
}
