#include "KLT_b200.h"

#include <cstdio>

#include "eorb_b200.h"

namespace EORB_SLAM
{
namespace b200
{
namespace
{
// one tracker workspace per calling thread (the reference tracks from the L1 and L2 event threads concurrently)
struct Workspace {
    eorb_lk* h = nullptr;
    int w = 0, hgt = 0, cap = 0, device = -1;
    ~Workspace() { if (h) eorb_lk_destroy(h); }
};
thread_local Workspace tls;

bool ensure(int device, int w, int hgt, int n)
{
    if (tls.h && tls.device == device && w <= tls.w && hgt <= tls.hgt && n <= tls.cap) return true;
    if (tls.h) { eorb_lk_destroy(tls.h); tls.h = nullptr; }
    const int cap = n < 1024 ? 1024 : 2 * n;
    if (eorb_lk_create(device, w, hgt, cap, &tls.h) != EORB_OK) {
        std::fprintf(stderr, "b200::calcOpticalFlowPyrLK: %s\n", eorb_last_error());
        tls.h = nullptr;
        return false;
    }
    tls.w = w; tls.hgt = hgt; tls.cap = cap; tls.device = device;
    return true;
}
} // namespace

bool calcOpticalFlowPyrLK(const cv::Mat& prevImg, const cv::Mat& nextImg, const std::vector<cv::Point2f>& prevPts,
                          std::vector<cv::Point2f>& nextPts, std::vector<unsigned char>& status, std::vector<float>& err,
                          int winSize, int maxLevel, int maxCount, double epsilon, bool useInitialFlow, int device)
{
    const int n = (int)prevPts.size();
    status.clear(); err.clear();
    if (prevImg.empty() || nextImg.empty() || n == 0 || prevImg.rows != nextImg.rows || prevImg.cols != nextImg.cols ||
        prevImg.type() != CV_8UC1 || nextImg.type() != CV_8UC1 || (useInitialFlow && (int)nextPts.size() != n)) {
        std::fprintf(stderr, "b200::calcOpticalFlowPyrLK: bad arguments\n");
        if (!useInitialFlow) nextPts.clear();
        return false;
    }
    if (!ensure(device, prevImg.cols, prevImg.rows, n)) { if (!useInitialFlow) nextPts.clear(); return false; }
    static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "cv::Point2f must be two packed floats");
    int rc = eorb_lk_set_ref(tls.h, prevImg.data, prevImg.cols, prevImg.rows, prevImg.step, &prevPts[0].x, n, winSize, maxLevel);
    if (rc == EORB_OK) {
        std::vector<cv::Point2f> init;
        if (useInitialFlow) init = nextPts;
        nextPts.resize(n); status.resize(n); err.resize(n);
        rc = eorb_lk_track(tls.h, nextImg.data, nextImg.step, useInitialFlow ? &init[0].x : nullptr, maxCount, epsilon, 1e-4f, &nextPts[0].x,
                           status.data(), err.data());
    }
    if (rc < 0) {
        std::fprintf(stderr, "b200::calcOpticalFlowPyrLK: %s\n", eorb_last_error());
        status.clear(); err.clear();
        if (!useInitialFlow) nextPts.clear();
        return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------- b200::ELK_Tracker
ELK_Tracker::ELK_Tracker(int kltWinSize, int maxLevel, int kltMaxItr, double kltEps, int device)
    : mHandle(nullptr), mCapW(0), mCapH(0), mCapN(0), mDevice(device), mPatchSz(kltWinSize), mMaxLevel(maxLevel), mMaxItr(kltMaxItr),
      mEps(kltEps) {}   // no GPU work here: the handle is created by the first setRefImage

ELK_Tracker::~ELK_Tracker() { if (mHandle) eorb_lk_destroy((eorb_lk*)mHandle); }

bool ELK_Tracker::ensure(int w, int hgt, int n)
{
    if (mHandle && w <= mCapW && hgt <= mCapH && n <= mCapN) return true;
    if (mHandle) { eorb_lk_destroy((eorb_lk*)mHandle); mHandle = nullptr; }
    const int cap = n < 1024 ? 1024 : 2 * n;
    eorb_lk* h = nullptr;
    if (eorb_lk_create(mDevice, w, hgt, cap, &h) != EORB_OK) {
        std::fprintf(stderr, "b200::ELK_Tracker: %s\n", eorb_last_error());
        return false;
    }
    mHandle = h; mCapW = w; mCapH = hgt; mCapN = cap;
    return true;
}

void ELK_Tracker::setRefImage(const cv::Mat& image, const std::vector<cv::KeyPoint>& refPts)
{
    static_assert(sizeof(cv::KeyPoint) == sizeof(eorb_keypoint), "cv::KeyPoint must be the 28-byte record");
    mRefKPoints.clear(); mLastTrackedKPts.clear();
    if (image.empty() || refPts.empty() || image.type() != CV_8UC1) {   // the reference asserts (KLT_Tracker.cpp:24)
        std::fprintf(stderr, "b200::ELK_Tracker::setRefImage: bad arguments\n");
        return;
    }
    if (!ensure(image.cols, image.rows, (int)refPts.size())) return;
    const int rc = eorb_lk_set_ref_keypoints((eorb_lk*)mHandle, image.data, image.cols, image.rows, image.step, 0,
                                             reinterpret_cast<const eorb_keypoint*>(refPts.data()), (int)refPts.size(), mPatchSz, mMaxLevel);
    if (rc != EORB_OK) { std::fprintf(stderr, "b200::ELK_Tracker::setRefImage: %s\n", eorb_last_error()); return; }
    mRefFrame = image.clone();
    mRefKPoints = refPts; mLastTrackedKPts = refPts;
}

void ELK_Tracker::setLastTrackedPts(const std::vector<cv::KeyPoint>& currTrackedPts)
{
    mLastTrackedKPts = currTrackedPts;
    if (mHandle && !mRefKPoints.empty())
        if (eorb_lk_set_last_tracked((eorb_lk*)mHandle, reinterpret_cast<const eorb_keypoint*>(currTrackedPts.data()), (int)currTrackedPts.size()) != EORB_OK)
            std::fprintf(stderr, "b200::ELK_Tracker::setLastTrackedPts: %s\n", eorb_last_error());
}

unsigned ELK_Tracker::run(const cv::Mat& image, std::vector<cv::KeyPoint>& trackedKPts, std::vector<int>& vMatches12,
                          std::vector<int>& vCntMatches, std::vector<float>& vPxDisp, bool firstOctaveOnly)
{
    if (mRefKPoints.empty() || mRefFrame.empty() || !mHandle) {   // KLT_Tracker.cpp:218-221
        std::fprintf(stderr, "trackAndMatchCurrImage: No reference info., did you forget to init. tracker??\n");
        return 0;
    }
    const size_t n = mRefKPoints.size();
    if (image.empty() || image.type() != CV_8UC1 || image.cols != mRefFrame.cols || image.rows != mRefFrame.rows) {
        std::fprintf(stderr, "b200::ELK_Tracker: bad image\n");
        return 0;
    }
    std::vector<unsigned char> matched(n);
    std::vector<float> disp(n);
    int counts[2] = {0, 0};
    trackedKPts.resize(n);
    const int rc = eorb_lk_track_and_match((eorb_lk*)mHandle, image.data, image.step, 0, mMaxItr, mEps, 1e-4f, firstOctaveOnly ? 1 : 0,
                                           reinterpret_cast<eorb_keypoint*>(trackedKPts.data()), matched.data(), disp.data(), counts);
    if (rc < 0) {
        std::fprintf(stderr, "b200::ELK_Tracker: %s\n", eorb_last_error());
        trackedKPts.clear();
        return 0;
    }
    // the vectors the reference updates in place (:113-133, 140-147, 166-176)
    if (vCntMatches.empty()) vCntMatches.resize(n, 1);
    vPxDisp.assign(disp.begin(), disp.begin() + counts[1]);
    vMatches12.resize(n, -1);
    unsigned nMatches = 0;
    for (size_t i = 0; i < n; i++)
        if (matched[i] & 1) { vCntMatches[i]++; vMatches12[i] = (int)i; nMatches++; }
    if (firstOctaveOnly)
        for (size_t i = 0; i < n; i++)
            if (vMatches12[i] >= 0 && vMatches12[i] < (int)n && mRefKPoints[i].octave > 0) { vMatches12[i] = -1; vCntMatches[i]--; nMatches--; }
    mLastTrackedKPts = trackedKPts;
    return nMatches;
}

unsigned ELK_Tracker::trackAndMatchCurrImage(const cv::Mat& currImage, std::vector<cv::KeyPoint>& trackedKPts, std::vector<int>& vMatches12,
                                             std::vector<int>& vCntMatches, std::vector<float>& vPxDisp)
{
    return run(currImage, trackedKPts, vMatches12, vCntMatches, vPxDisp, false);
}

unsigned ELK_Tracker::trackAndMatchCurrImage(const cv::Mat& currImage, std::vector<cv::KeyPoint>& trackedKPts, std::vector<int>& vMatches12)
{
    std::vector<int> vCntMatches;
    std::vector<float> vPxDisp;
    return run(currImage, trackedKPts, vMatches12, vCntMatches, vPxDisp, false);
}

unsigned ELK_Tracker::trackAndMatchCurrImageInit(const cv::Mat& currImage, std::vector<cv::KeyPoint>& trackedKPts, std::vector<int>& vMatches12,
                                                 std::vector<int>& vCntMatches, std::vector<float>& vPxDisp)
{
    return run(currImage, trackedKPts, vMatches12, vCntMatches, vPxDisp, true);
}
} // namespace b200
} // namespace EORB_SLAM
