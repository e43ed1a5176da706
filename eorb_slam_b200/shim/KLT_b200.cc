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
} // namespace b200
} // namespace EORB_SLAM
