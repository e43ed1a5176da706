// B200 replacement for the one OpenCV call of the reference's ELK_Tracker (src/Event/KLT_Tracker.cpp:63-70, 86-88):
//
//   cv::calcOpticalFlowPyrLK(mRefFrame, currImage, mRefPoints, kpts, status, err, Size(mPatchSz, mPatchSz), mMaxLevel,
//                            mLKCriteria [, OPTFLOW_USE_INITIAL_FLOW]);
//
// becomes (same containers, the TermCriteria fields passed as numbers):
//
//   EORB_SLAM::b200::calcOpticalFlowPyrLK(mRefFrame, currImage, mRefPoints, kpts, status, err, mPatchSz, mMaxLevel,
//                                         mLKCriteria.maxCount, mLKCriteria.epsilon, /*useInitialFlow=*/true|false);
//
// ELK_Tracker's own members stay as they are — or, to keep the tracker's state on the device between event frames, ELK_Tracker itself
// is replaced by b200::ELK_Tracker below: the same public members (include/Event/KLT_Tracker.h:27-85), the reference frame's pyramid,
// the reference keypoints and the last tracked points resident in HBM, refineTrackedPts / refineFirstOctaveLevel
// (KLT_Tracker.cpp:105-183) on the device, one device-to-host copy per frame.
#ifndef KLT_B200_H
#define KLT_B200_H

#include <vector>
#include <opencv2/core/core.hpp>

namespace EORB_SLAM
{
namespace b200
{
// prevImg / nextImg: CV_8UC1 of equal size.  nextPts is read when useInitialFlow is set (it must then have
// prevPts.size() entries, KLT_Tracker.cpp:63) and resized / overwritten otherwise.  status / err are resized.
// Returns false (and logs to stderr, leaving the outputs empty) when the library reports an error — the
// reference swallows cv::Exception the same way (KLT_Tracker.cpp:72-74).
bool calcOpticalFlowPyrLK(const cv::Mat& prevImg, const cv::Mat& nextImg, const std::vector<cv::Point2f>& prevPts,
                          std::vector<cv::Point2f>& nextPts, std::vector<unsigned char>& status, std::vector<float>& err,
                          int winSize, int maxLevel, int maxCount, double epsilon, bool useInitialFlow, int device = 0);

// Drop-in for EORB_SLAM::ELK_Tracker (include/Event/KLT_Tracker.h:27-99): construct with the four numbers the reference reads from
// EvParams (kltWinSize, maxLevel, kltMaxItr, kltEps; KLT_Tracker.cpp:14-20).
class ELK_Tracker {
public:
    ELK_Tracker(int kltWinSize, int maxLevel, int kltMaxItr, double kltEps, int device = 0);
    ~ELK_Tracker();
    ELK_Tracker(const ELK_Tracker&) = delete;
    ELK_Tracker& operator=(const ELK_Tracker&) = delete;

    void setRefImage(const cv::Mat& image, const std::vector<cv::KeyPoint>& refPts);

    unsigned trackAndMatchCurrImage(const cv::Mat& currImage, std::vector<cv::KeyPoint>& trackedKPts, std::vector<int>& vMatches12);
    unsigned trackAndMatchCurrImage(const cv::Mat& currImage, std::vector<cv::KeyPoint>& trackedKPts, std::vector<int>& vMatches12,
                                    std::vector<int>& vCntMatches, std::vector<float>& vPxDisp);
    unsigned trackAndMatchCurrImageInit(const cv::Mat& currImage, std::vector<cv::KeyPoint>& trackedKPts, std::vector<int>& vMatches12,
                                        std::vector<int>& vCntMatches, std::vector<float>& vPxDisp);

    std::vector<cv::KeyPoint> getRefPoints() { return mRefKPoints; }
    void getRefImageAndPoints(cv::Mat& im0, std::vector<cv::KeyPoint>& pts0) { im0 = mRefFrame.clone(); pts0 = mRefKPoints; }
    std::vector<cv::KeyPoint> getLastTrackedPts() { return mLastTrackedKPts; }
    void setLastTrackedPts(const std::vector<cv::KeyPoint>& currTrackedPts);

private:
    unsigned run(const cv::Mat& image, std::vector<cv::KeyPoint>& trackedKPts, std::vector<int>& vMatches12, std::vector<int>& vCntMatches,
                 std::vector<float>& vPxDisp, bool firstOctaveOnly);
    bool ensure(int w, int hgt, int n);
    void* mHandle;   // eorb_lk*
    int mCapW, mCapH, mCapN, mDevice;
    cv::Mat mRefFrame;
    std::vector<cv::KeyPoint> mRefKPoints, mLastTrackedKPts;
    const int mPatchSz, mMaxLevel, mMaxItr;
    const double mEps;
};
} // namespace b200
} // namespace EORB_SLAM
#endif
