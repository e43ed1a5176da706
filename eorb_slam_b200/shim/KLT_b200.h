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
// ELK_Tracker's own members (setRefImage, refineTrackedPts, ... — plain host bookkeeping) stay as they are.
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
} // namespace b200
} // namespace EORB_SLAM
#endif
