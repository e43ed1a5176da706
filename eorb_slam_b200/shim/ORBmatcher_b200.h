// B200 replacement for the Hamming hot loop of the reference's ORBmatcher (src/ORBmatcher.cc):
//   * ORBmatcher::DescriptorDistance  (:2360-2378)  — same signature, defined in ORBmatcher_b200.cc; drop the
//     reference's definition and link this one (MapPoint.cc:397, Frame.cc:944, MixedMatcher.cpp, ORBextractor.cc:1306
//     keep calling it unchanged);
//   * BruteForceBest2: the "for each query, scan candidates keeping best/second-best with strict '<', accept iff
//     best <= TH_LOW and best < ratio*second" loop that every Search* bottoms out in (:318-382, :741-772), in the
//     brute-force shape of Frame::ComputeStereoFishEyeMatches (Frame.cc:1225-1235), plus the rotation-histogram
//     consistency filter (:784-823, ComputeThreeMaxima :2314-2355).
#ifndef ORBMATCHER_B200_H
#define ORBMATCHER_B200_H

#include <vector>
#include <opencv2/core/core.hpp>

struct eorb_matcher;

namespace ORB_SLAM3
{
class BruteForceBest2
{
public:
    explicit BruteForceBest2(float nnratio = 0.6f, bool checkOri = true, int device = 0);
    ~BruteForceBest2();
    BruteForceBest2(const BruteForceBest2&) = delete;
    BruteForceBest2& operator=(const BruteForceBest2&) = delete;

    // trainDescs: N x 32 CV_8U (e.g. Frame::mDescriptors / KeyFrame::mDescriptors); stays resident in HBM
    bool SetTrainDescriptors(const cv::Mat& trainDescs);
    // queryDescs: M x 32 CV_8U.  vnMatches12[i] = matched train row or -1.  Angles (degrees, as in cv::KeyPoint)
    // are only used when checkOri is set.  Returns the number of matches, like the Search* methods.
    int Match(const cv::Mat& queryDescs, const std::vector<cv::KeyPoint>& queryKps, const std::vector<cv::KeyPoint>& trainKps,
              std::vector<int>& vnMatches12, std::vector<int>* bestDist = nullptr, int th = 50 /*TH_LOW*/);

protected:
    float mfNNratio;
    bool mbCheckOrientation;
    eorb_matcher* mpHandle;
};
} // namespace ORB_SLAM3
#endif
