// oracle/ref_mock/ORBmatcher.h — TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's include/ORBmatcher.h, found first on the include path when src/ORBextractor.cc (compiled unmodified
// into oracle/_ref) says #include "ORBmatcher.h".  The real header drags in g2o / DBoW2 / the whole map; this one declares, with the
// reference's signatures (include/ORBmatcher.h:35-114), the members whose BODIES the Makefile cuts out of src/ORBmatcher.cc at build
// time: DescriptorDistance :2360-2378 (_ref/gen_descriptor_distance.inc) and the tracking-thread searches, RadiusByViewingCos and
// ComputeThreeMaxima (_ref/gen_matcher.inc: :44-227, :276-478, :714-831, :1969-2187, :2189-2312, :2314-2355).
#pragma once
#include <set>
#include <vector>
#include <opencv2/core/core.hpp>
#include "Frame.h"

namespace ORB_SLAM3 {
class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3, const bool bFarPoints = false,
                           const float thFarPoints = 50.0f);
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist);
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);
    // keyframe-side searches of local mapping / loop closing (include/ORBmatcher.h:56-100), bodies cut into _ref/gen_matcher_kf.inc
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th,
                           float ratioHamming = 1.0);
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, const std::vector<KeyFrame*>& vpPointsKFs,
                           std::vector<MapPoint*>& vpMatched, std::vector<KeyFrame*>& vpMatchedKF, int th, float ratioHamming = 1.0);
    int Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th = 3.0, const bool bRight = false);
    int Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint);
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12, const cv::Mat& t12,
                     const float th);
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo,
                               const bool bCoarse = false);
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10);
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
protected:
    float RadiusByViewingCos(const float& viewCos);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
    float mfNNratio;
    bool mbCheckOrientation;
};
}  // namespace ORB_SLAM3
