// Drop-in replacement for the reference's include/ORBextractor.h (:33-47 ORBxParams, :62-136 ORBextractor):
// same namespace, class name, public methods and public data member, forwarding to the C ABI of libeorb_b200.so.
// Tracking / Frame / MixedFrame / EvFrame / EvBaseTracker link against it unchanged.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <vector>
#include <list>
#include <opencv2/core/core.hpp>

struct eorb_orb;

namespace ORB_SLAM3
{
#define DEF_STD_GAUSS_KER 2
#define DEF_DESC_LEN 32
#define DEF_IMAGE_WIDTH 752

struct ORBxParams {
    ORBxParams() : nfeatures(0), scaleFactor(1), nlevels(1), iniThFAST(10), minThFAST(7), edgeTh(19), patchSize(31) {}
    ORBxParams(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST);
    ORBxParams(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST, int _edgeTh, const cv::Size& imSz);

    int nfeatures;
    float scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    int edgeTh;
    int patchSize;
    cv::Size imSize;
};

class ORBextractor
{
public:
    ORBextractor(const ORBxParams& paramsORB);
    ~ORBextractor();
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image (mask is ignored, as in the reference).
    int operator()( cv::InputArray _image, cv::InputArray _mask,
                    std::vector<cv::KeyPoint>& _keypoints,
                    cv::OutputArray _descriptors, std::vector<int> &vLappingArea);

    // Only detect ORB features (no descriptors)
    int operator()( cv::InputArray _image, cv::InputArray _mask,
                    std::vector<cv::KeyPoint>& _keypoints, std::vector<int> &vLappingArea);

    int inline GetLevels() const { return nlevels; }
    float inline GetScaleFactor() const { return static_cast<float>(scaleFactor); }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Filled lazily (device -> host) after every call; Frame::ComputeStereoMatches reads it (Frame.cc:876,966-985).
    std::vector<cv::Mat> mvImagePyramid;

    int GetNumFeatures() const { return nfeatures; }

    void AssignKPtLevelByBestDesc(const cv::Mat& refDescs, const cv::Mat& trackedImage, std::vector<cv::KeyPoint>& trackedKPts);
    void ComputeTrackedKPtsDesc(const cv::Mat& trackedImage, const std::vector<cv::KeyPoint>& trackedKPts, cv::Mat& refDescs);

    // When false, mvImagePyramid is not downloaded after operator() (monocular callers never read it).
    bool mbDownloadPyramid = true;

protected:
    int extract(cv::InputArray image, std::vector<cv::KeyPoint>& kps, cv::Mat* desc, std::vector<int>& lap);
    void downloadPyramid();

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    eorb_orb* mpHandle;
    std::vector<unsigned char> mScratchDesc;
};

} //namespace ORB_SLAM3

#endif
