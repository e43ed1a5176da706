// Drop-in replacement for the reference's include/ORBextractor.h (:33-47 ORBxParams, :62-136 ORBextractor):
// same namespace, class name, public methods and public data member, forwarding to the C ABI of libeorb_b200.so.
// Tracking / Frame / MixedFrame / EvFrame / EvBaseTracker link against it unchanged.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <vector>
#include <list>
#include <map>
#include <mutex>
#include <thread>
#include <opencv2/core/core.hpp>

struct eorb_orb;
struct eorb_orb_params;

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

    // Device -> host copy of the pyramid after a call, only when mbDownloadPyramid is set: the one reader is
    // Frame::ComputeStereoMatches (Frame.cc:876,966-985), i.e. rectified-stereo frames.
    std::vector<cv::Mat> mvImagePyramid;

    int GetNumFeatures() const { return nfeatures; }

    void AssignKPtLevelByBestDesc(const cv::Mat& refDescs, const cv::Mat& trackedImage, std::vector<cv::KeyPoint>& trackedKPts);
    void ComputeTrackedKPtsDesc(const cv::Mat& trackedImage, const std::vector<cv::KeyPoint>& trackedKPts, cv::Mat& refDescs);

    // ---- additions of this implementation (not in the reference header) ----------------------------------------------
    // mvImagePyramid is downloaded after operator() only when set.  Off by default: monocular, RGB-D and event frames never read
    // it (Frame.cc:327), and the copy is 1.1 MB per 752x480 call; the rectified-stereo Frame constructor sets it on its two
    // extractors.
    bool mbDownloadPyramid = false;
    // CUDA device of the handles created after the call (default: $EORB_DEVICE, else 0).  Handles are created lazily, one per
    // calling thread and extractor, on first use: constructing an extractor touches no device, and two threads may drive the same
    // extractor object at once (the reference's left / right extractors run in two threads, Frame.cc:146-149).
    static void SetDevice(int device);
    static int GetDevice();

protected:
    int extract(cv::InputArray image, std::vector<cv::KeyPoint>& kps, std::vector<unsigned char>* desc, std::vector<int>& lap);
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

    eorb_orb* handle();                                   // the calling thread's device handle (created on first use)
    eorb_orb_params* mpParams;                            // the constructor's parameters, kept for lazy handle creation
    std::mutex mHandleMutex;
    std::map<std::thread::id, eorb_orb*> mHandles;
};

} //namespace ORB_SLAM3

#endif
