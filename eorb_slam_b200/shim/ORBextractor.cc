// Forwarding implementation of the reference's ORBextractor surface (src/ORBextractor.cc) over the C ABI.
#include "ORBextractor.h"

#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "eorb_b200.h"

namespace ORB_SLAM3
{
static_assert(sizeof(cv::KeyPoint) == sizeof(eorb_keypoint), "cv::KeyPoint must be the 28-byte POD the C ABI writes");

ORBxParams::ORBxParams(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST) :
        nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST),
        minThFAST(_minThFAST), edgeTh(19), patchSize(31) {}

ORBxParams::ORBxParams(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST,
                       const int _edgeTh, const cv::Size& imSz) :
        nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST),
        minThFAST(_minThFAST), edgeTh(_edgeTh), patchSize(31), imSize(imSz) {}

static int& defaultDevice()
{
    static int dev = [] { const char* e = std::getenv("EORB_DEVICE"); return e ? std::atoi(e) : 0; }();
    return dev;
}
void ORBextractor::SetDevice(int device) { defaultDevice() = device; }
int ORBextractor::GetDevice() { return defaultDevice(); }

ORBextractor::ORBextractor(const ORBxParams& p) :
        nfeatures(p.nfeatures), scaleFactor(p.scaleFactor), nlevels(p.nlevels), iniThFAST(p.iniThFAST),
        minThFAST(p.minThFAST), mpParams(new eorb_orb_params())
{
    eorb_orb_params& cp = *mpParams;
    cp.nfeatures = p.nfeatures; cp.scaleFactor = p.scaleFactor; cp.nlevels = p.nlevels;
    cp.iniThFAST = p.iniThFAST; cp.minThFAST = p.minThFAST; cp.edgeTh = p.edgeTh;
    cp.imW = p.imSize.width; cp.imH = p.imSize.height;
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    mvImagePyramid.resize(nlevels);
    // the constructor's tables are host arithmetic (ORBextractor.cc:420-489): no device is touched here
    if (eorb_orb_params_tables(&cp, nullptr, nullptr, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                               mvInvLevelSigma2.data(), mnFeaturesPerLevel.data()) != EORB_OK)
        std::fprintf(stderr, "ORBextractor(b200): %s\n", eorb_last_error());   // the reference has no error channel here either
}

ORBextractor::~ORBextractor()
{
    for (auto& kv : mHandles) eorb_orb_destroy(kv.second);
    delete mpParams;
}

eorb_orb* ORBextractor::handle()
{
    std::lock_guard<std::mutex> lk(mHandleMutex);
    auto it = mHandles.find(std::this_thread::get_id());
    if (it != mHandles.end()) return it->second;
    eorb_orb* h = nullptr;
    if (eorb_orb_create(mpParams, defaultDevice(), 1, &h) != EORB_OK) {
        std::fprintf(stderr, "ORBextractor(b200): %s\n", eorb_last_error());
        h = nullptr;                                           // no CPU fallback: the call returns -1 like an empty image
    }
    mHandles[std::this_thread::get_id()] = h;
    return h;
}

void ORBextractor::downloadPyramid()
{
    eorb_orb* mpHandle = handle();
    if (!mpHandle) return;
    for (int l = 0; l < nlevels; l++) {
        int w = 0, h = 0;
        if (eorb_orb_level_size(mpHandle, l, &w, &h) != EORB_OK) return;
        mvImagePyramid[l].create(h, w, CV_8UC1);
        eorb_orb_pyramid_level(mpHandle, 0, l, mvImagePyramid[l].data, mvImagePyramid[l].step);
    }
}

int ORBextractor::extract(cv::InputArray _image, std::vector<cv::KeyPoint>& _keypoints, std::vector<unsigned char>* desc, std::vector<int>& lap)
{
    if (_image.empty())
        return -1;                                             // ORBextractor.cc:1096
    eorb_orb* mpHandle = handle();
    if (!mpHandle)
        return -1;
    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);                           // ORBextractor.cc:1100
    const int cap = eorb_orb_max_keypoints_for_size(mpHandle, image.cols, image.rows);
    _keypoints = std::vector<cv::KeyPoint>(cap);
    if (desc) desc->resize((size_t)cap * DEF_DESC_LEN);
    int n = 0;
    const int lap0 = lap.size() > 0 ? lap[0] : 0, lap1 = lap.size() > 1 ? lap[1] : 0;
    int ret = eorb_orb_extract(mpHandle, image.data, image.cols, image.rows, image.step, lap0, lap1, desc ? 1 : 0,
                               reinterpret_cast<eorb_keypoint*>(_keypoints.data()), desc ? desc->data() : nullptr,
                               cap, &n);
    if (ret < -1) {
        std::fprintf(stderr, "ORBextractor(b200)::operator(): %s\n", eorb_last_error());
        _keypoints.clear();
        return -1;
    }
    _keypoints.resize(n);
    if (mbDownloadPyramid) downloadPyramid();
    return ret;
}

int ORBextractor::operator()( cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors, std::vector<int> &vLappingArea)
{
    std::vector<unsigned char> scratch;                        // per call: two threads may drive one extractor
    int ret = extract(_image, _keypoints, &scratch, vLappingArea);
    if (_image.empty() || ret < 0) return ret;
    const int n = (int)_keypoints.size();
    if (n == 0) {
        _descriptors.release();                                // ORBextractor.cc:1114-1115
    } else {
        _descriptors.create(n, DEF_DESC_LEN, CV_8U);
        cv::Mat d = _descriptors.getMat();
        for (int i = 0; i < n; i++) std::memcpy(d.ptr<unsigned char>(i), &scratch[(size_t)i * DEF_DESC_LEN], DEF_DESC_LEN);
    }
    return ret;
}

int ORBextractor::operator()( cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                              std::vector<int> &vLappingArea)
{
    return extract(_image, _keypoints, nullptr, vLappingArea);
}

void ORBextractor::AssignKPtLevelByBestDesc(const cv::Mat &refDescs, const cv::Mat &trackedImage, std::vector<cv::KeyPoint> &trackedKPts)
{
    eorb_orb* mpHandle = trackedImage.empty() ? nullptr : handle();
    if (trackedImage.empty() || !mpHandle)
        return;
    const int n = (int)trackedKPts.size();
    assert(trackedImage.type() == CV_8UC1 && refDescs.rows == n);
    std::vector<unsigned char> ref((size_t)n * DEF_DESC_LEN);
    for (int i = 0; i < n; i++) std::memcpy(&ref[(size_t)i * DEF_DESC_LEN], refDescs.ptr<unsigned char>(i), DEF_DESC_LEN);
    eorb_orb_assign_level_by_best_desc(mpHandle, ref.data(), trackedImage.data, trackedImage.cols, trackedImage.rows, trackedImage.step,
                                       reinterpret_cast<eorb_keypoint*>(trackedKPts.data()), n);
}

void ORBextractor::ComputeTrackedKPtsDesc(const cv::Mat &trackedImage, const std::vector<cv::KeyPoint> &trackedKPts, cv::Mat &refDescs)
{
    eorb_orb* mpHandle = trackedImage.empty() ? nullptr : handle();
    if (trackedImage.empty() || !mpHandle)
        return;
    const int n = (int)trackedKPts.size();
    assert(trackedImage.type() == CV_8UC1);
    refDescs.create(n, DEF_DESC_LEN, CV_8U);
    std::vector<unsigned char> d((size_t)n * DEF_DESC_LEN);
    eorb_orb_tracked_desc(mpHandle, trackedImage.data, trackedImage.cols, trackedImage.rows, trackedImage.step,
                          reinterpret_cast<const eorb_keypoint*>(trackedKPts.data()), n, d.data());
    for (int i = 0; i < n; i++) std::memcpy(refDescs.ptr<unsigned char>(i), &d[(size_t)i * DEF_DESC_LEN], DEF_DESC_LEN);
}

} //namespace ORB_SLAM3
