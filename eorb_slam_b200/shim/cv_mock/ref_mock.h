// Stand-ins for the few reference types the matcher / event / guided-matching shims mention, so that they compile without the
// reference tree (which does not travel with this repository).  A real build includes the reference headers:
//   include/ORBmatcher.h:35-114, include/Event/EventData.h:36-58, include/CameraModels/GeometricCamera.h:41-134
#pragma once
#include <set>
#include <tuple>
#include <vector>
#include <opencv2/core/core.hpp>
#include "../ORBVocabulary_b200.h"   // the DBoW2::FeatureVector stand-in (a real build has Thirdparty/DBoW2)

namespace ORB_SLAM3 {
class GeometricCamera {
public:
    explicit GeometricCamera(std::vector<float> p) : mvParameters(std::move(p)), mnType(mvParameters.size() >= 8 ? 1u : 0u) {}
    float getParameter(const int i) { return mvParameters[i]; }   // GeometricCamera.h:133
    virtual cv::Point2f project(const cv::Point3f& p) {           // GeometricCamera.h:94; the stand-in is the pinhole model (Pinhole.cpp:30-33)
        return cv::Point2f(mvParameters[0] * p.x / p.z + mvParameters[2], mvParameters[1] * p.y / p.z + mvParameters[3]);
    }
    virtual ~GeometricCamera() {}
    unsigned int GetType() { return mnType; }                     // GeometricCamera.h:144-147
    const unsigned int CAM_PINHOLE = 0;
    const unsigned int CAM_FISHEYE = 1;
protected:
    std::vector<float> mvParameters;   // fx, fy, cx, cy (Pinhole) [+ k1..k4 (KannalaBrandt8)]
    unsigned int mnType;
};

class MapPoint {   // what SearchByProjection reads (include/MapPoint.h: GetWorldPos, GetDescriptor, Observations)
public:
    MapPoint(const cv::Mat& pos, const cv::Mat& desc, int nObs) : mWorldPos(pos), mDescriptor(desc), mnObs(nObs) {}
    cv::Mat GetWorldPos() { return mWorldPos; }
    cv::Mat GetDescriptor() { return mDescriptor; }
    int Observations() { return mnObs; }
    bool isBad() { return mbBad; }
    // tracking fields Frame::isInFrustum fills (include/MapPoint.h:133-144)
    float mTrackProjX = 0.f, mTrackProjY = 0.f, mTrackDepth = 0.f, mTrackViewCos = 1.f, mTrackProjXR = 0.f;
    int mnTrackScaleLevel = 0;
    // what the relocalisation search reads (include/MapPoint.h:104-106)
    float GetMinDistanceInvariance() { return mfMinDistance; }
    float GetMaxDistanceInvariance() { return mfMaxDistance; }
    int PredictScale(const float& currentDist, class Frame* pF);
    // what the keyframe-side searches read and call (include/MapPoint.h:100-150); the map-graph calls are recorded by the stand-in
    int PredictScale(const float&, class KeyFrame*) { return mnPredictedLevel; }
    cv::Mat GetNormal() { return mNormal; }
    bool IsInKeyFrame(class KeyFrame*) { return mbInKF; }
    void Replace(MapPoint* p) { mpReplacedBy = p; }
    void AddObservation(class KeyFrame*, int idx) { mnAddedObs = idx; }
    std::tuple<int, int> GetIndexInKeyFrame(class KeyFrame*) { return std::tuple<int, int>(mnIdxInKF2, -1); }
    cv::Mat mNormal;
    bool mbInKF = false;
    MapPoint* mpReplacedBy = nullptr;
    int mnAddedObs = -1, mnIdxInKF2 = -1;
    float mfMinDistance = 0.f, mfMaxDistance = 1e30f;
    int mnPredictedLevel = 0;
    bool mbTrackInView = false, mbTrackInViewR = false;
    bool mbBad = false;
protected:
    cv::Mat mWorldPos, mDescriptor;   // 3x1 CV_32F, 1x32 CV_8U
    int mnObs;
};

class Frame {   // the accessors SearchForInitialization / SearchByProjection use (include/Frame.h:173, 193, 197, 230, 247, 370-373)
public:
    MapPoint* getMapPoint(int idx) const { return mvpMapPoints[idx]; }
    void setMapPoint(int idx, MapPoint* p) { mvpMapPoints[idx] = p; }
    bool getMPOutlier(int idx) const { return mvbOutlier[idx]; }
    std::vector<float> getAllORBScaleFactors() const { return mvScaleFactors; }
    cv::Mat mTcw;                                // 4x4 CV_32F
    int numKPtsLeft() const { return Nleft; }     // -1: monocular / rectified stereo / RGB-D; >= 0: fisheye rig (include/Frame.h)
    int Nleft = -1;
    std::vector<float> mvuRight;                 // right-image column per keypoint, negative = none (include/Frame.h:296)
    float mbf = 0.f, mb = 0.f;                   // baseline * fx, baseline in metres (include/Frame.h:268-271)
    GeometricCamera* mpCamera = nullptr;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<float> mvScaleFactors;
    int numAllKPts() const { return (int)mvKeysUn.size(); }
    cv::KeyPoint getUndistKPtMono(int idx) const { return mvKeysUn[idx]; }
    std::vector<cv::KeyPoint>& getAllUndistKPtsMono() { return mvKeysUn; }
    int getKPtLevelMono(int idx) const { return mvKeysUn[idx].octave; }
    cv::Mat& getAllORBDescMono() { return mDescriptors; }
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;                // include/Frame.h:332
};

inline int MapPoint::PredictScale(const float&, Frame*) { return mnPredictedLevel; }

class KeyFrame {   // what SearchByBoW reads (include/KeyFrame.h:343, 395, 400, 407, 523)
public:
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    int numAllKPts() const { return (int)mvKeysUn.size(); }
    cv::KeyPoint getUndistKPtMono(const int idx) const { return mvKeysUn[idx]; }
    cv::Mat getORBDescriptor(const int idx) const { return mDescriptors.row(idx); }
    DBoW2::FeatureVector mFeatVec;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors;
    // what the keyframe-side searches read (include/KeyFrame.h:343-446, 514-529)
    float fx = 0.f, fy = 0.f, cx = 0.f, cy = 0.f, mbf = 0.f;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
    GeometricCamera* mpCamera = nullptr;
    GeometricCamera* mpCamera2 = nullptr;
    std::vector<float> mvuRight, mvScaleFactors, mvInvLevelSigma2;
    cv::Mat mRcw, mtcw, mOw;                     // 3x3, 3x1, 3x1 CV_32F
    cv::Mat GetRotation() { return mRcw; }
    cv::Mat GetTranslation() { return mtcw; }
    cv::Mat GetCameraCenter() { return mOw; }
    int numAllKPtsLeft() const { return -1; }
    std::vector<cv::KeyPoint> getAllUndistKPtsMono() const { return mvKeysUn; }
    cv::Mat getAllORBDescriptors() const { return mDescriptors; }
    int getORBNLevels() const { return (int)mvScaleFactors.size(); }
    float getORBScaleFactor(const int level) const { return mvScaleFactors[level]; }
    float getORBInvLevelSigma2(const int level) const { return mvInvLevelSigma2[level]; }
    float getORBLevelSigma2(const int level) const { return 1.0f / mvInvLevelSigma2[level]; }
    bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    void AddMapPoint(MapPoint* p, const size_t& idx) { mvpMapPoints[idx] = p; }
    std::set<MapPoint*> GetMapPoints() { std::set<MapPoint*> s; for (MapPoint* p : mvpMapPoints) if (p) s.insert(p); return s; }
};

using std::vector; using std::pair;   // include/ORBmatcher.h declares Fuse / SearchForTriangulation with the bare names
class ORBmatcher {   // the members this path touches (ORBmatcher.h:39-42, 67-68, 96-113)
public:
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10);
    int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, float th, bool bMono);
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);
    int SearchByProjection(Frame &CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*> &sAlreadyFound, float th, int ORBdist);
    int SearchByProjection(Frame &F, const std::vector<MapPoint*> &vpMapPoints, float th=3,
            bool bFarPoints = false, float thFarPoints = 50.0f);
    // keyframe-side searches (ORBmatcher.h:56-100)
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*> &vpPoints, std::vector<MapPoint*> &vpMatched, int th, float ratioHamming=1.0);
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*> &vpPoints, const std::vector<KeyFrame*> &vpPointsKFs,
            std::vector<MapPoint*> &vpMatched, std::vector<KeyFrame*> &vpMatchedKF, int th, float ratioHamming=1.0);
    int SearchForTriangulation(KeyFrame *pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<pair<size_t, size_t> > &vMatchedPairs, bool bOnlyStereo, bool bCoarse = false);
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint *> &vpMatches12, const float &s12, const cv::Mat &R12, const cv::Mat &t12, float th);
    int Fuse(KeyFrame* pKF, const vector<MapPoint *> &vpMapPoints, float th=3.0, bool bRight = false);       // bare `vector`: as the reference writes it
    int Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*> &vpPoints, float th, vector<MapPoint *> &vpReplacePoint);
    explicit ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
protected:
    void ComputeThreeMaxima(std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3);
    float mfNNratio;
    bool mbCheckOrientation;
};
}  // namespace ORB_SLAM3

namespace EORB_SLAM {
struct EventData {   // EventData.h:36-58
    EventData() = default;
    EventData(double ts, float x, float y, bool p) : ts(ts), x(x), y(y), p(p) {}
    double ts = 0.0;
    float x = 0.f;
    float y = 0.f;
    bool p = false;
};
}  // namespace EORB_SLAM
