// oracle/ref_mock/Frame.h — TEST INFRASTRUCTURE ONLY.  Stand-ins for the reference's Frame / KeyFrame / MapPoint / DBoW2::FeatureVector
// (include/Frame.h, include/KeyFrame.h, include/MapPoint.h, Thirdparty/DBoW2/DBoW2/FeatureVector.h) holding exactly the members the
// matcher functions cut out of src/ORBmatcher.cc touch.  The GRID functions are the reference's own text: Frame::AssignFeaturesToGrid
// (src/Frame.cc:431-460), Frame::PosInGrid (:783-793) and Frame::GetFeaturesInArea (:709-777) are cut out of the source at build time
// (oracle/Makefile -> _ref/gen_matcher.inc); only their declarations live here.  MapPoint::PredictScale returns a level stored by the
// caller: its inputs (camera centre, distance invariance) are formed on the host side of the C ABI, outside the device path.
#pragma once
#include <cmath>
#include <map>
#include <set>
#include <tuple>
#include <vector>
#include <opencv2/core/core.hpp>
#include "GeometricCamera.h"

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

namespace DBoW2 {
typedef unsigned int NodeId;
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {};
}  // namespace DBoW2

namespace ORB_SLAM3 {
class Frame;
class KeyFrame;

class MapPoint {
public:
    // tracking fields Frame::isInFrustum fills (include/MapPoint.h:133-144)
    float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
    bool mbTrackInView = false, mbTrackInViewR = false;
    int mnTrackScaleLevel = 0, mnTrackScaleLevelR = -1;
    float mTrackViewCos = 1, mTrackViewCosR = 1;
    bool isBad() { return bad; }
    int Observations() { return nObs; }
    cv::Mat GetDescriptor() { return desc.clone(); }
    cv::Mat GetWorldPos() { return pos.clone(); }
    float GetMaxDistanceInvariance() { return maxDist; }
    float GetMinDistanceInvariance() { return minDist; }
    int PredictScale(const float&, Frame*) { return predictedLevel; }
    // keyframe-side searches (ORBmatcher.cc:480-712, :1407-1967): the map-graph calls are recorded, not performed
    int PredictScale(const float&, KeyFrame*) { return predictedLevel; }
    cv::Mat GetNormal() { return normal.clone(); }
    bool IsInKeyFrame(KeyFrame*) { return inKF; }
    static std::vector<std::pair<MapPoint*, MapPoint*> >& replaceLog() { static std::vector<std::pair<MapPoint*, MapPoint*> > v; return v; }
    void Replace(MapPoint* p) { replacedBy = p; replaceLog().push_back(std::make_pair(this, p)); }
    void AddObservation(KeyFrame*, int idx) { addedObsIdx = idx; }
    std::tuple<int, int> GetIndexInKeyFrame(KeyFrame*) { return std::tuple<int, int>(idxInOtherKF, -1); }
    float maxDist = 1e30f, minDist = 0.0f;
    cv::Mat normal;
    bool inKF = false;
    MapPoint* replacedBy = nullptr;
    int addedObsIdx = -1, idxInOtherKF = -1;
    bool bad = false;
    int nObs = 1, predictedLevel = 0, id = -1;
    cv::Mat desc, pos;
};

class Frame {
public:
    int N = 0, Nleft = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
    cv::Mat mDescriptors;                       // N x 32; rows [Nleft, N) are the right camera's when Nleft != -1
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<float> mvuRight;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
    std::vector<float> mvScaleFactors;
    float mbf = 0, mb = 0;
    cv::Mat mTcw, mTrl;
    GeometricCamera* mpCamera = nullptr;
    GeometricCamera* mpCamera2 = nullptr;
    DBoW2::FeatureVector mFeatVec;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];

    int numAllKPts() const { return N; }
    int numKPtsLeft() const { return Nleft; }
    cv::KeyPoint getUndistKPtMono(int idx) const { return mvKeysUn[idx]; }
    cv::KeyPoint getDistKPtMono(int idx) const { return mvKeys[idx]; }
    cv::KeyPoint getKPtRight(int idx) const { return mvKeysRight[idx]; }
    int getKPtLevelMono(int idx) const { return mvKeysUn[idx].octave; }       // Frame.cc: mvKeysUn[idx].octave
    cv::Mat getORBDescriptor(int idx) { return mDescriptors.row(idx); }
    MapPoint* getMapPoint(int idx) const { return mvpMapPoints[idx]; }
    void setMapPoint(int idx, MapPoint* p) { mvpMapPoints[idx] = p; }
    bool getMPOutlier(int idx) const { return mvbOutlier[idx]; }
    float getORBScaleFactor(const int level) const { return mvScaleFactors[level]; }

    void AssignFeaturesToGrid();
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    std::vector<std::size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1,
                                               const bool bRight = false) const;
};

class KeyFrame {
public:
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<cv::KeyPoint> mvKeysUn, mvKeysRight;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    int NLeft = -1;
    GeometricCamera* mpCamera2 = nullptr;
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    int numKPtsLeft() const { return NLeft; }
    int numAllKPtsLeft() const { return NLeft; }
    cv::KeyPoint getDistKPtMono(const int idx) const { return mvKeysUn[idx]; }
    cv::KeyPoint getUndistKPtMono(const int idx) const { return mvKeysUn[idx]; }
    cv::KeyPoint getKPtRight(const int idx) const { return mvKeysRight[idx]; }
    cv::Mat getORBDescriptor(const int idx) const { return mDescriptors.row(idx); }
    // keyframe-side searches: calibration, pose, grid (KeyFrame::GetFeaturesInArea / IsInImage are the reference's own text,
    // src/KeyFrame.cc:873-922, cut at build time), level tables, map-point slots
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
    GeometricCamera* mpCamera = nullptr;
    int N = 0;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;          // include/KeyFrame.h: const int
    int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    std::vector<std::vector<std::vector<std::size_t> > > mGrid, mGridRight;
    std::vector<float> mvuRight, mvScaleFactors, mvInvLevelSigma2;
    cv::Mat Rcw, tcw, Ow;
    cv::Mat GetRotation() { return Rcw.clone(); }
    cv::Mat GetTranslation() { return tcw.clone(); }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    cv::Mat GetRightRotation() { return Rcw.clone(); }
    cv::Mat GetRightTranslation() { return tcw.clone(); }
    cv::Mat GetRightCameraCenter() { return Ow.clone(); }
    float getORBScaleFactor(const int level) const { return mvScaleFactors[level]; }
    float getORBInvLevelSigma2(const int level) const { return mvInvLevelSigma2[level]; }
    float getORBLevelSigma2(const int level) const { return mvLevelSigma2[level]; }
    std::vector<float> mvLevelSigma2;
    int getKPtLevelMono(const int idx) const { return mvKeysUn[idx].octave; }        // KeyFrame.cc:1428-1431
    MapPoint* GetMapPoint(const std::size_t& idx) { return mvpMapPoints[idx]; }
    void AddMapPoint(MapPoint* p, const std::size_t& idx) { mvpMapPoints[idx] = p; }
    std::set<MapPoint*> GetMapPoints() { std::set<MapPoint*> s; for (MapPoint* p : mvpMapPoints) if (p) s.insert(p); return s; }
    std::vector<std::size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const bool bRight = false) const;
    bool IsInImage(const float& x, const float& y) const;
    // SearchByBoW(KeyFrame*, KeyFrame*, ...) (ORBmatcher.cc:833-846)
    const std::vector<cv::KeyPoint>& getAllUndistKPtsMono() const { return mvKeysUn; }
    const cv::Mat& getAllORBDescriptors() const { return mDescriptors; }
    int numAllKPts() const { return (int)mvKeysUn.size(); }
};
}  // namespace ORB_SLAM3
