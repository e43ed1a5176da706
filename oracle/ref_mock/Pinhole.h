// oracle/ref_mock/Pinhole.h — TEST INFRASTRUCTURE ONLY.  Stand-in for include/CameraModels/Pinhole.h: declarations of the four
// members the event path reaches (Pinhole.h:79, 81, 86, 90); their BODIES are the reference's own text, cut out of
// src/CameraModels/Pinhole.cpp:31-34, 36-40, 42-48, 60-63, 82-92 at build time (oracle/Makefile -> _ref/gen_event_deps.inc).
#pragma once
#include "GeometricCamera.h"
#define DEF_EC_DIST_COEF 3.84f   // include/CameraModels/Pinhole.h:36

namespace ORB_SLAM3 {
class Pinhole : public GeometricCamera {
public:
    explicit Pinhole(const std::vector<float> _vParameters) : GeometricCamera(_vParameters) {}
    cv::Point2f project(const cv::Point3f& p3D) override;
    cv::Point2f project(const cv::Mat& m3D) override;
    Eigen::Vector2d project(const Eigen::Vector3d& v3D) override;
    cv::Point3f unproject(const cv::Point2f& p2D) override;
    Eigen::Matrix<double, 2, 3> projectJac(const Eigen::Vector3d& v3D) override;
    // SearchForTriangulation's gate: bodies cut from src/CameraModels/Pinhole.cpp:128-132, 134-157, 175-180 (_ref/gen_matcher_kf.inc)
    cv::Mat toK() override;
    bool epipolarConstrain(GeometricCamera* pCamera2, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& R12, const cv::Mat& t12,
                           const float sigmaLevel, const float unc) override;
    cv::Mat SkewSymmetricMatrix(const cv::Mat& v);
};
}  // namespace ORB_SLAM3
