// oracle/ref_mock/GeometricCamera.h — TEST INFRASTRUCTURE ONLY.  Stand-in for include/CameraModels/GeometricCamera.h (boost
// serialization, TwoViewReconstruction ...): the four virtuals EventConversion.cc calls (GeometricCamera.h:94, 96, 101, 105).
#pragma once
#include <utility>
#include <vector>
#include <opencv2/core/core.hpp>
#include <Eigen/Geometry>

namespace ORB_SLAM3 {
class GeometricCamera {
public:
    GeometricCamera() {}
    explicit GeometricCamera(std::vector<float> p) : mvParameters(std::move(p)) {
        if (mvParameters.size() >= 4) {   // GeometricCamera.h:71-75
            mK = cv::Mat::zeros(3, 3, CV_32F);
            mK.at<float>(0, 0) = mvParameters[0]; mK.at<float>(0, 2) = mvParameters[2]; mK.at<float>(1, 1) = mvParameters[1];
            mK.at<float>(1, 2) = mvParameters[3]; mK.at<float>(2, 2) = 1.f;
        }
    }
    // SearchForTriangulation's gate (GeometricCamera.h:128-131); only the pinhole body is compiled into libref
    virtual cv::Mat toK() { return mK.clone(); }
    virtual bool epipolarConstrain(GeometricCamera*, const cv::KeyPoint&, const cv::KeyPoint&, const cv::Mat&, const cv::Mat&, const float, const float) { return false; }
    cv::Mat mK;
    virtual ~GeometricCamera() = default;
    virtual cv::Point2f project(const cv::Point3f& p3D) = 0;
    virtual cv::Point2f project(const cv::Mat& m3D) = 0;
    virtual Eigen::Vector2d project(const Eigen::Vector3d& v3D) = 0;
    virtual cv::Point3f unproject(const cv::Point2f& p2D) = 0;
    virtual Eigen::Matrix<double, 2, 3> projectJac(const Eigen::Vector3d& v3D) = 0;
    float getParameter(const int i) { return mvParameters[i]; }
protected:
    std::vector<float> mvParameters;
};
}  // namespace ORB_SLAM3
