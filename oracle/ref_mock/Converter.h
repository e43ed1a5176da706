// oracle/ref_mock/Converter.h — TEST INFRASTRUCTURE ONLY.  Stand-in for the reference's include/Converter.h (which drags in g2o):
// the three conversions EventConversion.cc calls (Converter.h:50-52), bodies cut out of src/Converter.cc:123-148 at build time,
// and the two g2o accessors ev2mci_gg_f_jac reads (vSE3->estimate().rotation() / .translation(), EventConversion.cc:555-556).
#pragma once
#include <opencv2/core/core.hpp>
#include <Eigen/Core>
#include <Eigen/Geometry>

namespace g2o {
struct SE3Quat {
    Eigen::Quaterniond r;
    Eigen::Vector3d t;
    const Eigen::Quaterniond& rotation() const { return r; }
    const Eigen::Vector3d& translation() const { return t; }
};
struct VertexSE3Expmap {
    SE3Quat est;
    const SE3Quat& estimate() const { return est; }
};
}  // namespace g2o

namespace ORB_SLAM3 {
class Converter {
public:
    static Eigen::Matrix<double, 3, 1> toVector3d(const cv::Mat& cvVector);
    static Eigen::Matrix<double, 3, 1> toVector3d(const cv::Point3f& cvPoint);
    static Eigen::Matrix<double, 3, 3> toMatrix3d(const cv::Mat& cvMat3);
};
}  // namespace ORB_SLAM3
