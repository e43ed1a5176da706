// oracle/ref_mock/KannalaBrandt8.h — TEST INFRASTRUCTURE ONLY.  Stand-in for include/CameraModels/KannalaBrandt8.h (boost serialization,
// TwoViewReconstruction): declarations of the members the event path reaches (KannalaBrandt8.h:66-72) and the `precision` member
// (:98, KB8_DEF_PRECISION :35).  Their BODIES are the reference's own text, cut out of src/CameraModels/KannalaBrandt8.cpp:86-103,
// 105-109, 111-129, 163-190 at build time (oracle/Makefile -> _ref/gen_event_deps.inc).  projectJac is not on this path.
#pragma once
#include <stdexcept>
#include "GeometricCamera.h"

namespace ORB_SLAM3 {
#define KB8_DEF_PRECISION 1e-6
class KannalaBrandt8 final : public GeometricCamera {
public:
    explicit KannalaBrandt8(const std::vector<float> _vParameters) : GeometricCamera(_vParameters), precision(KB8_DEF_PRECISION) {}
    cv::Point2f project(const cv::Point3f& p3D) override;
    cv::Point2f project(const cv::Mat& m3D) override;
    Eigen::Vector2d project(const Eigen::Vector3d& v3D) override;
    cv::Point3f unproject(const cv::Point2f& p2D) override;
    Eigen::Matrix<double, 2, 3> projectJac(const Eigen::Vector3d&) override { throw std::logic_error("KannalaBrandt8::projectJac is not compiled into libref"); }
private:
    const float precision;
};
}  // namespace ORB_SLAM3
