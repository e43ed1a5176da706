// oracle/ref_mock/ORBmatcher.h — TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's include/ORBmatcher.h, found first on the include path when src/ORBextractor.cc (compiled
// unmodified into oracle/_ref) says #include "ORBmatcher.h".  The real header drags in Frame/KeyFrame/MapPoint/g2o/DBoW2;
// ORBextractor.cc uses one static member of it (ORBextractor.cc:1305), declared here with the reference's signature
// (include/ORBmatcher.h:42).  Its BODY is the reference's own text: oracle/Makefile cuts ORBmatcher.cc:2360-2378 out of
// the reference source at build time into oracle/_ref/gen_descriptor_distance.inc.
#pragma once
#include <opencv2/core/core.hpp>

namespace ORB_SLAM3 {
class ORBmatcher {
public:
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
};
}  // namespace ORB_SLAM3
