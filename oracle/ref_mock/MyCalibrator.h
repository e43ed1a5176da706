// oracle/ref_mock/MyCalibrator.h — TEST INFRASTRUCTURE ONLY.  Stand-in for the reference's include/Utils/MyCalibrator.h: the one
// static member EventConversion.cc calls (MyCalibrator.h:34), and MyDepthMap (include/Utils/MyDataTypes.h:528), which only the
// dead depth-map overload of ev2mci_gg_f names (EventConversion.cc:450-531).  The BODY of isInImage is the reference's own text,
// cut out of src/Utils/MyCalibrator.cpp:36-39 at build time (oracle/Makefile -> _ref/gen_event_deps.inc).
#pragma once
#include <opencv2/core/core.hpp>

namespace EORB_SLAM {
class MyCalibrator {
public:
    static bool isInImage(float x, float y, int imWidth, int imHeight);
};
class MyDepthMap {
public:
    explicit MyDepthMap(float d = 1.f) : d_(d) {}
    float getDepthLinInterp(float, float) const { return d_; }
private:
    float d_;
};
}  // namespace EORB_SLAM
