// B200 replacement for ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:714-831), same signature: drop the
// reference's definition of this one member and link this file (Tracking.cc's MonocularInitialization keeps calling
// matcher.SearchForInitialization(mInitialFrame, mCurrentFrame, mvbPrevMatched, mvIniMatches, 100) unchanged).
// The Frame grid (AssignFeaturesToGrid / GetFeaturesInArea, src/Frame.cc:431-460, 709-793) is rebuilt on the device from
// F2's undistorted keypoints and the static image bounds, so Frame::mGrid is not read.
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "eorb_b200.h"
#ifdef EORB_SHIM_MOCK
#include "ref_mock.h"
#else
#include "ORBmatcher.h"
#include "Frame.h"
#endif

namespace ORB_SLAM3
{
#ifdef EORB_SHIM_MOCK
float Frame::mnMinX = 0.f, Frame::mnMaxX = 0.f, Frame::mnMinY = 0.f, Frame::mnMaxY = 0.f;
#endif

namespace {
// one handle per calling thread: matchers are stack objects in the reference (Tracking.cc), handles hold device buffers
eorb_guided* threadHandle()
{
    thread_local struct Holder { eorb_guided* h = nullptr; ~Holder() { eorb_guided_destroy(h); } } holder;
    if (!holder.h && eorb_guided_create(0, &holder.h) != EORB_OK) {
        std::fprintf(stderr, "ORBmatcher(b200)::SearchForInitialization: %s\n", eorb_last_error());
        holder.h = nullptr;
    }
    return holder.h;
}

void packFrame(Frame& F, std::vector<eorb_keypoint>& kps, std::vector<unsigned char>& desc)
{
    const int n = F.numAllKPts();
    kps.resize(n); desc.resize((size_t)n * 32);
    cv::Mat& D = F.getAllORBDescMono();
    for (int i = 0; i < n; i++) {
        const cv::KeyPoint kp = F.getUndistKPtMono(i);
        eorb_keypoint& k = kps[i];
        k.x = kp.pt.x; k.y = kp.pt.y; k.size = kp.size; k.angle = kp.angle; k.response = kp.response;
        k.octave = F.getKPtLevelMono(i); k.class_id = kp.class_id;
        std::memcpy(&desc[(size_t)i * 32], D.ptr<unsigned char>(i), 32);
    }
}
} // namespace

int ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize)
{
    const int n1 = F1.numAllKPts();
    vnMatches12 = std::vector<int>(n1, -1);
    eorb_guided* g = threadHandle();
    if (!g || n1 == 0 || (int)vbPrevMatched.size() < n1) return 0;   // no CPU fallback: no device -> no matches, error on stderr
    std::vector<eorb_keypoint> k1, k2;
    std::vector<unsigned char> d1, d2;
    packFrame(F1, k1, d1); packFrame(F2, k2, d2);
    const float bounds[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY};
    static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "cv::Point2f is two floats");
    int nmatches = 0;
    const int rc = eorb_guided_search_for_initialization(g, k1.data(), d1.data(), n1, k2.data(), d2.data(), (int)k2.size(), bounds,
                                                         reinterpret_cast<float*>(vbPrevMatched.data()), windowSize, mfNNratio,
                                                         mbCheckOrientation ? 1 : 0, vnMatches12.data(), &nmatches);
    if (rc != EORB_OK) {
        std::fprintf(stderr, "ORBmatcher(b200)::SearchForInitialization: %s\n", eorb_last_error());
        vnMatches12.assign(n1, -1);
        return 0;
    }
    return nmatches;
}
} // namespace ORB_SLAM3
