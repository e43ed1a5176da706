// oracle/ref_event_api.cc — C entry points of oracle/_ref/libref.so for the event-frame path.  TEST INFRASTRUCTURE ONLY.
//
// Everything computed here is the reference's src/Event/EventConversion.cc, compiled unmodified (see ref_api.cc for the
// build description).  The small functions it calls in files that cannot be compiled whole (boost / g2o / DBoW2 includes) are
// the reference's own text too, cut out at build time into _ref/gen_event_deps.inc:
//     MyCalibrator::isInImage        src/Utils/MyCalibrator.cpp:36-39
//     Converter::toVector3d (x2), toMatrix3d   src/Converter.cc:123-148
//     Pinhole::project (x2), unproject, projectJac   src/CameraModels/Pinhole.cpp:31-34, 42-48, 60-63, 82-92
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "EventConversion.h"   // the reference's include/Event/EventConversion.h
#include "Pinhole.h"
#include "KannalaBrandt8.h"
#include "oracle.h"

#include "gen_event_deps.inc"

namespace {
using EORB_SLAM::EventData;
using EORB_SLAM::EvImConverter;

static_assert(sizeof(EventData) == sizeof(orc_event), "EventData layout (include/Event/EventData.h:36-58)");

std::vector<EventData> toEvents(const orc_event* evs, int64_t n) {
    std::vector<EventData> v((size_t)n);
    for (int64_t i = 0; i < n; i++) v[(size_t)i] = EventData(evs[i].ts, evs[i].x, evs[i].y, evs[i].p != 0);
    return v;
}
cv::Mat run(const std::vector<EventData>& v, int w, int h, float sigma, int mode, const float* Tcw16, float depth, const float* K4,
            const float* se2, int se2_n, bool pol, bool normalized, int camModel = 0) {
    if (mode == 0) return EvImConverter::ev2im(v, w, h, pol, normalized);
    if (mode == 1) return EvImConverter::ev2im_gauss(v, w, h, sigma, pol, normalized);
    ORB_SLAM3::Pinhole pin(std::vector<float>(K4, K4 + 4));
    ORB_SLAM3::KannalaBrandt8 kb(camModel == 1 ? std::vector<float>(K4, K4 + 8) : std::vector<float>(8, 0.f));
    ORB_SLAM3::GeometricCamera& cam = camModel == 1 ? static_cast<ORB_SLAM3::GeometricCamera&>(kb) : static_cast<ORB_SLAM3::GeometricCamera&>(pin);
    if (mode == 2) {
        cv::Mat Tcw(4, 4, CV_32FC1);
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) Tcw.at<float>(r, c) = Tcw16[r * 4 + c];
        return EvImConverter::ev2mci_gg_f(v, &cam, Tcw, depth, (unsigned)w, (unsigned)h, sigma, pol, normalized);
    }
    cv::Mat params(se2_n, 1, CV_32FC1);
    for (int i = 0; i < se2_n; i++) params.at<float>(i, 0) = se2[i];
    return EvImConverter::ev2mci_gg_f(v, &cam, params, (unsigned)w, (unsigned)h, sigma, pol, normalized);
}
}  // namespace

extern "C" {

/* mode: 0 ev2im (EventConversion.cc:171-213), 1 ev2im_gauss (:216-269), 2 ev2mci_gg_f SE3 (:280-360), 3 ev2mci_gg_f SE2
   (:362-448).  img_f32 = the un-normalised float frame; when normalize != 0 the same call is repeated with normalized = true
   and its 8-bit frame goes to u8 (returns 1; ev2im returns a float frame unchanged when max <= min, reported as 0) */
int ref_ev_accumulate(const orc_event* evs, int64_t n, int w, int h, float sigma, int mode, const float* Tcw16, float depth,
                      const float* K4, const float* se2, int se2_n, int pol, int normalize, float* img_f32, uint8_t* u8) {
    if (mode < 0 || mode > 3) return -3;
    std::vector<EventData> v = toEvents(evs, n);
    cv::Mat f = run(v, w, h, sigma, mode, Tcw16, depth, K4, se2, se2_n, pol != 0, false);
    if (f.type() != CV_32FC1 || f.rows != h || f.cols != w) return -4;
    for (int y = 0; y < h; y++) std::memcpy(img_f32 + (size_t)y * w, f.ptr<float>(y), sizeof(float) * (size_t)w);
    if (!normalize) return 0;
    cv::Mat g = run(v, w, h, sigma, mode, Tcw16, depth, K4, se2, se2_n, pol != 0, true);
    if (g.type() != CV_8UC1) return 0;
    for (int y = 0; y < h; y++) std::memcpy(u8 + (size_t)y * w, g.ptr(y), (size_t)w);
    return 1;
}

/* the same with a camera model: cam_model 0 = Pinhole (cam8[0..3]), 1 = KannalaBrandt8 (cam8[0..7]; src/CameraModels/KannalaBrandt8.cpp) */
int ref_ev_accumulate_cam(const orc_event* evs, int64_t n, int w, int h, float sigma, int mode, const float* Tcw16, float depth,
                          const float* cam8, int cam_model, const float* se2, int se2_n, int pol, int normalize, float* img_f32, uint8_t* u8) {
    if (mode < 0 || mode > 3 || cam_model < 0 || cam_model > 1) return -3;
    std::vector<EventData> v = toEvents(evs, n);
    cv::Mat f = run(v, w, h, sigma, mode, Tcw16, depth, cam8, se2, se2_n, pol != 0, false, cam_model);
    if (f.type() != CV_32FC1 || f.rows != h || f.cols != w) return -4;
    for (int y = 0; y < h; y++) std::memcpy(img_f32 + (size_t)y * w, f.ptr<float>(y), sizeof(float) * (size_t)w);
    if (!normalize) return 0;
    cv::Mat g = run(v, w, h, sigma, mode, Tcw16, depth, cam8, se2, se2_n, pol != 0, true, cam_model);
    if (g.type() != CV_8UC1) return 0;
    for (int y = 0; y < h; y++) std::memcpy(u8 + (size_t)y * w, g.ptr(y), (size_t)w);
    return 1;
}

/* what: 0 measureImageFocusLocal(avg), 1 measureImageFocusGlobal, 2 imageMeanLocal(avg), 3 imageMean(global)  (:79-168) */
float ref_image_focus(const float* img, int w, int h, int what, int avg) {
    cv::Mat m(h, w, CV_32FC1, (void*)img, sizeof(float) * (size_t)w);
    switch (what) {
        case 0: return EvImConverter::measureImageFocusLocal(m, avg != 0);
        case 1: return EvImConverter::measureImageFocusGlobal(m);
        case 2: return EvImConverter::imageMeanLocal(m, avg != 0);
        default: return EvImConverter::imageMean(m, true, avg != 0);
    }
}

/* ev2mci_gg_f_jac (:533-662); Rt12 = row-major R (9) then t (3), double, as orc_ev_mci_jac */
void ref_ev_mci_jac(const orc_event* evs, int64_t n, int w, int h, float sigma, const double* Rt12, float medDepth, const float* K4,
                    int pol, int global, double* jac6) {
    std::vector<EventData> v = toEvents(evs, n);
    ORB_SLAM3::Pinhole cam(std::vector<float>(K4, K4 + 4));
    Eigen::Matrix3d R;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R(r, c) = Rt12[r * 3 + c];
    g2o::VertexSE3Expmap vtx;
    vtx.est.r = Eigen::Quaterniond(R);
    vtx.est.t << Rt12[9], Rt12[10], Rt12[11];
    Eigen::Matrix<double, 1, 6> j = EvImConverter::ev2mci_gg_f_jac(v, &cam, &vtx, medDepth, (unsigned)w, (unsigned)h, sigma, pol != 0, global != 0);
    for (int k = 0; k < 6; k++) jac6[k] = j[k];
}

/* CPU baseline of the event leg (bench.py): consecutive windows of `win` events, a thread pool over windows, each window
   through ev2im_gauss (mode 1) or ev2mci_gg_f SE3 (mode 2) with normalized = true as the trackers call it
   (EvImBuilder.cpp / EvAsynchTracker.cpp).  Returns a checksum (sum of all output bytes) so the work cannot be elided */
double ref_ev_accumulate_batch_mt(const orc_event* evs, int64_t n, int64_t win, int w, int h, float sigma, int mode, const float* Tcw16,
                                  float depth, const float* K4, int nthreads) {
    const int64_t nwin = win > 0 ? n / win : 0;
    if (nthreads < 1) nthreads = 1;
    std::atomic<int64_t> next{0};
    std::vector<double> sums((size_t)nthreads, 0.0);
    auto work = [&](int tid) {
        for (;;) {
            int64_t i = next.fetch_add(1);
            if (i >= nwin) break;
            std::vector<EventData> v = toEvents(evs + i * win, win);
            cv::Mat g = run(v, w, h, sigma, mode, Tcw16, depth, K4, nullptr, 0, false, true);
            double s = 0;
            for (int y = 0; y < g.rows; y++) { const uint8_t* p = g.ptr(y); for (int x = 0; x < g.cols; x++) s += p[x]; }
            sums[(size_t)tid] += s;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    double s = 0;
    for (double v : sums) s += v;
    return s;
}

}  // extern "C"
