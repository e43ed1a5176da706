#include "EventConversion_b200.h"

#include <cstdio>
#include <cstring>
#include <mutex>

#include "eorb_b200.h"

namespace EORB_SLAM
{
static_assert(sizeof(EventData) == sizeof(eorb_event), "EventData must be the 24-byte record the kernels read");

namespace {
// The reference calls these static methods from up to four threads at once (EvImBuilder.cpp:1165-1193); a handle
// is not re-entrant, so every calling thread gets its own converter (stream + workspace), created lazily.
struct ThreadConv {
    eorb_evconv* h = nullptr;
    long long cap = 0; int w = 0, hgt = 0;
    ~ThreadConv() { eorb_ev_destroy(h); }
    eorb_evconv* get(long long n, int W, int H) {
        if (!h || n > cap || W * H > w * hgt) {
            eorb_ev_destroy(h); h = nullptr;
            cap = n > (1 << 20) ? n : (1 << 20); w = W > 1280 ? W : 1280; hgt = H > 960 ? H : 960;
            if (eorb_ev_create(0, 1, cap, w, hgt, &h) != EORB_OK) {
                std::fprintf(stderr, "EvImConverter(b200): %s\n", eorb_last_error());
                h = nullptr;
            }
        }
        return h;
    }
};
thread_local ThreadConv t_conv;

cv::Mat run(const std::vector<EventData>& ev, eorb_ev_params& p, bool wantU8)
{
    const int W = p.width, H = p.height;
    cv::Mat f32 = cv::Mat::zeros(H, W, CV_32FC1);
    eorb_evconv* c = t_conv.get((long long)ev.size(), W, H);
    if (!c) return f32;
    if (wantU8) {
        cv::Mat u8 = cv::Mat::zeros(H, W, CV_8UC1);
        int rc = eorb_ev_accumulate(c, reinterpret_cast<const eorb_event*>(ev.data()), (int64_t)ev.size(), &p, nullptr, u8.data, nullptr);
        if (rc < EORB_EMPTY) std::fprintf(stderr, "EvImConverter(b200): %s\n", eorb_last_error());
        return u8;
    }
    int rc = eorb_ev_accumulate(c, reinterpret_cast<const eorb_event*>(ev.data()), (int64_t)ev.size(), &p, f32.ptr<float>(), nullptr, nullptr);
    if (rc < EORB_EMPTY) std::fprintf(stderr, "EvImConverter(b200): %s\n", eorb_last_error());
    return f32;
}

void baseParams(eorb_ev_params& p, int mode, unsigned w, unsigned h, float sigma, bool pol, int norm)
{
    std::memset(&p, 0, sizeof(p));
    p.mode = mode; p.width = (int)w; p.height = (int)h; p.sigma = sigma; p.pol = pol ? 1 : 0; p.normalize = norm;
    p.med_depth = 1.f;
    for (int i = 0; i < 4; i++) p.Tcw[5 * i] = 1.f;
}

void camParams(eorb_ev_params& p, ORB_SLAM3::GeometricCamera* cam)
{
    for (int i = 0; i < 4; i++) p.K[i] = cam->getParameter(i);   // fx, fy, cx, cy (Pinhole.cpp:30-62)
    if (cam->GetType() == cam->CAM_FISHEYE) {                    // KannalaBrandt8: k1..k4 follow (KannalaBrandt8.cpp:86-103)
        p.cam_model = 1;
        for (int i = 0; i < 4; i++) p.kb[i] = cam->getParameter(4 + i);
    }
}
}  // namespace

cv::Mat EvImConverter::ev2im(const std::vector<EventData> &vEvData, unsigned imWidth, unsigned imHeight, bool pol, bool normalized)
{
    eorb_ev_params p;
    baseParams(p, EORB_EV_NEAREST, imWidth, imHeight, 1.f, pol, normalized ? EORB_NORM_RUNNING : EORB_NORM_NONE);
    return run(vEvData, p, normalized);
}

cv::Mat EvImConverter::ev2im_gauss(const std::vector<EventData> &vEvData, unsigned imWidth, unsigned imHeight, float sigma,
                                   bool pol, bool normalized)
{
    eorb_ev_params p;
    baseParams(p, EORB_EV_GAUSS, imWidth, imHeight, sigma, pol, normalized ? EORB_NORM_RUNNING : EORB_NORM_NONE);
    return run(vEvData, p, normalized);
}

cv::Mat EvImConverter::ev2mci_gg_f(const std::vector<EventData> &vEvData, ORB_SLAM3::GeometricCamera* pCamera, const cv::Mat& Tcw,
                                   float medDepth, unsigned imWidth, unsigned imHeight, float imSigma, bool pol, bool normalized)
{
    eorb_ev_params p;
    baseParams(p, EORB_EV_SE3, imWidth, imHeight, imSigma, pol, normalized ? EORB_NORM_RUNNING : EORB_NORM_NONE);
    camParams(p, pCamera);
    p.med_depth = medDepth;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) p.Tcw[4 * r + c] = Tcw.at<float>(r, c);
    return run(vEvData, p, normalized);
}

cv::Mat EvImConverter::ev2mci_gg_f(const std::vector<EventData> &vEvData, ORB_SLAM3::GeometricCamera* pCamera, const cv::Mat& params2D,
                                   unsigned imWidth, unsigned imHeight, float sigma, bool pol, bool normalized)
{
    eorb_ev_params p;
    baseParams(p, EORB_EV_SE2, imWidth, imHeight, sigma, pol, normalized ? EORB_NORM_RUNNING : EORB_NORM_NONE);
    camParams(p, pCamera);
    p.se2_n = params2D.rows > 3 ? 4 : 3;
    for (int i = 0; i < p.se2_n; i++) p.se2[i] = params2D.at<float>(i, 0);
    return run(vEvData, p, normalized);
}

cv::Mat EvImConverter::ev2mci_gg_f_minmax_u8(const std::vector<EventData> &vEvData, ORB_SLAM3::GeometricCamera* pCamera, const cv::Mat& Tcw,
                                             float medDepth, unsigned imWidth, unsigned imHeight, float imSigma)
{
    eorb_ev_params p;
    baseParams(p, EORB_EV_SE3, imWidth, imHeight, imSigma, false, EORB_NORM_MINMAX);
    camParams(p, pCamera);
    p.med_depth = medDepth;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) p.Tcw[4 * r + c] = Tcw.at<float>(r, c);
    return run(vEvData, p, true);
}

namespace {
float focus(const cv::Mat& image, int what, bool avg)
{
    if (image.empty() || image.type() != CV_32FC1) {
        std::fprintf(stderr, "EvImConverter(b200): the contrast metric expects a non-empty CV_32FC1 frame\n");
        return 0.f;
    }
    eorb_evconv* c = t_conv.get(1, image.cols, image.rows);
    if (!c) return 0.f;
    float v = 0.f;
    int rc = eorb_ev_image_focus(c, image.ptr<float>(), image.cols, image.rows, image.step, what, avg ? 1 : 0, &v);
    if (rc < EORB_EMPTY) std::fprintf(stderr, "EvImConverter(b200): %s\n", eorb_last_error());
    return v;
}
} // namespace

float EvImConverter::measureImageFocus(const cv::Mat& image) { return focus(image, EORB_FOCUS_LOCAL_STD, true); }
float EvImConverter::measureImageFocusLocal(const cv::Mat& image, const bool avg) { return focus(image, EORB_FOCUS_LOCAL_STD, avg); }
float EvImConverter::measureImageFocusGlobal(const cv::Mat& image) { return focus(image, EORB_FOCUS_GLOBAL_STD, true); }
float EvImConverter::imageMeanLocal(const cv::Mat& image, const bool avg) { return focus(image, EORB_FOCUS_LOCAL_MEAN, avg); }
float EvImConverter::imageMean(const cv::Mat& image, const bool global, const bool avg)   // EventConversion.cc:152-161: cv::mean or the local form
{
    return global ? focus(image, 3 /* global mean */, true) : focus(image, EORB_FOCUS_LOCAL_MEAN, avg);
}


bool EvImConverter::ev2mci_gg_f_jac(const std::vector<EventData> &vEvData, ORB_SLAM3::GeometricCamera* pCamera, const double Rt12[12],
                                    const float medDepth, const unsigned imWidth, const unsigned imHeight, const float sigma,
                                    const bool pol, const bool global, double jac6[6])
{
    for (int k = 0; k < 6; k++) jac6[k] = 0.0;
    if (vEvData.empty()) {
        std::fprintf(stderr, "EvImConverter::ev2mci_gg_f_jac: no events.\n");
        return false;
    }
    eorb_evconv* c = t_conv.get((long long)vEvData.size(), (int)imWidth, (int)imHeight);
    if (!c) return false;
    eorb_ev_params p;
    baseParams(p, EORB_EV_SE3, imWidth, imHeight, sigma, pol, EORB_NORM_NONE);
    camParams(p, pCamera);
    int rc = eorb_ev_mci_jac(c, reinterpret_cast<const eorb_event*>(vEvData.data()), (int64_t)vEvData.size(), (int)imWidth, (int)imHeight, sigma,
                             Rt12, medDepth, p.K, pol ? 1 : 0, global ? 1 : 0, jac6);
    if (rc < EORB_EMPTY) { std::fprintf(stderr, "EvImConverter(b200): %s\n", eorb_last_error()); return false; }
    return rc == EORB_OK;
}

}// namespace EORB_SLAM
