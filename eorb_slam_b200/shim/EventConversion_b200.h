// Drop-in replacement for the live entry points of the reference's include/Event/EventConversion.h:40-81
// (EORB_SLAM::EvImConverter): same static method names, argument order and cv::Mat return convention
// (CV_32FC1 when normalized == false, CV_8UC1 when true), forwarding to the C ABI.
#ifndef EVENT_CONVERSION_B200_H
#define EVENT_CONVERSION_B200_H

#include <vector>
#include <opencv2/core/core.hpp>
#ifdef EORB_SHIM_MOCK
#include "ref_mock.h"
#else
#include "EventData.h"
#include "GeometricCamera.h"
#endif

namespace EORB_SLAM
{
    class EvImConverter
    {
    public:
        static cv::Mat ev2im(const std::vector<EventData> &vEvData, unsigned imWidth, unsigned imHeight,
                             bool pol = false, bool normalized = true);

        static cv::Mat ev2im_gauss(const std::vector<EventData> &vEvData, unsigned imWidth, unsigned imHeight,
                                   float sigma, bool pol = false, bool normalized = true);

        static cv::Mat ev2mci_gg_f(const std::vector<EventData> &vEvData, ORB_SLAM3::GeometricCamera* pCamera,
                                   const cv::Mat& Tcw, float medDepth, unsigned imWidth,
                                   unsigned imHeight, float imSigma, bool pol = false, bool normalized = true);

        static cv::Mat ev2mci_gg_f(const std::vector<EventData> &vEvData, ORB_SLAM3::GeometricCamera* pCamera,
                                   const cv::Mat& params2D, unsigned imWidth, unsigned imHeight, float sigma,
                                   bool pol = false, bool normalized = true);

        // contrast metric (EventConversion.h:45-50, EventConversion.cc:74-162); image: CV_32FC1 (the un-normalised
        // frame the callers pass, EvImBuilder.cpp:969-973).  Returns 0 and logs on error.
        static float measureImageFocus(const cv::Mat& image);
        static float measureImageFocusLocal(const cv::Mat& image, bool avg = true);
        static float measureImageFocusGlobal(const cv::Mat& image);
        static float imageMeanLocal(const cv::Mat& image, bool avg = true);
        static float imageMean(const cv::Mat& image, bool global = false, bool avg = true);

        // Jacobian of the contrast objective (EventConversion.h:65-68, EventConversion.cc:533-662).  The reference takes the
        // g2o vertex; here its estimate is passed as plain doubles so that the shim needs neither g2o nor Eigen:
        //   const auto& est = vSE3->estimate();  Eigen::Matrix3d R = est.rotation().toRotationMatrix();
        //   double Rt[12];  Eigen::Map<Eigen::Matrix<double,3,3,Eigen::RowMajor>>(Rt) = R;
        //   Eigen::Map<Eigen::Vector3d>(Rt + 9) = est.translation();
        //   double j[6];  EvImConverter::ev2mci_gg_f_jac(evs, cam, Rt, medDepth, W, H, sigma, pol, global, j);
        //   Eigen::Map<Eigen::Matrix<double,1,6>> jac(j);                       (MyOptimTypes.cpp:16)
        // Returns false and a zero Jacobian when there are no events or the library reports an error.
        static bool ev2mci_gg_f_jac(const std::vector<EventData> &vEvData, ORB_SLAM3::GeometricCamera* pCamera,
                                    const double Rt12[12], float medDepth, unsigned imWidth, unsigned imHeight,
                                    float sigma, bool pol, bool global, double jac6[6]);

        // cv::normalize(img, img, 255, 0, NORM_MINMAX, CV_8UC1) as applied by the callers (EvImBuilder.cpp:976,...)
        // fused on the device: returns the CV_8UC1 frame directly.
        static cv::Mat ev2mci_gg_f_minmax_u8(const std::vector<EventData> &vEvData, ORB_SLAM3::GeometricCamera* pCamera,
                                             const cv::Mat& Tcw, float medDepth, unsigned imWidth, unsigned imHeight, float imSigma);
    };
}// namespace EORB_SLAM
#endif
