#include "ORBmatcher_b200.h"

#include <cassert>
#include <cstdio>
#include <cstring>

#include "eorb_b200.h"
#ifdef EORB_SHIM_MOCK
#include "ref_mock.h"
#else
#include "ORBmatcher.h"
#endif

namespace ORB_SLAM3
{
// ORBmatcher.cc:36-38
#ifdef EORB_SHIM_MOCK
const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;
#endif

// Same contract as the reference (ORBmatcher.cc:2360-2378): two 1x32 CV_8U rows -> Hamming distance.
int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
{
    assert(a.rows == b.rows && a.cols == b.cols && !a.empty() && !b.empty());
    return eorb_descriptor_distance(a.ptr<unsigned char>(), b.ptr<unsigned char>());
}

static bool packRows(const cv::Mat& m, std::vector<unsigned char>& out)
{
    if (m.empty() || m.cols != 32 || m.type() != CV_8U) return false;
    out.resize((size_t)m.rows * 32);
    for (int i = 0; i < m.rows; i++) std::memcpy(&out[(size_t)i * 32], m.ptr<unsigned char>(i), 32);
    return true;
}

BruteForceBest2::BruteForceBest2(float nnratio, bool checkOri, int device) : mfNNratio(nnratio), mbCheckOrientation(checkOri), mpHandle(nullptr)
{
    if (eorb_matcher_create(device, &mpHandle) != EORB_OK) {
        std::fprintf(stderr, "BruteForceBest2(b200): %s\n", eorb_last_error());
        mpHandle = nullptr;
    }
}

BruteForceBest2::~BruteForceBest2() { eorb_matcher_destroy(mpHandle); }

bool BruteForceBest2::SetTrainDescriptors(const cv::Mat& trainDescs)
{
    std::vector<unsigned char> rows;
    if (!mpHandle) return false;
    if (trainDescs.empty()) return eorb_matcher_set_db_host(mpHandle, nullptr, 0, 0) == EORB_OK;
    if (!packRows(trainDescs, rows)) return false;
    return eorb_matcher_set_db_host(mpHandle, rows.data(), trainDescs.rows, 0) == EORB_OK;
}

int BruteForceBest2::Match(const cv::Mat& queryDescs, const std::vector<cv::KeyPoint>& queryKps, const std::vector<cv::KeyPoint>& trainKps,
                           std::vector<int>& vnMatches12, std::vector<int>* bestDist, int th)
{
    const int nq = queryDescs.rows;
    vnMatches12.assign(nq, -1);
    if (bestDist) bestDist->assign(nq, 256);
    std::vector<unsigned char> q;
    if (!mpHandle || nq == 0 || !packRows(queryDescs, q)) return 0;
    std::vector<eorb_match> res(nq);
    if (eorb_matcher_search(mpHandle, q.data(), nq, th, mfNNratio, res.data()) != EORB_OK) {
        std::fprintf(stderr, "BruteForceBest2(b200)::Match: %s\n", eorb_last_error());
        return 0;
    }
    int nmatches = 0;
    for (int i = 0; i < nq; i++) {
        if (bestDist) (*bestDist)[i] = res[i].best_dist;
        if (res[i].accepted) { vnMatches12[i] = res[i].best_idx; nmatches++; }
    }
    if (mbCheckOrientation && (int)queryKps.size() == nq && !trainKps.empty()) {
        std::vector<float> a1(nq), a2(trainKps.size());
        for (int i = 0; i < nq; i++) a1[i] = queryKps[i].angle;
        for (size_t i = 0; i < trainKps.size(); i++) a2[i] = trainKps[i].angle;
        nmatches = eorb_rotation_filter(a1.data(), a2.data(), vnMatches12.data(), nq);
    }
    return nmatches;
}
} // namespace ORB_SLAM3
