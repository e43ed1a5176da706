// bow_kernels.h — launch interface of the bag-of-words / undistortion kernels:
// DBoW2 TemplatedVocabulary::transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1258) and
// Frame::UndistortKeyPoints (src/Frame.cc:805-840, cv::undistortPoints).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eorb_b200.h"

namespace eorb {

#define EORB_BOW_MAX_FEATS 8192   // features per transform call (the reference extracts <= 5 * nFeatures + 3 * nlevels)

struct VocabDev {                 // the tree in HBM, flat (what loadFromTextFile builds, :1330-1417)
    const int* childStart;        // [nnodes + 1]
    const int* children;          // node ids, children of a node contiguous and in id order
    const uint8_t* desc;          // [nnodes][32], 16-byte aligned
    const double* weight;         // [nnodes]
    const uint32_t* wordId;       // [nnodes] (leaves)
    int nnodes, L;
};

struct BowOut {
    uint32_t* wordId; double* weight; uint32_t* nodeId;                     // per feature
    uint32_t* bowIds; double* bowVals; int* counts;                         // counts[0] = nbow, counts[1] = nfv
    uint32_t* fvNodes; int32_t* fvStart; uint32_t* fvFeats;
};

cudaError_t bow_configure();
// accumulate: TF_IDF / TF (addWeight) vs IDF / BINARY (addIfNotExist); norm: 0 none (then TF values are divided by the
// number of words), 1 L1, 2 L2 — ScoringObject::mustNormalize
cudaError_t launch_bow_transform(const VocabDev& v, const uint8_t* d_feats, int n, int levelsup, int accumulate, int norm, const BowOut& o,
                                 cudaStream_t st, long long* launches);
cudaError_t launch_undistort_keypoints(const eorb_keypoint* d_in, eorb_keypoint* d_out, int n, const float* K4, const float* dist5,
                                       cudaStream_t st);

}  // namespace eorb
