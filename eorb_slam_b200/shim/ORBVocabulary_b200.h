// B200 replacement for the two steps that follow extraction in ORB_SLAM3::Frame:
//   * ORBVocabularyB200: the parts of ORB_SLAM3::ORBVocabulary (= DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>,
//     include/ORBVocabulary.h) the front end uses — loadFromTextFile (TemplatedVocabulary.h:1330-1417, host parsing as in the
//     reference) and transform(features, BowVector, FeatureVector, levelsup) (:1127-1200) with the same argument types, so
//     Frame::ComputeBoW / KeyFrame::ComputeBoW (src/Frame.cc:796-803, src/KeyFrame.cc:205-213) only change the class name.
//     The tree lives in HBM; the descent and the vector assembly run on the device.
//   * b200::UndistortKeyPoints: the body of Frame::UndistortKeyPoints (src/Frame.cc:805-840).
#ifndef ORBVOCABULARY_B200_H
#define ORBVOCABULARY_B200_H

#include <map>
#include <string>
#include <vector>
#include <opencv2/core/core.hpp>

#ifdef EORB_SHIM_MOCK
namespace DBoW2 {   // Thirdparty/DBoW2/DBoW2/BowVector.h:23-26, 52; FeatureVector.h:21
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
class BowVector : public std::map<WordId, WordValue> {};
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {};
}
#else
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#endif

struct eorb_vocab;

namespace ORB_SLAM3
{
class ORBVocabularyB200
{
public:
    explicit ORBVocabularyB200(int device = 0) : mpHandle(nullptr), mDevice(device), mK(0), mL(0), mNodes(0) {}
    ~ORBVocabularyB200();
    ORBVocabularyB200(const ORBVocabularyB200&) = delete;
    ORBVocabularyB200& operator=(const ORBVocabularyB200&) = delete;

    bool loadFromTextFile(const std::string& filename);
    // the same tree from memory (node 0 = root; parent, is_leaf, 32-byte descriptor and weight per node, file order)
    bool loadFromArrays(int k, int L, int scoring, int weighting, const std::vector<int>& parent, const std::vector<unsigned char>& isLeaf,
                        const std::vector<unsigned char>& desc, const std::vector<double>& weight);
    bool empty() const { return mNodes <= 1; }
    unsigned int size() const { return mWords; }
    void transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const;

protected:
    eorb_vocab* mpHandle;
    int mDevice, mK, mL, mNodes;
    unsigned int mWords = 0;
};

namespace b200
{
// mvKeysUn = UndistortKeyPoints(mvKeys, mK, mDistCoef): K 3x3 CV_32F, distCoef 4x1 or 5x1 CV_32F (Frame.cc:805-840)
bool UndistortKeyPoints(const std::vector<cv::KeyPoint>& vKeys, const cv::Mat& K, const cv::Mat& distCoef, std::vector<cv::KeyPoint>& vKeysUn);
}
} // namespace ORB_SLAM3
#endif
