#include "ORBVocabulary_b200.h"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include "eorb_b200.h"

namespace ORB_SLAM3
{
ORBVocabularyB200::~ORBVocabularyB200() { eorb_vocab_destroy(mpHandle); }

bool ORBVocabularyB200::loadFromArrays(int k, int L, int scoring, int weighting, const std::vector<int>& parent,
                                       const std::vector<unsigned char>& isLeaf, const std::vector<unsigned char>& desc,
                                       const std::vector<double>& weight)
{
    eorb_vocab_destroy(mpHandle);
    mpHandle = nullptr; mNodes = 0; mWords = 0;
    const int n = (int)parent.size();
    if (n < 1 || (int)isLeaf.size() != n || (int)weight.size() != n || desc.size() != (size_t)n * 32) return false;
    if (eorb_vocab_create(mDevice, k, L, scoring, weighting, n, parent.data(), isLeaf.data(), desc.data(), weight.data(), &mpHandle) != EORB_OK) {
        std::fprintf(stderr, "ORBVocabularyB200: %s\n", eorb_last_error());
        mpHandle = nullptr;
        return false;
    }
    mK = k; mL = L; mNodes = n;
    for (int i = 1; i < n; i++) mWords += isLeaf[i] ? 1u : 0u;
    return true;
}

// The text format of ORB-SLAM's vocabulary (TemplatedVocabulary.h:1330-1417): header "k L scoring weighting", then one node
// per line: "parent isLeaf d0 ... d31 weight" (FORB::fromString: 32 decimal bytes).
bool ORBVocabularyB200::loadFromTextFile(const std::string& filename)
{
    std::ifstream f(filename.c_str());
    if (!f.is_open()) return false;
    std::string s;
    std::getline(f, s);
    std::stringstream ss(s);
    int k = -1, L = -1, n1 = -1, n2 = -1;
    ss >> k >> L >> n1 >> n2;
    if (k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) {
        std::fprintf(stderr, "Vocabulary loading failure: This is not a correct text file!\n");
        return false;
    }
    std::vector<int> parent(1, 0);
    std::vector<unsigned char> leaf(1, 0), desc(32, 0);
    std::vector<double> weight(1, 0.0);
    while (std::getline(f, s)) {
        if (s.empty()) continue;
        std::stringstream sn(s);
        int pid = 0, isLeaf = 0;
        sn >> pid >> isLeaf;
        unsigned char d[32];
        for (int i = 0; i < 32; i++) { int b = 0; sn >> b; d[i] = (unsigned char)b; }
        double w = 0.0;
        sn >> w;
        parent.push_back(pid); leaf.push_back(isLeaf > 0 ? 1 : 0); weight.push_back(w);
        desc.insert(desc.end(), d, d + 32);
    }
    return loadFromArrays(k, L, n1, n2, parent, leaf, desc, weight);
}

void ORBVocabularyB200::transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const
{
    v.clear();
    fv.clear();
    const int n = (int)features.size();
    if (empty() || !mpHandle || n == 0) return;
    std::vector<unsigned char> rows((size_t)n * 32);
    for (int i = 0; i < n; i++) std::memcpy(&rows[(size_t)i * 32], features[i].ptr<unsigned char>(), 32);
    std::vector<uint32_t> ids(n), nodes(n), feats(n);
    std::vector<double> vals(n);
    std::vector<int32_t> start(n + 1);
    int nbow = 0, nfv = 0;
    if (eorb_vocab_transform(mpHandle, rows.data(), n, levelsup, ids.data(), vals.data(), &nbow, nodes.data(), start.data(), feats.data(), &nfv,
                             nullptr, nullptr) != EORB_OK) {
        std::fprintf(stderr, "ORBVocabularyB200::transform: %s\n", eorb_last_error());
        return;
    }
    for (int q = 0; q < nbow; q++) v.insert(v.end(), std::make_pair(ids[q], vals[q]));          // ascending ids: O(1) hinted inserts
    for (int q = 0; q < nfv; q++)
        fv.insert(fv.end(), std::make_pair(nodes[q], std::vector<unsigned int>(feats.begin() + start[q], feats.begin() + start[q + 1])));
}

namespace b200
{
bool UndistortKeyPoints(const std::vector<cv::KeyPoint>& vKeys, const cv::Mat& K, const cv::Mat& distCoef, std::vector<cv::KeyPoint>& vKeysUn)
{
    const int n = (int)vKeys.size();
    if (distCoef.empty() || distCoef.at<float>(0, 0) == 0.0f) { vKeysUn = vKeys; return true; }   // Frame.cc:807-811
    vKeysUn.resize(n);
    if (n == 0) return true;
    const float K4[4] = {K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2)};
    float d5[5] = {0, 0, 0, 0, 0};
    const int nd = distCoef.rows * distCoef.cols;
    for (int i = 0; i < nd && i < 5; i++) d5[i] = distCoef.rows > 1 ? distCoef.at<float>(i, 0) : distCoef.at<float>(0, i);
    std::vector<eorb_keypoint> in(n), out(n);
    for (int i = 0; i < n; i++) {
        const cv::KeyPoint& kp = vKeys[i];
        in[i].x = kp.pt.x; in[i].y = kp.pt.y; in[i].size = kp.size; in[i].angle = kp.angle; in[i].response = kp.response;
        in[i].octave = kp.octave; in[i].class_id = kp.class_id;
    }
    if (eorb_undistort_keypoints(in.data(), n, K4, d5, out.data()) != EORB_OK) {
        std::fprintf(stderr, "b200::UndistortKeyPoints: %s\n", eorb_last_error());
        vKeysUn.clear();
        return false;
    }
    for (int i = 0; i < n; i++) { vKeysUn[i] = vKeys[i]; vKeysUn[i].pt.x = out[i].x; vKeysUn[i].pt.y = out[i].y; }
    return true;
}
} // namespace b200
} // namespace ORB_SLAM3
