// oracle/bow_oracle.cc — CPU ORACLE (test infrastructure, never linked into the product).
//
// Restates the two steps that follow extraction in the reference's Frame:
//   * DBoW2 TemplatedVocabulary<FORB::TDescriptor, FORB>::transform(features, BowVector, FeatureVector, levelsup)
//     (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1200), the single-feature tree descent (:1218-1258, strict '<':
//     the first child wins distance ties), FORB::distance (FORB.cpp:81-101), BowVector::addWeight / addIfNotExist /
//     normalize (BowVector.cpp) and FeatureVector::addFeature (FeatureVector.cpp), called from Frame::ComputeBoW
//     (src/Frame.cc:796-803) with levelsup = 4.  The tree layout is the one loadFromTextFile builds (:1330-1417): node id
//     = line number, children in increasing id order, word ids in order of appearance of the leaves.
//   * Frame::UndistortKeyPoints (src/Frame.cc:805-840): cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK) on the
//     keypoint positions.  OpenCV is not vendored; the arithmetic restated here is cvUndistortPoints' (5 fixed-point
//     iterations of the inverse Brown-Conrady model in double, R = I, P = K) and is pinned against cv2.undistortPoints
//     of the cv2 4.13.0 wheel in tests/test_oracle_bow.py.
#include "oracle.h"

#include <cmath>
#include <map>
#include <vector>

struct orc_vocab {
    int k, L, scoring, weighting;
    int nnodes;
    std::vector<int> parent, childStart, children;
    std::vector<uint8_t> leaf, desc;
    std::vector<double> weight;
    std::vector<uint32_t> wordId;
};

extern "C" {

orc_vocab* orc_vocab_create(int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent, const uint8_t* is_leaf,
                            const uint8_t* desc, const double* weight) {
    orc_vocab* v = new orc_vocab();
    v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting; v->nnodes = nnodes;
    v->parent.assign(parent, parent + nnodes);
    v->leaf.assign(is_leaf, is_leaf + nnodes);
    v->desc.assign(desc, desc + (size_t)nnodes * 32);
    v->weight.assign(weight, weight + nnodes);
    v->wordId.assign(nnodes, 0);
    std::vector<std::vector<int>> ch(nnodes);
    uint32_t nwords = 0;
    for (int nid = 1; nid < nnodes; nid++) {          // file order: children.push_back(nid), words numbered as they appear
        ch[parent[nid]].push_back(nid);
        if (is_leaf[nid]) v->wordId[nid] = nwords++;
    }
    v->childStart.assign(nnodes + 1, 0);
    for (int i = 0; i < nnodes; i++) {
        v->childStart[i] = (int)v->children.size();
        v->children.insert(v->children.end(), ch[i].begin(), ch[i].end());
    }
    v->childStart[nnodes] = (int)v->children.size();
    return v;
}

void orc_vocab_destroy(orc_vocab* v) { delete v; }

// mustNormalize of the scoring classes (ScoringObject.h): L1_NORM, CHI_SQUARE, KL, BHATTACHARYYA -> L1; L2_NORM -> L2; DOT_PRODUCT -> no
static int normOf(int scoring) { return scoring == 1 ? 2 : (scoring == 5 ? 0 : 1); }

int orc_vocab_transform(const orc_vocab* v, const uint8_t* feats, int n, int levelsup, uint32_t* word_id, double* word_w, uint32_t* node_id,
                        uint32_t* bow_ids, double* bow_vals, int* nbow, uint32_t* fv_nodes, int32_t* fv_start, uint32_t* fv_feats,
                        int* nfv) {
    std::map<uint32_t, double> bow;
    std::map<uint32_t, std::vector<uint32_t>> fv;
    const int norm = normOf(v->scoring);
    const bool must = norm != 0;
    const bool tf = v->weighting == 0 || v->weighting == 1;   // TF_IDF, TF accumulate; IDF, BINARY addIfNotExist
    if (v->nnodes > 1)
        for (int i = 0; i < n; i++) {
            const uint8_t* f = feats + (size_t)i * 32;
            const int nid_level = v->L - levelsup;
            uint32_t nid = 0;
            int final_id = 0, level = 0;
            do {
                ++level;
                const int cs = v->childStart[final_id], ce = v->childStart[final_id + 1];
                final_id = v->children[cs];
                double best = orc_descriptor_distance(f, &v->desc[(size_t)final_id * 32]);
                for (int c = cs + 1; c < ce; c++) {
                    const int id = v->children[c];
                    const double d = orc_descriptor_distance(f, &v->desc[(size_t)id * 32]);
                    if (d < best) { best = d; final_id = id; }
                }
                if (level == nid_level) nid = (uint32_t)final_id;
            } while (v->childStart[final_id] != v->childStart[final_id + 1]);   // isLeaf() == children.empty()
            const uint32_t id = v->wordId[final_id];
            const double w = v->weight[final_id];
            if (word_id) word_id[i] = id;
            if (word_w) word_w[i] = w;
            if (node_id) node_id[i] = nid;
            if (w > 0) {
                if (tf) bow[id] += w;   // addWeight: insert(id, w) or += w; identical because a fresh map value is 0.0 and 0.0 + w == w
                else bow.insert({id, w});
                fv[nid].push_back((uint32_t)i);
            }
        }
    if (tf && !bow.empty() && !must) {
        const double nd = (double)bow.size();
        for (auto& e : bow) e.second /= nd;
    }
    if (must) {
        double nrm = 0.0;
        if (norm == 1) for (auto& e : bow) nrm += std::fabs(e.second);
        else { for (auto& e : bow) nrm += e.second * e.second; nrm = std::sqrt(nrm); }
        if (nrm > 0.0) for (auto& e : bow) e.second /= nrm;
    }
    int k = 0;
    for (auto& e : bow) { bow_ids[k] = e.first; bow_vals[k] = e.second; k++; }
    *nbow = k;
    int s = 0, q = 0;
    for (auto& e : fv) {
        fv_nodes[q] = e.first; fv_start[q] = s;
        for (uint32_t i : e.second) fv_feats[s++] = i;
        q++;
    }
    fv_start[q] = s;
    *nfv = q;
    return 0;
}

// cv::undistortPoints(src, dst, K, dist(k1 k2 p1 p2 [k3]), noArray(), P = K): n points (x, y) float -> float.
// dist5: k1, k2, p1, p2, k3.  Arithmetic in double, 5 iterations, as cvUndistortPointsInternal does for float input.
void orc_undistort_points(const float* xy, int n, const float* K4 /*fx fy cx cy*/, const float* dist5, float* out_xy) {
    const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
    const double ifx = 1. / fx, ify = 1. / fy;
    const double k0 = dist5[0], k1 = dist5[1], p1 = dist5[2], p2 = dist5[3], k4 = dist5[4];
    for (int i = 0; i < n; i++) {
        double x = xy[2 * i], y = xy[2 * i + 1];
        const double u = x, v = y;
        x = (x - cx) * ifx;
        y = (y - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((0 * r2 + 0) * r2 + 0) * r2) / (1 + ((k4 * r2 + k1) * r2 + k0) * r2);
            if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }   // OpenCV >= 3.4.2: give up, keep the distorted point
            const double deltaX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x) + 0 * r2 + 0 * r2 * r2;
            const double deltaY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y + 0 * r2 + 0 * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        // R = I, then P = K
        const double xx = 1. * x + 0. * y + 0., yy = 0. * x + 1. * y + 0., ww = 1. / (0. * x + 0. * y + 1.);
        x = xx * ww; y = yy * ww;
        out_xy[2 * i] = (float)(x * fx + cx);
        out_xy[2 * i + 1] = (float)(y * fy + cy);
    }
}

}  // extern "C"
