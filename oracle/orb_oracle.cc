// oracle/orb_oracle.cc — CPU ORACLE (test infrastructure, never linked into the product).
//
// Restates reference src/ORBextractor.cc (ctor :420-489, IC_Angle :77-104, computeOrbDescriptor :108-157,
// DivideNode :500-556, DistributeOctTree :558-782, ComputeKeyPointsOctTree :784-902, operator() :1092-1238,
// ComputePyramid :1240-1265, AssignKPtLevelByBestDesc :1267-1314, ComputeTrackedKPtsDesc :1316-1363) plus the
// OpenCV primitives it calls (OpenCV is NOT vendored by the reference; build_eorb_slam.sh:136-139 pins 3.4.1;
// semantics here are pinned bit-exactly to cv2 4.13.0 by tests/golden/make_golden.py):
//   cv::resize INTER_LINEAR 8-bit  = 11-bit fixed point (imgproc resize.cpp, HResizeLinear/VResizeLinear)
//   cv::copyMakeBorder REFLECT_101, cv::GaussianBlur 5x5 sigma=2 8-bit (fixed-point 8.8 separable),
//   cv::FAST (9/16, cornerScore, 3x3 strict NMS), cv::fastAtan2, cvRound (round-half-even).
//
// Pins where the reference is not a function of its inputs (SURVEY.md §7 "Hard parts"):
//   * octree size ties are broken by NODE CREATION ORDER (the reference sorts pair<int,ExtractorNode*>, i.e.
//     by heap pointer): ascending creation sequence plays the role of ascending pointer value.
//   * px*b+py*a is evaluated WITHOUT FMA contraction (build with -ffp-contract=off), cos/sin = libm cosf/sinf.
//   * descriptor taps that fall outside the (borderless) blurred level (possible only when edgeTh < 19) are
//     read with REFLECT_101 indexing (the reference reads out of bounds there).
//   * orientation patches that run past the bordered level (possible only when the margin is below 8) are read with
//     REFLECT_101 indexing of the level (the reference reads past its buffer there).
//   * EDGE_THRESHOLD is per extractor instance (the reference has one mutable global, :73).
//   * levels whose FAST grid has 0 rows or columns yield no keypoints (the reference divides by zero).
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <list>
#include <thread>
#include <vector>

namespace {

const int kPatchSize = 31;
const int kHalfPatch = 15;
const float kWDenom = 30.f;

const int8_t kPattern[1024] = {
#include "brief_pattern_31.inc"
};

inline int cvRoundF(float v) { return (int)lrintf(v); }   // SSE cvtss2si, RNE
inline int cvRoundD(double v) { return (int)lrint(v); }
inline int cvFloorF(float v) { int i = (int)v; return i - (v < (float)i); }
inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
    return p;
}
inline uint8_t satU8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
inline short satShortF(float v) { int i = cvRoundF(v); return (short)(i < -32768 ? -32768 : (i > 32767 ? 32767 : i)); }

// ---------------------------------------------------------------- resize (O2)
void resizeLinearU8(const uint8_t* src, int sw, int sh, size_t sstride, uint8_t* dst, int dw, int dh, size_t dstride) {
    if (dw <= 0 || dh <= 0) return;
    // cv::resize: inv_scale = (double)dsize/ssize; scale = 1./inv_scale
    double inv_sx = (double)dw / sw, inv_sy = (double)dh / sh;
    double scale_x = 1. / inv_sx, scale_y = 1. / inv_sy;
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> ialpha(2 * dw), ibeta(2 * dh);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cvFloorF(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ialpha[2 * dx] = satShortF((1.f - fx) * 2048.f);
        ialpha[2 * dx + 1] = satShortF(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cvFloorF(fy);
        fy -= sy;
        yofs[dy] = sy;
        ibeta[2 * dy] = satShortF((1.f - fy) * 2048.f);
        ibeta[2 * dy + 1] = satShortF(fy * 2048.f);
    }
    std::vector<int> r0(dw), r1(dw);
    for (int dy = 0; dy < dh; dy++) {
        int sy0 = std::min(std::max(yofs[dy], 0), sh - 1);
        int sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
        const uint8_t* S0 = src + (size_t)sy0 * sstride;
        const uint8_t* S1 = src + (size_t)sy1 * sstride;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx];
            int sx1 = std::min(sx + 1, sw - 1);  // weight is 0 when clamped
            int a0 = ialpha[2 * dx], a1 = ialpha[2 * dx + 1];
            r0[dx] = S0[sx] * a0 + S0[sx1] * a1;
            r1[dx] = S1[sx] * a0 + S1[sx1] * a1;
        }
        int b0 = ibeta[2 * dy], b1 = ibeta[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++)
            D[dx] = satU8((((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2);
    }
}

// ---------------------------------------------------------------- blur (O7)
void gauss5x5(const uint8_t* src, int w, int h, size_t sstride, uint8_t* dst, size_t dstride) {
    static const int kw[5] = {39, 57, 64, 57, 39};  // getGaussianKernel(5,2) in 8.8 fixed point
    std::vector<int> tmp((size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t* S = src + (size_t)y * sstride;
        for (int x = 0; x < w; x++) {
            int s = 0;
            for (int k = -2; k <= 2; k++) s += kw[k + 2] * S[reflect101(x + k, w)];
            tmp[(size_t)y * w + x] = s;
        }
    }
    for (int y = 0; y < h; y++) {
        uint8_t* D = dst + (size_t)y * dstride;
        const int* R[5];
        for (int k = -2; k <= 2; k++) R[k + 2] = &tmp[(size_t)reflect101(y + k, h) * w];
        for (int x = 0; x < w; x++) {
            int s = 0;
            for (int k = 0; k < 5; k++) s += kw[k] * R[k][x];
            D[x] = satU8((s + 32768) >> 16);
        }
    }
}

// ---------------------------------------------------------------- FAST (O4)
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// largest m such that some 9-arc of the ring is entirely > v+m-1 ... i.e. max over arcs of min |v-p| with a
// common sign; corner at threshold t  <=>  m > t;  cv score = m-1  (cornerScore<16>).
inline int fastMaxArcMin(const uint8_t* p, const int* off) {
    int v = p[0];
    int d[25];
    for (int k = 0; k < 16; k++) d[k] = v - p[off[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int best = 0;
    for (int k = 0; k < 16; k++) {
        int mn = d[k], mx = d[k];
        for (int i = 1; i < 9; i++) { mn = std::min(mn, d[k + i]); mx = std::max(mx, d[k + i]); }
        best = std::max(best, mn);      // ring darker than centre
        best = std::max(best, -mx);     // ring brighter than centre
    }
    return best;
}

int fast9_16(const uint8_t* img, int w, int h, size_t stride, int threshold, bool nms,
             std::vector<int>& xs, std::vector<int>& ys, std::vector<int>& sc) {
    xs.clear(); ys.clear(); sc.clear();
    threshold = std::min(std::max(threshold, 0), 255);
    if (w < 7 || h < 7) return 0;
    int off[16];
    for (int k = 0; k < 16; k++) off[k] = kRingDy[k] * (int)stride + kRingDx[k];
    std::vector<uint8_t> score((size_t)w * h, 0);
    std::vector<uint8_t> corner((size_t)w * h, 0);
    for (int y = 3; y < h - 3; y++) {
        const uint8_t* row = img + (size_t)y * stride;
        for (int x = 3; x < w - 3; x++) {
            const uint8_t* p = row + x;
            int v = p[0];
            // cheap necessary test: any 9-arc contains ring[0] or ring[8]
            int d0 = v - p[off[0]], d8 = v - p[off[8]];
            if (std::abs(d0) <= threshold && std::abs(d8) <= threshold) continue;
            int m = fastMaxArcMin(p, off);
            if (m > threshold) {
                corner[(size_t)y * w + x] = 1;
                score[(size_t)y * w + x] = (uint8_t)(m - 1);
            }
        }
    }
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            size_t i = (size_t)y * w + x;
            if (!corner[i]) continue;
            int s = score[i];
            if (nms) {
                const uint8_t* c = &score[i];
                if (!(s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] &&
                      s > c[w - 1] && s > c[w] && s > c[w + 1]))
                    continue;
            }
            xs.push_back(x); ys.push_back(y); sc.push_back(s);
        }
    return (int)xs.size();
}

// ---------------------------------------------------------------- fastAtan2 (O6)
float fastAtan2f(float y, float x) {
    const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
    const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
    const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
    const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
    float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------- octree (O5)
struct Key { float x, y, resp; int idx; };
struct Node {
    std::vector<Key> keys;
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    std::list<Node>::iterator lit;
    bool noMore = false;
    long seq = 0;   // creation sequence: the oracle's stand-in for the heap pointer value
};

void divideNode(const Node& n, Node& n1, Node& n2, Node& n3, Node& n4) {
    const int halfX = (int)std::ceil(static_cast<float>(n.URx - n.ULx) / 2);
    const int halfY = (int)std::ceil(static_cast<float>(n.BRy - n.ULy) / 2);
    n1.ULx = n.ULx; n1.ULy = n.ULy;
    n1.URx = n.ULx + halfX; n1.URy = n.ULy;
    n1.BLx = n.ULx; n1.BLy = n.ULy + halfY;
    n1.BRx = n.ULx + halfX; n1.BRy = n.ULy + halfY;
    n2.ULx = n1.URx; n2.ULy = n1.URy;
    n2.URx = n.URx; n2.URy = n.URy;
    n2.BLx = n1.BRx; n2.BLy = n1.BRy;
    n2.BRx = n.URx; n2.BRy = n.ULy + halfY;
    n3.ULx = n1.BLx; n3.ULy = n1.BLy;
    n3.URx = n1.BRx; n3.URy = n1.BRy;
    n3.BLx = n.BLx; n3.BLy = n.BLy;
    n3.BRx = n1.BRx; n3.BRy = n.BLy;
    n4.ULx = n3.URx; n4.ULy = n3.URy;
    n4.URx = n2.BRx; n4.URy = n2.BRy;
    n4.BLx = n3.BRx; n4.BLy = n3.BRy;
    n4.BRx = n.BRx; n4.BRy = n.BRy;
    for (const Key& kp : n.keys) {
        if (kp.x < n1.URx) {
            if (kp.y < n1.BRy) n1.keys.push_back(kp); else n3.keys.push_back(kp);
        } else if (kp.y < n1.BRy) n2.keys.push_back(kp);
        else n4.keys.push_back(kp);
    }
    if (n1.keys.size() == 1) n1.noMore = true;
    if (n2.keys.size() == 1) n2.noMore = true;
    if (n3.keys.size() == 1) n3.noMore = true;
    if (n4.keys.size() == 1) n4.noMore = true;
}

std::vector<Key> distributeOctTree(const std::vector<Key>& in, int minX, int maxX, int minY, int maxY, int N) {
    typedef std::pair<std::pair<int, long>, Node*> SizeNode;   // ((size, seq), node)
    const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
    std::vector<Key> result;
    if (nIni <= 0) return result;   // reference: division by zero (height > 2*width); defined here as empty
    const float hX = static_cast<float>(maxX - minX) / nIni;
    std::list<Node> lNodes;
    std::vector<Node*> ini(nIni);
    long seq = 0;
    for (int i = 0; i < nIni; i++) {
        Node ni;
        ni.ULx = (int)(hX * static_cast<float>(i)); ni.ULy = 0;
        ni.URx = (int)(hX * static_cast<float>(i + 1)); ni.URy = 0;
        ni.BLx = ni.ULx; ni.BLy = maxY - minY;
        ni.BRx = ni.URx; ni.BRy = maxY - minY;
        ni.seq = seq++;
        lNodes.push_back(ni);
        ini[i] = &lNodes.back();
    }
    for (const Key& kp : in) {
        int slot = (int)(kp.x / hX);
        if (slot >= nIni) slot = nIni - 1;   // cannot happen for x < width; guards float round-up
        ini[slot]->keys.push_back(kp);
    }
    for (auto lit = lNodes.begin(); lit != lNodes.end();) {
        if (lit->keys.size() == 1) { lit->noMore = true; ++lit; }
        else if (lit->keys.empty()) lit = lNodes.erase(lit);
        else ++lit;
    }
    bool finish = false;
    std::vector<SizeNode> vSize;
    auto pushChild = [&](Node& c, bool count, int& nToExpand) {
        if (c.keys.empty()) return;
        c.seq = seq++;
        lNodes.push_front(c);
        if (c.keys.size() > 1) {
            if (count) nToExpand++;
            vSize.push_back(SizeNode(std::make_pair((int)c.keys.size(), lNodes.front().seq), &lNodes.front()));
            lNodes.front().lit = lNodes.begin();
        }
    };
    while (!finish) {
        int prevSize = (int)lNodes.size();
        auto lit = lNodes.begin();
        int nToExpand = 0;
        vSize.clear();
        while (lit != lNodes.end()) {
            if (lit->noMore) { ++lit; continue; }
            Node n1, n2, n3, n4;
            divideNode(*lit, n1, n2, n3, n4);
            pushChild(n1, true, nToExpand); pushChild(n2, true, nToExpand);
            pushChild(n3, true, nToExpand); pushChild(n4, true, nToExpand);
            lit = lNodes.erase(lit);
        }
        if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) {
            finish = true;
        } else if ((int)lNodes.size() + nToExpand * 3 > N) {
            while (!finish) {
                prevSize = (int)lNodes.size();
                std::vector<SizeNode> prev = vSize;
                vSize.clear();
                std::sort(prev.begin(), prev.end(),
                          [](const SizeNode& a, const SizeNode& b) { return a.first < b.first; });
                for (int j = (int)prev.size() - 1; j >= 0; j--) {
                    Node n1, n2, n3, n4;
                    divideNode(*prev[j].second, n1, n2, n3, n4);
                    int dummy = 0;
                    pushChild(n1, false, dummy); pushChild(n2, false, dummy);
                    pushChild(n3, false, dummy); pushChild(n4, false, dummy);
                    lNodes.erase(prev[j].second->lit);
                    if ((int)lNodes.size() >= N) break;
                }
                if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) finish = true;
            }
        }
    }
    result.reserve(lNodes.size());
    for (auto& nd : lNodes) {
        const Key* best = &nd.keys[0];
        float maxResp = best->resp;
        for (size_t k = 1; k < nd.keys.size(); k++)
            if (nd.keys[k].resp > maxResp) { best = &nd.keys[k]; maxResp = nd.keys[k].resp; }
        result.push_back(*best);
    }
    return result;
}

}  // namespace

// ==================================================================== extractor
struct orc_orb {
    orc_orb_params par;
    int nlevels, edge;
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> featuresPerLevel;
    int umax[16];
    // state of the last extraction
    std::vector<int> lw, lh;
    std::vector<std::vector<uint8_t>> pyr;       // bordered: (lw+2E) x (lh+2E)
    std::vector<std::vector<uint8_t>> blurred;   // lw x lh, filled only for levels with keypoints
    std::vector<std::vector<Key>> cand;          // FAST candidates per level (relative coords)
    std::vector<std::vector<orc_keypoint>> lkps; // per level after octree + orientation (level coords)
    int fallbackCells = 0;

    const uint8_t* levelPtr(int l) const { return pyr[l].data() + (size_t)edge * (lw[l] + 2 * edge) + edge; }
    size_t levelStride(int l) const { return (size_t)(lw[l] + 2 * edge); }
};

static void orbInit(orc_orb* o, const orc_orb_params* p) {
    o->par = *p;
    const int nl = p->nlevels;
    o->nlevels = nl;
    const double scaleFactor = p->scaleFactor;   // the member is a double holding the float (ORBextractor.h:123)
    o->scale.assign(nl, 1.f); o->sigma2.assign(nl, 1.f);
    for (int i = 1; i < nl; i++) {
        o->scale[i] = (float)(o->scale[i - 1] * scaleFactor);
        o->sigma2[i] = o->scale[i] * o->scale[i];
    }
    o->invScale.resize(nl); o->invSigma2.resize(nl);
    for (int i = 0; i < nl; i++) { o->invScale[i] = 1.0f / o->scale[i]; o->invSigma2[i] = 1.0f / o->sigma2[i]; }
    o->featuresPerLevel.assign(nl, 0);
    float factor = (float)(1.0f / scaleFactor);
    float nDesired = p->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
        o->featuresPerLevel[l] = cvRoundF(nDesired);
        sum += o->featuresPerLevel[l];
        nDesired *= factor;
    }
    o->featuresPerLevel[nl - 1] = std::max(p->nfeatures - sum, 0);
    // umax (ORBextractor.cc:463-478)
    int v, v0;
    int vmax = cvFloorF(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (v = 0; v < 16; v++) o->umax[v] = 0;
    for (v = 0; v <= vmax; ++v) o->umax[v] = cvRoundD(std::sqrt(hp2 - v * v));
    for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (o->umax[v0] == o->umax[v0 + 1]) ++v0;
        o->umax[v] = v0;
        ++v0;
    }
    if (p->edgeTh < 0) {
        float newEdge = 19 * ((float)p->imW / 752.f);
        int e = static_cast<int>(newEdge);
        e += (e % 2 - 1);
        o->edge = e;
    } else {
        o->edge = p->edgeTh;
    }
}

static void computePyramid(orc_orb* o, const uint8_t* img, int w, int h, size_t stride) {
    const int nl = o->nlevels, E = o->edge;
    o->lw.resize(nl); o->lh.resize(nl); o->pyr.resize(nl);
    for (int l = 0; l < nl; l++) {
        float sc = o->invScale[l];
        int sw = cvRoundF((float)w * sc), sh = cvRoundF((float)h * sc);
        o->lw[l] = sw; o->lh[l] = sh;
        size_t bs = (size_t)sw + 2 * E;
        o->pyr[l].assign(bs * ((size_t)sh + 2 * E), 0);
        uint8_t* inner = o->pyr[l].data() + (size_t)E * bs + E;
        if (sw <= 0 || sh <= 0) continue;
        if (l != 0) {
            resizeLinearU8(o->levelPtr(l - 1), o->lw[l - 1], o->lh[l - 1], o->levelStride(l - 1), inner, sw, sh, bs);
            std::vector<uint8_t> tmp((size_t)sw * sh);
            for (int y = 0; y < sh; y++) memcpy(&tmp[(size_t)y * sw], inner + (size_t)y * bs, sw);
            orc_copy_make_border_reflect101(tmp.data(), sw, sh, sw, o->pyr[l].data(), E, bs);
        } else {
            orc_copy_make_border_reflect101(img, w, h, stride, o->pyr[l].data(), E, bs);
        }
    }
}

static float icAngle(const uint8_t* center, int step, const int* umax) {
    int m01 = 0, m10 = 0;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * center[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vsum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int vp = center[u + v * step], vm = center[u - v * step];
            vsum += (vp - vm);
            m10 += u * (vp + vm);
        }
        m01 += v * vsum;
    }
    return fastAtan2f((float)m01, (float)m10);
}

static void computeKeyPointsOctTree(orc_orb* o) {
    const int nl = o->nlevels, E = o->edge;
    o->cand.assign(nl, {});
    o->lkps.assign(nl, {});
    o->fallbackCells = 0;
    std::vector<int> xs, ys, sc;
    for (int level = 0; level < nl; ++level) {
        const int W = o->lw[level], H = o->lh[level];
        const int minBX = E - 3, minBY = minBX;
        const int maxBX = W - E + 3, maxBY = H - E + 3;
        std::vector<Key>& toDist = o->cand[level];
        const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
        const int nCols = (int)(width / kWDenom), nRows = (int)(height / kWDenom);
        if (nCols > 0 && nRows > 0 && W > 0 && H > 0) {
            const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
            const uint8_t* base = o->levelPtr(level);
            const size_t st = o->levelStride(level);
            for (int i = 0; i < nRows; i++) {
                const float iniY = (float)(minBY + i * hCell);
                float maxY = iniY + hCell + 6;
                if (iniY >= maxBY - 3) continue;
                if (maxY > maxBY) maxY = (float)maxBY;
                for (int j = 0; j < nCols; j++) {
                    const float iniX = (float)(minBX + j * wCell);
                    float maxX = iniX + wCell + 6;
                    if (iniX >= maxBX - 3) continue;
                    if (maxX > maxBX) maxX = (float)maxBX;
                    const int x0 = (int)iniX, y0 = (int)iniY, x1 = (int)maxX, y1 = (int)maxY;
                    const uint8_t* roi = base + (size_t)y0 * st + x0;
                    int n = fast9_16(roi, x1 - x0, y1 - y0, st, o->par.iniThFAST, true, xs, ys, sc);
                    if (n == 0) {
                        o->fallbackCells++;
                        n = fast9_16(roi, x1 - x0, y1 - y0, st, o->par.minThFAST, true, xs, ys, sc);
                    }
                    for (int k = 0; k < n; k++) {
                        Key kp;
                        kp.x = (float)xs[k] + j * wCell;
                        kp.y = (float)ys[k] + i * hCell;
                        kp.resp = (float)sc[k];
                        kp.idx = (int)toDist.size();
                        toDist.push_back(kp);
                    }
                }
            }
        }
        std::vector<Key> sel;
        if (maxBY - minBY > 0 && maxBX - minBX > 0)
            sel = distributeOctTree(toDist, minBX, maxBX, minBY, maxBY, o->featuresPerLevel[level]);
        const int scaledPatch = (int)(kPatchSize * o->scale[level]);
        std::vector<orc_keypoint>& out = o->lkps[level];
        out.resize(sel.size());
        for (size_t i = 0; i < sel.size(); i++) {
            out[i].x = sel[i].x + minBX;
            out[i].y = sel[i].y + minBY;
            out[i].octave = level;
            out[i].size = (float)scaledPatch;
            out[i].response = sel[i].resp;
            out[i].angle = -1.f;
            out[i].class_id = -1;
        }
    }
    for (int level = 0; level < nl; ++level)
        for (auto& kp : o->lkps[level]) {
            if (E >= 8) {   // keypoints lie >= E px inside the level: the 15-px patch stays inside the E-px reflected border
                const uint8_t* c = o->levelPtr(level) + (size_t)cvRoundF(kp.y) * o->levelStride(level) + cvRoundF(kp.x);
                kp.angle = icAngle(c, (int)o->levelStride(level), o->umax);
            } else {
                // Pin for margins below 8 (e.g. the adaptive margin of images narrower than 316 px): the reference's patch runs
                // past its bordered buffer there (undefined reads).  Rule: REFLECT_101 indexing of the level, i.e. what a wide
                // enough border would hold; identical to the branch above whenever that one is defined.
                const int W = o->lw[level], H = o->lh[level], cx = cvRoundF(kp.x), cy = cvRoundF(kp.y);
                const uint8_t* L = o->levelPtr(level);
                const size_t st = o->levelStride(level);
                int m01 = 0, m10 = 0;
                for (int v = -kHalfPatch; v <= kHalfPatch; ++v) {
                    const int d = o->umax[v < 0 ? -v : v];
                    const uint8_t* row = L + (ptrdiff_t)reflect101(cy + v, H) * (ptrdiff_t)st;
                    for (int u = -d; u <= d; ++u) {
                        const int val = row[reflect101(cx + u, W)];
                        m10 += u * val; m01 += v * val;
                    }
                }
                kp.angle = fastAtan2f((float)m01, (float)m10);
            }
        }
}

static const float kFactorPI = (float)(M_PI / 180.f);

static void computeOrbDescriptor(const orc_keypoint& kpt, const uint8_t* img, int w, int h, uint8_t* desc) {
    float angle = (float)kpt.angle * kFactorPI;
    float a = cosf(angle), b = sinf(angle);
    const int cy = cvRoundF(kpt.y), cx = cvRoundF(kpt.x);
    const int8_t* pat = kPattern;
    auto get = [&](int idx) -> int {
        float px = (float)pat[2 * idx], py = (float)pat[2 * idx + 1];
        int r = cvRoundF(px * b + py * a);
        int c = cvRoundF(px * a - py * b);
        int yy = cy + r, xx = cx + c;
        if ((unsigned)yy >= (unsigned)h) yy = reflect101(yy, h);   // pin for margin < 19 (reference: OOB read)
        if ((unsigned)xx >= (unsigned)w) xx = reflect101(xx, w);
        return img[(size_t)yy * w + xx];
    };
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            int t0 = get(2 * k), t1 = get(2 * k + 1);
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

static void blurLevel(orc_orb* o, int level) {
    const int W = o->lw[level], H = o->lh[level];
    o->blurred[level].assign((size_t)W * H, 0);
    gauss5x5(o->levelPtr(level), W, H, o->levelStride(level), o->blurred[level].data(), W);
}

extern "C" {

int orc_cv_round_f(float v) { return cvRoundF(v); }
int orc_cv_round_d(double v) { return cvRoundD(v); }

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride, uint8_t* dst, int dw, int dh, size_t dstride) {
    resizeLinearU8(src, sw, sh, sstride, dst, dw, dh, dstride);
}

void orc_copy_make_border_reflect101(const uint8_t* src, int w, int h, size_t sstride, uint8_t* dst, int border, size_t dstride) {
    for (int y = -border; y < h + border; y++) {
        const uint8_t* S = src + (size_t)reflect101(y, h) * sstride;
        uint8_t* D = dst + (size_t)(y + border) * dstride;
        for (int x = -border; x < w + border; x++) D[x + border] = S[reflect101(x, w)];
    }
}

void orc_gauss5x5_s2_u8(const uint8_t* src, int w, int h, size_t sstride, uint8_t* dst, size_t dstride) {
    gauss5x5(src, w, h, sstride, dst, dstride);
}

int orc_fast9_16(const uint8_t* img, int w, int h, size_t stride, int threshold, int nms, int* xs, int* ys, int* scores, int cap) {
    std::vector<int> x, y, s;
    int n = fast9_16(img, w, h, stride, threshold, nms != 0, x, y, s);
    for (int i = 0; i < n && i < cap; i++) { xs[i] = x[i]; ys[i] = y[i]; scores[i] = s[i]; }
    return n;
}

float orc_fast_atan2(float y, float x) { return fastAtan2f(y, x); }

int orc_distribute_octtree(const float* kx, const float* ky, const float* kresp, int n, int minX, int maxX, int minY,
                           int maxY, int N, int* out_idx, int cap) {
    std::vector<Key> in(n);
    for (int i = 0; i < n; i++) { in[i].x = kx[i]; in[i].y = ky[i]; in[i].resp = kresp[i]; in[i].idx = i; }
    std::vector<Key> r = distributeOctTree(in, minX, maxX, minY, maxY, N);
    for (size_t i = 0; i < r.size() && (int)i < cap; i++) out_idx[i] = r[i].idx;
    return (int)r.size();
}

orc_orb* orc_orb_create(const orc_orb_params* p) {
    if (!p || p->nlevels < 1) return nullptr;
    orc_orb* o = new orc_orb();
    orbInit(o, p);
    return o;
}
void orc_orb_destroy(orc_orb* h) { delete h; }
int orc_orb_edge_threshold(const orc_orb* h) { return h->edge; }
int orc_orb_features_per_level(const orc_orb* h, int* out) {
    for (int i = 0; i < h->nlevels; i++) out[i] = h->featuresPerLevel[i];
    return h->nlevels;
}
int orc_orb_scale_factors(const orc_orb* h, float* s, float* is, float* s2, float* is2) {
    for (int i = 0; i < h->nlevels; i++) {
        if (s) s[i] = h->scale[i];
        if (is) is[i] = h->invScale[i];
        if (s2) s2[i] = h->sigma2[i];
        if (is2) is2[i] = h->invSigma2[i];
    }
    return h->nlevels;
}
int orc_orb_umax(const orc_orb* h, int* out16) { for (int i = 0; i < 16; i++) out16[i] = h->umax[i]; return 16; }

int orc_orb_extract(orc_orb* o, const uint8_t* img, int w, int hgt, size_t stride, int lap0, int lap1, int want_desc,
                    orc_keypoint* kps, uint8_t* desc, int cap, int* n_out) {
    if (n_out) *n_out = 0;
    if (!img || w <= 0 || hgt <= 0) return -1;   // _image.empty() -> -1 (ORBextractor.cc:1096)
    computePyramid(o, img, w, hgt, stride);
    computeKeyPointsOctTree(o);
    const int nl = o->nlevels;
    int nk = 0;
    for (int l = 0; l < nl; l++) nk += (int)o->lkps[l].size();
    if (n_out) *n_out = nk;
    o->blurred.assign(nl, {});
    if (nk > cap) return -2;
    int monoIndex = 0, stereoIndex = nk - 1;
    std::vector<uint8_t> d(32);
    for (int l = 0; l < nl; l++) {
        std::vector<orc_keypoint>& k = o->lkps[l];
        if (k.empty()) continue;
        if (want_desc) blurLevel(o, l);
        const float scale = o->scale[l];
        for (size_t i = 0; i < k.size(); i++) {
            orc_keypoint kp = k[i];
            if (want_desc) computeOrbDescriptor(kp, o->blurred[l].data(), o->lw[l], o->lh[l], d.data());
            if (l != 0) { kp.x *= scale; kp.y *= scale; }
            int dstIdx;
            if (kp.x >= (float)lap0 && kp.x <= (float)lap1) dstIdx = stereoIndex--;
            else dstIdx = monoIndex++;
            kps[dstIdx] = kp;
            if (want_desc && desc) memcpy(desc + (size_t)dstIdx * 32, d.data(), 32);
        }
    }
    return monoIndex;
}

int orc_orb_level_size(const orc_orb* h, int level, int* w, int* hgt) {
    if (level < 0 || level >= (int)h->lw.size()) return -1;
    *w = h->lw[level]; *hgt = h->lh[level];
    return 0;
}
int orc_orb_get_level(const orc_orb* h, int level, uint8_t* dst, size_t dstride) {
    if (level < 0 || level >= (int)h->lw.size()) return -1;
    for (int y = 0; y < h->lh[level]; y++) memcpy(dst + (size_t)y * dstride, h->levelPtr(level) + (size_t)y * h->levelStride(level), h->lw[level]);
    return 0;
}
int orc_orb_get_blurred(const orc_orb* h, int level, uint8_t* dst, size_t dstride) {
    if (level < 0 || level >= (int)h->blurred.size() || h->blurred[level].empty()) return -1;
    for (int y = 0; y < h->lh[level]; y++) memcpy(dst + (size_t)y * dstride, &h->blurred[level][(size_t)y * h->lw[level]], h->lw[level]);
    return 0;
}
int orc_orb_num_candidates(const orc_orb* h, int level) { return (int)h->cand[level].size(); }
int orc_orb_get_candidates(const orc_orb* h, int level, int* xs, int* ys, int* scores, int cap) {
    const auto& c = h->cand[level];
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) { xs[i] = (int)c[i].x; ys[i] = (int)c[i].y; scores[i] = (int)c[i].resp; }
    return (int)c.size();
}
int orc_orb_num_level_kps(const orc_orb* h, int level) { return (int)h->lkps[level].size(); }
int orc_orb_get_level_kps(const orc_orb* h, int level, int* xs, int* ys, int* scores, float* angles, int cap) {
    const auto& c = h->lkps[level];
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) {
        xs[i] = (int)c[i].x; ys[i] = (int)c[i].y; scores[i] = (int)c[i].response;
        if (angles) angles[i] = c[i].angle;
    }
    return (int)c.size();
}
int orc_orb_num_fallback_cells(const orc_orb* h) { return h->fallbackCells; }

long orc_orb_extract_batch_mt(const orc_orb_params* p, const uint8_t* imgs, int nframes, int w, int hgt, int nthreads,
                              int want_desc, int* n_per_frame) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next(0);
    std::atomic<long> total(0);
    auto worker = [&]() {
        orc_orb* o = orc_orb_create(p);   // one extractor per thread: instances are not re-entrant (ORBextractor.h:105)
        const int cap = p->nfeatures + 3 * p->nlevels + 64;
        std::vector<orc_keypoint> kps(cap);
        std::vector<uint8_t> desc((size_t)cap * 32);
        for (;;) {
            int f = next.fetch_add(1);
            if (f >= nframes) break;
            int n = 0;
            orc_orb_extract(o, imgs + (size_t)f * w * hgt, w, hgt, w, 0, 1000, want_desc, kps.data(), desc.data(), cap, &n);
            if (n_per_frame) n_per_frame[f] = n;
            total += n;
        }
        orc_orb_destroy(o);
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(worker);
    for (auto& t : th) t.join();
    return total.load();
}

int orc_orb_tracked_desc(orc_orb* o, const uint8_t* img, int w, int hgt, size_t stride, const orc_keypoint* kps, int n, uint8_t* desc) {
    if (!img || w <= 0 || hgt <= 0) return -1;
    computePyramid(o, img, w, hgt, stride);
    o->blurred.assign(o->nlevels, {});
    for (int l = 0; l < o->nlevels; l++) {
        bool any = false;
        for (int i = 0; i < n; i++) any |= (kps[i].octave == l);
        if (!any) continue;
        blurLevel(o, l);
        float scale = o->invScale[l];
        for (int i = 0; i < n; i++) {
            if (kps[i].octave != l) continue;
            orc_keypoint kp = kps[i];
            kp.x *= scale; kp.y *= scale;
            computeOrbDescriptor(kp, o->blurred[l].data(), o->lw[l], o->lh[l], desc + (size_t)i * 32);
        }
    }
    return 0;
}

int orc_orb_assign_level_by_best_desc(orc_orb* o, const uint8_t* ref_desc, const uint8_t* img, int w, int hgt, size_t stride, orc_keypoint* kps, int n) {
    if (!img || w <= 0 || hgt <= 0) return -1;
    computePyramid(o, img, w, hgt, stride);
    o->blurred.assign(o->nlevels, {});
    std::vector<int> minDist(n, INT32_MAX);
    uint8_t d[32];
    for (int l = 0; l < o->nlevels; l++) {
        blurLevel(o, l);
        float scale = o->invScale[l];
        for (int i = 0; i < n; i++) {
            orc_keypoint kp = kps[i];
            kp.x *= scale; kp.y *= scale;
            computeOrbDescriptor(kp, o->blurred[l].data(), o->lw[l], o->lh[l], d);
            int dist = orc_descriptor_distance(ref_desc + (size_t)i * 32, d);
            if (dist < minDist[i]) { minDist[i] = dist; kps[i].octave = l; }
        }
    }
    return 0;
}

}  // extern "C"
