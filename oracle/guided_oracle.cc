// oracle/guided_oracle.cc — CPU ORACLE (test infrastructure, never linked into the product).
//
// Restates the guided-matching callers of the Hamming kernel:
//   Frame::PosInGrid               src/Frame.cc:783-793   (64 x 48 grid, FRAME_GRID_COLS/ROWS include/Frame.h:45-46)
//   Frame::AssignFeaturesToGrid    src/Frame.cc:431-460   (cell lists in keypoint-index order)
//   Frame::GetFeaturesInArea       src/Frame.cc:709-777   (cells column by column, rows inside, list order inside)
//   ORBmatcher::SearchForInitialization  src/ORBmatcher.cc:714-831 (stateful sequential scan: vMatchedDistance /
//        vnMatches21 take-over, TH_LOW, nnratio, rotation histogram WITH the entries of matches that were taken over
//        later, ComputeThreeMaxima :2314-2355, vbPrevMatched update)
// The reference has no tests for these; tests/test_oracle_guided.py pins this file against an independent pure-Python
// restatement written from the same reference lines.
#include "oracle.h"

#include <climits>
#include <cmath>
#include <vector>

namespace {
const int GC = 64, GR = 48;   // FRAME_GRID_COLS, FRAME_GRID_ROWS

struct GridGeom {
    float minX, minY, wInv, hInv;
    explicit GridGeom(const float* b) {
        minX = b[0]; minY = b[1];
        // Frame.cc:165-166 / :268-269: static_cast<float>(FRAME_GRID_COLS)/(mnMaxX-mnMinX), all float
        wInv = (float)GC / (b[2] - b[0]);
        hInv = (float)GR / (b[3] - b[1]);
    }
};

bool posInGrid(const GridGeom& g, float x, float y, int& px, int& py) {
    px = (int)std::round((x - g.minX) * g.wInv);   // std::round(float): half away from zero
    py = (int)std::round((y - g.minY) * g.hInv);
    return !(px < 0 || px >= GC || py < 0 || py >= GR);
}

void featuresInArea(const orc_keypoint* kps, const GridGeom& g, const int* cellStart, const int* cellIdx, float x, float y, float r,
                    int minLevel, int maxLevel, std::vector<int>& out) {
    out.clear();
    const float fx = r, fy = r;
    const int nMinCellX = std::max(0, (int)std::floor((x - g.minX - fx) * g.wInv));
    if (nMinCellX >= GC) return;
    const int nMaxCellX = std::min(GC - 1, (int)std::ceil((x - g.minX + fx) * g.wInv));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)std::floor((y - g.minY - fy) * g.hInv));
    if (nMinCellY >= GR) return;
    const int nMaxCellY = std::min(GR - 1, (int)std::ceil((y - g.minY + fy) * g.hInv));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const int c = ix * GR + iy;
            for (int j = cellStart[c]; j < cellStart[c + 1]; j++) {
                const orc_keypoint& kp = kps[cellIdx[j]];
                if (bCheckLevels) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                const float dx = kp.x - x, dy = kp.y - y;
                if (std::fabs(dx) < fx && std::fabs(dy) < fy) out.push_back(cellIdx[j]);
            }
        }
}
}  // namespace

extern "C" {

int orc_frame_grid(const orc_keypoint* kps, int n, const float* bounds4, int* cell_start, int* cell_idx) {
    const GridGeom g(bounds4);
    std::vector<std::vector<int>> cells(GC * GR);
    int assigned = 0;
    for (int i = 0; i < n; i++) {
        int px, py;
        if (posInGrid(g, kps[i].x, kps[i].y, px, py)) { cells[px * GR + py].push_back(i); assigned++; }
    }
    int k = 0;
    for (int c = 0; c < GC * GR; c++) {
        cell_start[c] = k;
        for (int i : cells[c]) cell_idx[k++] = i;
    }
    cell_start[GC * GR] = k;
    return assigned;
}

int orc_features_in_area(const orc_keypoint* kps, int n, const float* bounds4, const int* cell_start, const int* cell_idx, float x,
                         float y, float r, int min_level, int max_level, int* out, int cap) {
    (void)n;
    std::vector<int> v;
    featuresInArea(kps, GridGeom(bounds4), cell_start, cell_idx, x, y, r, min_level, max_level, v);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = v[i];
    return (int)v.size();
}

int orc_search_for_initialization(const orc_keypoint* kps1, const uint8_t* desc1, int n1, const orc_keypoint* kps2, const uint8_t* desc2,
                                  int n2, const float* bounds4, float* prev_xy, int window_size, float nnratio, int check_ori,
                                  int32_t* matches12) {
    const int TH_LOW = 50, HISTO_LENGTH = 30;
    const GridGeom g(bounds4);
    std::vector<int> cellStart(GC * GR + 1), cellIdx(std::max(n2, 1));
    orc_frame_grid(kps2, n2, bounds4, cellStart.data(), cellIdx.data());

    int nmatches = 0;
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> vMatchedDistance(std::max(n2, 1), INT_MAX), vnMatches21(std::max(n2, 1), -1);
    std::vector<int> vIndices2;
    for (int i1 = 0; i1 < n1; i1++) {
        const int level1 = kps1[i1].octave;
        if (level1 > 0) continue;
        featuresInArea(kps2, g, cellStart.data(), cellIdx.data(), prev_xy[2 * i1], prev_xy[2 * i1 + 1], (float)window_size, level1, level1,
                       vIndices2);
        if (vIndices2.empty()) continue;
        const uint8_t* d1 = desc1 + (size_t)i1 * 32;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            const int dist = orc_descriptor_distance(d1, desc2 + (size_t)i2 * 32);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) { matches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                matches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (check_ori) {
                    float rot = kps1[i1].angle - kps2[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)std::round(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rotHist[bin].push_back(i1);
                }
            }
        }
    }
    if (check_ori) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i])
                if (matches12[idx1] >= 0) { matches12[idx1] = -1; nmatches--; }
        }
    }
    for (int i1 = 0; i1 < n1; i1++)
        if (matches12[i1] >= 0) { prev_xy[2 * i1] = kps2[matches12[i1]].x; prev_xy[2 * i1 + 1] = kps2[matches12[i1]].y; }
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, float th, bool bMono = true)
// (src/ORBmatcher.cc:1969-2150, monocular path: bForward = bBackward = false, mvuRight < 0, Nleft == -1).
// x3Dc = Rcw * x3Dw + tcw is formed by the CALLER (cv::Mat arithmetic stays with OpenCV); everything after it is restated:
// invzc < 0 test (:2007-2010), Pinhole::project in float (src/CameraModels/Pinhole.cpp:30-33), image-bounds test (:2014-2017),
// radius = th * mvScaleFactors[clamp(octave)] (:2023; Frame::checkORBLevel src/Frame.cc:1408-1415), GetFeaturesInArea with
// levels [octave-1, octave+1] (:2032), best distance among candidates whose slot is not held by a map point with observations
// (:2046-2048; a slot taken by a point WITHOUT observations can be overwritten and counted again, as the reference does),
// TH_HIGH (:2068), rotation histogram of the CLAIMED current-frame indices (:2073-2089) and the final un-setting of the
// claims in the non-maximal bins (:2141-2150, one nmatches-- per histogram entry).
// valid1[i] = LastFrame has a map point at i that is not an outlier; obs1[i] = that point's Observations().
// match_cur[i2] = last-frame index whose map point ends up in CurrentFrame.mvpMapPoints[i2], or -1.  z == 0 (inf / NaN
// projection; the reference's behaviour is undefined there) is treated like z < 0.
int orc_search_by_projection(const float* x3Dc, const uint8_t* valid1, const int32_t* obs1, const orc_keypoint* kps1,
                             const uint8_t* descMP, int n1, const orc_keypoint* kps2, const uint8_t* desc2, int n2, const float* bounds4,
                             const float* K4, const float* scale_factors, int nlevels, float th, int check_ori, int32_t* match_cur) {
    return orc_search_by_projection_ex(x3Dc, valid1, obs1, kps1, descMP, n1, kps2, desc2, n2, bounds4, K4, scale_factors, nlevels, th, check_ori,
                                       0, 0.f, nullptr, match_cur);
}

// The same with the rectified-stereo branches (Nleft == -1, mvuRight set): level_mode = 1 when bForward, 2 when bBackward
// (tlc.z against the baseline mb, :1989-1990; only for !bMono), which turn the level window [oct-1, oct+1] into [oct, inf) /
// [0, oct] (:2024-2029); u_right2 = CurrentFrame.mvuRight (n2 floats, null for a monocular frame) and mbf = CurrentFrame.mbf: a
// candidate with a right coordinate is dropped when |(u - mbf * invzc) - uRight| > radius (:2049-2055).
int orc_search_by_projection_ex(const float* x3Dc, const uint8_t* valid1, const int32_t* obs1, const orc_keypoint* kps1,
                                const uint8_t* descMP, int n1, const orc_keypoint* kps2, const uint8_t* desc2, int n2, const float* bounds4,
                                const float* K4, const float* scale_factors, int nlevels, float th, int check_ori, int level_mode, float mbf,
                                const float* u_right2, int32_t* match_cur) {
    const int TH_HIGH = 100, HISTO_LENGTH = 30;
    const GridGeom g(bounds4);
    std::vector<int> cellStart(GC * GR + 1), cellIdx(std::max(n2, 1));
    orc_frame_grid(kps2, n2, bounds4, cellStart.data(), cellIdx.data());
    for (int i = 0; i < n2; i++) match_cur[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> vIndices2;
    for (int i = 0; i < n1; i++) {
        if (!valid1[i]) continue;
        const float xc = x3Dc[3 * i], yc = x3Dc[3 * i + 1], zc = x3Dc[3 * i + 2];
        const float invzc = (float)(1.0 / zc);
        if (invzc < 0 || zc == 0) continue;
        const float u = K4[0] * xc / zc + K4[2], v = K4[1] * yc / zc + K4[3];
        if (u < bounds4[0] || u > bounds4[2]) continue;
        if (v < bounds4[1] || v > bounds4[3]) continue;
        const int oct = kps1[i].octave;
        const int lv = oct < 0 ? 0 : (oct >= nlevels ? nlevels - 1 : oct);
        const float radius = th * scale_factors[lv];
        if (level_mode == 1) featuresInArea(kps2, g, cellStart.data(), cellIdx.data(), u, v, radius, oct, -1, vIndices2);
        else if (level_mode == 2) featuresInArea(kps2, g, cellStart.data(), cellIdx.data(), u, v, radius, 0, oct, vIndices2);
        else featuresInArea(kps2, g, cellStart.data(), cellIdx.data(), u, v, radius, oct - 1, oct + 1, vIndices2);
        if (vIndices2.empty()) continue;
        const uint8_t* dMP = descMP + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            if (match_cur[i2] >= 0 && obs1[match_cur[i2]] > 0) continue;
            if (u_right2 && u_right2[i2] > 0) {
                const float ur = u - mbf * invzc;
                const float er = std::fabs(ur - u_right2[i2]);
                if (er > radius) continue;
            }
            const int dist = orc_descriptor_distance(dMP, desc2 + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            match_cur[bestIdx2] = i;
            nmatches++;
            if (check_ori) {
                float rot = kps1[i].angle - kps2[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (check_ori) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i != ind1 && i != ind2 && i != ind3)
                for (int idx : rotHist[i]) { match_cur[idx] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, float th, bool bFarPoints, float thFarPoints)
// (src/ORBmatcher.cc:44-218; Tracking::SearchLocalPoints, src/Tracking-1.cc:2379), monocular frame: Nleft == -1 and
// mvuRight < 0, so the right-camera blocks (:91-96, :139-141, :150-211) never run.  Per map point in list order: the
// mbTrackInView / far / isBad gates (:54-61), r = RadiusByViewingCos(mTrackViewCos) (:219-225, a float compared with the DOUBLE
// 0.998), r *= th when th != 1.0 (:70-71), GetFeaturesInArea(projX, projY, r * mvScaleFactors[clamp(level)], level-1, level)
// (:73-75), best / second-best over the candidates whose slot does not hold a map point with observations (:87-89) with the
// levels of both (:104-124), TH_HIGH and the ratio test only when both lie on the same level (:128-131), setMapPoint (:134).
// pts[i]: the MapPoint tracking fields Frame::isInFrustum fills; held2[i2] != 0: F already holds a point with observations
// at i2 on entry (may be null).  match_cur[i2] = index of the map point this call put at i2, or -1; returns nmatches.
int orc_search_by_projection_map_points(const orc_track_point* pts, const uint8_t* descMP, int n1, const orc_keypoint* kps2,
                                        const uint8_t* desc2, const uint8_t* held2, int n2, const float* bounds4,
                                        const float* scale_factors, int nlevels, float th, int far_points, float th_far, float nnratio,
                                        int32_t* match_cur) {
    return orc_search_by_projection_map_points_ex(pts, nullptr, descMP, n1, kps2, desc2, held2, nullptr, n2, bounds4, scale_factors, nlevels, th,
                                                  far_points, th_far, nnratio, match_cur);
}

// The same with the rectified-stereo test (:91-96): proj_xr[i] = mTrackProjXR of map point i, u_right2 = F.mvuRight; a candidate
// with a right coordinate is dropped when |mTrackProjXR - uRight| > r * mvScaleFactors[level].
int orc_search_by_projection_map_points_ex(const orc_track_point* pts, const float* proj_xr, const uint8_t* descMP, int n1,
                                           const orc_keypoint* kps2, const uint8_t* desc2, const uint8_t* held2, const float* u_right2, int n2,
                                           const float* bounds4, const float* scale_factors, int nlevels, float th, int far_points,
                                           float th_far, float nnratio, int32_t* match_cur) {
    const int TH_HIGH = 100;
    const GridGeom g(bounds4);
    std::vector<int> cellStart(GC * GR + 1), cellIdx(std::max(n2, 1));
    orc_frame_grid(kps2, n2, bounds4, cellStart.data(), cellIdx.data());
    for (int i = 0; i < n2; i++) match_cur[i] = -1;
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    std::vector<int> vIndices;
    for (int iMP = 0; iMP < n1; iMP++) {
        const orc_track_point& p = pts[iMP];
        if (!p.in_view) continue;
        if (far_points && p.depth > th_far) continue;
        if (p.bad) continue;
        const int nPredictedLevel = p.scale_level;
        float r = (p.view_cos > 0.998) ? 2.5f : 4.0f;
        if (bFactor) r *= th;
        const int lv = nPredictedLevel < 0 ? 0 : (nPredictedLevel >= nlevels ? nlevels - 1 : nPredictedLevel);
        featuresInArea(kps2, g, cellStart.data(), cellIdx.data(), p.proj_x, p.proj_y, r * scale_factors[lv], nPredictedLevel - 1,
                       nPredictedLevel, vIndices);
        if (vIndices.empty()) continue;
        const uint8_t* dMP = descMP + (size_t)iMP * 32;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : vIndices) {
            if (match_cur[idx] >= 0 ? pts[match_cur[idx]].observations > 0 : (held2 && held2[idx])) continue;
            if (u_right2 && proj_xr && u_right2[idx] > 0) {
                const float er = std::fabs(proj_xr[iMP] - u_right2[idx]);
                if (er > r * scale_factors[lv]) continue;
            }
            const int dist = orc_descriptor_distance(dMP, desc2 + (size_t)idx * 32);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kps2[idx].octave; bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = kps2[idx].octave; bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            match_cur[bestIdx] = iMP;
            nmatches++;
        }
    }
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, float th, int ORBdist)
// (src/ORBmatcher.cc:2189-2312; Tracking::Relocalization, src/Tracking-1.cc:2700-2730 after the PnP refinement).  Per keyframe map point
// in index order: present, not bad, not already found (valid1, formed by the caller); projection WITHOUT a depth-sign test (:2218,
// unlike the frame-to-frame search) and the image-bounds test (:2220-2223); the distance-invariance gate and PredictScale (:2226-2237)
// are host arithmetic on the map point and arrive as level1[i] (valid1[i] = 0 when the gate fails); window th * scale[level],
// levels [level-1, level+1] (:2240-2242); best distance over the candidates whose slot holds NO map point at all (:2253-2254: held on
// entry or set earlier in this call); accepted when <= ORBdist (:2266); rotation histogram of the keyframe keypoint's angle against the
// frame keypoint's (:2273-2281) and the three-maxima filter (:2287-2308).
int orc_search_by_projection_reloc(const float* x3Dc, const uint8_t* valid1, const int32_t* level1, const orc_keypoint* kps1,
                                   const uint8_t* descMP, int n1, const orc_keypoint* kps2, const uint8_t* desc2, const uint8_t* held2, int n2,
                                   const float* bounds4, const float* K4, const float* scale_factors, int nlevels, float th, int orb_dist,
                                   int check_ori, int32_t* match_cur) {
    const int HISTO_LENGTH = 30;
    const GridGeom g(bounds4);
    std::vector<int> cellStart(GC * GR + 1), cellIdx(std::max(n2, 1));
    orc_frame_grid(kps2, n2, bounds4, cellStart.data(), cellIdx.data());
    for (int i = 0; i < n2; i++) match_cur[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> vIndices2;
    for (int i = 0; i < n1; i++) {
        if (!valid1[i]) continue;
        const float xc = x3Dc[3 * i], yc = x3Dc[3 * i + 1], zc = x3Dc[3 * i + 2];
        const float u = K4[0] * xc / zc + K4[2], v = K4[1] * yc / zc + K4[3];
        if (u < bounds4[0] || u > bounds4[2]) continue;
        if (v < bounds4[1] || v > bounds4[3]) continue;
        const int lvl = level1[i];
        const int lv = lvl < 0 ? 0 : (lvl >= nlevels ? nlevels - 1 : lvl);
        const float radius = th * scale_factors[lv];
        featuresInArea(kps2, g, cellStart.data(), cellIdx.data(), u, v, radius, lvl - 1, lvl + 1, vIndices2);
        if (vIndices2.empty()) continue;
        const uint8_t* dMP = descMP + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            if (match_cur[i2] >= 0 || (held2 && held2[i2])) continue;
            const int dist = orc_descriptor_distance(dMP, desc2 + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= orb_dist) {
            match_cur[bestIdx2] = i;
            nmatches++;
            if (check_ori) {
                float rot = kps1[i].angle - kps2[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (check_ori) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i != ind1 && i != ind2 && i != ind3)
                for (int idx : rotHist[i]) { match_cur[idx] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (src/ORBmatcher.cc:276-478;
// Tracking::TrackReferenceKeyFrame src/Tracking-1.cc:1680, Relocalization :2625), monocular frame (Nleft == -1: the right-camera
// bookkeeping :345-363, :410-438 never fires).  The two FeatureVectors arrive in the CSR form of orc_vocab_transform (node ids
// ascending = std::map order, features of a node in ascending index = push_back order).  Walk of the common nodes (:296-447; the
// lower_bound jumps are a plain merge of two ascending lists), per keyframe feature with a good map point (:307-313) the best /
// second-best distance over the node's frame features that are still unmatched (:326-341), TH_LOW and the ratio test in float
// (:370-372), rotation histogram of the matched frame indices (:382-406) and the final filter (:449-470).
// valid_kf[i] = the keyframe has a map point at i that is not bad; match_f[i2] = keyframe feature index whose map point ends up in
// vpMapPointMatches[i2], or -1; returns nmatches.
int orc_search_by_bow(const orc_keypoint* kps_kf, const uint8_t* desc_kf, const uint8_t* valid_kf, const uint32_t* kf_nodes,
                      const int32_t* kf_start, const uint32_t* kf_feats, int nkf, const orc_keypoint* kps_f, const uint8_t* desc_f, int n2,
                      const uint32_t* f_nodes, const int32_t* f_start, const uint32_t* f_feats, int nf, float nnratio, int check_ori,
                      int32_t* match_f) {
    const int TH_LOW = 50, HISTO_LENGTH = 30;
    for (int i = 0; i < n2; i++) match_f[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int a = 0, b = 0;
    while (a < nkf && b < nf) {
        if (kf_nodes[a] == f_nodes[b]) {
            for (int iKF = kf_start[a]; iKF < kf_start[a + 1]; iKF++) {
                const unsigned realIdxKF = kf_feats[iKF];
                if (!valid_kf[realIdxKF]) continue;
                const uint8_t* dKF = desc_kf + (size_t)realIdxKF * 32;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (int iF = f_start[b]; iF < f_start[b + 1]; iF++) {
                    const unsigned realIdxF = f_feats[iF];
                    if (match_f[realIdxF] >= 0) continue;
                    const int dist = orc_descriptor_distance(dKF, desc_f + (size_t)realIdxF * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = (int)realIdxF; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 <= TH_LOW && (float)bestDist1 < nnratio * (float)bestDist2) {
                    match_f[bestIdxF] = (int)realIdxKF;
                    if (check_ori) {
                        float rot = kps_kf[realIdxKF].angle - kps_f[bestIdxF].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        if (bin >= 0 && bin < HISTO_LENGTH) rotHist[bin].push_back(bestIdxF);
                    }
                    nmatches++;
                }
            }
            a++; b++;
        } else if (kf_nodes[a] < f_nodes[b]) a++;
        else b++;
    }
    if (check_ori) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int idx : rotHist[i]) { match_f[idx] = -1; nmatches--; }
    }
    return nmatches;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) (src/ORBmatcher.cc:833-990; loop closing /
// place recognition), monocular keyframes.  Same walk of the common nodes as the keyframe-frame form, with three differences: the
// candidates of the second keyframe need a good map point too (:889-893), the acceptance is `bestDist1 < TH_LOW` (strict, :912), and the
// result is indexed by the FIRST keyframe's feature (vpMatches12[idx1] = map point of idx2, :916; the histogram holds idx1, :929).
// valid1 / valid2 = a map point that is not bad; match12[i1] = feature index of keyframe 2 or -1; returns nmatches.
int orc_search_by_bow_kf(const orc_keypoint* kps1, const uint8_t* desc1, const uint8_t* valid1, const uint32_t* nodes1, const int32_t* start1,
                         const uint32_t* feats1, int nn1, int n1, const orc_keypoint* kps2, const uint8_t* desc2, const uint8_t* valid2,
                         const uint32_t* nodes2, const int32_t* start2, const uint32_t* feats2, int nn2, int n2, float nnratio, int check_ori,
                         int32_t* match12) {
    const int TH_LOW = 50, HISTO_LENGTH = 30;
    for (int i = 0; i < n1; i++) match12[i] = -1;
    std::vector<bool> vbMatched2((size_t)n2, false);
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (nodes1[a] == nodes2[b]) {
            for (int i1 = start1[a]; i1 < start1[a + 1]; i1++) {
                const unsigned idx1 = feats1[i1];
                if (!valid1[idx1]) continue;
                const uint8_t* d1 = desc1 + (size_t)idx1 * 32;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int i2 = start2[b]; i2 < start2[b + 1]; i2++) {
                    const unsigned idx2 = feats2[i2];
                    if (vbMatched2[idx2] || !valid2[idx2]) continue;
                    const int dist = orc_descriptor_distance(d1, desc2 + (size_t)idx2 * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = (int)idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < TH_LOW && (float)bestDist1 < nnratio * (float)bestDist2) {
                    match12[idx1] = bestIdx2;
                    vbMatched2[(size_t)bestIdx2] = true;
                    if (check_ori) {
                        float rot = kps1[idx1].angle - kps2[bestIdx2].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        if (bin >= 0 && bin < HISTO_LENGTH) rotHist[bin].push_back((int)idx1);
                    }
                    nmatches++;
                }
            }
            a++; b++;
        } else if (nodes1[a] < nodes2[b]) a++;
        else b++;
    }
    if (check_ori) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int idx : rotHist[i]) { match12[idx] = -1; nmatches--; }
    }
    return nmatches;
}

// The matching core the keyframe-side searches share -- ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, ...)
// (src/ORBmatcher.cc:480-593 and :595-712), ORBmatcher::Fuse (:1407-1617 and :1619-1741), ORBmatcher::SearchBySim3 (:1743-1967):
// per map point i, in list order, the window the caller formed (x, y, r <= 0: the point failed a gate; KeyFrame::GetFeaturesInArea,
// src/KeyFrame.cc:873-917, has no level arguments: the level test `kpLevel < nPredictedLevel-1 || kpLevel > nPredictedLevel` sits in
// the candidate loop, :568-569 / :1526-1527 / :1713-1714 / :1846-1847), the smallest DescriptorDistance with strict `<`, accepted when
// <= th_high.  blocking: `if(vpMatched[idx]) continue;` (:563, :678) with vpMatched = held2 on entry + this call's own matches.
// inv_level_sigma2 != NULL: Fuse's reprojection gate (:1532-1558) in float, the product compared with the double constants.
int orc_search_windows(const orc_area_query* queries, const float* ur, const uint8_t* descMP, int n1, const orc_keypoint* kps2, const uint8_t* desc2,
                       const uint8_t* held2, const float* u_right2, int n2, const float* bounds4, const float* query_min_xy,
                       const float* inv_level_sigma2, int nlevels, int blocking, int th_high, int32_t* best_idx, int32_t* best_dist,
                       int32_t* match2) {
    GridGeom g(bounds4);      // the keypoints are binned with the Frame's float bounds (orc_frame_grid below) ...
    if (query_min_xy) { g.minX = query_min_xy[0]; g.minY = query_min_xy[1]; }   // ... the lookup subtracts the KeyFrame's int bounds
    std::vector<int> cellStart(GC * GR + 1), cellIdx((size_t)std::max(n2, 1));
    orc_frame_grid(kps2, n2, bounds4, cellStart.data(), cellIdx.data());
    std::vector<uint8_t> taken((size_t)std::max(n2, 1), 0);
    for (int i = 0; i < n2; i++) { taken[i] = held2 && held2[i] ? 1 : 0; if (match2) match2[i] = -1; }
    int nmatches = 0;
    std::vector<int> vIndices;
    for (int i = 0; i < n1; i++) {
        best_idx[i] = -1;
        if (best_dist) best_dist[i] = 256;
        const orc_area_query& q = queries[i];
        if (!(q.r > 0.0f)) continue;
        featuresInArea(kps2, g, cellStart.data(), cellIdx.data(), q.x, q.y, q.r, -1, -1, vIndices);
        const uint8_t* dMP = descMP + (size_t)i * 32;
        int bestDist = 256, bestIdx = -1;
        for (int idx : vIndices) {
            if (taken[idx]) continue;
            const orc_keypoint& kp = kps2[idx];
            const int kpLevel = kp.octave;
            if (q.min_level > 0 || q.max_level >= 0) {
                if (kpLevel < q.min_level) continue;
                if (q.max_level >= 0 && kpLevel > q.max_level) continue;
            }
            if (inv_level_sigma2) {
                const int lv = kpLevel < 0 ? 0 : (kpLevel >= nlevels ? nlevels - 1 : kpLevel);
                const float ex = q.x - kp.x, ey = q.y - kp.y;
                if (u_right2 && u_right2[idx] >= 0) {
                    const float er = (ur ? ur[i] : 0.0f) - u_right2[idx];
                    const float e2 = ex * ex + ey * ey + er * er;
                    if (e2 * inv_level_sigma2[lv] > 7.8) continue;
                } else {
                    const float e2 = ex * ex + ey * ey;
                    if (e2 * inv_level_sigma2[lv] > 5.99) continue;
                }
            }
            const int dist = orc_descriptor_distance(dMP, desc2 + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        // reported distance: the best candidate's (non-blocking searches); the claimed keypoint's, 256 without a claim (blocking searches)
        if (best_dist) best_dist[i] = (!blocking || (bestIdx >= 0 && bestDist <= th_high)) ? bestDist : 256;
        if (bestIdx >= 0 && bestDist <= th_high) {
            best_idx[i] = bestIdx;
            nmatches++;
            if (blocking) { taken[bestIdx] = 1; if (match2) match2[bestIdx] = i; }
        }
    }
    return nmatches;
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo, bCoarse) (src/ORBmatcher.cc:975-1214), keyframes with one
// pinhole camera each (mpCamera2 == NULL).  flags bit 0: the feature takes part (no map point, :1034-1038 / :1057-1059; stereo when
// bOnlyStereo, :1040-1044 / :1061-1065), bit 1: bStereo.  F = the matrix Pinhole::epipolarConstrain forms (Pinhole.cpp:137-140), row-major;
// the gate itself is :143-156.  The loops and the running `dist > TH_LOW || dist > bestDist` rule are the reference's (vbMatched2 is never
// set in this version of the function); match12[i1] = bestIdx2 or -1 after the rotation filter (:1177-1198).
int orc_search_for_triangulation(const orc_keypoint* kps1, const uint8_t* desc1, const uint8_t* flags1, int n1, const uint32_t* nodes1,
                                 const int32_t* start1, const uint32_t* feats1, int nn1, const orc_keypoint* kps2, const uint8_t* desc2,
                                 const uint8_t* flags2, int n2, const uint32_t* nodes2, const int32_t* start2, const uint32_t* feats2, int nn2,
                                 const float* F, const float* ep, const float* scale2, const float* sigma2, int nlevels, int coarse, int check_ori,
                                 int32_t* match12) {
    (void)n2;
    const int TH_LOW = 50, HISTO_LENGTH = 30;
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0;
    for (int i = 0; i < n1; i++) match12[i] = -1;
    std::vector<int> rotHist[30];
    auto lvl = [&](int o) { return o < 0 ? 0 : (o >= nlevels ? nlevels - 1 : o); };
    int q1 = 0, q2 = 0;
    while (q1 < nn1 && q2 < nn2) {
        if (nodes1[q1] == nodes2[q2]) {
            for (int a = start1[q1]; a < start1[q1 + 1]; a++) {
                const int idx1 = (int)feats1[a];
                if (!(flags1[idx1] & 1)) continue;
                const bool bStereo1 = (flags1[idx1] & 2) != 0;
                const orc_keypoint& kp1 = kps1[idx1];
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int b = start2[q2]; b < start2[q2 + 1]; b++) {
                    const int idx2 = (int)feats2[b];
                    if (!(flags2[idx2] & 1)) continue;
                    const bool bStereo2 = (flags2[idx2] & 2) != 0;
                    const int dist = orc_descriptor_distance(desc1 + (size_t)idx1 * 32, desc2 + (size_t)idx2 * 32);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const orc_keypoint& kp2 = kps2[idx2];
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ep[0] - kp2.x, distey = ep[1] - kp2.y;
                        if (distex * distex + distey * distey < 100 * scale2[lvl(kp2.octave)]) continue;
                    }
                    bool ok = coarse != 0;
                    if (!ok) {   // Pinhole::epipolarConstrain
                        const float la = kp1.x * F[0] + kp1.y * F[3] + F[6];
                        const float lb = kp1.x * F[1] + kp1.y * F[4] + F[7];
                        const float lc = kp1.x * F[2] + kp1.y * F[5] + F[8];
                        const float num = la * kp2.x + lb * kp2.y + lc;
                        const float den = la * la + lb * lb;
                        if (den != 0) {
                            const float dsqr = num * num / den;
                            ok = dsqr < 3.84f * sigma2[lvl(kp2.octave)];
                        }
                    }
                    if (ok) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    match12[idx1] = bestIdx2;
                    nmatches++;
                    if (check_ori) {
                        float rot = kp1.angle - kps2[bestIdx2].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(idx1);
                    }
                }
            }
            q1++; q2++;
        } else if (nodes1[q1] < nodes2[q2]) {
            while (q1 < nn1 && nodes1[q1] < nodes2[q2]) q1++;      // lower_bound
        } else {
            while (q2 < nn2 && nodes2[q2] < nodes1[q1]) q2++;
        }
    }
    if (check_ori) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx : rotHist[i]) { match12[idx] = -1; nmatches--; }
        }
    }
    return nmatches;
}

}  // extern "C"
