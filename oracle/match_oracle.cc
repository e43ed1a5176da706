// oracle/match_oracle.cc — CPU ORACLE (test infrastructure, never linked into the product).
//
// Restates reference src/ORBmatcher.cc: DescriptorDistance :2360-2378 (identical copy in
// Thirdparty/DBoW2/DBoW2/FORB.cpp:81-101), the sequential best/second-best scan with strict '<'
// (:741-770 SearchForInitialization, :318-378 SearchByBoW), the acceptance tests
// (:380-382  best<=TH_LOW && (float)best < nnratio*(float)second ;  :770-772  best<=TH_LOW && best < (float)second*nnratio)
// and the rotation histogram (:784-794, :800-823) + ComputeThreeMaxima (:2314-2355).
// The brute-force-over-a-database shape is the reference's own Frame.cc:1228-1235 (BFMatcher.knnMatch k=2 + ratio).
// Tie rule: lowest database index wins (sequential scan, strict '<'), pinned against cv2.BFMatcher in tests.
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <thread>
#include <vector>

extern "C" {

int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    const int32_t* pa = reinterpret_cast<const int32_t*>(a);
    const int32_t* pb = reinterpret_cast<const int32_t*>(b);
    int dist = 0;
    for (int i = 0; i < 8; i++, pa++, pb++) {
        unsigned int v = *pa ^ *pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

static inline int hamming256(const uint64_t* a, const uint64_t* b) {
    return __builtin_popcountll(a[0] ^ b[0]) + __builtin_popcountll(a[1] ^ b[1]) +
           __builtin_popcountll(a[2] ^ b[2]) + __builtin_popcountll(a[3] ^ b[3]);
}

// ratio_mode 0: (float)best < ratio*(float)second   (ORBmatcher.cc:382)
// ratio_mode 1: best < (float)second*ratio          (ORBmatcher.cc:772) — same value, kept for fidelity
void orc_hamming_best2(const uint8_t* q, int nq, const uint8_t* db, int64_t ndb, int th, float ratio, int ratio_mode,
                       orc_match* out, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    // Work item = a block of QB queries; the database is walked in tiles of TILE rows that stay in cache while the block's
    // queries scan them.  Every query still sees the rows in ascending order, so the result is the sequential scan's.
    const int QB = 16;
    const int64_t TILE = 2048;
    const int nblocks = (nq + QB - 1) / QB;
    std::atomic<int> next(0);
    auto worker = [&]() {
        for (;;) {
            const int b = next.fetch_add(1);
            if (b >= nblocks) break;
            const int q0 = b * QB, q1 = std::min(nq, q0 + QB);
            int best[QB], second[QB], idx[QB];
            for (int k = 0; k < QB; k++) { best[k] = 256; second[k] = 256; idx[k] = -1; }   // 256 == "no candidate" (SearchByBoW :318-320)
            for (int64_t t0 = 0; t0 < ndb; t0 += TILE) {
                const int64_t t1 = std::min(ndb, t0 + TILE);
                for (int i = q0; i < q1; i++) {
                    const uint64_t* qa = reinterpret_cast<const uint64_t*>(q + (size_t)i * 32);
                    int bs = best[i - q0], sc = second[i - q0], ix = idx[i - q0];
                    for (int64_t j = t0; j < t1; j++) {
                        int d = hamming256(qa, reinterpret_cast<const uint64_t*>(db + (size_t)j * 32));
                        if (d < bs) { sc = bs; bs = d; ix = (int)j; }
                        else if (d < sc) sc = d;
                    }
                    best[i - q0] = bs; second[i - q0] = sc; idx[i - q0] = ix;
                }
            }
            for (int i = q0; i < q1; i++) {
                orc_match m;
                const int bs = best[i - q0], sc = second[i - q0], ix = idx[i - q0];
                m.best_dist = bs; m.best_idx = ix; m.second_dist = sc;
                bool okr = ratio_mode ? ((float)bs < (float)sc * ratio) : ((float)bs < ratio * (float)sc);
                m.accepted = (ix >= 0 && bs <= th && okr) ? 1 : 0;
                out[i] = m;
            }
        }
    };
    std::vector<std::thread> t;
    for (int k = 0; k < nthreads; k++) t.emplace_back(worker);
    for (auto& x : t) x.join();
}

// Rotation-consistency filter exactly as written in the reference, including its quirk:
// factor = 1.0f/HISTO_LENGTH (HISTO_LENGTH=30), so bin = round(rot/30) only ever hits bins 0..12.
int orc_rotation_filter(const float* angle1, const float* angle2, int32_t* match12, int n1) {
    const int L = 30;
    const float factor = 1.0f / L;
    std::vector<int> hist[L];
    int nmatches = 0;
    for (int i1 = 0; i1 < n1; i1++) {
        if (match12[i1] < 0) continue;
        nmatches++;
        float rot = angle1[i1] - angle2[match12[i1]];
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)std::round(rot * factor);
        if (bin == L) bin = 0;
        hist[bin].push_back(i1);
    }
    int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = (int)hist[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
    for (int i = 0; i < L; i++) {
        if (i == ind1 || i == ind2 || i == ind3) continue;
        for (int idx1 : hist[i])
            if (match12[idx1] >= 0) { match12[idx1] = -1; nmatches--; }
    }
    return nmatches;
}

}  // extern "C"
