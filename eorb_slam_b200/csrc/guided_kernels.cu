// guided_kernels.cu — guided matching around the Hamming kernel.
//
//   G1 frame_grid_kernel        Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:431-460, 783-793): 64 x 48 cells,
//                               lists in keypoint-index order: a counting sort by cell in shared memory (one block),
//                               stable in the index through a rank pass over each cell's few members.
//                               Cell id = col * 48 + row, so the cells GetFeaturesInArea visits for one column are ONE
//                               contiguous CSR range in exactly the reference's order (columns outer, rows inner).
//   G2 guided_candidates_kernel one warp per frame-1 keypoint (level 0 only, ORBmatcher.cc:732-734): walks the window's
//                               columns 32 entries at a time (ballot compaction keeps the reference's visiting order),
//                               computes the Hamming distance of every candidate and keeps the 32 smallest
//                               (distance, position) keys sorted across the lanes (shuffle bitonic sort + merge).
//   G3 guided_resolve_kernel    the stateful part of ORBmatcher::SearchForInitialization (:747-796): candidates already
//                               matched with a smaller-or-equal distance are skipped, a better query takes a frame-2
//                               keypoint over.  That is a sequential dependency over the queries, so ONE warp replays the
//                               queries (compacted to those with candidates) in order — but each step is only "first two unfiltered entries of a sorted
//                               32-entry head" (a ballot) with the whole state (matched distance, owner, angles) in shared
//                               memory; the other seven warps stage the next 64 heads behind it (double buffer).  The full-list scan is kept as the
//                               slow path for heads that run dry.  Rotation histogram (:784-794, with the entries of
//                               matches that are taken over later, as in the reference), ComputeThreeMaxima
//                               (:2314-2355), the filter (:800-823) and the vbPrevMatched update (:826-828) follow in
//                               the same launch.
#include <cuda_runtime.h>
#include <stdint.h>

#include "guided_kernels.h"

namespace eorb {

typedef unsigned long long u64;
#define GUIDED_STAGE 64     // queries staged per round of the resolve kernel
#define GUIDED_ROW 33       // u64 entries per staged head row: 32 + 1 pad, so that lanes walking DIFFERENT rows hit different banks
#define GUIDED_NONE 0xffffu
#ifndef EORB_GUIDED_SPEC
#define EORB_GUIDED_SPEC 1
#endif
#ifndef EORB_GUIDED_FIXPOINT
#define EORB_GUIDED_FIXPOINT 1   // local-map resolve: fixed-point passes over a chunk instead of restarts at contested lanes
#endif
#define FULLMASK 0xffffffffu

// staging of one round of sorted heads (GUIDED_STAGE queries x 32 entries) by NTH threads: every thread requests all of its
// entries before it stores the first one (the loop form waited for each L2 round trip in turn, and the staging warps, not the
// ordered warp, set the pace of the resolve kernels)
template <int NTH>
__device__ __forceinline__ void guided_load_heads(u64* __restrict__ sp, const u64* __restrict__ top, const unsigned short* qlist, int r, int nact, int t) {
    constexpr int PER = (GUIDED_STAGE * 32 + NTH - 1) / NTH;
    u64 v[PER];
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int e = t + i * NTH, q = r * GUIDED_STAGE + (e >> 5);
        v[i] = (e < GUIDED_STAGE * 32 && q < nact) ? top[(size_t)qlist[q] * EORB_GUIDED_TOP + (e & 31)] : ~0ull;
    }
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int e = t + i * NTH;
        if (e < GUIDED_STAGE * 32) sp[(e >> 5) * GUIDED_ROW + (e & 31)] = v[i];
    }
}

// order-preserving compaction of the queries that have candidates into qlist (256 threads): four batches of 256 flags are
// requested before the first ballot / scan round, so the block waits for L2 once per 1024 queries instead of once per 256
__device__ __forceinline__ int guided_compact(const int* __restrict__ candCnt, int n1, unsigned short* qlist, int* sWarp, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    int nact = 0;
    for (int base = 0; base < n1; base += 1024) {
        bool act[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int i1 = base + u * 256 + tid; act[u] = i1 < n1 && candCnt[i1] > 0; }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i1 = base + u * 256 + tid;
            const unsigned bm = __ballot_sync(FULLMASK, act[u]);
            if (lane == 0) sWarp[warp] = __popc(bm);
            __syncthreads();
            int before = 0, total = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { const int v = sWarp[k]; total += v; if (k < warp) before += v; }
            if (act[u]) qlist[nact + before + __popc(bm & ((1u << lane) - 1u))] = (unsigned short)i1;
            nact += total;
            __syncthreads();
        }
    }
    return nact;
}

// ---- GetFeaturesInArea cell window (Frame.cc:722-744), all float like the reference
__device__ __forceinline__ bool area_cells(const GuidedGrid& g, float x, float y, float r, int& c0, int& c1, int& r0, int& r1) {
    c0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, g.minX), r), g.wInv)));
    if (c0 >= EORB_GRID_COLS) return false;
    c1 = min(EORB_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, g.minX), r), g.wInv)));
    if (c1 < 0) return false;
    r0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, g.minY), r), g.hInv)));
    if (r0 >= EORB_GRID_ROWS) return false;
    r1 = min(EORB_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, g.minY), r), g.hInv)));
    if (r1 < 0) return false;
    return true;
}

// candidate test of one list entry (Frame.cc:756-773)
__device__ __forceinline__ bool area_accept(const eorb_keypoint* __restrict__ kps, int i2, float x, float y, float r, bool check, int minLevel,
                                            int maxLevel) {
    const eorb_keypoint& kp = kps[i2];
    if (check) {
        const int level = kp.octave;
        if (level < minLevel) return false;
        if (maxLevel >= 0 && level > maxLevel) return false;
    }
    return fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r;
}

// ------------------------------------------------------------------------------------------------ G1
// Counting sort by cell, stable in the keypoint index (push_back order): cell populations by shared-memory atomics, block scan
// of the 3072 counters, an unordered scatter into the cell ranges, then every keypoint finds its rank among the (few) members
// of its cell.  O(n + sum of squared cell sizes); a bitonic sort of the (cell, index) keys took 88 us for 5000 keypoints.
__global__ void __launch_bounds__(1024) frame_grid_kernel(const eorb_keypoint* __restrict__ kps, int n, GuidedGrid g, int* __restrict__ cellStart,
                                                          int* __restrict__ cellIdx, int* __restrict__ assigned) {
    extern __shared__ __align__(16) unsigned char gsm[];
    int* cnt = reinterpret_cast<int*>(gsm);                       // [EORB_GRID_CELLS + 1] populations -> exclusive starts
    int* fill = cnt + EORB_GRID_CELLS + 1;                        // [EORB_GRID_CELLS] scatter cursors
    unsigned short* cellOf = reinterpret_cast<unsigned short*>(fill + EORB_GRID_CELLS);   // [n] cell of keypoint i (0xffff: outside the grid)
    unsigned short* tmp = cellOf + ((n + 1) & ~1);                // [n] members of every cell, unordered
    __shared__ int sWarp[33];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c <= EORB_GRID_CELLS; c += 1024) { cnt[c] = 0; if (c < EORB_GRID_CELLS) fill[c] = 0; }
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        // PosInGrid: round() of a float is std::round(float) = roundf, half away from zero
        const int px = (int)roundf(__fmul_rn(__fsub_rn(kps[i].x, g.minX), g.wInv));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(kps[i].y, g.minY), g.hInv));
        unsigned short c = 0xffffu;
        if (px >= 0 && px < EORB_GRID_COLS && py >= 0 && py < EORB_GRID_ROWS) { c = (unsigned short)(px * EORB_GRID_ROWS + py); atomicAdd(&cnt[c], 1); }
        cellOf[i] = c;
    }
    __syncthreads();
    // exclusive scan of the 3072 populations: three per thread, warp shuffles, one exchange through shared memory
    {
        const int c0 = tid * 3;
        const int v0 = cnt[c0], v1 = cnt[c0 + 1], v2 = cnt[c0 + 2];
        const int sum = v0 + v1 + v2;
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int up = __shfl_up_sync(FULLMASK, incl, o); if (lane >= o) incl += up; }
        if (lane == 31) sWarp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = sWarp[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int up = __shfl_up_sync(FULLMASK, wi, o); if (lane >= o) wi += up; }
            sWarp[lane] = wi - w;
            if (lane == 31) sWarp[32] = wi;
        }
        __syncthreads();
        const int off = sWarp[warp] + incl - sum;
        cnt[c0] = off; cnt[c0 + 1] = off + v0; cnt[c0 + 2] = off + v0 + v1;
        if (tid == 0) { cnt[EORB_GRID_CELLS] = sWarp[32]; *assigned = sWarp[32]; }
    }
    __syncthreads();
    for (int c = tid; c <= EORB_GRID_CELLS; c += 1024) cellStart[c] = cnt[c];
    for (int i = tid; i < n; i += 1024) {
        const unsigned c = cellOf[i];
        if (c != 0xffffu) tmp[cnt[c] + atomicAdd(&fill[c], 1)] = (unsigned short)i;
    }
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        const unsigned c = cellOf[i];
        if (c == 0xffffu) continue;
        const int b = cnt[c], e = cnt[c + 1];
        int rank = 0;
        for (int j = b; j < e; j++) rank += (int)tmp[j] < i;
        cellIdx[b + rank] = i;
    }
}

// ------------------------------------------------------------------------------------------------ GetFeaturesInArea batch
__global__ void __launch_bounds__(256) features_in_area_kernel(const eorb_keypoint* __restrict__ kps, GuidedGrid g, const int* __restrict__ cellStart,
                                                               const int* __restrict__ cellIdx, const eorb_area_query* __restrict__ qs, int nq,
                                                               int* __restrict__ count, int* __restrict__ out, int cap) {
    const int q = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= nq) return;
    const eorb_area_query Q = qs[q];
    const unsigned lt = (1u << lane) - 1u;
    int c0, c1, r0, r1, cnt = 0;
    if (area_cells(g, Q.x, Q.y, Q.r, c0, c1, r0, r1)) {
        const bool check = Q.min_level > 0 || Q.max_level >= 0;
        for (int ix = c0; ix <= c1; ix++) {
            const int b = cellStart[ix * EORB_GRID_ROWS + r0], e = cellStart[ix * EORB_GRID_ROWS + r1 + 1];
            for (int base = b; base < e; base += 32) {
                const int j = base + lane;
                int i2 = -1;
                bool ok = false;
                if (j < e) { i2 = cellIdx[j]; ok = area_accept(kps, i2, Q.x, Q.y, Q.r, check, Q.min_level, Q.max_level); }
                const unsigned m = __ballot_sync(FULLMASK, ok);
                const int pos = cnt + __popc(m & lt);
                if (ok && pos < cap) out[(size_t)q * cap + pos] = i2;
                cnt += __popc(m);
            }
        }
    }
    if (lane == 0) count[q] = cnt;
}

// ------------------------------------------------------------------------------------------------ G2
__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int m) { return __shfl_xor_sync(FULLMASK, v, m); }

__device__ __forceinline__ u64 warp_sort32(u64 v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const u64 o = shfl_xor_u64(v, j);
            const bool up = (lane & k) == 0, lower = (lane & j) == 0;
            v = (lower == up) ? (v < o ? v : o) : (v < o ? o : v);
        }
    return v;
}

// Queries: SearchForInitialization (qs == nullptr): frame-1 keypoint i1 at level 0, window `r0win` around prevXY[i1], levels
// [0, 0] (ORBmatcher.cc:732-736).  SearchByProjection (qs != nullptr): the window and level range guided_project_kernel
// prepared for last-frame keypoint i1 (r <= 0: no query); f1.desc then holds the map points' descriptors.
// Rectified stereo (uRight2 != nullptr; Nleft == -1 with mvuRight set): a candidate with a right-image column is dropped when it lies
// further than the window radius from the query's predicted right column qUr[i1] (ORBmatcher.cc:91-96, :2049-2055).  The test does
// not depend on the order of the queries, so it is part of the candidate filter.
// Keyframe-side searches (GuidedCandExtra, eorb_guided_search_windows): held2 != nullptr drops keypoints that are taken on entry (a
// static filter: nothing is claimed during a non-blocking search); chi2 != 0 replaces the right-column test by the reprojection gate
// of ORBmatcher::Fuse (ORBmatcher.cc:1532-1558): e2 * invLevelSigma2[octave] > 7.8 with a right column (mvuRight >= 0), > 5.99
// without, float products compared against the double constants.
__global__ void __launch_bounds__(256) guided_candidates_kernel(GuidedFrame f1, GuidedFrame f2, GuidedGrid g, const float* __restrict__ prevXY,
                                                                float r0win, const eorb_area_query* __restrict__ qs, GuidedWork w,
                                                                const float* __restrict__ uRight2, const float* __restrict__ qUr,
                                                                GuidedCandExtra cx) {
    const int i1 = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i1 >= f1.n) return;
    const unsigned lt = (1u << lane) - 1u;
    if (lane == 0) w.bin[i1] = -1;
    float x, y, r;
    int minL, maxL;
    bool live;
    if (qs) {
        const eorb_area_query q = qs[i1];
        x = q.x; y = q.y; r = q.r; minL = q.min_level; maxL = q.max_level;
        live = r > 0.0f;
    } else {
        const int level1 = f1.kps[i1].octave;
        x = prevXY[2 * i1]; y = prevXY[2 * i1 + 1]; r = r0win; minL = level1; maxL = level1;
        live = level1 <= 0;
    }
    const bool check = minL > 0 || maxL >= 0;              // bCheckLevels (Frame.cc:746)
    const bool stereo = uRight2 != nullptr && qUr != nullptr;
    const float ur = qUr ? qUr[i1] : 0.0f;
    auto stereoOk = [&](int i2) {
        if (cx.held2 && cx.held2[i2]) return false;
        if (cx.chi2) {
            const eorb_keypoint& kp = f2.kps[i2];
            const int lv = kp.octave < 0 ? 0 : (kp.octave > 31 ? 31 : kp.octave);
            const float ex = __fsub_rn(x, kp.x), ey = __fsub_rn(y, kp.y);
            float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
            const float u2 = uRight2 ? uRight2[i2] : -1.0f;
            if (u2 >= 0.0f) {
                const float er = __fsub_rn(ur, u2);
                e2 = __fadd_rn(e2, __fmul_rn(er, er));
                return !((double)__fmul_rn(e2, cx.invSigma2[lv]) > 7.8);
            }
            return !((double)__fmul_rn(e2, cx.invSigma2[lv]) > 5.99);
        }
        if (!stereo) return true;
        const float u2 = uRight2[i2];
        return !(u2 > 0.0f && fabsf(__fsub_rn(ur, u2)) > r);
    };
    int c0 = 0, c1 = -1, r0 = 0, r1 = 0;
    live = live && area_cells(g, x, y, r, c0, c1, r0, r1);
    // pass 1: count
    int cnt = 0;
    if (live)
        for (int ix = c0; ix <= c1; ix++) {
            const int b = w.cellStart[ix * EORB_GRID_ROWS + r0], e = w.cellStart[ix * EORB_GRID_ROWS + r1 + 1];
            for (int base = b; base < e; base += 32) {
                const int j = base + lane;
                bool ok = false;
                if (j < e) { const int i2 = w.cellIdx[j]; ok = area_accept(f2.kps, i2, x, y, r, check, minL, maxL) && stereoOk(i2); }
                cnt += __popc(__ballot_sync(FULLMASK, ok));
            }
        }
    int off = 0;
    if (lane == 0 && cnt > 0) off = atomicAdd(w.total, cnt);
    off = __shfl_sync(FULLMASK, off, 0);
    if (lane == 0) { w.candCnt[i1] = cnt; w.candOff[i1] = off; }
    if (cnt == 0 || off + cnt > w.candCap) return;   // overflow: the host sees total > candCap, grows the buffer and retries

    // pass 2: distances, full list in visiting order, sorted head of the 32 smallest (distance, position) keys
    uint32_t qd[8];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(f1.desc + (size_t)i1 * 32);
        const uint4 a = __ldg(qp), b = __ldg(qp + 1);
        qd[0] = a.x; qd[1] = a.y; qd[2] = a.z; qd[3] = a.w; qd[4] = b.x; qd[5] = b.y; qd[6] = b.z; qd[7] = b.w;
    }
    u64 top = ~0ull;
    int pos0 = 0;
    for (int ix = c0; ix <= c1; ix++) {
        const int b = w.cellStart[ix * EORB_GRID_ROWS + r0], e = w.cellStart[ix * EORB_GRID_ROWS + r1 + 1];
        for (int base = b; base < e; base += 32) {
            const int j = base + lane;
            int i2 = 0;
            bool ok = false;
            if (j < e) { i2 = w.cellIdx[j]; ok = area_accept(f2.kps, i2, x, y, r, check, minL, maxL) && stereoOk(i2); }
            const unsigned m = __ballot_sync(FULLMASK, ok);
            if (m == 0) continue;
            u64 key = ~0ull;
            if (ok) {
                const uint4* dp = reinterpret_cast<const uint4*>(f2.desc + (size_t)i2 * 32);
                const uint4 a = __ldg(dp), c = __ldg(dp + 1);
                const int dist = __popc(a.x ^ qd[0]) + __popc(a.y ^ qd[1]) + __popc(a.z ^ qd[2]) + __popc(a.w ^ qd[3]) + __popc(c.x ^ qd[4]) +
                                 __popc(c.y ^ qd[5]) + __popc(c.z ^ qd[6]) + __popc(c.w ^ qd[7]);
                const int pos = pos0 + __popc(m & lt);
                w.cand[off + pos] = ((uint32_t)dist << 16) | (uint32_t)i2;
                key = ((u64)dist << 32) | ((u64)pos << 16) | (u64)i2;
            }
            pos0 += __popc(m);
            if (!__any_sync(FULLMASK, key < __shfl_sync(FULLMASK, top, 31))) continue;   // nothing here beats the current 32nd smallest
            key = warp_sort32(key, lane);
            const u64 rev = __shfl_sync(FULLMASK, key, 31 - lane);
            top = top < rev ? top : rev;          // the 32 smallest of both, as a bitonic sequence
#pragma unroll
            for (int jj = 16; jj > 0; jj >>= 1) {
                const u64 o = shfl_xor_u64(top, jj);
                top = (lane & jj) == 0 ? (top < o ? top : o) : (top < o ? o : top);
            }
        }
    }
    w.top[(size_t)i1 * EORB_GUIDED_TOP + lane] = top;
}

// ------------------------------------------------------------------------------------------------ G3
// Only what truly depends on the order of the queries stays in the one-warp loop: the filter against the matched
// distances and the two state writes.  vnMatches12 / nmatches fall out of the final owner table (a frame-2 keypoint's
// last claimer owns it; a query claims at most once), and the rotation bin of a claim only depends on (query, claimed
// keypoint), so both are computed by the whole block after the loop.  Inside the loop matches12[i1] records the claim.
__global__ void __launch_bounds__(256) guided_resolve_kernel(GuidedFrame f1, GuidedFrame f2, float* __restrict__ prevXY, float nnratio, int checkOri,
                                                             GuidedWork w, int32_t* __restrict__ matches12, int* __restrict__ nmatchesOut) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int n1 = f1.n, n2 = f2.n, n2r = (n2 + 3) & ~3;
    u64* stop = reinterpret_cast<u64*>(sm);                               // [2][GUIDED_STAGE][32] double-buffered heads
    int* scnt = reinterpret_cast<int*>(stop + 2 * GUIDED_STAGE * GUIDED_ROW);     // [2][GUIDED_STAGE]
    int* soff = scnt + 2 * GUIDED_STAGE;                                  // [2][GUIDED_STAGE]
    int* hist = soff + 2 * GUIDED_STAGE;                                  // [32]
    unsigned short* md = reinterpret_cast<unsigned short*>(hist + 32);    // [n2r] matched distance (0xffff = INT_MAX)
    unsigned short* m21 = md + n2r;                                       // [n2r] owner in frame 1 (0xffff = none)
    unsigned short* qlist = m21 + n2r;                                    // [n1] queries that have candidates, ascending
    unsigned* ctab = reinterpret_cast<unsigned*>(qlist + ((n1 + 1) & ~1));  // [n2r] lowest lane of the current chunk claiming a slot
    __shared__ int sNm, sInd[3], sWarp[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (*w.total > w.candCap) {   // candidate buffer overflow: nothing below may run on partial lists
        if (tid == 0) *nmatchesOut = -1;
        return;
    }
    for (int i = tid; i < n2; i += 256) { md[i] = GUIDED_NONE; m21[i] = GUIDED_NONE; ctab[i] = 0xffffffffu; }
    for (int i = tid; i < n1; i += 256) matches12[i] = -1;
    if (tid < 32) hist[tid] = 0;
    if (tid == 0) sNm = 0;
    // compact the queries that have candidates (level-0 keypoints with a non-empty window), keeping their order
    const int nact = guided_compact(w.candCnt, n1, qlist, sWarp, tid);
    const int nrounds = (nact + GUIDED_STAGE - 1) / GUIDED_STAGE;

    // stage `r` = heads + list ranges of queries qlist[r*64 ..]; loaded by `nth` threads starting at thread `t0`
    auto loadStage = [&](int r, int t0, int nth) {
        u64* sp = stop + (r & 1) * GUIDED_STAGE * GUIDED_ROW;
        int* sc = scnt + (r & 1) * GUIDED_STAGE;
        int* so = soff + (r & 1) * GUIDED_STAGE;
        if (nth == 256) guided_load_heads<256>(sp, w.top, qlist, r, nact, tid - t0); else guided_load_heads<224>(sp, w.top, qlist, r, nact, tid - t0);
        for (int k = tid - t0; k < GUIDED_STAGE; k += nth) {
            const int q = r * GUIDED_STAGE + k;
            sc[k] = q < nact ? w.candCnt[qlist[q]] : 0;
            so[k] = q < nact ? w.candOff[qlist[q]] : 0;
        }
    };
    if (nrounds > 0) loadStage(0, 0, 256);
    __syncthreads();

    for (int r = 0; r < nrounds; r++) {
        if (warp != 0) {
            if (r + 1 < nrounds) loadStage(r + 1, 32, 224);   // warps 1..7 fetch the next stage behind the sequential warp
        } else {
            const u64* sp = stop + (r & 1) * GUIDED_STAGE * GUIDED_ROW;
            const int* sc = scnt + (r & 1) * GUIDED_STAGE;
            const int kend = min(GUIDED_STAGE, nact - r * GUIDED_STAGE);
            // one query, all 32 lanes on its head (and on its full list when the head runs dry)
            auto seqStep = [&](int k) {
                const u64 e = sp[k * GUIDED_ROW + lane];
                const uint32_t dist = (uint32_t)(e >> 32), i2 = e != ~0ull ? (uint32_t)e & 0xffffu : 0u;   // padding never indexes the state
                const uint32_t di = (dist << 16) | i2;
                const bool ok = e != ~0ull && (uint32_t)md[i2] > dist;   // vMatchedDistance[i2] <= dist -> skipped (:755)
                const unsigned mask = __ballot_sync(FULLMASK, ok);
                const unsigned m2 = mask & (mask - 1);
                int bestDist = 0x7fffffff, bestDist2 = 0x7fffffff, bestIdx = -1;
                if (m2 != 0 || sc[k] <= EORB_GUIDED_TOP) {
                    const uint32_t b1 = __shfl_sync(FULLMASK, di, __ffs(mask) - 1);
                    const uint32_t b2 = __shfl_sync(FULLMASK, di, __ffs(m2) - 1);
                    if (mask) { bestDist = (int)(b1 >> 16); bestIdx = (int)(b1 & 0xffffu); }
                    if (m2) bestDist2 = (int)(b2 >> 16);
                } else {   // the head ran dry: scan the whole list (order-independent thanks to the position in the key)
                    uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu;
                    const int c = sc[k], off = soff[(r & 1) * GUIDED_STAGE + k];
                    for (int p = lane; p < c; p += 32) {
                        const uint32_t ce = w.cand[off + p];
                        const uint32_t d = ce >> 16;
                        if ((uint32_t)md[ce & 0xffffu] > d) {
                            const uint32_t key = (d << 16) | (uint32_t)p;
                            if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const uint32_t a1 = __shfl_xor_sync(FULLMASK, k1, o), a2 = __shfl_xor_sync(FULLMASK, k2, o);
                        const uint32_t lo = min(k1, a1), hi = max(k1, a1);
                        k2 = min(hi, min(k2, a2)); k1 = lo;
                    }
                    if (k1 != 0xffffffffu) { bestDist = (int)(k1 >> 16); bestIdx = (int)(w.cand[off + (k1 & 0xffffu)] & 0xffffu); }
                    if (k2 != 0xffffffffu) bestDist2 = (int)(k2 >> 16);
                }
                if (bestDist <= 50 && (float)bestDist < __fmul_rn((float)bestDist2, nnratio)) {   // TH_LOW, mfNNratio (:770-772)
                    if (lane == 0) {
                        const int i1 = qlist[r * GUIDED_STAGE + k];
                        m21[bestIdx] = (unsigned short)i1;              // takes the keypoint over from an earlier query (:774-780)
                        md[bestIdx] = (unsigned short)bestDist;
                        matches12[i1] = bestIdx;                        // the claim; ownership is settled after the loop
                    }
                    __syncwarp();
                }
            };
#if EORB_GUIDED_SPEC
            // Speculative chunks: lane L evaluates query k + L on its own against the current state (first two unfiltered
            // entries of its sorted head, four entries fetched per step).  Accepting lanes post their lane number on the
            // slot they claim (ctab, atomicMin); a lane is DIRTY when an EARLIER lane of the chunk claims one of its two
            // entries' slots — only then can the order matter (conservative: the claim's distance is not looked at, a dirty
            // lane is simply evaluated again).  The clean prefix is committed at once — it cannot contain two claimers of
            // one slot — and the chunk restarts at the first dirty lane; a query whose head cannot decide (fewer than two
            // unfiltered entries of a longer list) is run by seqStep.  Lane 0 is never dirty: every round makes progress.
            int k = 0;
            while (k < kend) {
                const int q = k + lane;
                const bool have = q < kend;
                uint32_t b1 = 0xffffffffu, b2 = 0xffffffffu;
                bool slow = false;
                if (have) {
                    const u64* hp = sp + q * GUIDED_ROW;
                    int found = 0;
                    for (int en = 0; en < EORB_GUIDED_TOP && found < 2; en += 4) {
                        u64 he[4]; uint32_t mv[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) he[u] = hp[en + u];
#pragma unroll
                        for (int u = 0; u < 4; u++) mv[u] = he[u] != ~0ull ? (uint32_t)md[(uint32_t)he[u] & 0xffffu] : 0u;   // padding: never "unfiltered"
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const uint32_t d = (uint32_t)(he[u] >> 32);
                            if (found < 2 && he[u] != ~0ull && mv[u] > d) {
                                const uint32_t key = (d << 16) | ((uint32_t)he[u] & 0xffffu);
                                if (found == 0) b1 = key; else b2 = key;
                                found++;
                            }
                        }
                        if (he[3] == ~0ull) break;       // sorted: padding only at the end
                    }
                    slow = found < 2 && sc[q] > EORB_GUIDED_TOP;
                }
                const int bd1 = b1 != 0xffffffffu ? (int)(b1 >> 16) : 0x7fffffff, bd2 = b2 != 0xffffffffu ? (int)(b2 >> 16) : 0x7fffffff;
                const bool acc = have && !slow && bd1 <= 50 && (float)bd1 < __fmul_rn((float)bd2, nnratio);
                if (acc) atomicMin(&ctab[b1 & 0xffffu], (unsigned)lane);
                __syncwarp();
                bool dirty = false;
                if (have && b1 != 0xffffffffu && ctab[b1 & 0xffffu] < (unsigned)lane) dirty = true;
                if (have && b2 != 0xffffffffu && ctab[b2 & 0xffffu] < (unsigned)lane) dirty = true;
                const unsigned stopMask = __ballot_sync(FULLMASK, have && (dirty || slow));
                const int ncommit = stopMask ? __ffs(stopMask) - 1 : min(32, kend - k);
                __syncwarp();
                if (acc) {
                    ctab[b1 & 0xffffu] = 0xffffffffu;                     // leave the table clean for the next round
                    if (lane < ncommit) {
                        const int i1 = qlist[r * GUIDED_STAGE + q];
                        matches12[i1] = (int)(b1 & 0xffffu);
                        m21[b1 & 0xffffu] = (unsigned short)i1;
                        md[b1 & 0xffffu] = (unsigned short)(b1 >> 16);
                    }
                }
                __syncwarp();
                k += ncommit;
                if (stopMask && __shfl_sync(FULLMASK, (int)slow, ncommit & 31)) { seqStep(k); k++; }
            }
#else
            for (int k = 0; k < kend; k++) seqStep(k);
#endif
        }
        __syncthreads();
    }
    // claims -> rotation histogram (every claim counts, also the ones taken over later: rotHist keeps them, :793) and owners
    for (int i1 = tid; i1 < n1; i1 += 256) {
        const int f = matches12[i1];
        int bin = -1;
        if (f >= 0) {
            if (checkOri) {
                float rot = __fsub_rn(f1.kps[i1].angle, f2.kps[f].angle);
                if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));          // factor = 1.0f / HISTO_LENGTH (:723)
                if (bin == 30) bin = 0;
                if (bin >= 0 && bin < 30) atomicAdd(&hist[bin], 1); else bin = -1;
            }
            if ((int)m21[f] == i1) atomicAdd(&sNm, 1); else matches12[i1] = -1;
        }
        w.bin[i1] = (signed char)bin;
    }
    __syncthreads();
    if (tid == 0) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        if (checkOri) {   // ComputeThreeMaxima (:2314-2355)
            int max1 = 0, max2 = 0, max3 = 0;
            for (int i = 0; i < 30; i++) {
                const int s = hist[i];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
                else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
                else if (s > max3) { max3 = s; ind3 = i; }
            }
            if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
        }
        sInd[0] = ind1; sInd[1] = ind2; sInd[2] = ind3;
    }
    __syncthreads();
    if (checkOri)
        for (int i1 = tid; i1 < n1; i1 += 256) {
            const int b = w.bin[i1];
            if (b >= 0 && b != sInd[0] && b != sInd[1] && b != sInd[2] && matches12[i1] >= 0) { matches12[i1] = -1; atomicSub(&sNm, 1); }
        }
    __syncthreads();
    for (int i1 = tid; i1 < n1; i1 += 256) {
        const int m = matches12[i1];
        if (m >= 0) { prevXY[2 * i1] = f2.kps[m].x; prevXY[2 * i1 + 1] = f2.kps[m].y; }
    }
    if (tid == 0) *nmatchesOut = sNm;
}

// ------------------------------------------------------------------------------------------------ SearchByProjection
// P1 guided_project_kernel: per last-frame keypoint the window of ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th,
// bMono = true) (ORBmatcher.cc:1995-2032): camera-frame point (formed by the caller) -> invzc < 0 test, Pinhole::project in
// float (Pinhole.cpp:30-33), image-bounds test, radius = th * mvScaleFactors[clamp(octave)], levels [octave-1, octave+1].
__global__ void __launch_bounds__(256) guided_project_kernel(const float* __restrict__ x3Dc, const uint8_t* __restrict__ valid1,
                                                             const eorb_keypoint* __restrict__ kps1, int n1, GuidedProj pr,
                                                             eorb_area_query* __restrict__ qs, GuidedProjMode md) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n1) return;
    eorb_area_query q;
    q.x = 0.f; q.y = 0.f; q.r = -1.f; q.min_level = 0; q.max_level = -1;
    float ur = 0.0f;
    if (valid1[i]) {
        const float xc = x3Dc[3 * i], yc = x3Dc[3 * i + 1], zc = x3Dc[3 * i + 2];
        // frame-to-frame: invzc = 1.0 / zc < 0 (and the undefined z == 0) are skipped, NaN passes like in the reference (:2002-2006);
        // the relocalisation search has no depth-sign test (:2218)
        if (md.reloc || zc > 0.0f || zc != zc) {
            const float u = __fadd_rn(__fdiv_rn(__fmul_rn(pr.fx, xc), zc), pr.cx), v = __fadd_rn(__fdiv_rn(__fmul_rn(pr.fy, yc), zc), pr.cy);
            if (!(u < pr.minX || u > pr.maxX) && !(v < pr.minY || v > pr.maxY)) {
                const int oct = md.level1 ? md.level1[i] : kps1[i].octave;      // nPredictedLevel (:2237) / nLastOctave (:2016)
                const int lv = oct < 0 ? 0 : (oct >= pr.nlevels ? pr.nlevels - 1 : oct);
                q.x = u; q.y = v; q.r = __fmul_rn(pr.th, pr.scale[lv]);
                if (md.levelMode == 1) { q.min_level = oct; q.max_level = -1; }          // bForward: [oct, inf) (:2024-2025)
                else if (md.levelMode == 2) { q.min_level = 0; q.max_level = oct; }      // bBackward: [0, oct] (:2026-2027)
                else { q.min_level = oct - 1; q.max_level = oct + 1; }
                // ur = uv.x - mbf * invzc, invzc = (float)(1.0 / zc): the double quotient rounds to the same float as the float division
                ur = __fsub_rn(u, __fmul_rn(md.mbf, __fdiv_rn(1.0f, zc)));
            }
        }
    }
    qs[i] = q;
    if (md.qUr) md.qUr[i] = ur;
}

// P3 guided_resolve_proj_kernel: the order-dependent part (:2042-2070): a current-frame keypoint whose slot holds a map point
// WITH observations is skipped by every later query, so the best candidate of a query is the first unblocked entry of its
// sorted head.  Same structure as guided_resolve_kernel (one ordered warp, seven staging warps, full-list slow path); a
// query's claim is recorded in claim[i]; the owner of a slot is its LAST claimer; every claim enters the rotation histogram
// (:2073-2089) and a claim in a non-maximal bin un-sets its slot whoever owns it by then (:2141-2150).
// Relocalisation variant (ORBmatcher.cc:2189-2312): obs1 == nullptr (every set slot blocks, :2253), held2 = slots holding a point on
// entry, thHigh = ORBdist (:2266).
__global__ void __launch_bounds__(256) guided_resolve_proj_kernel(const eorb_keypoint* __restrict__ kps1, const int32_t* __restrict__ obs1, int n1,
                                                                  GuidedFrame f2, int checkOri, GuidedWork w, int32_t* __restrict__ claim,
                                                                  int32_t* __restrict__ matchCur, int* __restrict__ nmatchesOut,
                                                                  const uint8_t* __restrict__ held2, int thHigh) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int n2 = f2.n, n2r = (n2 + 3) & ~3;
    u64* stop = reinterpret_cast<u64*>(sm);                               // [2][GUIDED_STAGE][32]
    int* scnt = reinterpret_cast<int*>(stop + 2 * GUIDED_STAGE * GUIDED_ROW);     // [2][GUIDED_STAGE]
    int* soff = scnt + 2 * GUIDED_STAGE;                                  // [2][GUIDED_STAGE]
    int* sobs = soff + 2 * GUIDED_STAGE;                                  // [2][GUIDED_STAGE] observations of the query's map point
    int* hist = sobs + 2 * GUIDED_STAGE;                                  // [32]
    unsigned short* owner = reinterpret_cast<unsigned short*>(hist + 32); // [n2r] last claimer (0xffff = none)
    unsigned short* qlist = owner + n2r;                                  // [n1]
    unsigned* ctab = reinterpret_cast<unsigned*>(qlist + ((n1 + 1) & ~1));             // [n2r] lowest lane of the current chunk claiming a slot
    unsigned char* blk = reinterpret_cast<unsigned char*>(ctab + n2r);                 // [n2r] slot blocked / (later) un-set flag
    __shared__ int sNm, sInd[3], sWarp[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (*w.total > w.candCap) {
        if (tid == 0) *nmatchesOut = -1;
        return;
    }
    for (int i = tid; i < n2; i += 256) { owner[i] = GUIDED_NONE; blk[i] = (held2 && held2[i]) ? 1 : 0; ctab[i] = 0xffffffffu; }
    for (int i = tid; i < n1; i += 256) claim[i] = -1;
    if (tid < 32) hist[tid] = 0;
    if (tid == 0) sNm = 0;
    const int nact = guided_compact(w.candCnt, n1, qlist, sWarp, tid);
    const int nrounds = (nact + GUIDED_STAGE - 1) / GUIDED_STAGE;
    auto loadStage = [&](int r, int t0, int nth) {
        u64* sp = stop + (r & 1) * GUIDED_STAGE * GUIDED_ROW;
        if (nth == 256) guided_load_heads<256>(sp, w.top, qlist, r, nact, tid - t0); else guided_load_heads<224>(sp, w.top, qlist, r, nact, tid - t0);
        for (int k = tid - t0; k < GUIDED_STAGE; k += nth) {
            const int q = r * GUIDED_STAGE + k, o = (r & 1) * GUIDED_STAGE + k;
            scnt[o] = q < nact ? w.candCnt[qlist[q]] : 0;
            soff[o] = q < nact ? w.candOff[qlist[q]] : 0;
            sobs[o] = q < nact ? (obs1 ? obs1[qlist[q]] : 1) : 0;
        }
    };
    if (nrounds > 0) loadStage(0, 0, 256);
    __syncthreads();

    for (int r = 0; r < nrounds; r++) {
        if (warp != 0) {
            if (r + 1 < nrounds) loadStage(r + 1, 32, 224);
        } else {
            const u64* sp = stop + (r & 1) * GUIDED_STAGE * GUIDED_ROW;
            const int so = (r & 1) * GUIDED_STAGE;
            const int kend = min(GUIDED_STAGE, nact - r * GUIDED_STAGE);
            auto seqStep = [&](int k) {
                const u64 e = sp[k * GUIDED_ROW + lane];
                const uint32_t dist = (uint32_t)(e >> 32), i2 = e != ~0ull ? (uint32_t)e & 0xffffu : 0u;
                const bool ok = e != ~0ull && blk[i2] == 0;               // slot held by a point with observations -> skipped (:2046)
                const unsigned mask = __ballot_sync(FULLMASK, ok);
                int bestDist = 256, bestIdx = -1;
                if (mask != 0 || scnt[so + k] <= EORB_GUIDED_TOP) {
                    const uint32_t b1 = __shfl_sync(FULLMASK, (dist << 16) | i2, __ffs(mask) - 1);
                    if (mask) { bestDist = (int)(b1 >> 16); bestIdx = (int)(b1 & 0xffffu); }
                } else {   // every head entry is blocked: scan the whole list
                    uint32_t k1 = 0xffffffffu;
                    const int c = scnt[so + k], off = soff[so + k];
                    for (int p = lane; p < c; p += 32) {
                        const uint32_t ce = w.cand[off + p];
                        if (blk[ce & 0xffffu] == 0) k1 = min(k1, ((ce >> 16) << 16) | (uint32_t)p);
                    }
                    k1 = __reduce_min_sync(FULLMASK, k1);
                    if (k1 != 0xffffffffu) { bestDist = (int)(k1 >> 16); bestIdx = (int)(w.cand[off + (k1 & 0xffffu)] & 0xffffu); }
                }
                if (bestDist <= thHigh) {                                 // TH_HIGH (:2068) / ORBdist (:2266)
                    if (lane == 0) {
                        const int i1 = qlist[r * GUIDED_STAGE + k];
                        owner[bestIdx] = (unsigned short)i1;
                        blk[bestIdx] = sobs[so + k] > 0 ? 1 : 0;
                        claim[i1] = bestIdx;
                    }
                    __syncwarp();
                }
            };
#if EORB_GUIDED_SPEC
            // speculative chunks as in guided_resolve_kernel: lane L takes query k + L (first unblocked entry of its head); it is
            // dirty when an earlier lane of the chunk claims the slot it chose
            int k = 0;
            while (k < kend) {
                const int q = k + lane;
                const bool have = q < kend;
                uint32_t b1 = 0xffffffffu;
                bool slow = false;
                if (have) {
                    const u64* hp = sp + q * GUIDED_ROW;
                    for (int en = 0; en < EORB_GUIDED_TOP && b1 == 0xffffffffu; en += 4) {
                        u64 he[4]; uint32_t bv[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) he[u] = hp[en + u];
#pragma unroll
                        for (int u = 0; u < 4; u++) bv[u] = he[u] != ~0ull ? (uint32_t)blk[(uint32_t)he[u] & 0xffffu] : 1u;
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (b1 == 0xffffffffu && bv[u] == 0) b1 = ((uint32_t)(he[u] >> 32) << 16) | ((uint32_t)he[u] & 0xffffu);
                        if (he[3] == ~0ull) break;
                    }
                    slow = b1 == 0xffffffffu && scnt[so + q] > EORB_GUIDED_TOP;
                }
                const bool acc = have && !slow && b1 != 0xffffffffu && (int)(b1 >> 16) <= thHigh;
                if (acc) atomicMin(&ctab[b1 & 0xffffu], (unsigned)lane);
                __syncwarp();
                const bool dirty = have && b1 != 0xffffffffu && ctab[b1 & 0xffffu] < (unsigned)lane;
                const unsigned stopMask = __ballot_sync(FULLMASK, have && (dirty || slow));
                const int ncommit = stopMask ? __ffs(stopMask) - 1 : min(32, kend - k);
                __syncwarp();
                if (acc) {
                    ctab[b1 & 0xffffu] = 0xffffffffu;
                    if (lane < ncommit) {
                        const int i1 = qlist[r * GUIDED_STAGE + q];
                        claim[i1] = (int)(b1 & 0xffffu);
                        owner[b1 & 0xffffu] = (unsigned short)i1;
                        blk[b1 & 0xffffu] = sobs[so + q] > 0 ? 1 : 0;
                    }
                }
                __syncwarp();
                k += ncommit;
                if (stopMask && __shfl_sync(FULLMASK, (int)slow, ncommit & 31)) { seqStep(k); k++; }
            }
#else
            for (int k = 0; k < kend; k++) seqStep(k);
#endif
        }
        __syncthreads();
    }
    // claims -> nmatches and the rotation histogram; blk[] is reused as the "un-set by the rotation filter" flag
    for (int i = tid; i < n2; i += 256) blk[i] = 0;
    __syncthreads();
    for (int i1 = tid; i1 < n1; i1 += 256) {
        const int f = claim[i1];
        int bin = -1;
        if (f >= 0) {
            atomicAdd(&sNm, 1);
            if (checkOri) {
                float rot = __fsub_rn(kps1[i1].angle, f2.kps[f].angle);
                if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));
                if (bin == 30) bin = 0;
                if (bin >= 0 && bin < 30) atomicAdd(&hist[bin], 1); else bin = -1;
            }
        }
        w.bin[i1] = (signed char)bin;
    }
    __syncthreads();
    if (tid == 0) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        if (checkOri) {
            int max1 = 0, max2 = 0, max3 = 0;
            for (int i = 0; i < 30; i++) {
                const int s = hist[i];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
                else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
                else if (s > max3) { max3 = s; ind3 = i; }
            }
            if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
        }
        sInd[0] = ind1; sInd[1] = ind2; sInd[2] = ind3;
    }
    __syncthreads();
    if (checkOri)
        for (int i1 = tid; i1 < n1; i1 += 256) {
            const int b = w.bin[i1];
            if (b >= 0 && b != sInd[0] && b != sInd[1] && b != sInd[2]) { blk[claim[i1]] = 1; atomicSub(&sNm, 1); }   // one nmatches-- per entry
        }
    __syncthreads();
    for (int i = tid; i < n2; i += 256) matchCur[i] = (owner[i] != GUIDED_NONE && blk[i] == 0) ? (int)owner[i] : -1;
    if (tid == 0) *nmatchesOut = sNm;
}

// ------------------------------------------------------------------------------------------------ keyframe-side window searches
// W1 guided_best_kernel: the non-blocking searches of local mapping / loop closing pick, per map point on its own, the candidate with
// the smallest distance, the first one visited among equals (strict `dist < bestDist`: ORBmatcher::Fuse :1520-1571 and :1706-1724,
// SearchBySim3 :1838-1858 and :1918-1938) -- entry 0 of the sorted head.  bestDist[i] = that distance (256: no candidate),
// bestIdx[i] = its keypoint when bestDist <= thHigh, else -1.
__global__ void __launch_bounds__(256) guided_best_kernel(int n1, GuidedWork w, int thHigh, int32_t* __restrict__ bestIdx,
                                                          int32_t* __restrict__ bestDist, int* __restrict__ nmatchesOut) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const bool over = *w.total > w.candCap;
    int acc = 0;
    if (i < n1 && !over) {
        int d = 256, idx = -1;
        if (w.candCnt[i] > 0) {
            const u64 e = w.top[(size_t)i * EORB_GUIDED_TOP];
            d = (int)(e >> 32);
            if (d <= thHigh) { idx = (int)(e & 0xffffu); acc = 1; }
        }
        bestIdx[i] = idx;
        if (bestDist) bestDist[i] = d;
    }
    const int nb = __syncthreads_count(acc);
    if (threadIdx.x == 0 && nb) atomicAdd(nmatchesOut, nb);     // zeroed by the launcher; the candidate total sits next to it (w.total)
}

// W2 guided_claim_dist_kernel: distance of every claim of a blocking search (the reference's bestDist of an accepted point)
__global__ void __launch_bounds__(256) guided_claim_dist_kernel(const uint8_t* __restrict__ descMP, const uint8_t* __restrict__ desc2, int n1,
                                                                const int32_t* __restrict__ claim, int32_t* __restrict__ bestIdx,
                                                                int32_t* __restrict__ bestDist) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n1) return;
    const int c = claim[i];
    int d = 256;
    if (c >= 0) {
        const uint4* a = reinterpret_cast<const uint4*>(descMP + (size_t)i * 32);
        const uint4* b = reinterpret_cast<const uint4*>(desc2 + (size_t)c * 32);
        const uint4 a0 = __ldg(a), a1 = __ldg(a + 1), b0 = __ldg(b), b1 = __ldg(b + 1);
        d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) +
            __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
    }
    bestIdx[i] = c;
    if (bestDist) bestDist[i] = d;
}

// ------------------------------------------------------------------------------------------------ SearchByProjection (map points)
// L1 guided_mappoint_query_kernel: the window of ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints)
// (ORBmatcher.cc:54-75) per local map point: in-view / far / bad gates, RadiusByViewingCos (:219-225: the float cosine is compared
// with the DOUBLE constant 0.998), r *= th when th != 1.0, half size r * mvScaleFactors[clamp(level)], levels [level-1, level].
__global__ void __launch_bounds__(256) guided_mappoint_query_kernel(const eorb_track_point* __restrict__ pts, int n1, GuidedProj pr, int farPoints,
                                                                    float thFar, eorb_area_query* __restrict__ qs) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n1) return;
    eorb_area_query q;
    q.x = 0.f; q.y = 0.f; q.r = -1.f; q.min_level = 0; q.max_level = -1;
    const eorb_track_point p = pts[i];
    if (p.in_view && !(farPoints && p.depth > thFar) && !p.bad) {
        float r = ((double)p.view_cos > 0.998) ? 2.5f : 4.0f;
        if (pr.th != 1.0f) r = __fmul_rn(r, pr.th);
        const int lv = p.scale_level < 0 ? 0 : (p.scale_level >= pr.nlevels ? pr.nlevels - 1 : p.scale_level);
        q.x = p.proj_x; q.y = p.proj_y; q.r = __fmul_rn(r, pr.scale[lv]); q.min_level = p.scale_level - 1; q.max_level = p.scale_level;
    }
    qs[i] = q;
}

// L3 guided_resolve_map_kernel: the order-dependent part (:87-137).  A slot of F that holds a map point WITH observations (on
// entry: held2; later: a claim by such a point) is skipped by every later map point, so a point's best and second-best
// candidates are the first TWO unblocked entries of its (distance, visiting position)-sorted head: the strict "<" updates of
// :104-124 leave exactly that pair whatever the visiting order (a demoted best overwrites the second, a tie never replaces).
// Accepted when best <= TH_HIGH and not (same level and best > ratio * second) (:128-131); the owner of a slot is its LAST
// claimer (setMapPoint overwrites a point without observations), every claim counts in nmatches.  Same structure as
// guided_resolve_proj_kernel: one ordered warp with speculative chunks of 32 points, seven warps staging the next heads.
__global__ void __launch_bounds__(256) guided_resolve_map_kernel(const eorb_track_point* __restrict__ pts, int n1, GuidedFrame f2,
                                                                 const uint8_t* __restrict__ held2, float nnratio, GuidedWork w,
                                                                 int32_t* __restrict__ matchCur, int* __restrict__ nmatchesOut) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int n2 = f2.n, n2r = (n2 + 3) & ~3;
    u64* stop = reinterpret_cast<u64*>(sm);                               // [2][GUIDED_STAGE][32]
    int* scnt = reinterpret_cast<int*>(stop + 2 * GUIDED_STAGE * GUIDED_ROW);     // [2][GUIDED_STAGE]
    int* soff = scnt + 2 * GUIDED_STAGE;                                  // [2][GUIDED_STAGE]
    int* sobs = soff + 2 * GUIDED_STAGE;                                  // [2][GUIDED_STAGE] observations of the query's map point
    unsigned short* owner = reinterpret_cast<unsigned short*>(sobs + 2 * GUIDED_STAGE);   // [n2r] last claimer (0xffff = none)
    unsigned short* qlist = owner + n2r;                                  // [n1]
    unsigned* ctab = reinterpret_cast<unsigned*>(qlist + ((n1 + 1) & ~1));             // [n2r] lowest lane of the current chunk claiming a slot
    unsigned char* blk = reinterpret_cast<unsigned char*>(ctab + n2r);                 // [n2r] slot holds a point with observations
    signed char* lvl = reinterpret_cast<signed char*>(blk + n2r);                      // [n2r] octave of the frame's keypoints
    __shared__ int sNm, sWarp[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (*w.total > w.candCap) {
        if (tid == 0) *nmatchesOut = -1;
        return;
    }
    for (int i = tid; i < n2; i += 256) {
        owner[i] = GUIDED_NONE; ctab[i] = 0xffffffffu;
        blk[i] = held2 ? (held2[i] != 0) : 0;
        const int o = f2.kps[i].octave;
        lvl[i] = (signed char)(o < -100 ? -100 : (o > 100 ? 100 : o));
    }
    if (tid == 0) sNm = 0;
    const int nact = guided_compact(w.candCnt, n1, qlist, sWarp, tid);
    const int nrounds = (nact + GUIDED_STAGE - 1) / GUIDED_STAGE;
    auto loadStage = [&](int r, int t0, int nth) {
        u64* sp = stop + (r & 1) * GUIDED_STAGE * GUIDED_ROW;
        if (nth == 256) guided_load_heads<256>(sp, w.top, qlist, r, nact, tid - t0); else guided_load_heads<224>(sp, w.top, qlist, r, nact, tid - t0);
        for (int k = tid - t0; k < GUIDED_STAGE; k += nth) {
            const int q = r * GUIDED_STAGE + k, o = (r & 1) * GUIDED_STAGE + k;
            scnt[o] = q < nact ? w.candCnt[qlist[q]] : 0;
            soff[o] = q < nact ? w.candOff[qlist[q]] : 0;
            sobs[o] = q < nact ? pts[qlist[q]].observations : 0;
        }
    };
    if (nrounds > 0) loadStage(0, 0, 256);
    __syncthreads();

    // (:128-131) with b = dist << 16 | slot, 0xffffffff = none
    auto accepted = [&](uint32_t b1, uint32_t b2) -> bool {
        if (b1 == 0xffffffffu) return false;
        const int d1 = (int)(b1 >> 16);
        if (d1 > 100) return false;                                       // TH_HIGH
        const int l1 = lvl[b1 & 0xffffu];
        const int d2 = b2 != 0xffffffffu ? (int)(b2 >> 16) : 256, l2 = b2 != 0xffffffffu ? (int)lvl[b2 & 0xffffu] : -1;
        return !(l1 == l2 && (float)d1 > __fmul_rn(nnratio, (float)d2));
    };

    for (int r = 0; r < nrounds; r++) {
        if (warp != 0) {
            if (r + 1 < nrounds) loadStage(r + 1, 32, 224);
        } else {
            const u64* sp = stop + (r & 1) * GUIDED_STAGE * GUIDED_ROW;
            const int so = (r & 1) * GUIDED_STAGE;
            const int kend = min(GUIDED_STAGE, nact - r * GUIDED_STAGE);
            // whole warp on one query: head by ballot, or the full list when the head holds fewer than two unblocked entries
            auto seqStep = [&](int k) {
                const u64 e = sp[k * GUIDED_ROW + lane];
                const uint32_t dist = (uint32_t)(e >> 32), i2 = e != ~0ull ? (uint32_t)e & 0xffffu : 0u;
                const bool ok = e != ~0ull && blk[i2] == 0;
                unsigned mask = __ballot_sync(FULLMASK, ok);
                uint32_t b1 = 0xffffffffu, b2 = 0xffffffffu;
                if (__popc(mask) >= 2 || scnt[so + k] <= EORB_GUIDED_TOP) {
                    const uint32_t v = (dist << 16) | i2;
                    if (mask) { b1 = __shfl_sync(FULLMASK, v, __ffs(mask) - 1); mask &= mask - 1; }
                    if (mask) b2 = __shfl_sync(FULLMASK, v, __ffs(mask) - 1);
                } else {
                    uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu;          // two smallest (dist << 16 | position) keys of this lane
                    const int c = scnt[so + k], off = soff[so + k];
                    for (int p = lane; p < c; p += 32) {
                        const uint32_t ce = w.cand[off + p];
                        if (blk[ce & 0xffffu] == 0) {
                            const uint32_t key = ((ce >> 16) << 16) | (uint32_t)p;
                            if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
                        }
                    }
                    const uint32_t m1 = __reduce_min_sync(FULLMASK, k1);
                    const uint32_t m2 = __reduce_min_sync(FULLMASK, k1 == m1 ? k2 : k1);
                    if (m1 != 0xffffffffu) b1 = (m1 & 0xffff0000u) | (w.cand[off + (m1 & 0xffffu)] & 0xffffu);
                    if (m2 != 0xffffffffu) b2 = (m2 & 0xffff0000u) | (w.cand[off + (m2 & 0xffffu)] & 0xffffu);
                }
                if (accepted(b1, b2)) {
                    if (lane == 0) {
                        const int i1 = qlist[r * GUIDED_STAGE + k];
                        owner[b1 & 0xffffu] = (unsigned short)i1;
                        blk[b1 & 0xffffu] = sobs[so + k] > 0 ? 1 : 0;
                        sNm++;
                    }
                }
                __syncwarp();
            };
#if EORB_GUIDED_SPEC
#if EORB_GUIDED_FIXPOINT
            // lane L takes query k + L.  Its answer depends on the lower lanes only through "is this slot claimed by a lower lane
            // whose point has observations", so the chunk is solved as a fixed point: every lane picks the first two head entries
            // that are neither blocked nor claimed (table ctab = lowest blocking claimer of a slot, rebuilt from the previous
            // pass), until a pass changes nothing.  Lane j is final after pass j + 1 by induction, the fixed point is the
            // sequential answer, and the number of passes is the longest chain of contested slots in the chunk (2-4) instead of
            // one restart per contested lane.  A lane whose head runs dry stops the chunk: the lanes below it commit, it takes
            // the full-list step.
            int k = 0;
            while (k < kend) {
                const int q = k + lane;
                const bool have = q < kend;
                const bool blocking = have && sobs[so + q] > 0;
                uint32_t b1 = 0xffffffffu, b2 = 0xffffffffu, p1 = 0xfffffffeu, p2 = 0xfffffffeu;
                bool slow = false, acc = false, pacc = false, pslow = false;
                int pass = 0;
                for (;; pass++) {
                    b1 = 0xffffffffu; b2 = 0xffffffffu; slow = false;
                    if (have) {
                        const u64* hp = sp + q * GUIDED_ROW;
                        for (int en = 0; en < EORB_GUIDED_TOP && b2 == 0xffffffffu; en += 4) {
                            u64 he[4]; uint32_t bv[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) he[u] = hp[en + u];
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const uint32_t sl = (uint32_t)he[u] & 0xffffu;
                                bv[u] = he[u] != ~0ull ? ((uint32_t)blk[sl] | (uint32_t)(ctab[sl] < (unsigned)lane)) : 1u;
                            }
#pragma unroll
                            for (int u = 0; u < 4; u++)
                                if (bv[u] == 0 && b2 == 0xffffffffu) {
                                    const uint32_t v = ((uint32_t)(he[u] >> 32) << 16) | ((uint32_t)he[u] & 0xffffu);
                                    if (b1 == 0xffffffffu) b1 = v; else b2 = v;
                                }
                            if (he[3] == ~0ull) break;
                        }
                        slow = b2 == 0xffffffffu && scnt[so + q] > EORB_GUIDED_TOP && (b1 == 0xffffffffu || (b1 >> 16) <= 100u);
                    }
                    acc = have && !slow && accepted(b1, b2);
                    const bool changed = have && (b1 != p1 || b2 != p2 || acc != pacc || slow != pslow);
                    if (!__any_sync(FULLMASK, changed) || pass >= 40) break;
                    if (pacc && blocking) ctab[p1 & 0xffffu] = 0xffffffffu;          // rebuild the claim table from this pass
                    __syncwarp();
                    if (acc && blocking) atomicMin(&ctab[b1 & 0xffffu], (unsigned)lane);
                    __syncwarp();
                    p1 = b1; p2 = b2; pacc = acc; pslow = slow;
                }
                // here (b1, b2, acc, slow) == the previous pass and ctab holds exactly its blocking claims
                const unsigned stopMask = pass >= 40 ? 0xfffffffeu : __ballot_sync(FULLMASK, have && slow);
                const int ncommit = stopMask ? __ffs(stopMask) - 1 : min(32, kend - k);
                if (pacc && blocking) ctab[p1 & 0xffffu] = 0xffffffffu;
                __syncwarp();
                const bool commit = acc && lane < ncommit;
                // several points without observations may claim one slot in a chunk (the last one owns it); at most one blocking
                // claim per slot, and it is the last
                if (commit) atomicMin(&ctab[b1 & 0xffffu], (unsigned)(31 - lane));
                __syncwarp();
                if (commit && ctab[b1 & 0xffffu] == (unsigned)(31 - lane)) owner[b1 & 0xffffu] = qlist[r * GUIDED_STAGE + q];
                if (commit && blocking) blk[b1 & 0xffffu] = 1;
                __syncwarp();
                if (commit) ctab[b1 & 0xffffu] = 0xffffffffu;
                const int nc = __popc(__ballot_sync(FULLMASK, commit));
                if (lane == 0) sNm += nc;
                __syncwarp();
                k += ncommit;
                if (stopMask && pass < 40 && __shfl_sync(FULLMASK, (int)slow, ncommit & 31)) { seqStep(k); k++; }
            }
#else
            // lane L takes query k + L: the first two unblocked entries of its head.  It is dirty when an earlier lane of the
            // chunk claims either of them (conservative for claims by points without observations); the clean prefix commits.
            int k = 0;
            while (k < kend) {
                const int q = k + lane;
                const bool have = q < kend;
                uint32_t b1 = 0xffffffffu, b2 = 0xffffffffu;
                bool slow = false;
                if (have) {
                    const u64* hp = sp + q * GUIDED_ROW;
                    for (int en = 0; en < EORB_GUIDED_TOP && b2 == 0xffffffffu; en += 4) {
                        u64 he[4]; uint32_t bv[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) he[u] = hp[en + u];
#pragma unroll
                        for (int u = 0; u < 4; u++) bv[u] = he[u] != ~0ull ? (uint32_t)blk[(uint32_t)he[u] & 0xffffu] : 1u;
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (bv[u] == 0 && b2 == 0xffffffffu) {
                                const uint32_t v = ((uint32_t)(he[u] >> 32) << 16) | ((uint32_t)he[u] & 0xffffu);
                                if (b1 == 0xffffffffu) b1 = v; else b2 = v;
                            }
                        if (he[3] == ~0ull) break;
                    }
                    slow = b2 == 0xffffffffu && scnt[so + q] > EORB_GUIDED_TOP && (b1 == 0xffffffffu || (b1 >> 16) <= 100u);
                }
                const bool acc = have && !slow && accepted(b1, b2);
                if (acc) atomicMin(&ctab[b1 & 0xffffu], (unsigned)lane);
                __syncwarp();
                const bool dirty = have && ((b1 != 0xffffffffu && ctab[b1 & 0xffffu] < (unsigned)lane) ||
                                            (b2 != 0xffffffffu && ctab[b2 & 0xffffu] < (unsigned)lane));
                const unsigned stopMask = __ballot_sync(FULLMASK, have && (dirty || slow));
                const int ncommit = stopMask ? __ffs(stopMask) - 1 : min(32, kend - k);
                __syncwarp();
                const bool commit = acc && lane < ncommit;
                if (acc) ctab[b1 & 0xffffu] = 0xffffffffu;
                if (commit) {
                    owner[b1 & 0xffffu] = qlist[r * GUIDED_STAGE + q];
                    blk[b1 & 0xffffu] = sobs[so + q] > 0 ? 1 : 0;
                }
                const int nc = __popc(__ballot_sync(FULLMASK, commit));
                if (lane == 0) sNm += nc;
                __syncwarp();
                k += ncommit;
                if (stopMask && __shfl_sync(FULLMASK, (int)slow, ncommit & 31)) { seqStep(k); k++; }
            }
#endif
#else
            for (int k = 0; k < kend; k++) seqStep(k);
#endif
        }
        __syncthreads();
    }
    for (int i = tid; i < n2; i += 256) matchCur[i] = owner[i] != GUIDED_NONE ? (int)owner[i] : -1;
    if (tid == 0) *nmatchesOut = sNm;
}

// Register path of SearchByBoW for one vocabulary node (one warp): the node's frame features sit in registers, S per lane
// (feature p of the node's list <-> slot p / 32 of lane p % 32: index, descriptor, free flag); the keyframe side is loaded 32
// features at a time, one per lane, and broadcast by shuffles, so the ordered loop over the keyframe features touches no memory.
// S = 1 covers the usual ~10 features per node; S = 2 / 4 / 8 cover crowded nodes up to 256 features (a coarse vocabulary level or a
// repetitive scene), which used to fall to the memory-resident loop below (509 us for 10 nodes x 100 features against 19.5 us).
// Ties: the key is (distance << 16) | list position, so the smallest position wins like the reference's strict "<" scan
// (ORBmatcher.cc:318-378).  Returns the number of matches made (the caller adds it to work[32]).
template <int S>
__device__ __forceinline__ int bow_node_registers(const GuidedBowSide& kf, const uint8_t* __restrict__ validKF, const GuidedBowSide& f, int w,
                                                  int fb, int fe, float nnratio, int checkOri, int32_t* matchF, int* work, int lane,
                                                  const uint8_t* __restrict__ validF, int thLow) {
    auto rotBin = [&](int ik, int jf) -> int {
        float rot = __fsub_rn(kf.kps[ik].angle, f.kps[jf].angle);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));
        if (bin == 30) bin = 0;
        return (bin >= 0 && bin < 30) ? bin : -1;
    };
    int jf[S], mine[S];
    uint32_t fd[S][8];
    unsigned freeMask = 0;
#pragma unroll
    for (int s = 0; s < S; s++) {
        const int p = fb + s * 32 + lane;
        const bool has = p < fe;
        jf[s] = has ? (int)f.feats[p] : 0;
        mine[s] = -1;
#pragma unroll
        for (int j = 0; j < 8; j++) fd[s][j] = 0;
        if (has) {
            const uint4* dp = reinterpret_cast<const uint4*>(f.desc + (size_t)jf[s] * 32);
            const uint4 x = __ldg(dp), y = __ldg(dp + 1);
            fd[s][0] = x.x; fd[s][1] = x.y; fd[s][2] = x.z; fd[s][3] = x.w; fd[s][4] = y.x; fd[s][5] = y.y; fd[s][6] = y.z; fd[s][7] = y.w;
            if (!validF || validF[jf[s]]) freeMask |= 1u << s;   // keyframe-keyframe form: a candidate needs a good map point too
        }
    }
    int nm = 0;
    for (int a0 = kf.start[w], ke = kf.start[w + 1]; a0 < ke; a0 += 32) {
        const int cnt = min(32, ke - a0);
        int ikL = -1;
        uint32_t kd[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (lane < cnt) {
            const int t = (int)kf.feats[a0 + lane];
            if (validKF[t]) {
                ikL = t;
                const uint4* qp = reinterpret_cast<const uint4*>(kf.desc + (size_t)t * 32);
                const uint4 x = __ldg(qp), y = __ldg(qp + 1);
                kd[0] = x.x; kd[1] = x.y; kd[2] = x.z; kd[3] = x.w; kd[4] = y.x; kd[5] = y.y; kd[6] = y.z; kd[7] = y.w;
            }
        }
        unsigned todo = __ballot_sync(FULLMASK, ikL >= 0);
        while (todo) {
            const int sl = __ffs(todo) - 1;
            todo &= todo - 1;
            uint32_t q[8];
#pragma unroll
            for (int j = 0; j < 8; j++) q[j] = __shfl_sync(FULLMASK, kd[j], sl);
            uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu;     // this lane's best and second-best key over its free slots
#pragma unroll
            for (int s = 0; s < S; s++) {
                int dist = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) dist += __popc(q[j] ^ fd[s][j]);
                const uint32_t key = (freeMask >> s) & 1u ? (((uint32_t)dist << 16) | (uint32_t)(s * 32 + lane)) : 0xffffffffu;
                if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
            }
            const uint32_t m1 = __reduce_min_sync(FULLMASK, k1);
            if (m1 == 0xffffffffu) break;                      // every frame feature of the node is matched
            const uint32_t m2 = __reduce_min_sync(FULLMASK, k1 == m1 ? k2 : k1);
            const int d1 = (int)(m1 >> 16), d2 = m2 != 0xffffffffu ? (int)(m2 >> 16) : 256;
            if (d1 <= thLow && (float)d1 < __fmul_rn(nnratio, (float)d2)) {
                const int ik = __shfl_sync(FULLMASK, ikL, sl);
                const int pos = (int)(m1 & 0xffffu);
                if (lane == (pos & 31)) {
#pragma unroll
                    for (int s = 0; s < S; s++)
                        if (s == (pos >> 5)) { freeMask &= ~(1u << s); mine[s] = ik; }
                }
                nm++;
            }
        }
    }
#pragma unroll
    for (int s = 0; s < S; s++)
        if (mine[s] >= 0) {
            matchF[jf[s]] = mine[s];
            if (checkOri) { const int bin = rotBin(mine[s], jf[s]); if (bin >= 0) atomicAdd(&work[bin], 1); }
        }
    return nm;
}

// ------------------------------------------------------------------------------------------------ SearchByBoW
// ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) (ORBmatcher.cc:276-478), monocular.  A frame feature belongs to exactly one
// vocabulary node, so the "already matched" state (:330-331) never crosses nodes: the nodes are independent, only the keyframe
// features INSIDE a node are order-dependent.  One warp per common node (binary search of the keyframe's node in the frame's
// ascending node list = the lower_bound walk of :296-447), its keyframe features in list order, the lanes striding over the
// node's frame features: two smallest (distance << 16 | position) keys per lane, two warp min-reductions -> best (first of the
// smallest, like the strict "<") and second-best distance.  TH_LOW + float ratio test (:370-372), then rotation histogram
// (:382-406) in global memory; the block that finishes last (ticket counter) runs ComputeThreeMaxima and the filter (:449-470).
// Four nodes per block: a frame has ~100 nodes of ~10 features, so the ordered lists are short and the nodes spread over the SMs.
#define BOW_WARPS 4     // warps (= vocabulary nodes) per block
// work[0..29] rotation histogram, work[32] nmatches, work[33] blocks finished (zeroed before the launch); matchF preset to -1
__global__ void __launch_bounds__(BOW_WARPS * 32) search_by_bow_kernel(GuidedBowSide kf, const uint8_t* __restrict__ validKF, GuidedBowSide f,
                                                                        float nnratio, int checkOri, int32_t* matchF, int* work,
                                                                        int* __restrict__ nmatchesOut, const uint8_t* __restrict__ validF,
                                                                        int thLow, int32_t* __restrict__ match12) {
    __shared__ int sInd[3], sLast;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    volatile int32_t* taken = matchF;
    auto rotBin = [&](int ik, int jf) -> int {
        float rot = __fsub_rn(kf.kps[ik].angle, f.kps[jf].angle);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));
        if (bin == 30) bin = 0;
        return (bin >= 0 && bin < 30) ? bin : -1;
    };
    const int w = blockIdx.x * BOW_WARPS + warp;
    if (w < kf.nnodes) {
        const uint32_t node = kf.nodes[w];
        int lo = 0, hi = f.nnodes;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (f.nodes[mid] < node) lo = mid + 1; else hi = mid; }
        if (lo < f.nnodes && f.nodes[lo] == node) {
            const int fb = f.start[lo], fe = f.start[lo + 1];
            int nm = 0;
            if (fe - fb <= 32) nm = bow_node_registers<1>(kf, validKF, f, w, fb, fe, nnratio, checkOri, matchF, work, lane, validF, thLow);
            else if (fe - fb <= 64) nm = bow_node_registers<2>(kf, validKF, f, w, fb, fe, nnratio, checkOri, matchF, work, lane, validF, thLow);
            else if (fe - fb <= 128) nm = bow_node_registers<4>(kf, validKF, f, w, fb, fe, nnratio, checkOri, matchF, work, lane, validF, thLow);
            else if (fe - fb <= 256) nm = bow_node_registers<8>(kf, validKF, f, w, fb, fe, nnratio, checkOri, matchF, work, lane, validF, thLow);
            else
            for (int a = kf.start[w], ke = kf.start[w + 1]; a < ke; a++) {   // nodes above 256 frame features: descriptors stay in memory
                const int ik = (int)kf.feats[a];
                if (!validKF[ik]) continue;
                uint32_t qd[8];
                {
                    const uint4* qp = reinterpret_cast<const uint4*>(kf.desc + (size_t)ik * 32);
                    const uint4 x = __ldg(qp), y = __ldg(qp + 1);
                    qd[0] = x.x; qd[1] = x.y; qd[2] = x.z; qd[3] = x.w; qd[4] = y.x; qd[5] = y.y; qd[6] = y.z; qd[7] = y.w;
                }
                uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu;
                for (int p = fb + lane; p < fe; p += 32) {
                    const int jf = (int)f.feats[p];
                    if (taken[jf] >= 0 || (validF && !validF[jf])) continue;
                    const uint4* dp = reinterpret_cast<const uint4*>(f.desc + (size_t)jf * 32);
                    const uint4 x = __ldg(dp), y = __ldg(dp + 1);
                    const int dist = __popc(x.x ^ qd[0]) + __popc(x.y ^ qd[1]) + __popc(x.z ^ qd[2]) + __popc(x.w ^ qd[3]) + __popc(y.x ^ qd[4]) +
                                     __popc(y.y ^ qd[5]) + __popc(y.z ^ qd[6]) + __popc(y.w ^ qd[7]);
                    const uint32_t key = ((uint32_t)dist << 16) | (uint32_t)(p - fb);
                    if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
                }
                const uint32_t m1 = __reduce_min_sync(FULLMASK, k1);
                const uint32_t m2 = __reduce_min_sync(FULLMASK, k1 == m1 ? k2 : k1);
                if (m1 == 0xffffffffu) continue;
                const int d1 = (int)(m1 >> 16), d2 = m2 != 0xffffffffu ? (int)(m2 >> 16) : 256;
                if (d1 <= thLow && (float)d1 < __fmul_rn(nnratio, (float)d2)) {    // TH_LOW (<= 50, or < 50 = <= 49 in the keyframe-keyframe form), mfNNratio
                    if (lane == 0) {
                        const int jf = (int)f.feats[fb + (int)(m1 & 0xffffu)];
                        taken[jf] = ik;
                        nm++;
                        if (checkOri) { const int bin = rotBin(ik, jf); if (bin >= 0) atomicAdd(&work[bin], 1); }
                    }
                    __syncwarp();
                }
            }
            if (lane == 0 && nm > 0) atomicAdd(&work[32], nm);
        }
    }
    // the block that finishes last filters the matches (:449-470)
    __threadfence();
    __syncthreads();
    if (tid == 0) sLast = (atomicAdd(&work[33], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!sLast) return;
    __threadfence();
    if (tid == 0) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        if (checkOri) {
            int max1 = 0, max2 = 0, max3 = 0;
            for (int i = 0; i < 30; i++) {
                const int s = __ldcg(&work[i]);
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
                else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
                else if (s > max3) { max3 = s; ind3 = i; }
            }
            if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
        }
        sInd[0] = ind1; sInd[1] = ind2; sInd[2] = ind3;
    }
    __syncthreads();
    if (checkOri)
        for (int i = tid; i < f.n; i += BOW_WARPS * 32) {
            const int ik = __ldcg(&matchF[i]);
            if (ik < 0) continue;
            const int b = rotBin(ik, i);
            if (b >= 0 && b != sInd[0] && b != sInd[1] && b != sInd[2]) { matchF[i] = -1; atomicSub(&work[32], 1); }
        }
    __syncthreads();
    if (match12)   // keyframe-keyframe form: the result is indexed by the first keyframe's feature (vpMatches12[idx1], ORBmatcher.cc:916)
        for (int i = tid; i < f.n; i += BOW_WARPS * 32) {
            const int ik = __ldcg(&matchF[i]);
            if (ik >= 0) match12[ik] = i;
        }
    if (tid == 0) *nmatchesOut = atomicAdd(&work[32], 0);
}

// ------------------------------------------------------------------------------------------------ launches
static size_t resolveSmem(int n1, int n2) {
    const size_t n2r = (size_t)((n2 + 3) & ~3);
    return (size_t)2 * GUIDED_STAGE * GUIDED_ROW * 8 + 2 * GUIDED_STAGE * 4 * 2 + 32 * 4 + n2r * 2 * 2 + (size_t)((n1 + 1) & ~1) * 2 + n2r * 4 + 16;
}

static size_t gridSmem(int n) { return (size_t)(2 * EORB_GRID_CELLS + 1) * 4 + 2 * (size_t)((n + 1) & ~1) * 2 + 16; }

static size_t resolveProjSmem(int n1, int n2) {
    const size_t n2r = (size_t)((n2 + 3) & ~3);
    return (size_t)2 * GUIDED_STAGE * GUIDED_ROW * 8 + 2 * GUIDED_STAGE * 4 * 3 + 32 * 4 + n2r * 2 + (size_t)((n1 + 1) & ~1) * 2 + n2r * 4 + n2r + 16;
}

static size_t resolveMapSmem(int n1, int n2) {
    const size_t n2r = (size_t)((n2 + 3) & ~3);
    return (size_t)2 * GUIDED_STAGE * GUIDED_ROW * 8 + 2 * GUIDED_STAGE * 4 * 3 + n2r * 2 + (size_t)((n1 + 1) & ~1) * 2 + n2r * 4 + n2r * 2 + 16;
}

// ------------------------------------------------------------------------------------------------ SearchForTriangulation
// ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo, bCoarse) (src/ORBmatcher.cc:975-1214; LocalMapping::
// CreateNewMapPoints, Tracking's keyframe insertion), keyframes with ONE pinhole camera each (mpCamera2 == NULL).  In this version of the
// function vbMatched2 is never set (:1008 declares it, nothing writes it), so the features of KF1 do not interact: one warp per KF1 feature,
// the lanes stride over the KF2 features of the same vocabulary node.  The running rule `dist > TH_LOW ||
// dist > bestDist -> skip, ... -> bestIdx2 = idx2, bestDist = dist` (:1066-1140) keeps the LAST feature among those with the smallest
// distance that pass the static gates, i.e. the minimum of (dist << 16 | 0xffff - position).  Gates: flag bit 0 (no map point, and stereo
// when bOnlyStereo), the epipole distance for monocular pairs (:1077-1085), Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:
// 134-157) with the fundamental matrix the camera forms from R12 / t12 passed in by the caller -- float arithmetic in the reference's order.
__global__ void __launch_bounds__(BOW_WARPS * 32) search_triangulation_kernel(GuidedBowSide k1, const uint8_t* __restrict__ flags1, GuidedBowSide k2,
                                                                               const uint8_t* __restrict__ flags2, GuidedTriGeom tg, int checkOri,
                                                                               int nEntries1, int32_t* __restrict__ match12,
                                                                               signed char* __restrict__ binOf, int* __restrict__ work) {
    // one warp per FeatureVector entry of the first keyframe (= one feature; a feature lies under exactly one node), the lanes stride over the
    // second keyframe's features of the same node: the work is a few hundred 8-POPC distances, so the critical path is what matters
    const int lane = threadIdx.x & 31, a = blockIdx.x * BOW_WARPS + (threadIdx.x >> 5);
    if (a >= nEntries1) return;
    const int i1 = (int)k1.feats[a];
    const unsigned f1 = flags1[i1];
    if (!(f1 & 1u)) return;
    int lo = 0, hi = k1.nnodes;                        // node of entry a: the last q with start[q] <= a
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (k1.start[mid] <= a) lo = mid; else hi = mid; }
    const uint32_t node = k1.nodes[lo];
    lo = 0; hi = k2.nnodes;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (k2.nodes[mid] < node) lo = mid + 1; else hi = mid; }
    if (lo >= k2.nnodes || k2.nodes[lo] != node) return;
    const int fb = k2.start[lo], fe = k2.start[lo + 1];
    const eorb_keypoint kp1 = k1.kps[i1];
    uint32_t q[8];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(k1.desc + (size_t)i1 * 32);
        const uint4 u = __ldg(qp), v = __ldg(qp + 1);
        q[0] = u.x; q[1] = u.y; q[2] = u.z; q[3] = u.w; q[4] = v.x; q[5] = v.y; q[6] = v.z; q[7] = v.w;
    }
    // epipolar line in the second image l = x1' F12 = [a b c] (Pinhole.cpp:143-145)
    const float la = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, tg.F[0]), __fmul_rn(kp1.y, tg.F[3])), tg.F[6]);
    const float lb = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, tg.F[1]), __fmul_rn(kp1.y, tg.F[4])), tg.F[7]);
    const float lc = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, tg.F[2]), __fmul_rn(kp1.y, tg.F[5])), tg.F[8]);
    const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
    uint32_t best = 0xffffffffu;
    for (int j = fb + lane; j < fe; j += 32) {
        const int i2 = (int)k2.feats[j];
        const unsigned f2 = flags2[i2];
        if (!(f2 & 1u)) continue;
        const uint4* dp = reinterpret_cast<const uint4*>(k2.desc + (size_t)i2 * 32);
        const uint4 u = __ldg(dp), v = __ldg(dp + 1);
        const int dist = __popc(u.x ^ q[0]) + __popc(u.y ^ q[1]) + __popc(u.z ^ q[2]) + __popc(u.w ^ q[3]) + __popc(v.x ^ q[4]) + __popc(v.y ^ q[5]) +
                         __popc(v.z ^ q[6]) + __popc(v.w ^ q[7]);
        if (dist > 50) continue;                                       // TH_LOW
        const eorb_keypoint kp2 = k2.kps[i2];
        if (!((f1 | f2) & 2u)) {                                       // neither side stereo: not too close to the epipole (:1077-1085)
            const float dx = __fsub_rn(tg.ep[0], kp2.x), dy = __fsub_rn(tg.ep[1], kp2.y);
            const int lv = kp2.octave < 0 ? 0 : (kp2.octave > 31 ? 31 : kp2.octave);
            if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.0f, tg.scale2[lv])) continue;
        }
        if (!tg.coarse) {
            if (den == 0.0f) continue;
            const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, kp2.x), __fmul_rn(lb, kp2.y)), lc);
            const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
            const int lv = kp2.octave < 0 ? 0 : (kp2.octave > 31 ? 31 : kp2.octave);
            if (!(dsqr < __fmul_rn(3.84f, tg.sigma2[lv]))) continue;   // DEF_EC_DIST_COEF * unc (include/CameraModels/Pinhole.h:36)
        }
        best = min(best, ((uint32_t)dist << 16) | (uint32_t)(0xffff - (j - fb)));
    }
    best = __reduce_min_sync(FULLMASK, best);
    if (best == 0xffffffffu || lane != 0) return;
    const int i2 = (int)k2.feats[fb + (0xffff - (int)(best & 0xffffu))];
    match12[i1] = i2;
    int bin = -1;
    if (checkOri) {
        float rot = __fsub_rn(kp1.angle, k2.kps[i2].angle);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));
        if (bin == 30) bin = 0;
        if (bin >= 0 && bin < 30) atomicAdd(&work[bin], 1); else bin = -1;
    }
    binOf[i1] = (signed char)bin;
    atomicAdd(&work[32], 1);
}

// rotation consistency (:1177-1198): matches in the bins that are not among the three maxima are taken back
__global__ void __launch_bounds__(256) triangulation_rot_kernel(int n1, int checkOri, int32_t* __restrict__ match12, const signed char* __restrict__ binOf,
                                                                const int* __restrict__ work, int* __restrict__ nmatchesOut) {
    __shared__ int sInd[3], sDrop;
    const int tid = threadIdx.x;
    if (tid == 0) {
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < 30; i++) {
            const int s = work[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
        sInd[0] = ind1; sInd[1] = ind2; sInd[2] = ind3; sDrop = 0;
    }
    __syncthreads();
    if (checkOri)
        for (int i = tid; i < n1; i += 256) {
            if (match12[i] < 0) continue;
            const int b = binOf[i];
            if (b >= 0 && b != sInd[0] && b != sInd[1] && b != sInd[2]) { match12[i] = -1; atomicAdd(&sDrop, 1); }
        }
    __syncthreads();
    if (tid == 0) *nmatchesOut = work[32] - sDrop;
}

cudaError_t launch_search_triangulation(const GuidedBowSide& k1, const uint8_t* d_flags1, const GuidedBowSide& k2, const uint8_t* d_flags2,
                                        const GuidedTriGeom& tg, int checkOri, int nEntries1, int32_t* d_match12, signed char* d_binOf,
                                        int* d_work /* 64 ints */, int* d_nmatches, cudaStream_t st, long long* launches) {
    cudaError_t e = cudaMemsetAsync(d_match12, 0xff, (size_t)(k1.n > 0 ? k1.n : 1) * sizeof(int32_t), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(d_work, 0, 64 * sizeof(int), st);
    if (e != cudaSuccess) return e;
    if (k1.nnodes > 0 && k2.nnodes > 0 && nEntries1 > 0) {
        search_triangulation_kernel<<<(nEntries1 + BOW_WARPS - 1) / BOW_WARPS, BOW_WARPS * 32, 0, st>>>(k1, d_flags1, k2, d_flags2, tg, checkOri, nEntries1,
                                                                                                       d_match12, d_binOf, d_work);
        (*launches)++;
    }
    triangulation_rot_kernel<<<1, 256, 0, st>>>(k1.n, checkOri, d_match12, d_binOf, d_work, d_nmatches);
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t guided_configure() {
    {
        const cudaError_t e0 = cudaFuncSetAttribute(guided_resolve_map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)resolveMapSmem(EORB_GUIDED_MAX_KPS, EORB_GUIDED_MAX_KPS));
        if (e0 != cudaSuccess) return e0;
    }
    cudaError_t e = cudaFuncSetAttribute(guided_resolve_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)resolveProjSmem(EORB_GUIDED_MAX_KPS, EORB_GUIDED_MAX_KPS));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(guided_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)resolveSmem(EORB_GUIDED_MAX_KPS, EORB_GUIDED_MAX_KPS));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(frame_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gridSmem(EORB_GUIDED_MAX_KPS));
}

cudaError_t launch_frame_grid(const eorb_keypoint* d_kps, int n, GuidedGrid g, int* d_cellStart, int* d_cellIdx, int* d_assigned, cudaStream_t st) {
    frame_grid_kernel<<<1, 1024, gridSmem(n), st>>>(d_kps, n, g, d_cellStart, d_cellIdx, d_assigned);
    return cudaGetLastError();
}

cudaError_t launch_features_in_area(const eorb_keypoint* d_kps, GuidedGrid g, const int* d_cellStart, const int* d_cellIdx,
                                    const eorb_area_query* d_q, int nq, int* d_count, int* d_out, int capPerQuery, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    features_in_area_kernel<<<(nq + 7) / 8, 256, 0, st>>>(d_kps, g, d_cellStart, d_cellIdx, d_q, nq, d_count, d_out, capPerQuery);
    return cudaGetLastError();
}

cudaError_t launch_search_init(const GuidedFrame& f1, const GuidedFrame& f2, GuidedGrid g, float* d_prevXY, int window, float nnratio, int checkOri,
                               const GuidedWork& w, int32_t* d_matches12, int* d_nmatches, cudaStream_t st, long long* launches) {
    cudaError_t e = cudaMemsetAsync(w.total, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    e = launch_frame_grid(f2.kps, f2.n, g, w.cellStart, w.cellIdx, w.assigned, st);
    if (e != cudaSuccess) return e;
    (*launches)++;
    if (f1.n > 0) {
        guided_candidates_kernel<<<(f1.n + 7) / 8, 256, 0, st>>>(f1, f2, g, d_prevXY, (float)window, nullptr, w, nullptr, nullptr, GuidedCandExtra{});
        (*launches)++;
    }
    guided_resolve_kernel<<<1, 256, resolveSmem(f1.n, f2.n), st>>>(f1, f2, d_prevXY, nnratio, checkOri, w, d_matches12, d_nmatches);
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_search_proj(const float* d_x3Dc, const uint8_t* d_valid1, const int32_t* d_obs1, const eorb_keypoint* d_kps1,
                               const uint8_t* d_descMP, int n1, const GuidedFrame& f2, GuidedGrid g, const GuidedProj& pr, int checkOri,
                               const GuidedWork& w, int32_t* d_claim, int32_t* d_matchCur, int* d_nmatches, cudaStream_t st, long long* launches,
                               const GuidedProjMode& md, const float* d_uRight2, const uint8_t* d_held2, int thHigh) {
    cudaError_t e = cudaMemsetAsync(w.total, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    e = launch_frame_grid(f2.kps, f2.n, g, w.cellStart, w.cellIdx, w.assigned, st);
    if (e != cudaSuccess) return e;
    (*launches)++;
    if (n1 > 0) {
        guided_project_kernel<<<(n1 + 255) / 256, 256, 0, st>>>(d_x3Dc, d_valid1, d_kps1, n1, pr, w.q, md);
        GuidedFrame f1{d_kps1, d_descMP, n1};
        guided_candidates_kernel<<<(n1 + 7) / 8, 256, 0, st>>>(f1, f2, g, nullptr, 0.f, w.q, w, d_uRight2, d_uRight2 ? md.qUr : nullptr, GuidedCandExtra{});
        (*launches) += 2;
    }
    guided_resolve_proj_kernel<<<1, 256, resolveProjSmem(n1, f2.n), st>>>(d_kps1, d_obs1, n1, f2, checkOri, w, d_claim, d_matchCur, d_nmatches, d_held2,
                                                                          thHigh);
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_search_map_points(const eorb_track_point* d_pts, const uint8_t* d_descMP, int n1, const GuidedFrame& f2,
                                     const uint8_t* d_held2, GuidedGrid g, const GuidedProj& pr, int farPoints, float thFar, float nnratio,
                                     const GuidedWork& w, int32_t* d_matchCur, int* d_nmatches, cudaStream_t st, long long* launches,
                                     const float* d_projXR, const float* d_uRight2) {
    cudaError_t e = cudaMemsetAsync(w.total, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    e = launch_frame_grid(f2.kps, f2.n, g, w.cellStart, w.cellIdx, w.assigned, st);
    if (e != cudaSuccess) return e;
    (*launches)++;
    if (n1 > 0) {
        guided_mappoint_query_kernel<<<(n1 + 255) / 256, 256, 0, st>>>(d_pts, n1, pr, farPoints, thFar, w.q);
        GuidedFrame f1{nullptr, d_descMP, n1};
        guided_candidates_kernel<<<(n1 + 7) / 8, 256, 0, st>>>(f1, f2, g, nullptr, 0.f, w.q, w, d_uRight2, d_projXR, GuidedCandExtra{});
        (*launches) += 2;
    }
    guided_resolve_map_kernel<<<1, 256, resolveMapSmem(n1, f2.n), st>>>(d_pts, n1, f2, d_held2, nnratio, w, d_matchCur, d_nmatches);
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_search_by_bow(const GuidedBowSide& kf, const uint8_t* d_validKF, const GuidedBowSide& f, float nnratio, int checkOri,
                                 int32_t* d_matchF, int* d_work, int* d_nmatches, cudaStream_t st, long long* launches,
                                 const uint8_t* d_validF, int32_t* d_match12) {
    cudaError_t e = cudaMemsetAsync(d_matchF, 0xff, (size_t)f.n * sizeof(int32_t), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(d_work, 0, 64 * sizeof(int), st);
    if (e != cudaSuccess) return e;
    if (d_match12) {
        e = cudaMemsetAsync(d_match12, 0xff, (size_t)kf.n * sizeof(int32_t), st);
        if (e != cudaSuccess) return e;
    }
    const int blocks = kf.nnodes > 0 ? (kf.nnodes + BOW_WARPS - 1) / BOW_WARPS : 1;
    // the keyframe-keyframe form (d_match12 given) accepts bestDist1 < TH_LOW, the keyframe-frame form bestDist1 <= TH_LOW
    search_by_bow_kernel<<<blocks, BOW_WARPS * 32, 0, st>>>(kf, d_validKF, f, nnratio, checkOri, d_matchF, d_work, d_nmatches, d_validF,
                                                           d_match12 ? 49 : 50, d_match12);
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_search_windows(const eorb_area_query* d_q, const float* d_qUr, const uint8_t* d_descMP, int n1, const GuidedFrame& f2,
                                  const uint8_t* d_held2, const float* d_uRight2, GuidedGrid g, GuidedGrid gq, const GuidedCandExtra& cx, int blocking,
                                  int thHigh, const GuidedWork& w, int32_t* d_claim, int32_t* d_bestIdx, int32_t* d_bestDist, int32_t* d_match2, int* d_nmatches,
                                  cudaStream_t st, long long* launches) {
    cudaError_t e = cudaMemsetAsync(w.total, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(d_nmatches, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    e = launch_frame_grid(f2.kps, f2.n, g, w.cellStart, w.cellIdx, w.assigned, st);
    if (e != cudaSuccess) return e;
    (*launches)++;
    GuidedFrame f1{nullptr, d_descMP, n1};
    GuidedCandExtra c2 = cx;
    c2.held2 = blocking ? nullptr : d_held2;      // the blocking resolve keeps its own table of taken slots
    if (n1 > 0) {
        // the right-image columns only enter through Fuse's reprojection gate; without it they are ignored (no rectified-stereo column test here)
        guided_candidates_kernel<<<(n1 + 7) / 8, 256, 0, st>>>(f1, f2, gq, nullptr, 0.f, d_q, w, cx.chi2 ? d_uRight2 : nullptr,
                                                               cx.chi2 ? d_qUr : nullptr, c2);
        (*launches)++;
    }
    if (blocking) {
        // vpMatched[idx] != NULL skips the keypoint whatever the point (ORBmatcher.cc:563, :678): the relocalisation form of the ordered resolve
        guided_resolve_proj_kernel<<<1, 256, resolveProjSmem(n1, f2.n), st>>>(nullptr, nullptr, n1, f2, 0, w, d_claim, d_match2, d_nmatches, d_held2, thHigh);
        (*launches)++;
        if (n1 > 0) {
            guided_claim_dist_kernel<<<(n1 + 255) / 256, 256, 0, st>>>(d_descMP, f2.desc, n1, d_claim, d_bestIdx, d_bestDist);
            (*launches)++;
        }
    } else if (n1 > 0) {
        guided_best_kernel<<<(n1 + 255) / 256, 256, 0, st>>>(n1, w, thHigh, d_bestIdx, d_bestDist, d_nmatches);
        (*launches)++;
    }
    return cudaGetLastError();
}

}  // namespace eorb
