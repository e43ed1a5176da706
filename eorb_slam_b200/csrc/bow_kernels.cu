// bow_kernels.cu — the two steps that follow extraction in the reference's Frame.
//
//   B1 bow_descend_kernel   TemplatedVocabulary::transform(feature, word, weight, nid, levelsup) (:1218-1258): one warp per
//                           descriptor; at every level lane c takes child c (two 128-bit loads of its 256-bit centre, 8 POPC),
//                           the warp keeps the minimum of (distance << 8 | child position) — strict '<' in the reference's
//                           scan means the first child wins a tie — and descends until a node has no children.
//   B2 bow_assemble_kernel  BowVector (block 0) and FeatureVector (block 1) of transform(features, v, fv, levelsup)
//                           (:1127-1200): std::map insertion becomes a shared-memory bitonic sort of (id << 32 | feature)
//                           keys; a word's value is the sum of its features' weights in feature order (addWeight) or the first
//                           one (addIfNotExist); the L1 / L2 norm is accumulated by ONE thread in word order (BowVector.cpp:
//                           normalize iterates the map), so every double is bit-identical to the reference's.
//   U1 undistort_keypoints_kernel  cv::undistortPoints(mat, mat, K, dist, noArray(), K) (Frame.cc:805-840): 5 iterations of
//                           the inverse Brown-Conrady model in double without FMA contraction (the cv2 wheel's result, bit
//                           for bit), position replaced, the other cv::KeyPoint fields copied (Frame.cc:833-838).
#include <cuda_runtime.h>
#include <stdint.h>

#include "bow_kernels.h"

namespace eorb {

typedef unsigned long long u64;
#define FULLMASK 0xffffffffu

// ------------------------------------------------------------------------------------------------ B1
__global__ void __launch_bounds__(256) bow_descend_kernel(VocabDev v, const uint8_t* __restrict__ feats, int n, int nidLevel, uint32_t* __restrict__ wordId,
                                                          double* __restrict__ weight, uint32_t* __restrict__ nodeId) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    uint32_t q[8];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(feats + (size_t)i * 32);
        const uint4 a = __ldg(qp), b = __ldg(qp + 1);
        q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
    }
    int node = 0, level = 0;
    uint32_t nid = 0;
    int cs = __ldg(v.childStart), ce = __ldg(v.childStart + 1);
    while (ce > cs) {
        ++level;
        uint32_t best = 0xffffffffu;
        for (int base = cs; base < ce; base += 32) {       // k <= 20 in DBoW2: one round
            const int c = base + lane;
            uint32_t key = 0xffffffffu;
            if (c < ce) {
                const int id = __ldg(v.children + c);
                const uint4* dp = reinterpret_cast<const uint4*>(v.desc + (size_t)id * 32);
                const uint4 a = __ldg(dp), b = __ldg(dp + 1);
                const int d = __popc(a.x ^ q[0]) + __popc(a.y ^ q[1]) + __popc(a.z ^ q[2]) + __popc(a.w ^ q[3]) + __popc(b.x ^ q[4]) +
                              __popc(b.y ^ q[5]) + __popc(b.z ^ q[6]) + __popc(b.w ^ q[7]);
                key = ((uint32_t)d << 16) | (uint32_t)(c - cs);
            }
            best = min(best, __reduce_min_sync(FULLMASK, key));
        }
        node = __ldg(v.children + cs + (int)(best & 0xffffu));
        if (level == nidLevel) nid = (uint32_t)node;
        cs = __ldg(v.childStart + node); ce = __ldg(v.childStart + node + 1);
    }
    if (lane == 0) {
        wordId[i] = __ldg(v.wordId + node);
        weight[i] = __ldg(v.weight + node);
        nodeId[i] = nid;
    }
}

// ------------------------------------------------------------------------------------------------ B2
__device__ __forceinline__ void block_sort_u64(u64* keys, int np2, int tid) {
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < np2; i += 1024) {
                const int o = i ^ j;
                if (o > i) {
                    const u64 a = keys[i], b = keys[o];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[o] = a; }
                }
            }
            __syncthreads();
        }
}

// exclusive prefix over the block's per-thread counts (1024 threads); returns this thread's offset, *total = sum
__device__ __forceinline__ int block_exclusive_scan(int v, int tid, int* sWarp, int* total) {
    const int lane = tid & 31, warp = tid >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(FULLMASK, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = sWarp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(FULLMASK, wi, o);
            if (lane >= o) wi += up;
        }
        sWarp[lane] = wi - w;
        if (lane == 31) sWarp[32] = wi;
    }
    __syncthreads();
    *total = sWarp[32];
    return sWarp[warp] + incl - v;
}

__global__ void __launch_bounds__(1024) bow_assemble_kernel(int n, int accumulate, int norm, BowOut o) {
    extern __shared__ __align__(16) unsigned char sm[];
    int np2 = 2;
    while (np2 < n) np2 <<= 1;
    u64* keys = reinterpret_cast<u64*>(sm);               // [np2]
    double* svals = reinterpret_cast<double*>(keys + np2); // [np2] (block 0)
    __shared__ int sWarp[33];
    __shared__ double sNorm;
    const int tid = threadIdx.x;
    const bool isBow = blockIdx.x == 0;
    const uint32_t* ids = isBow ? o.wordId : o.nodeId;
    for (int i = tid; i < np2; i += 1024) keys[i] = (i < n && o.weight[i] > 0) ? (((u64)ids[i] << 32) | (u64)i) : ~0ull;   // w > 0: not stopped
    __syncthreads();
    block_sort_u64(keys, np2, tid);
    // every thread owns a contiguous run of sorted positions; a position is a head when its id differs from its predecessor's
    const int per = np2 / 1024 > 0 ? np2 / 1024 : 1;
    const int j0 = tid * per, j1 = min(j0 + per, np2);
    int heads = 0, valid = 0;
    for (int j = j0; j < j1 && j < np2; j++) {
        const u64 kx = keys[j];
        if (kx == ~0ull) break;
        valid++;
        if (j == 0 || (uint32_t)(keys[j - 1] >> 32) != (uint32_t)(kx >> 32)) heads++;
    }
    int totalHeads = 0, totalValid = 0;
    int q = block_exclusive_scan(heads, tid, sWarp, &totalHeads);
    __syncthreads();
    block_exclusive_scan(valid, tid, sWarp, &totalValid);
    for (int j = j0; j < j1 && j < np2; j++) {
        const u64 kx = keys[j];
        if (kx == ~0ull) break;
        const uint32_t id = (uint32_t)(kx >> 32);
        if (!isBow) o.fvFeats[j] = (uint32_t)kx;
        if (j == 0 || (uint32_t)(keys[j - 1] >> 32) != id) {
            if (isBow) {
                double val = o.weight[(uint32_t)kx];
                if (accumulate)
                    for (int t = j + 1; t < np2 && keys[t] != ~0ull && (uint32_t)(keys[t] >> 32) == id; t++) val = __dadd_rn(val, o.weight[(uint32_t)keys[t]]);
                o.bowIds[q] = id;
                svals[q] = val;
            } else {
                o.fvNodes[q] = id;
                o.fvStart[q] = j;
            }
            q++;
        }
    }
    __syncthreads();
    if (!isBow) {
        if (tid == 0) { o.fvStart[totalHeads] = totalValid; o.counts[1] = totalHeads; }
        return;
    }
    const int nbow = totalHeads;
    if (tid == 0) {
        double nrm = 0.0;
        if (norm == 1) { for (int t = 0; t < nbow; t++) nrm = __dadd_rn(nrm, fabs(svals[t])); }
        else if (norm == 2) { for (int t = 0; t < nbow; t++) nrm = __dadd_rn(nrm, __dmul_rn(svals[t], svals[t])); nrm = sqrt(nrm); }
        else nrm = accumulate ? (double)nbow : 0.0;       // TF / TF_IDF without normalisation: values / number of words (:1162-1168)
        sNorm = nrm;
        o.counts[0] = nbow;
    }
    __syncthreads();
    const double nrm = sNorm;
    for (int t = tid; t < nbow; t += 1024) o.bowVals[t] = nrm > 0.0 ? __ddiv_rn(svals[t], nrm) : svals[t];
}

// ------------------------------------------------------------------------------------------------ U1
struct UndistConst { double fx, fy, cx, cy, ifx, ify, k0, k1, p1, p2, k4; };

__global__ void __launch_bounds__(256) undistort_keypoints_kernel(const eorb_keypoint* __restrict__ in, eorb_keypoint* __restrict__ out, int n, UndistConst c) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    eorb_keypoint kp = in[i];
    const double u = (double)kp.x, v = (double)kp.y;
    double x = __dmul_rn(__dsub_rn(u, c.cx), c.ifx), y = __dmul_rn(__dsub_rn(v, c.cy), c.ify);
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        const double xx = __dmul_rn(x, x), yy = __dmul_rn(y, y);
        const double r2 = __dadd_rn(xx, yy);
        // icdist = 1 / (1 + ((k4 r2 + k1) r2 + k0) r2), Horner without contraction
        const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(c.k4, r2), c.k1), r2), c.k0), r2));
        const double icdist = __ddiv_rn(1.0, den);
        if (icdist < 0) { x = __dmul_rn(__dsub_rn(u, c.cx), c.ifx); y = __dmul_rn(__dsub_rn(v, c.cy), c.ify); break; }
        // deltaX = 2 p1 x y + p2 (r2 + 2 x x);  deltaY = p1 (r2 + 2 y y) + 2 p2 x y   (left-to-right products as in the source)
        const double dX = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, c.p1), x), y), __dmul_rn(c.p2, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x))));
        const double dY = __dadd_rn(__dmul_rn(c.p1, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))), __dmul_rn(__dmul_rn(__dmul_rn(2.0, c.p2), x), y));
        x = __dmul_rn(__dsub_rn(x0, dX), icdist);
        y = __dmul_rn(__dsub_rn(y0, dY), icdist);
    }
    kp.x = (float)__dadd_rn(__dmul_rn(c.fx, x), c.cx);
    kp.y = (float)__dadd_rn(__dmul_rn(c.fy, y), c.cy);
    out[i] = kp;
}

// ------------------------------------------------------------------------------------------------ launches
cudaError_t bow_configure() {
    return cudaFuncSetAttribute(bow_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EORB_BOW_MAX_FEATS * 16);
}

cudaError_t launch_bow_transform(const VocabDev& v, const uint8_t* d_feats, int n, int levelsup, int accumulate, int norm, const BowOut& o,
                                 cudaStream_t st, long long* launches) {
    if (n > 0) {
        bow_descend_kernel<<<(n + 7) / 8, 256, 0, st>>>(v, d_feats, n, v.L - levelsup, o.wordId, o.weight, o.nodeId);
        (*launches)++;
    }
    int np2 = 2;
    while (np2 < n) np2 <<= 1;
    bow_assemble_kernel<<<2, 1024, (size_t)np2 * 16, st>>>(n, accumulate, norm, o);
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_undistort_keypoints(const eorb_keypoint* d_in, eorb_keypoint* d_out, int n, const float* K4, const float* dist5, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    UndistConst c;
    c.fx = K4[0]; c.fy = K4[1]; c.cx = K4[2]; c.cy = K4[3];
    c.ifx = 1. / c.fx; c.ify = 1. / c.fy;
    c.k0 = dist5[0]; c.k1 = dist5[1]; c.p1 = dist5[2]; c.p2 = dist5[3]; c.k4 = dist5[4];
    undistort_keypoints_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_in, d_out, n, c);
    return cudaGetLastError();
}

}  // namespace eorb
