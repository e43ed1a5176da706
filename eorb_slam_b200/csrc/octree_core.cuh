// octree_core.cuh — level-synchronous restatement of ORBextractor::DistributeOctTree
// (reference src/ORBextractor.cc:558-782, ExtractorNode::DivideNode :500-556).
//
// The reference walks a std::list serially: every node with >1 keys is split into up to four children that
// are push_front-ed, until the node count reaches the quota N; the last rounds split the largest nodes first
// (sorted by (size, pointer)) and stop mid-way as soon as N is reached.  Here one thread block owns one
// (frame, level) problem and replays exactly that sequence with data-parallel steps:
//
//   * a key never moves: it carries the list position of its node (knode[]); one pass of shared-memory
//     atomics gives the 4 child populations of every candidate node;
//   * the list is rebuilt each round by prefix sums: children of the i-th processed node, created in order
//     n1..n4, get creation index g and list position T-1-g (that is what repeated push_front produces),
//     untouched nodes follow in their old order;
//   * the "largest first, stop at N" rounds sort the candidates by (size desc, creation desc) with a bitonic
//     network, prefix-sum the growth (children-1) along that order and split exactly the prefix the serial
//     loop would have split before its `break` (:749-750);
//   * size ties: the reference compares heap pointers; the pin (shared with oracle/orb_oracle.cc) is the
//     creation order, later-created = larger.
//
// The same source compiles as plain C++ (EORB_HOST_MODEL) with one "thread", so tests/ can check the
// formulation against the serial oracle on thousands of random cases without a GPU.
#pragma once
#include <stdint.h>

#ifdef EORB_HOST_MODEL
#define OCT_DEV inline
#define OCT_TID 0
#define OCT_NT 1
#define OCT_SYNC() ((void)0)
static inline int oct_atomic_add(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned oct_atomic_max(unsigned* p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
static inline int oct_atomic_min(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
#else
#define OCT_DEV __device__ __forceinline__
#define OCT_TID ((int)threadIdx.x)
#define OCT_NT ((int)blockDim.x)
#define OCT_SYNC() __syncthreads()
__device__ __forceinline__ int oct_atomic_add(int* p, int v) { return atomicAdd(p, v); }
__device__ __forceinline__ unsigned oct_atomic_max(unsigned* p, unsigned v) { return atomicMax(p, v); }
__device__ __forceinline__ int oct_atomic_min(int* p, int v) { return atomicMin(p, v); }
#endif

namespace eorb {

#ifndef OCT_MAX_THREADS
#define OCT_MAX_THREADS 512   // largest block the kernels are launched with (scratch sizing); launch sets use 128 threads
#endif

#ifndef OCT_KU
#define OCT_KU 2   // keys per thread and step of the key passes, loads first (measured: 1 -> 0.48, 2 -> 0.38, 4 -> 0.39, 8 -> 0.49 us/frame)
#endif

struct OctBox { short x0, y0, x1, y1; };

// packed candidate: x | y << 12 | score << 24   (x, y relative to (minBorderX, minBorderY), < 4096)
OCT_DEV int oct_key_x(uint32_t k) { return (int)(k & 0xFFFu); }
OCT_DEV int oct_key_y(uint32_t k) { return (int)((k >> 12) & 0xFFFu); }
OCT_DEV int oct_key_score(uint32_t k) { return (int)(k >> 24); }

// shared-memory footprint (bytes) for node capacity nc (keep in sync with oct_carve)
static inline size_t oct_smem_bytes(int nc) {
    int p2 = 1; while (p2 < nc) p2 <<= 1;
    size_t b = 0;
    b += 2 * sizeof(OctBox) * (size_t)nc;        // box[2]
    b += 2 * sizeof(int) * (size_t)nc * 3;       // cnt[2], seq[2], cand[2]
    b += sizeof(int) * (size_t)nc * 4;           // cc
    b += sizeof(int) * (size_t)nc * 5;           // nch, split, gfirst, newpos, scanA
    b += sizeof(unsigned long long) * (size_t)p2;// sort keys
    b += sizeof(int) * (OCT_MAX_THREADS + 16);   // scan scratch + scalars
    return (b + 15) & ~(size_t)15;
}

struct OctSmem {
    OctBox* box[2];
    int* cnt[2];
    int* seq[2];
    int* cand[2];
    int* cc;
    int* nch;
    int* split;
    int* gfirst;
    int* newpos;
    int* scanA;
    unsigned long long* skey;
    int* scratch;   // OCT_MAX_THREADS
    int* scal;      // 16 scalars
    int p2;
};

OCT_DEV void oct_carve(OctSmem& s, unsigned char* base, int nc) {
    int p2 = 1; while (p2 < nc) p2 <<= 1;
    s.p2 = p2;
    unsigned char* p = base;
    s.skey = (unsigned long long*)p; p += sizeof(unsigned long long) * (size_t)p2;
    s.box[0] = (OctBox*)p; p += sizeof(OctBox) * (size_t)nc;
    s.box[1] = (OctBox*)p; p += sizeof(OctBox) * (size_t)nc;
    for (int i = 0; i < 2; i++) { s.cnt[i] = (int*)p; p += sizeof(int) * (size_t)nc; }
    for (int i = 0; i < 2; i++) { s.seq[i] = (int*)p; p += sizeof(int) * (size_t)nc; }
    for (int i = 0; i < 2; i++) { s.cand[i] = (int*)p; p += sizeof(int) * (size_t)nc; }
    s.cc = (int*)p; p += sizeof(int) * (size_t)nc * 4;
    s.nch = (int*)p; p += sizeof(int) * (size_t)nc;
    s.split = (int*)p; p += sizeof(int) * (size_t)nc;
    s.gfirst = (int*)p; p += sizeof(int) * (size_t)nc;
    s.newpos = (int*)p; p += sizeof(int) * (size_t)nc;
    s.scanA = (int*)p; p += sizeof(int) * (size_t)nc;
    s.scratch = (int*)p; p += sizeof(int) * OCT_MAX_THREADS;
    s.scal = (int*)p;
}

// in-place exclusive scan of a[0..n) by the whole block; returns the total (same value in every thread)
// NTC: the block size when it is a compile-time constant above 128 (selects the shuffle form), 0 otherwise
template <int NTC = 0>
OCT_DEV int oct_exclusive_scan(int* a, int n, int* scratch) {
    const int tid = OCT_TID, nt = OCT_NT;
    const int per = (n + nt - 1) / nt;
    const int b = tid * per, e = (b + per < n) ? b + per : n;
    int sum = 0;
    for (int i = b; i < e; i++) sum += a[i];
#ifdef EORB_HOST_MODEL
    int total = sum;
    int off = 0;
#else
    int total, off;
    if (NTC > 128) {
        // big blocks (the single-frame call: one 512-thread block per level, where the barriers ARE the latency): warp-shuffle scan of
        // the per-thread sums and a serial pass over the <= 16 warp totals, two barriers instead of 2 * log2(nt).  (With the 128-thread
        // blocks of a launch set the same change was measured slower, 0.48 -> 0.53 us/frame, so those keep the Hillis-Steele form.)
        const int lane = tid & 31, wp = tid >> 5, nw = nt >> 5;
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) scratch[wp] = incl;
        OCT_SYNC();
        int before = 0;
        total = 0;
        for (int j = 0; j < nw; j++) { const int t = scratch[j]; total += t; if (j < wp) before += t; }
        off = before + incl - sum;
    } else {
        // Hillis-Steele over nt partials
        scratch[tid] = sum;
        OCT_SYNC();
        for (int d = 1; d < nt; d <<= 1) {
            int v = (tid >= d) ? scratch[tid - d] : 0;
            OCT_SYNC();
            scratch[tid] += v;
            OCT_SYNC();
        }
        total = scratch[nt - 1];
        off = scratch[tid] - sum;
    }
#endif
    for (int i = b; i < e; i++) { int v = a[i]; a[i] = off; off += v; }
    OCT_SYNC();
    return total;
}

// descending bitonic sort of s.skey[0..p2)
OCT_DEV void oct_bitonic_desc(unsigned long long* k, int p2) {
#ifdef EORB_HOST_MODEL
    for (int i = 1; i < p2; i++) {   // insertion sort (host model only)
        unsigned long long v = k[i]; int j = i - 1;
        while (j >= 0 && k[j] < v) { k[j + 1] = k[j]; j--; }
        k[j + 1] = v;
    }
#else
    for (int size = 2; size <= p2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = OCT_TID; i < (p2 >> 1); i += OCT_NT) {
                int lo = 2 * i - (i & (stride - 1));
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                unsigned long long a = k[lo], b = k[hi];
                if ((a < b) == desc) { k[lo] = b; k[hi] = a; }
            }
            OCT_SYNC();
        }
    }
#endif
}

OCT_DEV int oct_child_of(uint32_t key, const OctBox& b) {
    const int midx = b.x0 + ((b.x1 - b.x0 + 1) >> 1);   // UL.x + ceil((UR.x-UL.x)/2)
    const int midy = b.y0 + ((b.y1 - b.y0 + 1) >> 1);
    const int x = oct_key_x(key), y = oct_key_y(key);
    return (x < midx) ? ((y < midy) ? 0 : 2) : ((y < midy) ? 1 : 3);
}

OCT_DEV OctBox oct_child_box(const OctBox& b, int c) {
    const short midx = (short)(b.x0 + ((b.x1 - b.x0 + 1) >> 1));
    const short midy = (short)(b.y0 + ((b.y1 - b.y0 + 1) >> 1));
    OctBox r;
    r.x0 = (c & 1) ? midx : b.x0; r.x1 = (c & 1) ? b.x1 : midx;
    r.y0 = (c & 2) ? midy : b.y0; r.y1 = (c & 2) ? b.y1 : midy;
    return r;
}

// Runs the distribution.  keys[n]: packed candidates in the reference's order (cell-row-major, then pixel
// row-major).  knode[n]: scratch.  out[<= nodeCap]: selected keys in final list order.  Returns the count.
// All threads of the block must call; smem must hold oct_smem_bytes(nodeCap).
template <int NTC = 0>
OCT_DEV int oct_distribute(const uint32_t* keys, uint16_t* knode, int n, int width, int height, int nIni, float hX,
                           int N, int nodeCap, unsigned char* smem, uint32_t* out) {
    OctSmem s;
    oct_carve(s, smem, nodeCap);
    const int tid = OCT_TID, nt = OCT_NT;
    (void)width;
    if (n <= 0 || nIni <= 0) return 0;
    int cur = 0;
    // ---- roots (:562-582)
    for (int i = tid; i < nIni; i += nt) {
        OctBox b;
        b.x0 = (short)(int)(hX * (float)i); b.y0 = 0;
        b.x1 = (short)(int)(hX * (float)(i + 1)); b.y1 = (short)height;
        s.box[cur][i] = b; s.cnt[cur][i] = 0; s.seq[cur][i] = i;
    }
    OCT_SYNC();
    for (int k0 = tid; k0 < n; k0 += OCT_KU * nt) {
        uint32_t ky[OCT_KU];
#pragma unroll
        for (int j = 0; j < OCT_KU; j++) { const int k = k0 + j * nt; ky[j] = k < n ? keys[k] : 0u; }
#pragma unroll
        for (int j = 0; j < OCT_KU; j++) {
            const int k = k0 + j * nt;
            if (k < n) {
                int slot = (int)((float)oct_key_x(ky[j]) / hX);
                if (slot >= nIni) slot = nIni - 1;
                knode[k] = (uint16_t)slot;
                oct_atomic_add(&s.cnt[cur][slot], 1);
            }
        }
    }
    OCT_SYNC();
    // drop empty roots (:590-603)
    for (int i = tid; i < nIni; i += nt) s.scanA[i] = s.cnt[cur][i] > 0 ? 1 : 0;
    OCT_SYNC();
    int L = oct_exclusive_scan<NTC>(s.scanA, nIni, s.scratch);
    for (int i = tid; i < nIni; i += nt) {
        if (s.cnt[cur][i] > 0) {
            int np = s.scanA[i];
            s.box[cur ^ 1][np] = s.box[cur][i]; s.cnt[cur ^ 1][np] = s.cnt[cur][i];
            s.seq[cur ^ 1][np] = i; s.cand[cur ^ 1][np] = s.cnt[cur][i] > 1;
        }
    }
    OCT_SYNC();
    for (int k = tid; k < n; k += nt) knode[k] = (uint16_t)s.scanA[knode[k]];
    OCT_SYNC();
    cur ^= 1;

    bool finalPhase = false;
    for (int iter = 0; iter < 64; iter++) {
        const int prevL = L;
        OctBox* box = s.box[cur]; int* cnt = s.cnt[cur]; int* seq = s.seq[cur]; int* cand = s.cand[cur];
        // ---- pass 1: child populations of every candidate node
        for (int i = tid; i < 4 * L; i += nt) s.cc[i] = 0;
        OCT_SYNC();
        // key passes are unrolled by 4 with the loads first: the passes wait on global memory (keys, node positions), and
        // the shared-memory atomics in the body keep the compiler from overlapping the loads of consecutive keys itself
        for (int k0 = tid; k0 < n; k0 += OCT_KU * nt) {
            int pp[OCT_KU]; uint32_t ky[OCT_KU];
#pragma unroll
            for (int j = 0; j < OCT_KU; j++) { const int k = k0 + j * nt; pp[j] = k < n ? (int)knode[k] : 0; ky[j] = k < n ? keys[k] : 0u; }
#pragma unroll
            for (int j = 0; j < OCT_KU; j++) {
                const int k = k0 + j * nt, p = pp[j];
                if (k < n && cand[p]) oct_atomic_add(&s.cc[4 * p + oct_child_of(ky[j], box[p])], 1);
            }
        }
        OCT_SYNC();
        for (int p = tid; p < L; p += nt) {
            int c = 0;
            if (cand[p]) c = (s.cc[4 * p] > 0) + (s.cc[4 * p + 1] > 0) + (s.cc[4 * p + 2] > 0) + (s.cc[4 * p + 3] > 0);
            s.nch[p] = c;
            s.split[p] = 0;
        }
        OCT_SYNC();
        int T;
        if (!finalPhase) {
            // every candidate is split, processing order = list order (:620-683)
            for (int p = tid; p < L; p += nt) { s.scanA[p] = s.nch[p]; s.split[p] = cand[p]; }
            OCT_SYNC();
            T = oct_exclusive_scan<NTC>(s.scanA, L, s.scratch);
            for (int p = tid; p < L; p += nt) s.gfirst[p] = s.scanA[p];
            OCT_SYNC();
        } else {
            // largest first, later-created first among equals, stop once the list holds N nodes (:696-751)
            for (int i = tid; i < s.p2; i += nt) {
                unsigned long long key = 0;
                if (i < L && cand[i])
                    key = ((unsigned long long)(unsigned)cnt[i] << 40) | ((unsigned long long)(unsigned)seq[i] << 16) |
                          (unsigned long long)i;
                s.skey[i] = key;
            }
            if (tid == 0) s.scal[0] = 0x7fffffff;
            OCT_SYNC();
            oct_bitonic_desc(s.skey, s.p2);
            // m = number of candidates = first zero key
            for (int q = tid; q < L; q += nt) s.scanA[q] = s.skey[q] ? s.nch[(int)(s.skey[q] & 0xFFFFu)] : 0;
            OCT_SYNC();
            int total = oct_exclusive_scan<NTC>(s.scanA, L, s.scratch);
            (void)total;
            for (int q = tid; q < L; q += nt) {
                if (s.skey[q]) {
                    int p = (int)(s.skey[q] & 0xFFFFu);
                    int incl = s.scanA[q] + s.nch[p];
                    if (L + incl - (q + 1) >= N) oct_atomic_min(&s.scal[0], q + 1);
                }
            }
            OCT_SYNC();
            int K = s.scal[0];   // 0x7fffffff: split every candidate
            T = 0;
            for (int q = tid; q < L; q += nt) {
                if (s.skey[q] && q < K) {
                    int p = (int)(s.skey[q] & 0xFFFFu);
                    s.split[p] = 1;
                    s.gfirst[p] = s.scanA[q];
                }
            }
            OCT_SYNC();
            // T = children of the split prefix
            for (int p = tid; p < L; p += nt) s.scanA[p] = s.split[p] ? s.nch[p] : 0;
            OCT_SYNC();
            T = oct_exclusive_scan<NTC>(s.scanA, L, s.scratch);
        }
        // ---- untouched nodes keep their order behind the new children
        for (int p = tid; p < L; p += nt) s.newpos[p] = s.split[p] ? 0 : 1;
        OCT_SYNC();
        const int U = oct_exclusive_scan<NTC>(s.newpos, L, s.scratch);
        const int Lnew = T + U;
        if (Lnew > nodeCap) return -1;   // cannot happen (bounded by max(N+2, 4*nIni)); guards smem
        OctBox* nbox = s.box[cur ^ 1]; int* ncnt = s.cnt[cur ^ 1]; int* nseq = s.seq[cur ^ 1]; int* ncand = s.cand[cur ^ 1];
        for (int p = tid; p < L; p += nt) {
            if (s.split[p]) {
                int g = s.gfirst[p];
                for (int c = 0; c < 4; c++) {
                    int pop = s.cc[4 * p + c];
                    if (pop > 0) {
                        int np = T - 1 - g;
                        nbox[np] = oct_child_box(box[p], c);
                        ncnt[np] = pop; nseq[np] = g; ncand[np] = pop > 1;
                        s.cc[4 * p + c] = np;
                        g++;
                    }
                }
            } else {
                int np = T + s.newpos[p];
                nbox[np] = box[p]; ncnt[np] = cnt[p]; nseq[np] = seq[p]; ncand[np] = 0;
                s.newpos[p] = np;
            }
        }
        OCT_SYNC();
        for (int k0 = tid; k0 < n; k0 += OCT_KU * nt) {
            int pp[OCT_KU]; uint32_t ky[OCT_KU];
#pragma unroll
            for (int j = 0; j < OCT_KU; j++) { const int k = k0 + j * nt; pp[j] = k < n ? (int)knode[k] : 0; ky[j] = k < n ? keys[k] : 0u; }
#pragma unroll
            for (int j = 0; j < OCT_KU; j++) {
                const int k = k0 + j * nt, p = pp[j];
                if (k < n) knode[k] = (uint16_t)(s.split[p] ? s.cc[4 * p + oct_child_of(ky[j], box[p])] : s.newpos[p]);
            }
        }
        // nToExpand = children with more than one key
        for (int p = tid; p < Lnew; p += nt) s.scanA[p] = 0;
        OCT_SYNC();
        cur ^= 1;
        L = Lnew;
        for (int p = tid; p < L; p += nt) s.scanA[p] = s.cand[cur][p];
        OCT_SYNC();
        const int nToExpand = oct_exclusive_scan<NTC>(s.scanA, L, s.scratch);
        if (L >= N || L == prevL) break;
        if (!finalPhase && L + 3 * nToExpand > N) finalPhase = true;
    }
    // ---- best key per node: max response, first in input order wins ties (:764-778)
    unsigned* best = (unsigned*)s.scanA;
    for (int p = tid; p < L; p += nt) best[p] = 0;
    OCT_SYNC();
    for (int k0 = tid; k0 < n; k0 += OCT_KU * nt) {
        int pp[OCT_KU]; uint32_t ky[OCT_KU];
#pragma unroll
        for (int j = 0; j < OCT_KU; j++) { const int k = k0 + j * nt; pp[j] = k < n ? (int)knode[k] : 0; ky[j] = k < n ? keys[k] : 0u; }
#pragma unroll
        for (int j = 0; j < OCT_KU; j++) {
            const int k = k0 + j * nt;
            if (k < n) oct_atomic_max(&best[pp[j]], ((unsigned)oct_key_score(ky[j]) << 20) | (unsigned)(0xFFFFF - k));
        }
    }
    OCT_SYNC();
    for (int p = tid; p < L; p += nt) out[p] = keys[0xFFFFF - (int)(best[p] & 0xFFFFFu)];
    OCT_SYNC();
    return L;
}

}  // namespace eorb
