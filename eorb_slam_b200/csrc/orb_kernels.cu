// orb_kernels.cu — hand-written sm_100a kernels of the ORB extractor path
// (reference src/ORBextractor.cc; one kernel per reference stage):
//   K1 pyr_resize_kernel     ComputePyramid            :1240-1265  (cv::resize INTER_LINEAR, 11-bit fixed point)
//   K2 fast_cells_kernel     ComputeKeyPointsOctTree   :784-878    (grid FAST 9/16 + cell-local NMS + threshold fallback; orb_fast.cu)
//   K3 octree_kernel         DistributeOctTree         :558-782    (level-synchronous node splitting)
//   K7 orb_index_kernel      operator() ordering       :1152-1172  (level-major order, lapping split from both ends)
//   K5 blur_kernel           GaussianBlur 5x5 s=2      :1141-1142
//   K4+K6 orient_desc_kernel IC_Angle :77-104, computeOrbDescriptor :108-157
// Integer stages are bit-exact with the CPU oracle; float stages use explicit no-FMA intrinsics.
#include <cuda_runtime.h>
#include <mutex>
#include <stdint.h>

#include "../../include/eorb_b200.h"
#include "eorb_math.cuh"
#include "fast_score.cuh"
#include "octree_core.cuh"
#include "orb_kernels.h"
#include "tma_utils.cuh"

namespace eorb {

__device__ const signed char d_brief_pattern[1024] = {
#include "brief_pattern_31.inc"
};
// the same pattern as floats: K6 multiplies the coordinates by cos / sin, this saves the int -> float conversions
__device__ const float d_brief_pattern_f[1024] = {
#include "brief_pattern_31.inc"
};

__device__ __forceinline__ const uint8_t* level_ptr(const OrbArgs& a, const LevelPlan& lp, int level, int f, int& pitch) {
    if (level == 0) {
        pitch = (int)a.lvl0Pitch;
        return a.lvl0 + (size_t)f * (size_t)a.lvl0FrameStride;
    }
    pitch = lp.pitch;
    return a.pyr + (size_t)f * (size_t)a.plan->pyrBytesPerFrame + (size_t)lp.off;
}

// ------------------------------------------------------------------------------------------------ K1
// cv::resize INTER_LINEAR (11-bit fixed point) of level l-1 into level l.  One warp owns 128 destination columns
// x EORB_PYR_BAND destination rows.  A lane owns 4 destination columns; its source taps are per-lane constants:
// two pairs of aligned 32-bit words per source row (columns 0-1 and 2-3), a byte-permute selector per column
// that extracts the two adjacent taps, and the two 11-bit weights packed as 16-bit halves so that ONE dp2a gives
// S = a0*p[sx] + a1*p[sx+1].  Both source rows of every destination row are interpolated in straight-line code
// (no row cache: at s = 1.2 it would save 0.8 of 2 row interpolations, but the warp-uniform branch per row stops the
// loads of consecutive rows from overlapping — measured 0.92 -> 1.19 us/frame), the
// vertical step (b*(S>>4))>>16 is one IMAD.HI per tap, and the loop is unrolled so the loads of the next
// destination row are in flight while the current one is combined.
#ifndef EORB_PYR_HOIST
#define EORB_PYR_HOIST 1
#endif
#ifndef EORB_PYR_BAND
#define EORB_PYR_BAND 16   // measured: 8 -> 0.924, 16 -> 0.904 us/frame
#endif

// rowA / rowB point at the first word of the lane's two 8-byte source windows (columns 0-1 and 2-3)
__device__ __forceinline__ void pyr_hrow(const uint8_t* __restrict__ rowA, const uint8_t* __restrict__ rowB, const unsigned* sel,
                                         const unsigned* wt, unsigned* h) {
    const unsigned A0 = __ldg(reinterpret_cast<const unsigned*>(rowA));
    const unsigned A1 = __ldg(reinterpret_cast<const unsigned*>(rowA + 4));
    const unsigned B0 = __ldg(reinterpret_cast<const unsigned*>(rowB));
    const unsigned B1 = __ldg(reinterpret_cast<const unsigned*>(rowB + 4));
    h[0] = __dp2a_lo(wt[0], __byte_perm(A0, A1, sel[0]), 0u) >> 4;
    h[1] = __dp2a_lo(wt[1], __byte_perm(A0, A1, sel[1]), 0u) >> 4;
    h[2] = __dp2a_lo(wt[2], __byte_perm(B0, B1, sel[2]), 0u) >> 4;
    h[3] = __dp2a_lo(wt[3], __byte_perm(B0, B1, sel[3]), 0u) >> 4;
}

#define PYR_LOAD4(W, PA, PB)                                                   \
    do {                                                                       \
        W[0] = __ldg(reinterpret_cast<const unsigned*>(PA));                   \
        W[1] = __ldg(reinterpret_cast<const unsigned*>(PA + 4));               \
        W[2] = __ldg(reinterpret_cast<const unsigned*>(PB));                   \
        W[3] = __ldg(reinterpret_cast<const unsigned*>(PB + 4));               \
    } while (0)
__device__ __forceinline__ void pyr_hcalc(const unsigned* w, const unsigned* sel, const unsigned* wt, unsigned* h) {
    h[0] = __dp2a_lo(wt[0], __byte_perm(w[0], w[1], sel[0]), 0u) >> 4;
    h[1] = __dp2a_lo(wt[1], __byte_perm(w[0], w[1], sel[1]), 0u) >> 4;
    h[2] = __dp2a_lo(wt[2], __byte_perm(w[2], w[3], sel[2]), 0u) >> 4;
    h[3] = __dp2a_lo(wt[3], __byte_perm(w[2], w[3], sel[3]), 0u) >> 4;
}

// (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2 for four columns, packed into one word.
// bs0 = b0 << 16, bs1 = b1 << 16: (b * x) >> 16 == umulhi(b << 16, x) for the non-negative operands here.
__device__ __forceinline__ unsigned pyr_vrow(const unsigned* __restrict__ h0, const unsigned* __restrict__ h1, unsigned bs0, unsigned bs1) {
    unsigned v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) v[j] = (__umulhi(bs0, h0[j]) + __umulhi(bs1, h1[j]) + 2u) >> 2;
    const unsigned lo = __byte_perm(v[0], v[1], 0x0040);
    const unsigned hi = __byte_perm(v[2], v[3], 0x0040);
    return __byte_perm(lo, hi, 0x5410);
}

#ifndef EORB_PYR_MINB
#define EORB_PYR_MINB 8
#endif
// one block = 128 destination columns x 4 warps x BAND rows of `level`, at tile (bx, by) of frame f
// `ready()` is called by every warp after its per-lane tap constants are formed and before the first source row is read (the chained
// single-launch pyramid waits for the source level there, so the table loads overlap the wait)
struct PyrNoWait { __device__ __forceinline__ void operator()() const {} };
template <int BAND, class Ready>
__device__ __forceinline__ void pyr_resize_block(const OrbArgs& a, int level, int bx, int by, int f, Ready ready) {
    const OrbPlan& P = *a.plan;
    const int dw = P.lv[level].w, dh = P.lv[level].h, dpitch = P.lv[level].pitch;
    const int sw = P.lv[level - 1].w;
    const int xtabOff = P.lv[level].xtabOff, ytabOff = P.lv[level].ytabOff;
    const long long doff = P.lv[level].off;
    const long long pyrBytes = P.pyrBytesPerFrame;
    const int lane = threadIdx.x;
    const int dx0 = bx * 128 + lane * 4;
    const int y0 = (by * (int)blockDim.y + (int)threadIdx.y) * BAND;
    if (y0 >= dh || bx * 128 >= dw) return;     // warp-uniform
    const int y1 = min(y0 + BAND, dh);
    int sp;
    const uint8_t* __restrict__ src = level_ptr(a, P.lv[level - 1], level - 1, f, sp);
    uint8_t* __restrict__ dst = a.pyr + (size_t)f * (size_t)pyrBytes + (size_t)doff + dx0;
    const int4* __restrict__ ytab = a.ytab + ytabOff;
    const bool active = dx0 < dw;
    if (sw < 8) {   // degenerate source width: plain per-pixel evaluation (the windows below need two whole words)
        ready();
        if (active)
            for (int dy = y0; dy < y1; dy++) {
                const int4 yt = __ldg(&ytab[dy]);
                const uint8_t* r0 = src + (size_t)yt.x * sp;
                const uint8_t* r1 = src + (size_t)yt.y * sp;
                for (int j = 0; j < 4 && dx0 + j < dw; j++) {
                    const short4 xt = __ldg(&a.xtab[xtabOff + dx0 + j]);
                    const int S0 = resize_hsum(r0[xt.x], r0[xt.y], xt.z, xt.w), S1 = resize_hsum(r1[xt.x], r1[xt.y], xt.z, xt.w);
                    dst[(size_t)dy * dpitch + j] = (uint8_t)resize_vsum(S0, S1, yt.z >> 16, yt.w >> 16);
                }
            }
        return;
    }
    // per-lane constant taps: window bases (aligned words; a window is the word and its right neighbour, moved one
    // word left at the right edge of the row so that both words exist), byte selectors, packed weights
    int baseA, baseB;
    unsigned sel[4], wt[4];
    {
        const int lastBase = ((sw + 3) & ~3) - 8;          // last window start with both words inside the row
        int sx[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const short4 xt = __ldg(&a.xtab[xtabOff + min(dx0 + j, dw - 1)]);   // sx, sx+1 (clamped), a0, a1
            sx[j] = xt.x;
            wt[j] = (unsigned)(unsigned short)xt.z | ((unsigned)(unsigned short)xt.w << 16);
        }
        baseA = min(sx[0] & ~3, lastBase); baseB = min(sx[2] & ~3, lastBase);
        // byte positions inside the 8-byte window; a second tap that would fall outside (only at the right edge of
        // the row, where its weight is 0) is kept inside 0..7
        const int o0 = sx[0] - baseA, o1 = sx[1] - baseA, o2 = sx[2] - baseB, o3 = sx[3] - baseB;
        sel[0] = (unsigned)(min(o0, 7) | (min(o0 + 1, 7) << 4));
        sel[1] = (unsigned)(min(o1, 7) | (min(o1 + 1, 7) << 4));
        sel[2] = (unsigned)(min(o2, 7) | (min(o2 + 1, 7) << 4));
        sel[3] = (unsigned)(min(o3, 7) | (min(o3 + 1, 7) << 4));
    }
    const uint8_t* __restrict__ laneA = src + baseA;
    const uint8_t* __restrict__ laneB = src + baseB;
    uint8_t* dp = dst + (size_t)y0 * dpitch;
    ready();
#if EORB_PYR_HOIST
    // two destination rows per step, all sixteen source words requested before the first interpolation
    for (int dy = y0; dy < y1; dy += 2) {
        const bool two = dy + 1 < y1;
        const int4 ya = __ldg(&ytab[dy]);                  // sy0, sy1 (clamped), b0 << 16, b1 << 16 — warp-uniform
        const int4 yb = __ldg(&ytab[two ? dy + 1 : dy]);
        const uint8_t* pa0 = laneA + (size_t)(unsigned)ya.x * (unsigned)sp; const uint8_t* pb0 = laneB + (size_t)(unsigned)ya.x * (unsigned)sp;
        const uint8_t* pa1 = laneA + (size_t)(unsigned)ya.y * (unsigned)sp; const uint8_t* pb1 = laneB + (size_t)(unsigned)ya.y * (unsigned)sp;
        const uint8_t* pa2 = laneA + (size_t)(unsigned)yb.x * (unsigned)sp; const uint8_t* pb2 = laneB + (size_t)(unsigned)yb.x * (unsigned)sp;
        const uint8_t* pa3 = laneA + (size_t)(unsigned)yb.y * (unsigned)sp; const uint8_t* pb3 = laneB + (size_t)(unsigned)yb.y * (unsigned)sp;
        unsigned w[4][4];
        PYR_LOAD4(w[0], pa0, pb0); PYR_LOAD4(w[1], pa1, pb1); PYR_LOAD4(w[2], pa2, pb2); PYR_LOAD4(w[3], pa3, pb3);
        unsigned h0[4], h1[4];
        pyr_hcalc(w[0], sel, wt, h0); pyr_hcalc(w[1], sel, wt, h1);
        const unsigned o0 = pyr_vrow(h0, h1, (unsigned)ya.z, (unsigned)ya.w);
        if (active) *reinterpret_cast<unsigned*>(dp) = o0;
        dp += dpitch;
        pyr_hcalc(w[2], sel, wt, h0); pyr_hcalc(w[3], sel, wt, h1);
        const unsigned o1 = pyr_vrow(h0, h1, (unsigned)yb.z, (unsigned)yb.w);
        if (active && two) *reinterpret_cast<unsigned*>(dp) = o1;
        dp += dpitch;
    }
#else
#pragma unroll 2
    for (int dy = y0; dy < y1; dy++) {
        const int4 yt = __ldg(&ytab[dy]);                  // sy0, sy1 (clamped), b0 << 16, b1 << 16 — warp-uniform
        const size_t r0 = (size_t)(unsigned)yt.x * (unsigned)sp, r1 = (size_t)(unsigned)yt.y * (unsigned)sp;
        unsigned h0[4], h1[4];
        pyr_hrow(laneA + r0, laneB + r0, sel, wt, h0);
        pyr_hrow(laneA + r1, laneB + r1, sel, wt, h1);
        const unsigned o = pyr_vrow(h0, h1, (unsigned)yt.z, (unsigned)yt.w);
        if (active) *reinterpret_cast<unsigned*>(dp) = o;
        dp += dpitch;
    }
#endif
}

__global__ void __launch_bounds__(128, EORB_PYR_MINB) pyr_resize_kernel(OrbArgs a, int level) {
    pyr_resize_block<EORB_PYR_BAND>(a, level, blockIdx.x, blockIdx.y, blockIdx.z, PyrNoWait());
}

// K1c: EVERY level in one launch, for small batches (the single-frame call is launch bound: seven dependent launches cost 47 us for
// 8 us of work).  Blocks are laid out level by level (blockIdx.x) per frame (blockIdx.y); a block of level l waits until all tiles of
// level l-1 of its frame have been published (done[f][l-1], release / acquire through __threadfence + a volatile poll), computes its
// tile with pyr_resize_block and publishes itself.  A block only ever waits for blocks of LOWER linear index, which the hardware has
// dispatched before it, so the wait cannot deadlock however many blocks are resident.  done[] is zeroed by the host before the launch.
#ifndef EORB_PYR_CHAIN_BAND
#define EORB_PYR_CHAIN_BAND 4   // destination rows per warp in the single-launch pyramid: short bands = many blocks = short critical path per level
#endif
#define EORB_PYR_CHAIN_ROWS 64   // block rows (16 destination rows each) a level can have in the chain kernel: levels up to 1024 rows
__global__ void __launch_bounds__(128, EORB_PYR_MINB) pyr_chain_kernel(OrbArgs a, int* __restrict__ done) {
    const OrbPlan& P = *a.plan;
    const int f = blockIdx.y;
    int t = blockIdx.x, level = 1, gx = 1, tiles = 0;
    for (;; level++) {
        gx = (P.lv[level].w + 127) >> 7;
        tiles = gx * ((P.lv[level].h + 4 * EORB_PYR_CHAIN_BAND - 1) / (4 * EORB_PYR_CHAIN_BAND));
        if (t < tiles || level + 1 >= P.nlevels) break;
        t -= tiles;
    }
    if (t >= tiles) return;
    // One counter per (level, block row): a warp of level L waits only for the block rows of level L - 1 that hold its source rows
    // (two or three of them), so the levels pipeline instead of running one after the other.  Source blocks always have lower block
    // indices than their readers, so a waiting block never waits for one that cannot have been scheduled.
    int* flags = done + (size_t)f * EORB_MAX_LEVELS * EORB_PYR_CHAIN_ROWS;
    const int by = t / gx;
    int need = 0, br0 = 0, br1 = -1;
    if (level > 1) {
        need = (P.lv[level - 1].w + 127) >> 7;
        const int dh = P.lv[level].h;
        const int y0 = (by * 4 + (int)threadIdx.y) * EORB_PYR_CHAIN_BAND;
        if (y0 < dh) {
            const int y1 = min(y0 + EORB_PYR_CHAIN_BAND, dh);
            const int4* __restrict__ ytab = a.ytab + P.lv[level].ytabOff;
            const int syA = __ldg(&ytab[y0]).x, syB = __ldg(&ytab[y1 - 1]).y;      // first / last source row (clamped into the level)
            const int lastRow = (P.lv[level - 1].h + 4 * EORB_PYR_CHAIN_BAND - 1) / (4 * EORB_PYR_CHAIN_BAND) - 1;
            br0 = min(max(syA / (4 * EORB_PYR_CHAIN_BAND), 0), lastRow);
            br1 = min(max(syB / (4 * EORB_PYR_CHAIN_BAND), 0), lastRow);
        }
    }
    const volatile int* fl = flags + (size_t)(level - 1) * EORB_PYR_CHAIN_ROWS;
    auto ready = [&]() {   // per warp: lane 0 polls the counters of its source block rows (acquire), the warp follows
        if (need > 0) {
            if (threadIdx.x == 0)
                for (int b = br0; b <= br1; b++) {
                    unsigned spins = 0;
                    while (fl[b] < need && ++spins < (1u << 24)) __nanosleep(20);   // bounded: a wrong result fails a test, a hang costs the box
                }
            __syncwarp();
            __threadfence();
        }
    };
    pyr_resize_block<EORB_PYR_CHAIN_BAND>(a, level, t % gx, by, f, ready);
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        __threadfence();
        atomicAdd(flags + (size_t)level * EORB_PYR_CHAIN_ROWS + by, 1);
    }
}

static int pyr_chain_tiles(const OrbPlan& hp) {
    int t = 0;
    for (int l = 1; l < hp.nlevels; l++) t += ((hp.lv[l].w + 127) / 128) * ((hp.lv[l].h + 4 * EORB_PYR_CHAIN_BAND - 1) / (4 * EORB_PYR_CHAIN_BAND));
    return t;
}

// ------------------------------------------------------------------------------------------------ K3
// One block per (level, frame).  Gathers the level's per-cell candidate lists into the reference's order
// (cell-row-major, pixel-row-major inside a cell), then runs the shared level-synchronous distribution.
// NT threads per block: 128 for launch sets (many blocks hide each other's barriers), NT for small batches, where the
// level-0 block of a single frame is the critical path and its passes over ~2000 keys are NT-wide.
template <int NT>
__global__ void __launch_bounds__(NT) octree_kernel(OrbArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_arr[NT];
    __shared__ int s_scr[NT];
    const OrbPlan& P = *a.plan;
    const int level = blockIdx.x, f = blockIdx.y;
    const LevelPlan& lp = P.lv[level];
    const int tid = threadIdx.x;
    uint32_t* okeys = a.okeys + (size_t)f * P.slotsPerFrame + lp.slotBase;
    uint16_t* knode = a.knode + (size_t)f * P.slotsPerFrame + lp.slotBase;
    const uint16_t* counts = a.cellCount + (size_t)f * P.nCells + lp.cellBase;
    const uint32_t* cand = a.cand + (size_t)f * P.slotsPerFrame;
    int n = 0;
    for (int c0 = 0; c0 < lp.nCells; c0 += NT) {
        const int ci = c0 + tid;
        const int cnt = (ci < lp.nCells) ? (int)counts[ci] : 0;
        s_arr[tid] = cnt;
        __syncthreads();
        const int tot = oct_exclusive_scan<NT>(s_arr, NT, s_scr);
        if (cnt > 0) {
            const uint32_t* src = cand + a.cells[lp.cellBase + ci].slotOff;
            uint32_t* dst = okeys + n + s_arr[tid];
            for (int j = 0; j < cnt; j += 4) {       // loads first (a cell holds ~5 candidates; the copies wait on L2)
                uint32_t v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = j + u < cnt ? src[j + u] : 0u;
#pragma unroll
                for (int u = 0; u < 4; u++) if (j + u < cnt) dst[j + u] = v[u];
            }
        }
        n += tot;
        __syncthreads();
    }
    __syncthreads();
    uint32_t* out = a.sel + (size_t)f * P.selPerFrame + lp.selBase;
    int L = 0;
    if (lp.maxBX > lp.minBX && lp.maxBY > lp.minBY)
        L = oct_distribute<NT>(okeys, knode, n, lp.maxBX - lp.minBX, lp.maxBY - lp.minBY, lp.nIni, lp.hX, lp.quota,
                           lp.nodeCap, smem_raw, out);
    if (tid == 0) {
        a.selCount[(size_t)f * P.nlevels + level] = L < 0 ? 0 : L;
        a.candCount[(size_t)f * P.nlevels + level] = n;
    }
}

// ------------------------------------------------------------------------------------------------ K7
// One block per frame: final position of every selected keypoint.  The reference walks levels in order and
// keypoints in octree-list order, writing keypoints inside the lapping area from the END of the output and
// the others from the front (:1152-1172); positions are prefix counts of the "inside" flag.
template <int NT>
__global__ void __launch_bounds__(NT) orb_index_kernel(OrbArgs a) {
    __shared__ int s_arr[NT];
    __shared__ int s_scr[NT];
    __shared__ int s_start[EORB_MAX_LEVELS + 1];
    const OrbPlan& P = *a.plan;
    const int f = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        int acc = 0;
        for (int l = 0; l < P.nlevels; l++) { s_start[l] = acc; acc += a.selCount[(size_t)f * P.nlevels + l]; }
        s_start[P.nlevels] = acc;
    }
    __syncthreads();
    const int nk = s_start[P.nlevels];
    const uint32_t* sel = a.sel + (size_t)f * P.selPerFrame;
    int* dstIdx = a.dstIdx + (size_t)f * P.selPerFrame;
    const float lap0 = (float)a.lap0, lap1 = (float)a.lap1;
    int stereoRun = 0;
    for (int g0 = 0; g0 < nk; g0 += NT) {
        const int g = g0 + tid;
        int flag = 0, slot = 0;
        if (g < nk) {
            int l = 0;
            while (g >= s_start[l + 1]) l++;
            const LevelPlan& lp = P.lv[l];
            slot = lp.selBase + (g - s_start[l]);
            float x = (float)(oct_key_x(sel[slot]) + lp.minBX);
            if (l != 0) x = fmul(x, lp.scale);
            flag = (x >= lap0 && x <= lap1) ? 1 : 0;
            a.kpList[(size_t)f * P.selPerFrame + g] = (uint32_t)slot | ((uint32_t)l << 24);   // work list of K4+K6
        }
        s_arr[tid] = flag;
        __syncthreads();
        const int tot = oct_exclusive_scan<NT>(s_arr, NT, s_scr);
        if (g < nk) {
            const int stereoBefore = stereoRun + s_arr[tid];
            dstIdx[slot] = flag ? (nk - 1 - stereoBefore) : (g - stereoBefore);
        }
        stereoRun += tot;
        __syncthreads();
    }
    if (tid == 0) { a.outN[f] = nk; a.outMono[f] = nk - stereoRun; }
}

// ------------------------------------------------------------------------------------------------ K5
// 5x5 sigma-2 Gaussian on every level (8.8 fixed-point separable weights [39 57 64 57 39], REFLECT_101).
// One warp owns a strip of 128 columns x EORB_BLUR_BAND rows and streams down the rows.  Per source row a lane
// loads ONE aligned 32-bit word (4 pixels) and gets its neighbours' words by shuffle; the five byte windows the
// horizontal pass needs are cut out with byte-permutes whose selectors are per-lane constants computed once:
// they already contain the REFLECT_101 mapping of the image's left/right border (and of a partial last word), so
// the row loop is straight-line code with no edge branches.  Horizontal sums: unsigned dp4a on packed weights
// (the 5th tap by a second dp4a with a one-hot weight word).  A 5-row ring of horizontal sums lives in registers
// (loop unrolled by 5, no moves).  Horizontal sums are < 2^16 and the vertical sum < 2^24, so byte 2 of the
// accumulator IS the result ((s + 32768) >> 16).  (The vertical pass works on packed row pairs, see the kernel.)
// EORB_BLUR_BAND (orb_plan.h) is a multiple of 4: full bands run the unrolled ring loop without a tail

// REFLECT_101 of a row index that overshoots [0, len) by at most 2 (single bounce when len >= 3)
__device__ __forceinline__ int reflect101_near(int p, int len) {
    if (len < 3) return reflect101(p, len);
    p = p < 0 ? -p : p;
    return p >= len ? 2 * len - 2 - p : p;
}

struct BlurLane {
    unsigned selQ0, selQ1, selW, selQ3, selR;   // byte-permute selectors (see blur_hrow)
    bool loadL, loadR;                          // lane 0 / lane 31 fetch their outer neighbour word themselves
};

// window byte index of image column x (after REFLECT_101) relative to column `base`; clamped into the 8-byte window
template <bool TINY>
__device__ __forceinline__ unsigned blur_sel(int x, int w, int base) {
    int r;
    if (TINY) {
        r = reflect101(x, w);
    } else {   // w >= 8 and -2 <= x <= w + 132: a single bounce
        r = x < 0 ? -x : x;
        r = r >= w ? 2 * w - 2 - r : r;
    }
    return (unsigned)min(max(r - base, 0), 7);
}

template <bool TINY>
__device__ __forceinline__ void blur_selectors(BlurLane& bl, int x0, int w) {
    const int b = x0 - 4;
    unsigned s[6];   // columns x0-2 .. x0+3 in the (L,W) window
#pragma unroll
    for (int j = 0; j < 6; j++) s[j] = blur_sel<TINY>(x0 - 2 + j, w, b);
    unsigned t[5];   // columns x0+1 .. x0+5 in the (W,R) window
#pragma unroll
    for (int j = 0; j < 5; j++) t[j] = blur_sel<TINY>(x0 + 1 + j, w, x0);
    bl.selQ0 = s[0] | (s[1] << 4) | (s[2] << 8) | (s[3] << 12);
    bl.selQ1 = s[1] | (s[2] << 4) | (s[3] << 8) | (s[4] << 12);
    bl.selW = s[2] | (s[3] << 4) | (s[4] << 8) | (s[5] << 12);
    bl.selQ3 = t[0] | (t[1] << 4) | (t[2] << 8) | (t[3] << 12);
    bl.selR = t[3] | (t[4] << 4);
}

// W = the lane's own word of the source row, X = the outer neighbour word lane 0 / lane 31 fetched themselves
__device__ __forceinline__ void blur_hrow(unsigned W, unsigned X, const BlurLane& bl, unsigned* h) {
    unsigned L = __shfl_up_sync(0xffffffffu, W, 1);
    unsigned R = __shfl_down_sync(0xffffffffu, W, 1);
    if (bl.loadL) L = X;
    if (bl.loadR) R = X;
    const unsigned WT = 0x39403927u;                   // bytes (39, 57, 64, 57)
    const unsigned q0 = __byte_perm(L, W, bl.selQ0);   // x0-2 .. x0+1
    const unsigned q1 = __byte_perm(L, W, bl.selQ1);   // x0-1 .. x0+2
    const unsigned Wc = __byte_perm(L, W, bl.selW);    // x0   .. x0+3
    const unsigned q3 = __byte_perm(W, R, bl.selQ3);   // x0+1 .. x0+4
    const unsigned Rc = __byte_perm(W, R, bl.selR);    // x0+4, x0+5 in bytes 0, 1
    h[0] = __dp4a(Wc, 0x00270000u, __dp4a(q0, WT, 0u));   // + 39 * p[x0+2]
    h[1] = __dp4a(Wc, 0x27000000u, __dp4a(q1, WT, 0u));   // + 39 * p[x0+3]
    h[2] = __dp4a(Rc, 0x00000027u, __dp4a(Wc, WT, 0u));   // + 39 * p[x0+4]
    h[3] = __dp4a(Rc, 0x00002700u, __dp4a(q3, WT, 0u));   // + 39 * p[x0+5]
}

#ifndef EORB_BLUR_GROUP8
#define EORB_BLUR_GROUP8 1   // measured: 0.748 -> 0.707 us/frame (64 registers); 80 registers: 0.724
#endif
#ifndef EORB_BLUR_MINB
#define EORB_BLUR_MINB 8
#endif
__global__ void __launch_bounds__(128, EORB_BLUR_MINB) blur_kernel(OrbArgs a) {
    const OrbPlan& P = *a.plan;
    const int f = blockIdx.y;
    const int lane = threadIdx.x;
    const int task = blockIdx.x * blockDim.y + threadIdx.y;
    if (task >= P.blurTasksTotal) return;
    int level = 0;
    while (level + 1 < P.nlevels && task >= P.lv[level + 1].blurTaskBase) level++;
    const LevelPlan& lp = P.lv[level];
    const int w = lp.w, hgt = lp.h, bp = lp.bpitch;
    const int strips = (w + 127) >> 7;
    const int t = task - lp.blurTaskBase;
    const int band = t / strips, strip = t - band * strips;
    const int x0 = strip * 128 + lane * 4;
    const int y0 = band * EORB_BLUR_BAND;
    const int y1 = min(y0 + EORB_BLUR_BAND, hgt);
    int sp;
    const uint8_t* __restrict__ src = level_ptr(a, lp, level, f, sp);
    uint8_t* __restrict__ dst = a.blur + (size_t)f * (size_t)P.blurBytesPerFrame + (size_t)lp.blurOff;

    if (hgt < 3) {   // degenerate level: rows bounce more than once; plain per-pixel evaluation
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < min(x0 + 4, w); x++) {
                int hs[5];
#pragma unroll
                for (int r = 0; r < 5; r++) {
                    const uint8_t* row = src + (size_t)reflect101(y + r - 2, hgt) * sp;
                    hs[r] = gauss5_h(row[reflect101(x - 2, w)], row[reflect101(x - 1, w)], row[x], row[reflect101(x + 1, w)], row[reflect101(x + 2, w)]);
                }
                dst[(size_t)y * bp + x] = (uint8_t)gauss5_v(hs[0], hs[1], hs[2], hs[3], hs[4]);
            }
        return;
    }

    // per-lane constants: own word offset and the REFLECT_101-aware selectors (windows: (L,W) starts at column
    // x0-4, (W,R) at column x0); idle lanes (x0 >= w) read the row's last word and produce values nobody stores
    const int lastWord = ((w + 3) & ~3) - 4;
    const int off = min(x0, lastWord);
    const bool store = x0 < w;
    BlurLane bl;
    {
        bl.loadL = (lane == 0) && off >= 4;
        bl.loadR = (lane == 31) && off + 4 <= lastWord;
        if (w < 8) blur_selectors<true>(bl, x0, w);
        else blur_selectors<false>(bl, x0, w);
    }

    // the source pointer walks the REFLECT_101 row sequence: from row index yy-1 to yy it moves one row DOWN when
    // 1 <= yy <= hgt-1 and one row UP otherwise (single bounce: hgt >= 3 and the overshoot is at most 2 rows)
    const uint8_t* rp = src + (size_t)reflect101(y0 - 2, hgt) * sp + off;
    uint8_t* dp = dst + (size_t)y0 * bp + off;
    int yy = y0 - 2;                       // row index rp currently stands for
    const long long spl = sp;
#define NEXTROW() do { yy++; rp += ((unsigned)(yy - 1) < (unsigned)(hgt - 1)) ? spl : -spl; } while (0)
    // loads are split from the arithmetic: the four words of a group of rows are requested before the first row is
    // combined (0.85 -> 0.75 us/frame; requesting the NEXT group too was measured neutral at 72 registers, slower at 64)
    const bool edgeLane = bl.loadL || bl.loadR;
    const int eoff = bl.loadL ? -4 : 4;
#define LOADROW(Wv, Xv) do { Wv = __ldg(reinterpret_cast<const unsigned*>(rp)); Xv = edgeLane ? __ldg(reinterpret_cast<const unsigned*>(rp + eoff)) : 0u; NEXTROW(); } while (0)
#define HROW(dstv) do { unsigned w_, x_; LOADROW(w_, x_); blur_hrow(w_, x_, bl, dstv); } while (0)
    // Vertical pass on PAIRS of rows: horizontal sums are < 2^16, so two vertically adjacent sums share one register
    // (lo = upper row) and a 5-tap column costs  dp2a(P(y-2,y-1), {39,57}) + dp2a(P(y,y+1), {64,57}) + 39 * h(y+2):
    // one pack + two IDP.2A + one IMAD per pixel.  Every pair P(k,k+1) is packed once and used twice; the ring holds four
    // pairs and the newest raw row, and returns to its starting assignment every 4 rows (no register moves).
#define PACK(P, E, L)                                                                               \
    do {                                                                                            \
        _Pragma("unroll") for (int j_ = 0; j_ < 4; j_++) P[j_] = __byte_perm(E[j_], L[j_], 0x5410);   \
    } while (0)
#define OUT(P1, P2, HN)                                                                             \
    do {                                                                                            \
        unsigned acc_[4];                                                                           \
        _Pragma("unroll") for (int j_ = 0; j_ < 4; j_++)                                              \
            acc_[j_] = 39u * HN[j_] + __dp2a_lo(P2[j_], 0x00003940u, __dp2a_lo(P1[j_], 0x00003927u, 32768u)); \
        const unsigned lo_ = __byte_perm(acc_[0], acc_[1], 0x0062);                                 \
        const unsigned hi_ = __byte_perm(acc_[2], acc_[3], 0x0062);                                 \
        const unsigned o_ = __byte_perm(lo_, hi_, 0x5410);                                          \
        if (store) *reinterpret_cast<unsigned*>(dp) = o_;                                           \
        dp += bp;                                                                                   \
    } while (0)
    unsigned pa[4], pb[4], pc[4], pd[4], hx[4], hy[4];
    {
        unsigned h0[4], h1[4];
        HROW(h0); HROW(h1); PACK(pa, h0, h1);     // P(y0-2, y0-1)
        HROW(h0); PACK(pb, h1, h0);               // P(y0-1, y0)
        HROW(hy); PACK(pc, h0, hy);               // P(y0,   y0+1); hy = newest raw row
    }
    int y = y0;
#if EORB_BLUR_GROUP8
    for (; y + 8 <= y1; y += 8) {          // eight rows in flight per warp
        unsigned w0, w1, w2, w3, w4, w5, w6, w7, x0_, x1_, x2_, x3_, x4_, x5_, x6_, x7_;
        LOADROW(w0, x0_); LOADROW(w1, x1_); LOADROW(w2, x2_); LOADROW(w3, x3_);
        LOADROW(w4, x4_); LOADROW(w5, x5_); LOADROW(w6, x6_); LOADROW(w7, x7_);
        blur_hrow(w0, x0_, bl, hx); PACK(pd, hy, hx); OUT(pa, pc, hx);
        blur_hrow(w1, x1_, bl, hy); PACK(pa, hx, hy); OUT(pb, pd, hy);
        blur_hrow(w2, x2_, bl, hx); PACK(pb, hy, hx); OUT(pc, pa, hx);
        blur_hrow(w3, x3_, bl, hy); PACK(pc, hx, hy); OUT(pd, pb, hy);
        blur_hrow(w4, x4_, bl, hx); PACK(pd, hy, hx); OUT(pa, pc, hx);
        blur_hrow(w5, x5_, bl, hy); PACK(pa, hx, hy); OUT(pb, pd, hy);
        blur_hrow(w6, x6_, bl, hx); PACK(pb, hy, hx); OUT(pc, pa, hx);
        blur_hrow(w7, x7_, bl, hy); PACK(pc, hx, hy); OUT(pd, pb, hy);
    }
#endif
    for (; y + 4 <= y1; y += 4) {
        unsigned w0, w1, w2, w3, x0_, x1_, x2_, x3_;
        LOADROW(w0, x0_); LOADROW(w1, x1_); LOADROW(w2, x2_); LOADROW(w3, x3_);
        blur_hrow(w0, x0_, bl, hx); PACK(pd, hy, hx); OUT(pa, pc, hx);
        blur_hrow(w1, x1_, bl, hy); PACK(pa, hx, hy); OUT(pb, pd, hy);
        blur_hrow(w2, x2_, bl, hx); PACK(pb, hy, hx); OUT(pc, pa, hx);
        blur_hrow(w3, x3_, bl, hy); PACK(pc, hx, hy); OUT(pd, pb, hy);
    }
    if (y < y1) {                          // last band of a level: up to 3 rows left
        HROW(hx); PACK(pd, hy, hx); OUT(pa, pc, hx);
        if (y + 1 < y1) {
            HROW(hy); PACK(pa, hx, hy); OUT(pb, pd, hy);
            if (y + 2 < y1) { HROW(hx); OUT(pc, pa, hx); }
        }
    }
#undef NEXTROW
#undef LOADROW
#undef HROW
#undef OUT
#undef PACK
}

// ------------------------------------------------------------------------------------------------ K5t
// The same blur with the strip's source rows staged by TMA (A/B of the design north_star names; EORB_BLUR_TMA=1 selects it, see
// DESIGN.md for the measurement).  One warp = one block owns 128 columns x EORB_BLUR_TBAND rows: ONE cp.async.bulk.tensor tile of
// 160 columns (the strip, 16 bytes either side: TMA's inner coordinate must be a multiple of 16 bytes) x (band + 4) rows of the
// level's {x, y, frame} map, zero fill outside the level.  REFLECT_101 needs no extra data: mirrored rows lie inside the tile (they
// are rows 1, 2 / h-2, h-3 of the level) and are reached by index; mirrored columns are handled by the per-lane selectors exactly
// as in blur_kernel, which never read a byte outside the level.  The arithmetic (blur_hrow, row-pair ring) is blur_kernel's.
#define EORB_BLUR_TBOXW 160
// W = own word, L / R = the neighbour words read from the tile (instead of two shuffles and the edge lanes' conditional loads)
__device__ __forceinline__ void blur_hrow_lds(unsigned L, unsigned W, unsigned R, const BlurLane& bl, unsigned* h) {
    const unsigned WT = 0x39403927u;
    const unsigned q0 = __byte_perm(L, W, bl.selQ0);
    const unsigned q1 = __byte_perm(L, W, bl.selQ1);
    const unsigned Wc = __byte_perm(L, W, bl.selW);
    const unsigned q3 = __byte_perm(W, R, bl.selQ3);
    const unsigned Rc = __byte_perm(W, R, bl.selR);
    h[0] = __dp4a(Wc, 0x00270000u, __dp4a(q0, WT, 0u));
    h[1] = __dp4a(Wc, 0x27000000u, __dp4a(q1, WT, 0u));
    h[2] = __dp4a(Rc, 0x00000027u, __dp4a(Wc, WT, 0u));
    h[3] = __dp4a(Rc, 0x00002700u, __dp4a(q3, WT, 0u));
}
template <int EORB_BLUR_TBAND, bool LDSNB>
__global__ void __launch_bounds__(32) blur_tma_kernel(OrbArgs a, const __grid_constant__ CUtensorMap tmL0, int tasksTotal) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const OrbPlan& P = *a.plan;
    const int f = blockIdx.y, lane = threadIdx.x;
    int task = blockIdx.x;
    if (task >= tasksTotal) return;
    int level = 0, strips = 0;
    for (;; level++) {          // tasks of a level: ceil(h / TBAND) bands x ceil(w / 128) strips, levels in order
        strips = (P.lv[level].w + 127) >> 7;
        const int nt = ((P.lv[level].h + EORB_BLUR_TBAND - 1) / EORB_BLUR_TBAND) * strips;
        if (task < nt || level + 1 >= P.nlevels) break;
        task -= nt;
    }
    const LevelPlan& lp = P.lv[level];
    const int w = lp.w, hgt = lp.h, bp = lp.bpitch;
    const int band = task / strips, strip = task - band * strips;
    const int x0t = strip * 128, x0 = x0t + lane * 4;
    const int y0 = band * EORB_BLUR_TBAND;
    const int y1 = min(y0 + EORB_BLUR_TBAND, hgt);
    const unsigned bar = smem_u32(smem_raw + (EORB_BLUR_TBAND + 4) * EORB_BLUR_TBOXW);
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, (unsigned)((EORB_BLUR_TBAND + 4) * EORB_BLUR_TBOXW));
        const CUtensorMap* tm = level == 0 ? &tmL0 : a.blurMaps + level;
        tma_load_3d(smem_u32(smem_raw), tm, x0t - 16, y0 - 2, f, bar);
    }
    uint8_t* __restrict__ dst = a.blur + (size_t)f * (size_t)P.blurBytesPerFrame + (size_t)lp.blurOff;
    const int lastWord = ((w + 3) & ~3) - 4;
    const int off = min(x0, lastWord);
    const bool store = x0 < w;
    BlurLane bl;
    bl.loadL = (lane == 0) && off >= 4;
    bl.loadR = (lane == 31) && off + 4 <= lastWord;
    if (w < 8) blur_selectors<true>(bl, x0, w);
    else blur_selectors<false>(bl, x0, w);
    // tile row of level row r = r - (y0 - 2); tile byte of level column c = c - (x0t - 16)
    const unsigned char* tcol = smem_raw + (off - x0t + 16) - (y0 - 2) * EORB_BLUR_TBOXW;
    const bool edgeLane = bl.loadL || bl.loadR;
    const int eoff = bl.loadL ? -4 : 4;
    uint8_t* dp = dst + (size_t)y0 * bp + off;
    int yy = y0 - 2;
    __syncwarp();
    mbar_wait(bar, 0);
#define TLOADROW(Wv, Xv) do { const unsigned char* rp_ = tcol + reflect101_near(yy, hgt) * EORB_BLUR_TBOXW; \
        Wv = *reinterpret_cast<const unsigned*>(rp_); Xv = edgeLane ? *reinterpret_cast<const unsigned*>(rp_ + eoff) : 0u; yy++; } while (0)
#define THROW(dstv) do { if (LDSNB) { const unsigned char* rp_ = tcol + reflect101_near(yy, hgt) * EORB_BLUR_TBOXW; yy++;              \
            blur_hrow_lds(*reinterpret_cast<const unsigned*>(rp_ - 4), *reinterpret_cast<const unsigned*>(rp_),                      \
                          *reinterpret_cast<const unsigned*>(rp_ + 4), bl, dstv); }                                                   \
        else { unsigned w_, x_; TLOADROW(w_, x_); blur_hrow(w_, x_, bl, dstv); } } while (0)
#define PACK(P, E, L)                                                                               \
    do {                                                                                            \
        _Pragma("unroll") for (int j_ = 0; j_ < 4; j_++) P[j_] = __byte_perm(E[j_], L[j_], 0x5410);   \
    } while (0)
#define OUT(P1, P2, HN)                                                                             \
    do {                                                                                            \
        unsigned acc_[4];                                                                           \
        _Pragma("unroll") for (int j_ = 0; j_ < 4; j_++)                                              \
            acc_[j_] = 39u * HN[j_] + __dp2a_lo(P2[j_], 0x00003940u, __dp2a_lo(P1[j_], 0x00003927u, 32768u)); \
        const unsigned lo_ = __byte_perm(acc_[0], acc_[1], 0x0062);                                 \
        const unsigned hi_ = __byte_perm(acc_[2], acc_[3], 0x0062);                                 \
        const unsigned o_ = __byte_perm(lo_, hi_, 0x5410);                                          \
        if (store) *reinterpret_cast<unsigned*>(dp) = o_;                                           \
        dp += bp;                                                                                   \
    } while (0)
    unsigned pa[4], pb[4], pc[4], pd[4], hx[4], hy[4];
    {
        unsigned h0[4], h1[4];
        THROW(h0); THROW(h1); PACK(pa, h0, h1);
        THROW(h0); PACK(pb, h1, h0);
        THROW(hy); PACK(pc, h0, hy);
    }
    int y = y0;
    for (; y + 4 <= y1; y += 4) {
        THROW(hx); PACK(pd, hy, hx); OUT(pa, pc, hx);
        THROW(hy); PACK(pa, hx, hy); OUT(pb, pd, hy);
        THROW(hx); PACK(pb, hy, hx); OUT(pc, pa, hx);
        THROW(hy); PACK(pc, hx, hy); OUT(pd, pb, hy);
    }
    if (y < y1) {
        THROW(hx); PACK(pd, hy, hx); OUT(pa, pc, hx);
        if (y + 1 < y1) {
            THROW(hy); PACK(pa, hx, hy); OUT(pb, pd, hy);
            if (y + 2 < y1) { THROW(hx); OUT(pc, pa, hx); }
        }
    }
#undef TLOADROW
#undef THROW
#undef OUT
#undef PACK
}

static int blur_tma_band(int variant) { return (variant == 2 || variant == 3) ? 64 : 32; }
static int blur_tma_tasks(const OrbPlan& hp, int band) {
    int t = 0;
    for (int l = 0; l < hp.nlevels; l++) t += ((hp.lv[l].h + band - 1) / band) * ((hp.lv[l].w + 127) >> 7);
    return t;
}
int blur_tma_box_w() { return EORB_BLUR_TBOXW; }
int blur_tma_box_h(int variant) { return blur_tma_band(variant) + 4; }

// ------------------------------------------------------------------------------------------------ K4 + K6
// One warp per selected keypoint (taken from the compact list K7 writes), EORB_KP_GROUP keypoints per block.
// Orientation (IC_Angle): the 31x31 patch is read as aligned 32-bit words, 9 words per row; lane <-> (row, word)
// tasks, 9 steps per keypoint.  The circular mask and the u / v moment weights of every (alignment, row, word)
// are a table of packed s8x4 pairs, so a task is one word load, one table load and two mixed-sign dp4a
// (m10 += sum u*I, m01 += sum v*I); integer moments reduced with shuffles, cv::fastAtan2 restated without FMA.
// sin/cos: the block's angles meet in shared memory and threads 0..G-1 evaluate them together, one pass of the
// double-precision sincos per G keypoints instead of one per keypoint.  Descriptor: lane L evaluates pattern pair 32*j+L in round j;
// __ballot_sync yields descriptor word j directly (bit k of byte i = pair 8i+k).
#define EORB_KP_GROUP 4
#define EORB_IC_WORDS 9                    // aligned words covering 31 columns at any alignment
#define EORB_IC_TASKS 288                  // 32 rows x 9 words (row 31 is padding with zero weights)

__device__ __forceinline__ int dp4a_u8_s8(unsigned data, int weights, int acc) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(data), "r"(weights), "r"(acc));
    return d;
}

#ifndef EORB_OD_MINB
#define EORB_OD_MINB 11
#endif
#ifndef EORB_OD_BATCH
#define EORB_OD_BATCH 4
#endif

// Both patches of a keypoint are staged by TMA (L2 -> shared memory past the L1 / LSU path; ncu: the kernel was bound by L1/TEX data-pipe
// wavefronts, 87 %, not by issue):
//   BRIEF  the rotated pattern reaches 19 pixels from the keypoint (safe path); the box starts at the 16-byte aligned column at or
//          before x - 19 (TMA's inner coordinate must be a multiple of 16 bytes), so 39 rows x (39 + 15 <= 64) columns cover it.  The
//          box is 80 bytes wide: a row pitch of 20 words spreads the 32 scattered byte gathers of one instruction over the banks
//          (with 64 bytes only (row & 1) and the column select the bank).
//   IC     31 rows x 48 columns of the level itself from the 16-byte aligned column at or before x - 15: the nine aligned words of
//          every row lie inside (offset <= 12, 12 + 36 = 48); the padding task row 31 reads the tile's own 48 bytes of slack.
#ifndef EORB_BRIEF_BOXW
#define EORB_BRIEF_BOXW 80
#endif
#define EORB_BRIEF_BOXH 39
#define EORB_BRIEF_TILE ((EORB_BRIEF_BOXW * EORB_BRIEF_BOXH + 127) / 128 * 128)   // TMA destination alignment
#define EORB_IC_BOXW 48
#define EORB_IC_BOXH 31
#define EORB_IC_TILE 1536                  // 48 * 31 = 1488 + one more row for the padding tasks
int brief_tma_box_w() { return EORB_BRIEF_BOXW; }
int brief_tma_box_h() { return EORB_BRIEF_BOXH; }
int ic_tma_box_w() { return EORB_IC_BOXW; }
int ic_tma_box_h() { return EORB_IC_BOXH; }

__global__ void __launch_bounds__(32 * EORB_KP_GROUP, EORB_OD_MINB) orient_desc_kernel(OrbArgs a, const __grid_constant__ CUtensorMap icMap0,
                                                                                      const __grid_constant__ OdLevels L) {
    __shared__ __align__(128) uint8_t s_tile[EORB_KP_GROUP][EORB_BRIEF_TILE];
    __shared__ __align__(128) uint8_t s_ic[EORB_KP_GROUP][EORB_IC_TILE];
    __shared__ __align__(8) unsigned long long s_bar[EORB_KP_GROUP][2];
    __shared__ float s_angle[EORB_KP_GROUP], s_cos[EORB_KP_GROUP], s_sin[EORB_KP_GROUP];
    const OrbPlan& P = *a.plan;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.y;
    const int nk = min(a.outN[f], L.selPerFrame);
    const int kidx = blockIdx.x * EORB_KP_GROUP + warp;
    if (blockIdx.x * EORB_KP_GROUP >= nk) return;      // block-uniform
    const bool valid = kidx < nk;
    const unsigned FULL = 0xffffffffu;

    // per-level constants come with the kernel parameters (constant bank), not from the plan in global memory
    int slot = 0, level = 0, x = 0, y = 0, lw = 0, lh = 0;
    uint32_t key = 0;
    if (valid) {
        const uint32_t e = a.kpList[(size_t)f * L.selPerFrame + kidx];
        slot = e & 0xffffff; level = e >> 24;
        key = a.sel[(size_t)f * L.selPerFrame + slot];
        const int4 g = L.geo[level];
        x = oct_key_x(key) + g.x; y = oct_key_y(key) + g.y; lw = g.z; lh = g.w;
    }

    // ---- the keypoint's two patches start their way into shared memory now
    int dst = -1;
    bool staged = false, stagedIc = false;
    const bool inside = (x >= 15) && (y >= 15) && (x + 15 < lw) && (y + 15 < lh);
    const bool safe = (x >= 19) && (y >= 19) && (x + 19 < lw) && (y + 19 < lh);
    if (valid) {
        dst = a.dstIdx[(size_t)f * L.selPerFrame + slot];
        staged = a.briefMaps != nullptr && a.wantDesc && safe && dst >= 0 && dst < a.cap;
        stagedIc = a.icMaps != nullptr && inside;
        if (lane == 0 && (staged || stagedIc)) {
            const unsigned bar0 = smem_u32(&s_bar[warp][0]), bar1 = smem_u32(&s_bar[warp][1]);
            mbar_init(bar0, 1);
            mbar_init(bar1, 1);
            mbar_fence_init();
            if (stagedIc) {
                mbar_expect_tx(bar0, EORB_IC_BOXW * EORB_IC_BOXH);
                tma_load_3d(smem_u32(s_ic[warp]), level == 0 ? &icMap0 : a.icMaps + level, (x - 15) & ~15, y - 15, f, bar0);
            }
            if (staged) {
                mbar_expect_tx(bar1, EORB_BRIEF_BOXW * EORB_BRIEF_BOXH);
                tma_load_3d(smem_u32(s_tile[warp]), a.briefMaps + level, (x - 19) & ~15, y - 19, f, bar1);
            }
        }
        __syncwarp();
    }
    const LevelPlan& lp = P.lv[level];   // only the paths that are not staged read it

    // ---- orientation (one warp per keypoint)
    if (valid) {
        int m10 = 0, m01 = 0;
        if (stagedIc) {
            // task i = it*32 + lane <-> (row r = i / 9, word k = i % 9) of the 31 x 9 aligned words covering the patch; the circular
            // mask and the u / v weights of every (alignment, task) come from the table
            const int xs = x - 15;
            const int2* __restrict__ tab = a.icTab + (xs & 3) * EORB_IC_TASKS + lane;
            const unsigned tbase = smem_u32(s_ic[warp]) + (unsigned)((xs & ~3) - (xs & ~15)) + 4u * (unsigned)lane;
            int2 t[EORB_IC_TASKS / 32];
#pragma unroll
            for (int it = 0; it < EORB_IC_TASKS / 32; it++) t[it] = __ldg(tab + it * 32);
            mbar_wait(smem_u32(&s_bar[warp][0]), 0);
#pragma unroll
            for (int it = 0; it < EORB_IC_TASKS / 32; it++) {
                const int r = ((lane + 32 * it) * 57) >> 9;                       // i / 9 for i < 512
                const unsigned data = lds_u32(tbase + 128 * it + (unsigned)(r * (EORB_IC_BOXW - 36)));   // r * 48 + 4 * (i - 9 r)
                m10 = dp4a_u8_s8(data, t[it].x, m10);
                m01 = dp4a_u8_s8(data, t[it].y, m01);
            }
        } else if (inside) {
            int sp;
            const uint8_t* img = level_ptr(a, lp, level, f, sp);
            const int xs = x - 15, xa = xs & ~3;
            const int wAligned = (lp.w + 3) & ~3;
            const int2* __restrict__ tab = a.icTab + (xs & 3) * EORB_IC_TASKS + lane;
            const uint8_t* base = img + (size_t)(y - 15) * sp + xa;
            // r and the word's byte offset are formed arithmetically from the lane (three integer instructions, independent of every
            // load: taking them from the table entry was measured SLOWER, 1.00 -> 1.13 us/frame, because the pixel load then waits for
            // the table load).  Words past the row's last aligned word and the padding row 31 only meet zero weights, so their
            // addresses are clamped instead of predicated.
            const int k4Max = wAligned - 4 - xa;
            const int lane4 = 4 * lane;
#pragma unroll
            for (int it = 0; it < EORB_IC_TASKS / 32; it++) {
                const int2 t = __ldg(tab + it * 32);
                int r = ((lane + 32 * it) * 57) >> 9;                  // i / 9 for i < 512
                const int k4 = lane4 + 128 * it - 36 * r;              // 4 * (i % 9)
                if (it == EORB_IC_TASKS / 32 - 1) r = min(r, 30);
                const unsigned off = (unsigned)(r * sp + min(k4, k4Max));
                const unsigned data = __ldg(reinterpret_cast<const unsigned*>(base + off));
                m10 = dp4a_u8_s8(data, t.x, m10);
                m01 = dp4a_u8_s8(data, t.y, m01);
            }
        } else if (lane < 31) {   // patch crosses the level border (margin < 15): REFLECT_101, byte by byte
            int sp;
            const uint8_t* img = level_ptr(a, lp, level, f, sp);
            const int u = lane - 15;
            const int au = u < 0 ? -u : u;
            const int xx = reflect101(x + u, lp.w);
            int colsum = 0;
            for (int v = -15; v <= 15; v++) {
                const int av = v < 0 ? -v : v;
                if (au <= P.umax[av]) {
                    const int val = __ldg(img + (size_t)reflect101(y + v, lp.h) * sp + xx);
                    colsum += val; m01 += v * val;
                }
            }
            m10 = u * colsum;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m10 += __shfl_xor_sync(FULL, m10, o);
            m01 += __shfl_xor_sync(FULL, m01, o);
        }
        if (lane == 0) s_angle[warp] = fast_atan2_deg((float)m01, (float)m10);
    } else if (lane == 0) {
        s_angle[warp] = 0.f;
    }
    __syncthreads();

    // ---- sin / cos of the block's angles in one pass: thread g <-> keypoint g (one double-precision sincos per
    //      EORB_KP_GROUP keypoints instead of one per keypoint)
    if (threadIdx.x < EORB_KP_GROUP) {
        const float factorPI = (float)(3.14159265358979323846 / 180.f);
        double sd, cd;
        sincos((double)fmul(s_angle[threadIdx.x], factorPI), &sd, &cd);
        s_cos[threadIdx.x] = (float)cd; s_sin[threadIdx.x] = (float)sd;
    }
    __syncthreads();
    if (!valid) return;

    // ---- keypoint record + steered BRIEF on the blurred level
    const float angle = s_angle[warp], ca = s_cos[warp], sa = s_sin[warp];
    if (a.levelAngle && lane == 0) a.levelAngle[(size_t)f * L.selPerFrame + slot] = angle;
    if (dst < 0 || dst >= a.cap) return;
    if (lane == 0) {
        const float2 sc = L.sc[level];
        eorb_keypoint kp;
        const float xf = (float)x, yf = (float)y;
        kp.x = level ? fmul(xf, sc.x) : xf;
        kp.y = level ? fmul(yf, sc.x) : yf;
        kp.size = sc.y;
        kp.angle = angle;
        kp.response = (float)oct_key_score(key);
        kp.octave = level;
        kp.class_id = -1;
        a.outKps[(size_t)f * a.cap + dst] = kp;
    }
    if (!a.wantDesc) return;
    uint32_t myword = 0;
    if (staged) {
        // sample address in the tile: row / column leave the 1.5 * 2^23 rounding trick as integers biased by 0x4B400000; the bias times
        // (tile pitch + 1), the patch origin (-19, -19) and the tile's alignment offset are ONE per-warp constant folded into the base
        const unsigned base = smem_u32(s_tile[warp]) + (unsigned)((x - 19) & 15) - (0x4B400000u - 19u) * (unsigned)(EORB_BRIEF_BOXW + 1);
        mbar_wait(smem_u32(&s_bar[warp][1]), 0);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float4 pt = __ldg(reinterpret_cast<const float4*>(d_brief_pattern_f) + 32 * j + lane);
            unsigned r0, c0, r1, c1;
            brief_offset_biased(pt.x, pt.y, ca, sa, r0, c0);
            brief_offset_biased(pt.z, pt.w, ca, sa, r1, c1);
            const unsigned t0 = lds_u8(r0 * EORB_BRIEF_BOXW + c0 + base), t1 = lds_u8(r1 * EORB_BRIEF_BOXW + c1 + base);
            const uint32_t word = __ballot_sync(FULL, t0 < t1);
            if (lane == j) myword = word;
        }
    } else if (safe) {
        // not staged (EORB_BRIEF_TMA=0 or no tensor map): the samples are gathered from global memory, EORB_OD_BATCH rounds requested
        // before the first comparison; same biased 32-bit offsets, from the patch's top-left corner
        const int bp = lp.bpitch;
        const uint8_t* Bc = a.blur + (size_t)f * (size_t)P.blurBytesPerFrame + (size_t)lp.blurOff + (size_t)y * bp + x;
        const uint8_t* Bo = Bc - 19 * bp - 19;
        const unsigned kBias = (0x4B400000u - 19u) * (unsigned)(bp + 1);
#pragma unroll
        for (int j0 = 0; j0 < 8; j0 += EORB_OD_BATCH) {
            unsigned o0[EORB_OD_BATCH], o1[EORB_OD_BATCH];
#pragma unroll
            for (int u = 0; u < EORB_OD_BATCH; u++) {
                const float4 pt = __ldg(reinterpret_cast<const float4*>(d_brief_pattern_f) + 32 * (j0 + u) + lane);
                unsigned r0, c0, r1, c1;
                brief_offset_biased(pt.x, pt.y, ca, sa, r0, c0);
                brief_offset_biased(pt.z, pt.w, ca, sa, r1, c1);
                o0[u] = r0 * (unsigned)bp + c0 - kBias; o1[u] = r1 * (unsigned)bp + c1 - kBias;
            }
            int t0[EORB_OD_BATCH], t1[EORB_OD_BATCH];
#pragma unroll
            for (int u = 0; u < EORB_OD_BATCH; u++) { t0[u] = __ldg(Bo + o0[u]); t1[u] = __ldg(Bo + o1[u]); }
#pragma unroll
            for (int u = 0; u < EORB_OD_BATCH; u++) {
                const uint32_t word = __ballot_sync(FULL, t0[u] < t1[u]);
                if (lane == j0 + u) myword = word;
            }
        }
    } else {   // margin < 19: the reference reads out of bounds; pinned to REFLECT_101 (see oracle)
        const int bp = lp.bpitch;
        const uint8_t* B = a.blur + (size_t)f * (size_t)P.blurBytesPerFrame + (size_t)lp.blurOff;
#pragma unroll 1
        for (int j = 0; j < 8; j++) {
            const float4 pt = __ldg(reinterpret_cast<const float4*>(d_brief_pattern_f) + 32 * j + lane);
            int r0, c0, r1, c1;
            brief_offset_f(pt.x, pt.y, ca, sa, r0, c0);
            brief_offset_f(pt.z, pt.w, ca, sa, r1, c1);
            const int t0 = __ldg(B + (size_t)reflect101(y + r0, lp.h) * bp + reflect101(x + c0, lp.w));
            const int t1 = __ldg(B + (size_t)reflect101(y + r1, lp.h) * bp + reflect101(x + c1, lp.w));
            const uint32_t word = __ballot_sync(FULL, t0 < t1);
            if (lane == j) myword = word;
        }
    }
    if (lane < 8) reinterpret_cast<uint32_t*>(a.outDesc + ((size_t)f * a.cap + dst) * 32)[lane] = myword;
}

// ------------------------------------------------------------------------------------------------ O10
// Descriptors for caller-supplied keypoints (ComputeTrackedKPtsDesc :1316-1363 and the inner loop of
// AssignKPtLevelByBestDesc :1267-1314): one warp per (keypoint, level); position scaled by
// mvInvScaleFactor[level], the keypoint's own angle is used.  mode 0: only kp.octave == level, writes desc;
// mode 1: every level, writes the Hamming distance to refDesc into dist[level*n + i].
__global__ void __launch_bounds__(256) tracked_desc_kernel(OrbArgs a, const eorb_keypoint* kps, int n, int mode,
                                                          const float* invScale, const uint8_t* refDesc,
                                                          uint8_t* descOut, int* distOut) {
    const OrbPlan& P = *a.plan;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    const int level = blockIdx.y;
    if (i >= n) return;
    const eorb_keypoint kp = kps[i];
    if (mode == 0 && kp.octave != level) return;
    const LevelPlan& lp = P.lv[level];
    const float sc = invScale[level];
    const int x = round_rne(fmul(kp.x, sc)), y = round_rne(fmul(kp.y, sc));
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float ang = fmul(kp.angle, factorPI);
    const float ca = (float)cos((double)ang), sa = (float)sin((double)ang);
    const uint8_t* B = a.blur + (size_t)lp.blurOff;
    const int bp = lp.bpitch;
    const unsigned FULL = 0xffffffffu;
    uint32_t myword = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const char4 pt = __ldg(reinterpret_cast<const char4*>(d_brief_pattern) + 32 * j + lane);
        int r0, c0, r1, c1;
        brief_offset(pt.x, pt.y, ca, sa, r0, c0);
        brief_offset(pt.z, pt.w, ca, sa, r1, c1);
        const int y0 = reflect101(y + r0, lp.h), x0 = reflect101(x + c0, lp.w);
        const int y1 = reflect101(y + r1, lp.h), x1 = reflect101(x + c1, lp.w);
        const int t0 = __ldg(B + (size_t)y0 * bp + x0);
        const int t1 = __ldg(B + (size_t)y1 * bp + x1);
        const uint32_t word = __ballot_sync(FULL, t0 < t1);
        if (lane == j) myword = word;
    }
    if (mode == 0) {
        if (lane < 8) reinterpret_cast<uint32_t*>(descOut + (size_t)i * 32)[lane] = myword;
    } else {
        int d = 0;
        if (lane < 8) d = __popc(myword ^ reinterpret_cast<const uint32_t*>(refDesc + (size_t)i * 32)[lane]);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) d += __shfl_xor_sync(FULL, d, o);
        if (lane == 0) distOut[(size_t)level * n + i] = d;
    }
}

// ------------------------------------------------------------------------------------------------ self-test
// Device evaluation of the shared scalar arithmetic on caller-supplied inputs; capi.cu compares it with the
// HOST compilation of the same header (guards against toolchain miscompiles such as the VIMNMX3 one).
__global__ void selftest_math_kernel(const int* __restrict__ fastIn, int nFast, int* __restrict__ fastOut,
                                     const float* __restrict__ atanIn, int nAtan, float* __restrict__ atanOut,
                                     int* __restrict__ briefOut) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nFast) {
        int ring[16];
#pragma unroll
        for (int k = 0; k < 16; k++) ring[k] = fastIn[17 * i + 1 + k];
        fastOut[i] = fast_max_arc_min_packed(fastIn[17 * i], ring);   // the device form used by K2
    }
    if (i < nAtan) {
        const float y = atanIn[2 * i], x = atanIn[2 * i + 1];
        const float ang = fast_atan2_deg(y, x);
        atanOut[i] = ang;
        const float rad = fmul(ang, (float)(3.14159265358979323846 / 180.f));
        const float ca = (float)cos((double)rad), sa = (float)sin((double)rad);
        int r, c;
        brief_offset_f((float)((i % 27) - 13), (float)(((i / 27) % 27) - 13), ca, sa, r, c);   // the device form used by K6
        briefOut[2 * i] = r; briefOut[2 * i + 1] = c;
    }
}

cudaError_t launch_selftest_math(const int* fastIn, int nFast, int* fastOut, const float* atanIn, int nAtan, float* atanOut,
                                 int* briefOut, cudaStream_t st) {
    const int n = nFast > nAtan ? nFast : nAtan;
    if (n <= 0) return cudaSuccess;
    selftest_math_kernel<<<(n + 255) / 256, 256, 0, st>>>(fastIn, nFast, fastOut, atanIn, nAtan, atanOut, briefOut);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ launches
static inline unsigned cdiv(unsigned a, unsigned b) { return (a + b - 1) / b; }

static cudaError_t launch_pyramid_level(const OrbArgs& a, const OrbPlan& hp, int l, int nframes, cudaStream_t st, const CUtensorMap* pyrMaps) {
    if (pyrMaps && hp.lv[l].pyrTW > 0) return launch_pyr_tma(a, hp, l, nframes, pyrMaps[l], st);
    dim3 blk(32, 4), grd(cdiv(hp.lv[l].w, 128), cdiv(hp.lv[l].h, 4 * EORB_PYR_BAND), nframes);
    pyr_resize_kernel<<<grd, blk, 0, st>>>(a, l);
    return cudaGetLastError();
}

cudaError_t launch_orb_pipeline(const OrbArgs& a, const OrbPlan& hp, int nframes, const CUtensorMap& tm0, cudaStream_t st,
                                long long* launches, cudaEvent_t* ev, const CUtensorMap* pyrMaps, const OrbFork* fork, const CUtensorMap* icMap0) {
    const bool forkBlur = fork && fork->side && !ev && a.wantDesc && hp.blurTasksTotal > 0;
    // ev (optional, EORB_ORB_STAGES+1 events): recorded around every stage for the per-kernel timings of bench.py
    if (ev) cudaEventRecord(ev[0], st);
    // K1: pyramid, level by level (each level is resized from the previous one)
    bool chainFits = true;                                // a counter per block row: levels of at most 16 * EORB_PYR_CHAIN_ROWS rows
    for (int l = 1; l < hp.nlevels; l++) chainFits = chainFits && (hp.lv[l].h + 4 * EORB_PYR_CHAIN_BAND - 1) / (4 * EORB_PYR_CHAIN_BAND) <= EORB_PYR_CHAIN_ROWS;
    if (a.pyrDone && nframes <= 8 && hp.nlevels > 1 && chainFits) {   // one launch for all levels (pyr_chain_kernel)
        cudaError_t e = cudaMemsetAsync(a.pyrDone, 0, (size_t)nframes * EORB_MAX_LEVELS * EORB_PYR_CHAIN_ROWS * sizeof(int), st);
        if (e != cudaSuccess) return e;
        pyr_chain_kernel<<<dim3(pyr_chain_tiles(hp), nframes), dim3(32, 4), 0, st>>>(a, a.pyrDone);
        (*launches)++;
    } else
    for (int l = 1; l < hp.nlevels; l++) {
        if (hp.lv[l].w <= 0 || hp.lv[l].h <= 0) continue;
        cudaError_t e = launch_pyramid_level(a, hp, l, nframes, st, pyrMaps);
        if (e != cudaSuccess) return e;
        (*launches)++;
    }
    if (ev) cudaEventRecord(ev[1], st);
    const bool blurTma = a.blurMaps != nullptr && pyrMaps != nullptr;
    auto launchBlur = [&](cudaStream_t s) {
        if (blurTma) {
            const int band = blur_tma_band(a.blurVariant), nt = blur_tma_tasks(hp, band);
            const size_t sm = (size_t)(band + 4) * EORB_BLUR_TBOXW + 16;
            const dim3 g(nt, nframes);
            switch (a.blurVariant) {
                case 2: blur_tma_kernel<64, false><<<g, 32, sm, s>>>(a, pyrMaps[0], nt); break;
                case 3: blur_tma_kernel<64, true><<<g, 32, sm, s>>>(a, pyrMaps[0], nt); break;
                case 4: blur_tma_kernel<32, true><<<g, 32, sm, s>>>(a, pyrMaps[0], nt); break;
                default: blur_tma_kernel<32, false><<<g, 32, sm, s>>>(a, pyrMaps[0], nt); break;
            }
        } else {
            dim3 blk(32, 4), grd(cdiv(hp.blurTasksTotal, 4), nframes);
            blur_kernel<<<grd, blk, 0, s>>>(a);
        }
    };
    if (forkBlur) {   // K5 beside K2 / K3 / K7
        cudaEventRecord(fork->forked, st);
        cudaStreamWaitEvent(fork->side, fork->forked, 0);
        launchBlur(fork->side);
        (*launches)++;
        cudaEventRecord(fork->joined, fork->side);
    }
    // K2: FAST over every cell of every level
    if (hp.nCells > 0) {
        cudaError_t e = launch_fast_cells(a, hp, nframes, tm0, st);
        if (e != cudaSuccess) return e;
        (*launches)++;
    }
    if (ev) cudaEventRecord(ev[2], st);
    // K3: octree distribution per (level, frame)
    {
        dim3 grd(hp.nlevels, nframes);
        if (nframes <= 8) octree_kernel<OCT_MAX_THREADS><<<grd, OCT_MAX_THREADS, hp.octSmemBytes, st>>>(a);
        else octree_kernel<128><<<grd, 128, hp.octSmemBytes, st>>>(a);
        (*launches)++;
    }
    if (ev) cudaEventRecord(ev[3], st);
    // K7: output order
    if (nframes <= 8) orb_index_kernel<OCT_MAX_THREADS><<<nframes, OCT_MAX_THREADS, 0, st>>>(a);
    else orb_index_kernel<128><<<nframes, 128, 0, st>>>(a);
    (*launches)++;
    if (ev) cudaEventRecord(ev[4], st);
    // K5: blur (only needed for descriptors)
    if (forkBlur) {
        cudaStreamWaitEvent(st, fork->joined, 0);
    } else if (a.wantDesc && hp.blurTasksTotal > 0) {
        launchBlur(st);
        (*launches)++;
    }
    if (ev) cudaEventRecord(ev[5], st);
    // K4 + K6
    {
        dim3 grd(cdiv(hp.selPerFrame, EORB_KP_GROUP), nframes);
        OdLevels L;
        memset(&L, 0, sizeof(L));
        L.selPerFrame = hp.selPerFrame;
        for (int l = 0; l < hp.nlevels && l < EORB_MAX_LEVELS; l++) {
            L.geo[l] = make_int4(hp.lv[l].minBX, hp.lv[l].minBY, hp.lv[l].w, hp.lv[l].h);
            L.sc[l] = make_float2(hp.lv[l].scale, hp.lv[l].sizeF);
        }
        CUtensorMap ic0;
        if (icMap0) ic0 = *icMap0; else memset(&ic0, 0, sizeof(ic0));
        orient_desc_kernel<<<grd, 32 * EORB_KP_GROUP, 0, st>>>(a, ic0, L);
        (*launches)++;
    }
    if (ev) cudaEventRecord(ev[6], st);
    return cudaGetLastError();
}

cudaError_t launch_pyramid_and_blur(const OrbArgs& a, const OrbPlan& hp, cudaStream_t st, long long* launches, const CUtensorMap* pyrMaps) {
    for (int l = 1; l < hp.nlevels; l++) {
        if (hp.lv[l].w <= 0 || hp.lv[l].h <= 0) continue;
        cudaError_t e = launch_pyramid_level(a, hp, l, 1, st, pyrMaps);
        if (e != cudaSuccess) return e;
        (*launches)++;
    }
    if (hp.blurTasksTotal > 0) {
        dim3 blk(32, 4), grd(cdiv(hp.blurTasksTotal, 4), 1);
        blur_kernel<<<grd, blk, 0, st>>>(a);
        (*launches)++;
    }
    return cudaGetLastError();
}

cudaError_t launch_tracked_desc(const OrbArgs& a, const OrbPlan& hp, const eorb_keypoint* d_kps, int n, int mode,
                                const float* d_invScale, const uint8_t* d_refDesc, uint8_t* d_desc, int* d_dist,
                                cudaStream_t st, long long* launches) {
    if (n <= 0) return cudaSuccess;
    dim3 grd(cdiv(n, 8), hp.nlevels);
    tracked_desc_kernel<<<grd, 256, 0, st>>>(a, d_kps, n, mode, d_invScale, d_refDesc, d_desc, d_dist);
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t orb_kernels_configure(const OrbPlan& hp) {
    cudaError_t e = fast_cells_configure(hp);
    if (e != cudaSuccess) return e;
    // per-kernel attribute shared by every extractor of the process: set once to the plan limit (see fast_cells_configure)
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(octree_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(octree_kernel<OCT_MAX_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
    return e;
}

}  // namespace eorb
