// orb_fast.cu — K2: grid FAST-9/16 with cell-local NMS and the iniThFAST -> minThFAST fallback
// (reference: ComputeKeyPointsOctTree grid loop src/ORBextractor.cc:784-878 calling cv::FAST at :832,851).
//
// One warp per FAST grid cell, no block-level barrier.
//   stage    the cell ROI (<= 66x66 incl. the 3-px apron) arrives in shared memory as ONE TMA tile
//            (cp.async.bulk.tensor.3d on a per-level {x, y, frame} tensor map, zero fill outside the level),
//            completion on the warp's own mbarrier; the score map is cleared while the copy is in flight.
//            Measured on B200: the innermost TMA coordinate must be a multiple of 16 bytes (x = 4 raises "illegal
//            instruction", x = 0 / 16 / -16 work), so the box starts at x0 & ~15 and the ROI sits at column
//            aoff = x0 & 15 of the tile.
//   phase 1  necessary condition, 8 pixels per lane per step on packed bytes: a 9-arc of the 16-ring always
//            contains one of {N, S} and one of {E, W}, so a corner at threshold t needs
//            (|N-v| > t or |S-v| > t) and (|E-v| > t or |W-v| > t).  |.| is one VABSDIFF4 per 4 pixels; "> t" is
//            bit 7 of ((a + (127 - t)) | a) per byte (a carry out of a byte can only turn the next byte's test
//            into ">= t", i.e. the filter stays a superset).  Survivors are compacted in row-major order.
//   phase 2  exact score m = max over the 16 arcs of min(|v - p|) with a common sign for survivors only; both
//            polarities ride in one register as s16x2 (lo = p - v, hi = v - p) through a MIN3/MAX3 network
//            (VIMNMX3.S16x2): 32 + 8 instructions instead of two scalar networks.
//   phase 3  cell-local strict 3x3 NMS on the score map and ordered emission (corner at t <=> m > t,
//            OpenCV score = m - 1).  If the cell produced nothing at iniThFAST the three phases re-run at
//            minThFAST (:849-852); only a few per cent of the cells take that path.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "../../include/eorb_b200.h"
#include "eorb_math.cuh"
#include "fast_score.cuh"
#include "orb_kernels.h"
#include "tma_utils.cuh"

namespace eorb {

// ---- phase-1 helper: bit 7 of byte j set  <=>  pixel j may be a corner at the threshold encoded in K ---------------
__device__ __forceinline__ unsigned fast_compass4(unsigned C, unsigned N, unsigned S, unsigned E, unsigned W, unsigned K) {
    const unsigned aN = __vabsdiffu4(N, C), aS = __vabsdiffu4(S, C), aE = __vabsdiffu4(E, C), aW = __vabsdiffu4(W, C);
    const unsigned ns = (aN + K) | aN | (aS + K) | aS;
    const unsigned ew = (aE + K) | aE | (aW + K) | aW;
    return ns & ew & 0x80808080u;
}

// plan constants of the kernel, passed by value (constant bank) instead of being re-read from the plan in HBM
struct FastConst {
    int nCells, slotsPerFrame;
    int smemPerWarp, mapOff, listOff, taskOff, barOff;
    int TS, tileRows, MS;
    int tA, tB;
};

__global__ void __launch_bounds__(EORB_FAST_WARPS * 32) fast_cells_kernel(OrbArgs a, const __grid_constant__ CUtensorMap tm0, FastConst K0) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cell = blockIdx.x * EORB_FAST_WARPS + warp;
    const int f = blockIdx.y;
    if (cell >= K0.nCells) return;
    // The 32-byte CellPlan is fetched as two 128-bit words and unpacked by hand: besides saving six loads this avoids a
    // ptxas 12.9 miscompile seen with one-warp blocks, where the struct copy's `y0` (high half of word 0) was replaced
    // by the high word of an unrelated 64-bit product before it reached the TMA coordinate (every cell row then saw
    // tile row 0; caught by the parity tests).
    CellPlan c;
    {
        const uint4* cp = reinterpret_cast<const uint4*>(a.cells + cell);
        const uint4 c0 = __ldg(cp), c1 = __ldg(cp + 1);
        c.x0 = (short)(c0.x & 0xffffu); c.y0 = (short)(c0.x >> 16);
        c.w = (short)(c0.y & 0xffffu);  c.h = (short)(c0.y >> 16);
        c.level = (short)(c0.z & 0xffffu); c._pad = 0;
        c.slotOff = (int)c0.w;
        c.ox = (short)(c1.x & 0xffffu); c.oy = (short)(c1.x >> 16);
        c.slotCap = (int)c1.y;
        c.aoff = (unsigned char)(c1.z & 0xffu); c.p0 = (unsigned char)((c1.z >> 8) & 0xffu);
        c.np = (unsigned char)((c1.z >> 16) & 0xffu); c.rps = (unsigned char)(c1.z >> 24);
        c.firstMask = (unsigned char)(c1.w & 0xffu); c.lastMask = (unsigned char)((c1.w >> 8) & 0xffu);
        c.rcpNpM1 = (unsigned short)(c1.w >> 16);
    }
    unsigned char* ws = smem_raw + (size_t)warp * K0.smemPerWarp;
    const uint8_t* tile = ws;
    uint8_t* smap = ws + K0.mapOff;
    uint16_t* list = reinterpret_cast<uint16_t*>(ws + K0.listOff);
    uint32_t* tlist = reinterpret_cast<uint32_t*>(ws + K0.taskOff);
    const unsigned bar = smem_u32(ws + K0.barOff);
    const int TS = K0.TS, MS = K0.MS;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;

    uint32_t* slots = a.cand + (size_t)f * K0.slotsPerFrame + c.slotOff;
    uint16_t* countOut = a.cellCount + (size_t)f * K0.nCells + cell;
    const int w = c.w, h = c.h;
    const int cw = w - 6, ch = h - 6;
    if (cw <= 0 || ch <= 0) { if (lane == 0) *countOut = 0; return; }

    // ---- stage: one TMA tile per cell
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, (unsigned)(TS * K0.tileRows));
        const CUtensorMap* tm = c.level == 0 ? &tm0 : a.tmaps + c.level;
        tma_load_3d(smem_u32(tile), tm, c.x0 & ~15, c.y0, f, bar);
    }
    {
        const int mapVecs = ((ch + 2) * MS + 15) >> 4;   // the map region is padded to 16 bytes in the plan
        for (int i = lane; i < mapVecs; i += 32) reinterpret_cast<uint4*>(smap)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
    mbar_wait(bar, 0);

    const int tA = K0.tA, tB = K0.tB;
    // interior = tile columns [aoff + 3, aoff + w - 3); phase 1 walks it in aligned groups of 8 columns.
    // Fixed lane -> (row-in-step, group) assignment: a step covers rps rows x np groups in row-major lane order
    // (the walk constants come precomputed with the cell).
    const int aoff = c.aoff, np = c.np, rps = c.rps;
    const int lr = (lane * ((int)c.rcpNpM1 + 1)) >> 16, lp = lane - lr * np;
    const bool laneUsed = lr < rps;
    unsigned vm = lp == 0 ? (unsigned)c.firstMask : 0xffu;        // interior columns of this lane's group
    if (lp == np - 1) vm &= (unsigned)c.lastMask;
    const int colBase = ((int)c.p0 + lp) * 8;
    const uint8_t* qLane = tile + (lr + 3) * TS + colBase;

    int cnt = 0;
    for (int pass = 0; pass < 2; pass++) {
        const int t = pass ? tB : tA;
        const unsigned K = (unsigned)(127 - min(t, 127)) * 0x01010101u;

        // ---- phase 1a: packed compass test; groups with at least one flagged pixel go to the group list
        int ntask = 0;
        {
            const uint8_t* q = qLane;
            for (int rbase = 0; rbase < ch; rbase += rps, q += rps * TS) {
                const int rr = rbase + lr;
                unsigned m8 = 0;
                if (laneUsed && rr < ch) {
                    const uint2 C = *reinterpret_cast<const uint2*>(q);
                    const uint2 N = *reinterpret_cast<const uint2*>(q + 3 * TS);
                    const uint2 S = *reinterpret_cast<const uint2*>(q - 3 * TS);
                    const unsigned L = *reinterpret_cast<const unsigned*>(q - 4);
                    const unsigned R = *reinterpret_cast<const unsigned*>(q + 8);
                    const unsigned f0 = fast_compass4(C.x, N.x, S.x, __byte_perm(C.x, C.y, 0x6543), __byte_perm(L, C.x, 0x4321), K);
                    const unsigned f1 = fast_compass4(C.y, N.y, S.y, __byte_perm(C.y, R, 0x6543), __byte_perm(C.x, C.y, 0x4321), K);
                    // gather the four bit-7 flags of each word into a nibble (multiplier places bits 7,15,23,31 at 28..31)
                    m8 = (((f0 * 0x00204081u) >> 28) | (((f1 * 0x00204081u) >> 28) << 4)) & vm;
                }
                const unsigned bal = __ballot_sync(FULL, m8 != 0);
                if (m8 != 0) tlist[ntask + __popc(bal & lt)] = (m8 << 16) | ((unsigned)rr << 7) | (unsigned)colBase;
                ntask += __popc(bal);
            }
        }
        __syncwarp();
        // ---- phase 1b: expand the flagged groups into the pixel list (row-major order is preserved)
        int nsurv = 0;
        for (int base = 0; base < ntask; base += 32) {
            const unsigned e = base + lane < ntask ? tlist[base + lane] : 0u;
            const unsigned m8 = e >> 16, code = e & 0xffffu;
            const int k = __popc(m8);
            int incl = k;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += up;
            }
            uint16_t* lp16 = list + nsurv + incl - k;
            nsurv += __shfl_sync(FULL, incl, 31);
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (m8 & (1u << j)) *lp16++ = (uint16_t)(code + j);
        }
        __syncwarp();

        // ---- phase 2: exact arc score for survivors; corners (m > t) are compacted IN PLACE at the head of the list
        // (still row-major: a chunk only overwrites entries at or before its own, which it has already read)
        int ncorn = 0;
        for (int base = 0; base < nsurv; base += 32) {
            const int s = base + lane;
            bool corner = false;
            int code = 0;
            if (s < nsurv) {
                code = list[s];
                const int rr = code >> 7, col = code & 127;
                const uint8_t* p = tile + (rr + 3) * TS + col;
                const int v = p[0];
                int ring[16];
                ring[0] = p[3 * TS];       ring[1] = p[3 * TS + 1];   ring[2] = p[2 * TS + 2];    ring[3] = p[TS + 3];
                ring[4] = p[3];            ring[5] = p[-TS + 3];      ring[6] = p[-2 * TS + 2];   ring[7] = p[-3 * TS + 1];
                ring[8] = p[-3 * TS];      ring[9] = p[-3 * TS - 1];  ring[10] = p[-2 * TS - 2];  ring[11] = p[-TS - 3];
                ring[12] = p[-3];          ring[13] = p[TS - 3];      ring[14] = p[2 * TS - 2];   ring[15] = p[3 * TS - 1];
                const int m = fast_max_arc_min_packed(v, ring);
                corner = m > t;
                if (corner) smap[(rr + 1) * MS + col - aoff - 2] = (uint8_t)m;   // map column = interior x + 1
            }
            const unsigned cm = __ballot_sync(FULL, corner);
            if (corner) list[ncorn + __popc(cm & lt)] = (uint16_t)code;
            ncorn += __popc(cm);
        }
        __syncwarp();

        // ---- phase 3: strict 3x3 NMS over the corners + ordered emission
        cnt = 0;
        for (int base = 0; base < ncorn; base += 32) {
            const int s = base + lane;
            bool keep = false;
            uint32_t packed = 0;
            if (s < ncorn) {
                const int code = list[s];
                const int rr = code >> 7, col = code & 127;
                const uint8_t* q = smap + (rr + 1) * MS + col - aoff - 2;
                const int m = q[0];
                // every stored score is > t, everything else is 0: strict maximum over the 8 neighbours, and the
                // OpenCV score m - 1 must beat a non-corner's 0
                const int n0 = max(max((int)q[-MS - 1], (int)q[-MS]), (int)q[-MS + 1]);
                const int n1 = max(max((int)q[-1], (int)q[1]), 1);
                const int n2 = max(max((int)q[MS - 1], (int)q[MS]), (int)q[MS + 1]);
                keep = m > max(max(n0, n1), n2);
                packed = (uint32_t)(col - aoff + c.ox) | ((uint32_t)(rr + 3 + c.oy) << 12) | ((uint32_t)(m - 1) << 24);
            }
            const unsigned msk = __ballot_sync(FULL, keep);
            if (keep) slots[cnt + __popc(msk & lt)] = packed;
            cnt += __popc(msk);
        }
        if (cnt > 0 || tB >= tA) break;
        __syncwarp();
    }
    if (lane == 0) *countOut = (uint16_t)cnt;
}

cudaError_t launch_fast_cells(const OrbArgs& a, const OrbPlan& hp, int nframes, const CUtensorMap& tm0, cudaStream_t st) {
    dim3 grd((hp.nCells + EORB_FAST_WARPS - 1) / EORB_FAST_WARPS, nframes);
    FastConst k;
    k.nCells = hp.nCells; k.slotsPerFrame = hp.slotsPerFrame;
    k.smemPerWarp = hp.cellSmemPerWarp; k.mapOff = hp.cellMapOff; k.listOff = hp.cellListOff; k.taskOff = hp.cellTaskOff; k.barOff = hp.cellBarOff;
    k.TS = hp.cellTileStride; k.tileRows = hp.cellTileRows; k.MS = hp.cellMapStride;
    k.tA = hp.iniTh < 0 ? 0 : (hp.iniTh > 255 ? 255 : hp.iniTh);
    k.tB = hp.minTh < 0 ? 0 : (hp.minTh > 255 ? 255 : hp.minTh);
    fast_cells_kernel<<<grd, EORB_FAST_WARPS * 32, (size_t)hp.cellSmemPerWarp * EORB_FAST_WARPS, st>>>(a, tm0, k);
    return cudaGetLastError();
}

cudaError_t fast_cells_configure(const OrbPlan& hp) {
    // The attribute belongs to the KERNEL, not to an extractor: several extractors with different plans live in one process
    // (image ORB, event L1 / L2: Tracking.cc:115-137, EvBaseTracker.cpp:163), so it is set once to the largest size a plan can
    // ask for (orbBuildPlan rejects plans above 200 KB) instead of to this plan's size, which a later, smaller plan would undo.
    (void)hp;
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(fast_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
    return e;
}

}  // namespace eorb
