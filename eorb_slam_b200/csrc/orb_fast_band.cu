// orb_fast_band.cu — K2b: grid FAST-9/16 on whole runs of cells (reference: ComputeKeyPointsOctTree grid loop src/ORBextractor.cc:784-878
// calling cv::FAST at :832,851).  Same three phases and the same packed arithmetic as fast_cells_kernel (orb_fast.cu); what changes is
// the work unit.  The per-cell kernel spends a third of its phase-1 lane-steps on alignment slack (a ~31-36 pixel cell interior walked
// in aligned 8-pixel groups uses 67-76 % of the lanes) and a sixth of its instructions on the per-cell prologue
// (profiles/r02_fast_phase_attribution.md).  Here one 4-warp block owns a SEGMENT of up to 5 horizontally adjacent cells:
//
//   stage    ONE TMA tile for the union of the cells' ROIs (cp.async.bulk.tensor.3d, zero fill outside the level); while it is in
//            flight the block clears the score map and builds a column table (cell index + "first / last column of its cell" flags).
//   phase 1  the union interior is walked as a 1-D sequence of (row, 8-pixel group) tasks, 32 per warp-step, so every lane works
//            except on the last step; warp w owns the rows [w*ch/4, (w+1)*ch/4).  Flagged groups -> the warp's group list ->
//            the warp's pixel list (row-major inside the warp's rows).
//   phase 2  exact score per survivor into the block's score map; corners compacted in place.
//   phase 3  after a block barrier (neighbour rows belong to other warps): strict 3x3 NMS that treats the columns of a neighbouring
//            CELL as absent (cv::FAST ran per cell ROI, so a cell's border pixels have no neighbours outside it), kept corners
//            compacted in place and counted per (warp, cell); a second barrier turns the counts into per-cell offsets (warps own
//            ascending row ranges, so "warp order, then list order" is the row-major order of each cell) and the kept corners are
//            written to their cells' slots.
//   fallback a cell that stays empty at iniThFAST re-runs the phases at minThFAST over its own columns only (:849-852).
//
// Outputs are identical to fast_cells_kernel's: per-cell slot ranges and counts (tests compare candidate sets AND order).
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

__device__ __forceinline__ unsigned band_compass4(unsigned C, unsigned N, unsigned S, unsigned E, unsigned W, unsigned K) {
    const unsigned aN = __vabsdiffu4(N, C), aS = __vabsdiffu4(S, C), aE = __vabsdiffu4(E, C), aW = __vabsdiffu4(W, C);
    const unsigned ns = (aN + K) | aN | (aS + K) | aS;
    const unsigned ew = (aE + K) | aE | (aW + K) | aW;
    return ns & ew & 0x80808080u;
}

struct FastBandConst {
    int nSegs, nCells, slotsPerFrame;
    int TS, tileRows, MS;
    int mapOff, lutOff, cellOff, listOff, listPerWarp, taskOff, taskPerWarp, barOff;
    int tA, tB;
};

// per-block cell table in shared memory
struct BandCells {
    int slotOff[8];
    int cs[8], ce[8];          // interior tile columns [cs, ce) of cell ci
    int cntW[4][8];            // kept corners per (warp, cell) of the current pass
    int runW[4][8];            // emission cursor per (warp, cell)
    int off[4][8];             // first output position of warp w inside cell ci
    int tot[8];
    int needB[8];
};

static_assert(sizeof(BandCells) <= EORB_FAST_BAND_CELL_BYTES, "the plan reserves EORB_FAST_BAND_CELL_BYTES for the cell table");

#define BAND_WARPS 4

__global__ void __launch_bounds__(BAND_WARPS * 32) fast_band_kernel(OrbArgs a, const __grid_constant__ CUtensorMap tm0, FastBandConst K0) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int seg = blockIdx.x, f = blockIdx.y;
    if (seg >= K0.nSegs) return;
    SegPlan sp;
    {
        const uint4* p = reinterpret_cast<const uint4*>(a.segs + seg);
        const uint4 c0 = __ldg(p), c1 = __ldg(p + 1);
        sp.x0 = (short)(c0.x & 0xffffu); sp.y0 = (short)(c0.x >> 16);
        sp.w = (short)(c0.y & 0xffffu);  sp.h = (short)(c0.y >> 16);
        sp.level = (short)(c0.z & 0xffffu); sp.ncell = (short)(c0.z >> 16);
        sp.firstCell = (int)c0.w;
        sp.ox = (short)(c1.x & 0xffffu); sp.oy = (short)(c1.x >> 16);
        sp.wCell = (short)(c1.y & 0xffffu);
        sp.aoff = (unsigned char)(c1.z & 0xffu); sp.p0 = (unsigned char)((c1.z >> 8) & 0xffu); sp.np = (unsigned char)((c1.z >> 16) & 0xffu);
        sp.firstMask = (unsigned char)(c1.w & 0xffu); sp.lastMask = (unsigned char)((c1.w >> 8) & 0xffu);
        sp.rcpNp = (unsigned short)(c1.w >> 16);
    }
    const uint8_t* tile = smem_raw;
    uint8_t* smap = smem_raw + K0.mapOff;
    uint8_t* lut = smem_raw + K0.lutOff;
    BandCells& bc = *reinterpret_cast<BandCells*>(smem_raw + K0.cellOff);
    uint16_t* list = reinterpret_cast<uint16_t*>(smem_raw + K0.listOff) + (size_t)warp * K0.listPerWarp;
    uint32_t* tlist = reinterpret_cast<uint32_t*>(smem_raw + K0.taskOff) + (size_t)warp * K0.taskPerWarp;
    const unsigned bar = smem_u32(smem_raw + K0.barOff);
    const int TS = K0.TS, MS = K0.MS;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    const int ncell = sp.ncell;
    const int ch = sp.h - 6, cwU = sp.w - 6;
    const int aoff = sp.aoff;
    uint16_t* countOut = a.cellCount + (size_t)f * K0.nCells + sp.firstCell;
    if (ch <= 0 || cwU <= 0) {
        if (tid < ncell) countOut[tid] = 0;
        return;
    }

    // ---- stage
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, (unsigned)(TS * K0.tileRows));
        const CUtensorMap* tm = sp.level == 0 ? &tm0 : a.bandMaps + sp.level;
        tma_load_3d(smem_u32(tile), tm, sp.x0 & ~15, sp.y0, f, bar);
    }
    if (tid < 8) {
        int so = 0, cs = 0, ce = 0;
        if (tid < ncell) {
            const uint4 c0 = __ldg(reinterpret_cast<const uint4*>(a.cells + sp.firstCell + tid));
            const int cw = (int)(short)(c0.y & 0xffffu) - 6;
            so = (int)c0.w;
            cs = aoff + 3 + tid * sp.wCell; ce = cs + max(cw, 0);
        }
        bc.slotOff[tid] = so; bc.cs[tid] = cs; bc.ce[tid] = ce; bc.tot[tid] = 0; bc.needB[tid] = 0;
    }
    {
        const int mapVecs = ((ch + 2) * MS + 15) >> 4;
        for (int i = tid; i < mapVecs; i += BAND_WARPS * 32) reinterpret_cast<uint4*>(smap)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    for (int c = tid; c < TS; c += BAND_WARPS * 32) {   // column table: cell index | 0x40 first column | 0x80 last column; 0x3f outside
        unsigned v = 0x3fu;
        for (int ci = 0; ci < ncell; ci++)
            if (c >= bc.cs[ci] && c < bc.ce[ci]) v = (unsigned)ci | (c == bc.cs[ci] ? 0x40u : 0u) | (c == bc.ce[ci] - 1 ? 0x80u : 0u);
        lut[c] = (uint8_t)v;
    }
    __syncthreads();
    mbar_wait(bar, 0);

    const int rowsW = (ch + BAND_WARPS - 1) / BAND_WARPS;
    const int rBeg = min(warp * rowsW, ch), rEnd = min(rBeg + rowsW, ch);
    uint32_t* slotsF = a.cand + (size_t)f * K0.slotsPerFrame;
    const int tA = K0.tA, tB = K0.tB;

    // one pass of the three phases over the tile columns described by (p0, np, firstMask, lastMask, rcp) at threshold t
    auto runPass = [&](int t, int p0, int np, unsigned firstMask, unsigned lastMask, unsigned rcp) {
        const unsigned K = (unsigned)(127 - min(t, 127)) * 0x01010101u;
        if (tid < BAND_WARPS * 8) { (&bc.cntW[0][0])[tid] = 0; (&bc.runW[0][0])[tid] = 0; }
        // ---- phase 1a
        int ntask = 0;
        {
            const int total = (rEnd - rBeg) * np;
            for (int id0 = 0; id0 < total; id0 += 32) {
                const int id = id0 + lane;
                unsigned m8 = 0;
                int rr = 0, colBase = 0;
                if (id < total) {
                    const int row = (int)(((unsigned)id * rcp) >> 16), grp = id - row * np;
                    rr = rBeg + row; colBase = (p0 + grp) * 8;
                    const uint8_t* q = tile + (rr + 3) * TS + colBase;
                    const uint2 C = *reinterpret_cast<const uint2*>(q);
                    const uint2 N = *reinterpret_cast<const uint2*>(q + 3 * TS);
                    const uint2 S = *reinterpret_cast<const uint2*>(q - 3 * TS);
                    const unsigned L = *reinterpret_cast<const unsigned*>(q - 4);
                    const unsigned R = *reinterpret_cast<const unsigned*>(q + 8);
                    const unsigned f0 = band_compass4(C.x, N.x, S.x, __byte_perm(C.x, C.y, 0x6543), __byte_perm(L, C.x, 0x4321), K);
                    const unsigned f1 = band_compass4(C.y, N.y, S.y, __byte_perm(C.y, R, 0x6543), __byte_perm(C.x, C.y, 0x4321), K);
                    unsigned vm = grp == 0 ? firstMask : 0xffu;
                    if (grp == np - 1) vm &= lastMask;
                    m8 = (((f0 * 0x00204081u) >> 28) | (((f1 * 0x00204081u) >> 28) << 4)) & vm;
                }
                const unsigned bal = __ballot_sync(FULL, m8 != 0);
                if (m8 != 0) tlist[ntask + __popc(bal & lt)] = (m8 << 16) | ((unsigned)rr << 8) | (unsigned)colBase;
                ntask += __popc(bal);
            }
        }
        __syncwarp();
        // ---- phase 1b
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
        // ---- phase 2
        int ncorn = 0;
        for (int base = 0; base < nsurv; base += 32) {
            const int s = base + lane;
            bool corner = false;
            int code = 0;
            if (s < nsurv) {
                code = list[s];
                const int rr = code >> 8, col = code & 255;
                const uint8_t* p = tile + (rr + 3) * TS + col;
                const int v = p[0];
                int ring[16];
                ring[0] = p[3 * TS];       ring[1] = p[3 * TS + 1];   ring[2] = p[2 * TS + 2];    ring[3] = p[TS + 3];
                ring[4] = p[3];            ring[5] = p[-TS + 3];      ring[6] = p[-2 * TS + 2];   ring[7] = p[-3 * TS + 1];
                ring[8] = p[-3 * TS];      ring[9] = p[-3 * TS - 1];  ring[10] = p[-2 * TS - 2];  ring[11] = p[-TS - 3];
                ring[12] = p[-3];          ring[13] = p[TS - 3];      ring[14] = p[2 * TS - 2];   ring[15] = p[3 * TS - 1];
                const int m = fast_max_arc_min_packed(v, ring);
                corner = m > t;
                if (corner) smap[(rr + 1) * MS + col - aoff - 2] = (uint8_t)m;   // map column = union interior x + 1
            }
            const unsigned cm = __ballot_sync(FULL, corner);
            if (corner) list[ncorn + __popc(cm & lt)] = (uint16_t)code;
            ncorn += __popc(cm);
        }
        __syncthreads();   // every warp's scores are in the map
        // ---- phase 3a: NMS inside the corner's own cell, kept corners compacted in place, counted per (warp, cell)
        int nkept = 0;
        for (int base = 0; base < ncorn; base += 32) {
            const int s = base + lane;
            bool keep = false;
            int code = 0;
            unsigned ci = 15u;
            if (s < ncorn) {
                code = list[s];
                const int rr = code >> 8, col = code & 255;
                const unsigned lv = lut[col];
                const uint8_t* q = smap + (rr + 1) * MS + col - aoff - 2;
                const int m = q[0];
                int l0 = q[-MS - 1], l1 = q[-1], l2 = q[MS - 1];
                int r0 = q[-MS + 1], r1 = q[1], r2 = q[MS + 1];
                if (lv & 0x40u) { l0 = 0; l1 = 0; l2 = 0; }          // first column of its cell: no left neighbours
                if (lv & 0x80u) { r0 = 0; r1 = 0; r2 = 0; }          // last column: no right neighbours
                const int n0 = max(max(l0, (int)q[-MS]), r0);
                const int n1 = max(max(l1, r1), 1);                  // the OpenCV score m - 1 must beat a non-corner's 0
                const int n2 = max(max(l2, (int)q[MS]), r2);
                keep = m > max(max(n0, n1), n2);
                if (keep) ci = lv & 7u;
            }
            const unsigned msk = __ballot_sync(FULL, keep);
            const unsigned peers = __match_any_sync(FULL, ci);
            __syncwarp();
            if (keep) {
                list[nkept + __popc(msk & lt)] = (uint16_t)code;
                if ((peers & lt) == 0) bc.cntW[warp][ci] += __popc(peers);   // the lowest lane of every cell group
            }
            nkept += __popc(msk);
            __syncwarp();
        }
        __syncthreads();
        if (tid < 8) {
            int acc = 0;
#pragma unroll
            for (int w = 0; w < BAND_WARPS; w++) { bc.off[w][tid] = acc; acc += bc.cntW[w][tid]; }
            bc.tot[tid] = acc;
        }
        __syncthreads();
        // ---- phase 3b: ordered emission into the cells' slot ranges
        for (int base = 0; base < nkept; base += 32) {
            const int s = base + lane;
            const bool have = s < nkept;
            unsigned ci = 15u;
            uint32_t packed = 0;
            if (have) {
                const int code = list[s];
                const int rr = code >> 8, col = code & 255;
                ci = lut[col] & 7u;
                const int m = smap[(rr + 1) * MS + col - aoff - 2];
                packed = (uint32_t)(col - aoff + sp.ox) | ((uint32_t)(rr + 3 + sp.oy) << 12) | ((uint32_t)(m - 1) << 24);
            }
            const unsigned peers = __match_any_sync(FULL, ci);
            int pos = 0;
            if (have) pos = bc.off[warp][ci] + bc.runW[warp][ci] + __popc(peers & lt);
            __syncwarp();
            if (have) {
                slotsF[bc.slotOff[ci] + pos] = packed;
                if ((peers & lt) == 0) bc.runW[warp][ci] += __popc(peers);
            }
            __syncwarp();
        }
        __syncthreads();   // tot[] is read by everybody after the pass; the lists are reused by the next pass
    };

    runPass(tA, sp.p0, sp.np, sp.firstMask, sp.lastMask, sp.rcpNp);
    const bool fallback = tB < tA;
    if (tid < ncell) {
        const int c = bc.tot[tid];
        const bool again = fallback && c == 0 && bc.ce[tid] > bc.cs[tid];
        bc.needB[tid] = again ? 1 : 0;
        if (!again) countOut[tid] = (uint16_t)c;
    }
    __syncthreads();
    if (!fallback) return;
    for (int ci = 0; ci < ncell; ci++) {
        if (!bc.needB[ci]) continue;                         // block-uniform
        const int cs = bc.cs[ci], ce = bc.ce[ci];
        const int p0 = cs >> 3, np = ((ce - 1) >> 3) - p0 + 1;
        const unsigned firstMask = (0xffu << (cs & 7)) & 0xffu;
        const int lastBits = ce - 8 * (p0 + np - 1);
        const unsigned lastMask = lastBits >= 8 ? 0xffu : ((1u << lastBits) - 1u);
        const unsigned rcp = (65536u + (unsigned)np - 1u) / (unsigned)np;
        __syncthreads();
        runPass(tB, p0, np, firstMask, lastMask, rcp);
        if (tid == 0) countOut[ci] = (uint16_t)bc.tot[ci];
        __syncthreads();
    }
}

cudaError_t launch_fast_band(const OrbArgs& a, const OrbPlan& hp, int nframes, const CUtensorMap& tmBand0, cudaStream_t st) {
    FastBandConst k;
    k.nSegs = hp.nSegs; k.nCells = hp.nCells; k.slotsPerFrame = hp.slotsPerFrame;
    k.TS = hp.bandTS; k.tileRows = hp.bandRows; k.MS = hp.bandMS;
    k.mapOff = hp.bandMapOff; k.lutOff = hp.bandLutOff; k.cellOff = hp.bandCellOff; k.listOff = hp.bandListOff; k.listPerWarp = hp.bandListPerWarp;
    k.taskOff = hp.bandTaskOff; k.taskPerWarp = hp.bandTaskPerWarp; k.barOff = hp.bandBarOff;
    k.tA = hp.iniTh < 0 ? 0 : (hp.iniTh > 255 ? 255 : hp.iniTh);
    k.tB = hp.minTh < 0 ? 0 : (hp.minTh > 255 ? 255 : hp.minTh);
    fast_band_kernel<<<dim3(hp.nSegs, nframes), BAND_WARPS * 32, (size_t)hp.bandSmem, st>>>(a, tmBand0, k);
    return cudaGetLastError();
}

cudaError_t fast_band_configure() {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(fast_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
    return e;
}

}  // namespace eorb
