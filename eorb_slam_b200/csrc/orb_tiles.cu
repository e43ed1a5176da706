// orb_tiles.cu — K1t: the pyramid resize with TMA-staged source tiles (reference ComputePyramid, src/ORBextractor.cc:1240-1265:
// cv::resize INTER_LINEAR of level l-1 into level l, 11-bit fixed point; same arithmetic as pyr_resize_kernel in orb_kernels.cu).
//
// The source footprint of a destination tile (about 1.2 TW + 17 by 1.2 RS + 2 bytes at s = 1.2) arrives in shared memory as ONE
// cp.async.bulk.tensor tile of the {x, y, frame} map of level l-1 (zero fill outside the level: taps that fall there carry weight
// 0, cv::resize clamps sx to sw-1 with fx = 0; clamped ROW indices come from the y table, so the bottom edge never touches the
// fill).  A lane owns 4 destination columns = two 8-byte source windows, a byte-permute selector and a packed 11-bit weight pair
// per column (one dp2a = a0*p[sx] + a1*p[sx+1]).  At s = 1.2 five of six consecutive destination rows share a source row (sy1 of
// row dy == sy0 of row dy+1): the horizontally interpolated row is kept in registers, so 1.17 instead of 2 source rows are
// interpolated per destination row -- with global loads the branch stopped loads from overlapping (pyr_resize_kernel's comment);
// from shared memory the latency is short and the reuse pays.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eorb_b200.h"
#include "eorb_math.cuh"
#include "orb_kernels.h"
#include "tma_utils.cuh"

namespace eorb {

// (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2 for four columns, packed into one word (bs = b << 16, so that
// (b * x) >> 16 is one multiply-high).  ncu: the fmaheavy pipe (every IMAD / IMAD.HI / IDP) is this kernel's busiest unit (77 %);
// folding the additions into 64-bit multiply-add addends was measured slower (0.64 -> 0.74 us/frame: IMAD.WIDE costs two slots).
__device__ __forceinline__ unsigned pyrt_vrow(const unsigned* __restrict__ h0, const unsigned* __restrict__ h1, unsigned bs0, unsigned bs1) {
    unsigned v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) v[j] = (__umulhi(bs0, h0[j]) + __umulhi(bs1, h1[j]) + 2u) >> 2;
    const unsigned lo = __byte_perm(v[0], v[1], 0x0040);
    const unsigned hi = __byte_perm(v[2], v[3], 0x0040);
    return __byte_perm(lo, hi, 0x5410);
}

// One block (4 warps) owns a destination tile of TW x TH pixels: one TMA copy of its source footprint, then every warp walks a
// contiguous band of TH/4 destination rows.  About 16 blocks are resident per SM, so the copy of one block's tile overlaps the
// arithmetic of the others (a per-warp double-buffered strip walk was measured slower: 0.79 vs 0.67 us/frame, its 4 KB copies
// could not be issued far enough ahead at 28 resident warps).
__global__ void __launch_bounds__(128) pyr_tma_kernel(OrbArgs a, const __grid_constant__ CUtensorMap tm, PyrTileConst K) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint8_t* tile = smem_raw;
    const unsigned bar = smem_u32(smem_raw + K.barOff);
    const int f = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dx0t = blockIdx.x * K.TW, dy0t = blockIdx.y * K.TH;
    const short4* __restrict__ xtab = a.xtab + K.xtabOff;
    // y taps of the TMA path: {sy0 * BW, sy1 * BW, b0 << 16, b1 << 16} -- tile-row BYTE offsets, so a row address is one add on the
    // ALU pipe instead of a multiply on the (saturated) fmaheavy pipe
    const int4* __restrict__ ytab = a.ytabT + K.ytabOff;
    // tile origin in the source level: the first tap column rounded down to 16 bytes (TMA inner-coordinate rule), the first tap row
    const int sx0 = (int)__ldg(&xtab[dx0t]).x & ~15;
    const int sy0 = __ldg(&a.ytab[K.ytabOff + dy0t]).x;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, (unsigned)(K.BW * K.BH));
        tma_load_3d(smem_u32(tile), &tm, sx0, sy0, f, bar);
    }
    // ---- per-lane constants (while the tile is in flight)
    const int dx0 = dx0t + lane * 4;
    const bool store = lane * 4 < min(K.TW, K.dw - dx0t);
    int baseA, baseB;
    unsigned sel[4], wt[4];
    {
        int sx[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const short4 xt = __ldg(&xtab[min(dx0 + j, K.dw - 1)]);   // sx, sx+1 (clamped), a0, a1
            sx[j] = max((int)xt.x - sx0, 0);
            wt[j] = (unsigned)(unsigned short)xt.z | ((unsigned)(unsigned short)xt.w << 16);
        }
        // idle lanes (columns past the level) repeat the last column: keep their windows inside the tile
        baseA = min(sx[0] & ~3, K.BW - 8); baseB = min(sx[2] & ~3, K.BW - 8);
        const int o0 = sx[0] - baseA, o1 = sx[1] - baseA, o2 = sx[2] - baseB, o3 = sx[3] - baseB;
        sel[0] = (unsigned)(min(o0, 7) | (min(o0 + 1, 7) << 4));
        sel[1] = (unsigned)(min(o1, 7) | (min(o1 + 1, 7) << 4));
        sel[2] = (unsigned)(min(o2, 7) | (min(o2 + 1, 7) << 4));
        sel[3] = (unsigned)(min(o3, 7) | (min(o3 + 1, 7) << 4));
    }
    const int band = K.TH >> 2;
    const int y0 = dy0t + warp * band;
    const int y1 = min(min(y0 + band, dy0t + K.TH), K.dh);
    const int BW = K.BW, dpitch = K.dpitch;
    uint8_t* dp = a.pyr + (size_t)f * (size_t)K.pyrBytes + (size_t)K.doff + (size_t)y0 * dpitch + dx0;
    // rows of the tile are addressed by their source row index: the window bases and the tile origin are folded into two pointers
    const uint8_t* tA = tile + baseA - sy0 * BW;
    const uint8_t* tB = tile + (baseB == baseA ? baseA : baseA + 4) + 4 - sy0 * BW;   // the third word of the lane's 12-byte footprint
    const bool sameWin = baseB == baseA;
    __syncthreads();                                       // the barrier is initialised before anybody polls it
    mbar_wait(bar, 0);
    if (y0 >= y1) return;
    // the four columns' taps lie in 12 consecutive bytes (scale <= 2): three loads; the second window is words (0,1) or (1,2)
#define PYRT_H(R, H) do {                                                                              \
        const uint8_t* pa_ = tA + (R); const uint8_t* pb_ = tB + (R);                                     \
        const unsigned A0 = *reinterpret_cast<const unsigned*>(pa_), A1 = *reinterpret_cast<const unsigned*>(pa_ + 4); \
        const unsigned B1 = *reinterpret_cast<const unsigned*>(pb_);                                      \
        const unsigned B0 = sameWin ? A0 : A1;                                                            \
        H[0] = __dp2a_lo(wt[0], __byte_perm(A0, A1, sel[0]), 0u) >> 4;                                  \
        H[1] = __dp2a_lo(wt[1], __byte_perm(A0, A1, sel[1]), 0u) >> 4;                                  \
        H[2] = __dp2a_lo(wt[2], __byte_perm(B0, B1, sel[2]), 0u) >> 4;                                  \
        H[3] = __dp2a_lo(wt[3], __byte_perm(B0, B1, sel[3]), 0u) >> 4;                                  \
    } while (0)
    unsigned ha[4], hb[4];
    const int4* yp = ytab + y0;
    int4 yt = __ldg(yp);                                   // sy0, sy1 (clamped), b0 << 16, b1 << 16 -- warp-uniform
    int n = y1 - y0;
    PYRT_H(yt.x, ha);
    // two destination rows per trip: ha / hb swap roles, the kept row never moves between registers
    for (;;) {
        PYRT_H(yt.y, hb);
        {
            const unsigned o = pyrt_vrow(ha, hb, (unsigned)yt.z, (unsigned)yt.w);
            if (store) *reinterpret_cast<unsigned*>(dp) = o;
            dp += dpitch;
        }
        if (--n == 0) break;
        {
            const int kept = yt.y;
            yt = __ldg(++yp);
            if (yt.x != kept) PYRT_H(yt.x, hb);            // warp-uniform; one row in six at s = 1.2
        }
        PYRT_H(yt.y, ha);
        {
            const unsigned o = pyrt_vrow(hb, ha, (unsigned)yt.z, (unsigned)yt.w);
            if (store) *reinterpret_cast<unsigned*>(dp) = o;
            dp += dpitch;
        }
        if (--n == 0) break;
        {
            const int kept = yt.y;
            yt = __ldg(++yp);
            if (yt.x != kept) PYRT_H(yt.x, ha);
        }
    }
#undef PYRT_H
}

cudaError_t launch_pyr_tma(const OrbArgs& a, const OrbPlan& hp, int level, int nframes, const CUtensorMap& tmSrc, cudaStream_t st) {
    const LevelPlan& lp = hp.lv[level];
    PyrTileConst k;
    k.TW = lp.pyrTW; k.TH = lp.pyrTH; k.BW = lp.pyrBW; k.BH = lp.pyrBH;
    k.barOff = (lp.pyrBW * lp.pyrBH + 15) & ~15;
    k.dw = lp.w; k.dh = lp.h; k.dpitch = lp.pitch;
    k.xtabOff = lp.xtabOff; k.ytabOff = lp.ytabOff;
    k.doff = lp.off; k.pyrBytes = hp.pyrBytesPerFrame;
    dim3 grd((lp.w + lp.pyrTW - 1) / lp.pyrTW, (lp.h + lp.pyrTH - 1) / lp.pyrTH, nframes);
    pyr_tma_kernel<<<grd, 128, (size_t)k.barOff + 16, st>>>(a, tmSrc, k);
    return cudaGetLastError();
}

}  // namespace eorb
