// lk_kernels.cu — pyramidal Lucas-Kanade tracking of event-frame keypoints.
// Reference call: ELK_Tracker::trackCurrImage (src/Event/KLT_Tracker.cpp:49-91) -> cv::calcOpticalFlowPyrLK with
// win 23, maxLevel 1, 10 iterations, eps 0.03 (Examples/Event/EvETHZ.yaml:205-208).  Restates OpenCV's algorithm
// (see oracle/lk_oracle.cc for the step-by-step description and the pin against cv2):
//   lk_pyrdown_kernel  cv::pyrDown u8: [1 4 6 4 1] x [1 4 6 4 1], (sum + 128) >> 8, REFLECT_101
//   lk_scharr_kernel   calcSharrDeriv: (Ix, Iy) int16, REFLECT_101 inside the level
//   lk_track_kernel    LKTrackerInvoker: ONE BLOCK (128 threads) PER POINT walks all pyramid levels in one launch; the
//                      bilinear patch (14-bit fixed point) lives in shared memory as int16 triples (I, Ix, Iy), the
//                      current frame's window is staged in shared memory once per level; the five sums
//                      A11 A12 A22 b1 b2 are accumulated as EXACT integers (int64, shuffle + smem reduction) and
//                      converted to float once — order independent, hence bit-equal to the oracle, and within one
//                      float rounding of whichever lane order OpenCV's SIMD build uses; the scalar float algebra
//                      uses explicit no-FMA intrinsics in the oracle's operation order.
#include <cuda_runtime.h>
#include <stdint.h>

#include "eorb_math.cuh"
#include "lk_kernels.h"

namespace eorb {

__global__ void __launch_bounds__(256) lk_pyrdown_kernel(const uint8_t* __restrict__ src, int w, int h, int pitch, uint8_t* __restrict__ dst,
                                                         int dw, int dh, int dpitch) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const int wt[5] = {1, 4, 6, 4, 1};
    int xs[5];
#pragma unroll
    for (int j = 0; j < 5; j++) xs[j] = reflect101(2 * x + j - 2, w);
    int acc = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const uint8_t* r = src + (size_t)reflect101(2 * y + k - 2, h) * pitch;
        int rs = 0;
#pragma unroll
        for (int j = 0; j < 5; j++) rs += wt[j] * (int)__ldg(r + xs[j]);
        acc += wt[k] * rs;
    }
    dst[(size_t)y * dpitch + x] = (uint8_t)((acc + 128) >> 8);
}

__global__ void __launch_bounds__(256) lk_scharr_kernel(const uint8_t* __restrict__ src, int w, int h, int pitch, short2* __restrict__ dst) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t* r0 = src + (size_t)reflect101(y - 1, h) * pitch;
    const uint8_t* r1 = src + (size_t)y * pitch;
    const uint8_t* r2 = src + (size_t)reflect101(y + 1, h) * pitch;
    const int xl = reflect101(x - 1, w), xr = reflect101(x + 1, w);
    // vertical pass at the three columns: t0 = smoothing, t1 = difference
    const int t0l = ((int)__ldg(r0 + xl) + (int)__ldg(r2 + xl)) * 3 + (int)__ldg(r1 + xl) * 10;
    const int t0r = ((int)__ldg(r0 + xr) + (int)__ldg(r2 + xr)) * 3 + (int)__ldg(r1 + xr) * 10;
    const int t1l = (int)__ldg(r2 + xl) - (int)__ldg(r0 + xl);
    const int t1c = (int)__ldg(r2 + x) - (int)__ldg(r0 + x);
    const int t1r = (int)__ldg(r2 + xr) - (int)__ldg(r0 + xr);
    dst[(size_t)y * w + x] = make_short2((short)(t0r - t0l), (short)((t1r + t1l) * 3 + t1c * 10));
}

__device__ __forceinline__ int lk_descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

__device__ __forceinline__ void lk_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    const float oa = fsub(1.f, a), ob = fsub(1.f, b);
    w00 = round_rne(fmul(fmul(oa, ob), 16384.f));
    w01 = round_rne(fmul(fmul(a, ob), 16384.f));
    w10 = round_rne(fmul(fmul(oa, b), 16384.f));
    w11 = 16384 - w00 - w01 - w10;
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// bilinear sample of a u8 level with the REFLECT_101 border OpenCV pads its pyramid with, scaled by 32
__device__ __forceinline__ int lk_sample32(const LkLevelDev& L, const uint8_t* __restrict__ img, int X, int Y, int w00, int w01, int w10, int w11) {
    const int x0 = reflect101(X, L.w), x1 = reflect101(X + 1, L.w);
    const uint8_t* r0 = img + (size_t)reflect101(Y, L.h) * L.pitch;
    const uint8_t* r1 = img + (size_t)reflect101(Y + 1, L.h) * L.pitch;
    return lk_descale((int)__ldg(r0 + x0) * w00 + (int)__ldg(r0 + x1) * w01 + (int)__ldg(r1 + x0) * w10 + (int)__ldg(r1 + x1) * w11, 14 - 5);
}

__device__ __forceinline__ short2 lk_deriv_at(const LkLevelDev& L, int X, int Y) {   // zero border
    if (X < 0 || X >= L.w || Y < 0 || Y >= L.h) return make_short2(0, 0);
    return __ldg(L.dI + (size_t)Y * L.w + X);
}

#define LK_THREADS 128       // one block per point: the 23x23 patch is ~4 pixels per thread, so a Newton step is short
#define LK_MARGIN 4          // the current-frame window is staged with this many spare pixels on every side

// Stage the (win + 1 + 2*LK_MARGIN)^2 neighbourhood of the current frame around window origin (ox, oy) into shared
// memory with the REFLECT_101 border applied, all loads independent (one memory latency instead of one per Newton
// step and pixel).  Region origin = (ox - LK_MARGIN, oy - LK_MARGIN).
__device__ __forceinline__ void lk_stage_region(const LkLevelDev& L, const uint8_t* __restrict__ img, int ox, int oy, int S, uint8_t* reg) {
    const int x0 = ox - LK_MARGIN, y0 = oy - LK_MARGIN;
    for (int i = threadIdx.x; i < S * S; i += LK_THREADS) {
        const int y = i / S, x = i - y * S;
        reg[i] = __ldg(img + (size_t)reflect101(y0 + y, L.h) * L.pitch + reflect101(x0 + x, L.w));
    }
}

__device__ __forceinline__ int lk_sample32_smem(const uint8_t* reg, int S, int rx, int ry, int w00, int w01, int w10, int w11) {
    const uint8_t* r0 = reg + ry * S + rx;
    return lk_descale((int)r0[0] * w00 + (int)r0[1] * w01 + (int)r0[S] * w10 + (int)r0[S + 1] * w11, 14 - 5);
}

// exact block-wide sums of up to three int64 values; every thread gets the totals.  `slot` alternates between
// consecutive calls so that one __syncthreads per call is enough.
__device__ __forceinline__ void lk_block_sum(long long& a, long long& b, long long& c, long long (*scratch)[LK_THREADS / 32][3], int slot) {
    a = warp_sum_ll(a); b = warp_sum_ll(b); c = warp_sum_ll(c);
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { scratch[slot][warp][0] = a; scratch[slot][warp][1] = b; scratch[slot][warp][2] = c; }
    __syncthreads();
    a = b = c = 0;
#pragma unroll
    for (int w = 0; w < LK_THREADS / 32; w++) { a += scratch[slot][w][0]; b += scratch[slot][w][1]; c += scratch[slot][w][2]; }
}

__global__ void __launch_bounds__(LK_THREADS) lk_track_kernel(LkLevels LV, LkParams prm, const float2* __restrict__ prevPts,
                                                              float2* __restrict__ nextPts, int n, uint8_t* __restrict__ status,
                                                              float* __restrict__ err) {
    extern __shared__ short s_patch[];   // win*win x (I, Ix, Iy), then the staged window of the current frame
    __shared__ long long s_scratch[2][LK_THREADS / 32][3];
    const int tid = threadIdx.x;
    const int p = blockIdx.x;
    const int win = prm.win, area = win * win;
    short* pI = s_patch;
    short* pIx = pI + area;
    short* pIy = pIx + area;
    const int S = win + 1 + 2 * LK_MARGIN;
    uint8_t* reg = reinterpret_cast<uint8_t*>(s_patch + (size_t)area * 3);
    const float halfWin = fmul((float)(win - 1), 0.5f);
    const float FLT_SCALE = 1.f / (1 << 20);
    const float2 pp = prevPts[p];
    float2 np = nextPts[p];          // only meaningful with useInitialFlow
    bool ok = true;                  // status[p]
    float errv = 0.f;
    int slot = 0;
    long long z = 0;

    // every decision below depends only on block-uniform values (the sums are exact integers), so all threads take
    // the same path through the level / iteration structure
    for (int level = LV.maxLevel; level >= 0; level--) {
        const LkLevelDev& L = LV.lv[level];
        const float sc = (float)(1. / (double)(1 << level));
        float px = fmul(pp.x, sc), py = fmul(pp.y, sc);
        float nx, ny;
        if (level == LV.maxLevel) {
            if (prm.useInitialFlow) { nx = fmul(np.x, sc); ny = fmul(np.y, sc); }
            else { nx = px; ny = py; }
        } else {
            nx = fmul(np.x, 2.f); ny = fmul(np.y, 2.f);
        }
        np = make_float2(nx, ny);
        px = fsub(px, halfWin); py = fsub(py, halfWin);
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -win || ipx >= L.w || ipy < -win || ipy >= L.h) {
            if (level == 0) { ok = false; errv = 0.f; }
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(fsub(px, (float)ipx), fsub(py, (float)ipy), w00, w01, w10, w11);
        long long sA11 = 0, sA12 = 0, sA22 = 0;
        __syncthreads();                 // the previous level is done with the patch and the staged window
#pragma unroll 2
        for (int i = tid; i < area; i += LK_THREADS) {
            const int y = i / win, x = i - y * win;
            const int X = ipx + x, Y = ipy + y;
            const int ival = lk_sample32(L, L.I, X, Y, w00, w01, w10, w11);
            const short2 d00 = lk_deriv_at(L, X, Y), d01 = lk_deriv_at(L, X + 1, Y), d10 = lk_deriv_at(L, X, Y + 1), d11 = lk_deriv_at(L, X + 1, Y + 1);
            const int ixval = lk_descale((int)d00.x * w00 + (int)d01.x * w01 + (int)d10.x * w10 + (int)d11.x * w11, 14);
            const int iyval = lk_descale((int)d00.y * w00 + (int)d01.y * w01 + (int)d10.y * w10 + (int)d11.y * w11, 14);
            pI[i] = (short)ival; pIx[i] = (short)ixval; pIy[i] = (short)iyval;
            sA11 += (long long)ixval * ixval; sA12 += (long long)ixval * iyval; sA22 += (long long)iyval * iyval;
        }
        lk_block_sum(sA11, sA12, sA22, s_scratch, slot); slot ^= 1;     // also publishes the patch (one barrier inside)
        const float A11 = fmul(__ll2float_rn(sA11), FLT_SCALE);
        const float A12 = fmul(__ll2float_rn(sA12), FLT_SCALE);
        const float A22 = fmul(__ll2float_rn(sA22), FLT_SCALE);
        float D = fsub(fmul(A11, A22), fmul(A12, A12));
        const float dif = fsub(A11, A22);
        const float minEig = fdiv(fsub(fadd(A22, A11), __fsqrt_rn(fadd(fmul(dif, dif), fmul(fmul(4.f, A12), A12)))), (float)(2 * win * win));
        if (minEig < prm.minEigThreshold || D < 1.1920929e-07f) {
            if (level == 0) ok = false;
            continue;
        }
        D = fdiv(1.f, D);
        nx = fsub(nx, halfWin); ny = fsub(ny, halfWin);
        float pdx = 0.f, pdy = 0.f;
        int rox = 0, roy = 0;            // window origin the staged region was centred on
        bool staged = false;
        for (int j = 0; j < prm.maxIter; j++) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -win || inx >= L.w || iny < -win || iny >= L.h) {
                if (level == 0) ok = false;
                break;
            }
            if (!staged || inx < rox - LK_MARGIN || inx > rox + LK_MARGIN || iny < roy - LK_MARGIN || iny > roy + LK_MARGIN) {
                __syncthreads();
                lk_stage_region(L, L.J, inx, iny, S, reg);
                rox = inx; roy = iny; staged = true;
                __syncthreads();
            }
            lk_weights(fsub(nx, (float)inx), fsub(ny, (float)iny), w00, w01, w10, w11);
            long long sb1 = 0, sb2 = 0;
            const int bx = inx - rox + LK_MARGIN, by = iny - roy + LK_MARGIN;
#pragma unroll 2
            for (int i = tid; i < area; i += LK_THREADS) {
                const int y = i / win, x = i - y * win;
                const int diff = lk_sample32_smem(reg, S, bx + x, by + y, w00, w01, w10, w11) - (int)pI[i];
                sb1 += (long long)diff * (int)pIx[i]; sb2 += (long long)diff * (int)pIy[i];
            }
            z = 0;
            lk_block_sum(sb1, sb2, z, s_scratch, slot); slot ^= 1;
            const float b1 = fmul(__ll2float_rn(sb1), FLT_SCALE), b2 = fmul(__ll2float_rn(sb2), FLT_SCALE);
            const float dx = fmul(fsub(fmul(A12, b2), fmul(A22, b1)), D), dy = fmul(fsub(fmul(A12, b1), fmul(A11, b2)), D);
            nx = fadd(nx, dx); ny = fadd(ny, dy);
            np = make_float2(fadd(nx, halfWin), fadd(ny, halfWin));
            if (__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= prm.epsilon2) break;
            if (j > 0 && (double)fabsf(fadd(dx, pdx)) < 0.01 && (double)fabsf(fadd(dy, pdy)) < 0.01) {
                np.x = fsub(np.x, fmul(dx, 0.5f)); np.y = fsub(np.y, fmul(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (ok && level == 0) {   // L1 patch error at the final position
            const float ex = fsub(np.x, halfWin), ey = fsub(np.y, halfWin);
            const int iex = (int)floorf(ex), iey = (int)floorf(ey);
            if (iex < -win || iex >= L.w || iey < -win || iey >= L.h) { ok = false; continue; }
            lk_weights(fsub(ex, (float)iex), fsub(ey, (float)iey), w00, w01, w10, w11);
            long long se = 0, z1 = 0, z2 = 0;
            const bool inReg = staged && iex >= rox - LK_MARGIN && iex <= rox + LK_MARGIN && iey >= roy - LK_MARGIN && iey <= roy + LK_MARGIN;
            const int bx = iex - rox + LK_MARGIN, by = iey - roy + LK_MARGIN;
            for (int i = tid; i < area; i += LK_THREADS) {
                const int y = i / win, x = i - y * win;
                const int v = inReg ? lk_sample32_smem(reg, S, bx + x, by + y, w00, w01, w10, w11)
                                    : lk_sample32(L, L.J, iex + x, iey + y, w00, w01, w10, w11);
                const int diff = v - (int)pI[i];
                se += diff < 0 ? -diff : diff;
            }
            lk_block_sum(se, z1, z2, s_scratch, slot); slot ^= 1;
            errv = fdiv(fmul(__ll2float_rn(se), 1.f), (float)(32 * win * win));
        }
    }
    if (tid == 0) {
        nextPts[p] = np;
        status[p] = ok ? 1 : 0;
        if (err) err[p] = errv;
    }
}

// ------------------------------------------------------------------------------------------------ post-LK bookkeeping
// ELK_Tracker::refineTrackedPts (src/Event/KLT_Tracker.cpp:105-155) and refineFirstOctaveLevel (:157-183) for one LK result, ONE block:
//   tracked[i] = KeyPoint(curr[i], ref[i].size, ref[i].angle, ref[i].response, ref[i].octave, ref[i].class_id)        (:138)
//   matched[i] bit 0 = status[i] == 1 && 0 <= x < W && 0 <= y < H (isInImage :99-102, the test of :140); bit 1 = bit 0 and not un-matched
//                by the first-octave filter
//   pxDisp     = the matched points' sqrtf(dx*dx + dy*dy) in index order (push_back :147) -- an ordered compaction
//   counts[0]  = nMatches after the optional first-octave filter (ref octave > 0 un-matches, :171-175), counts[1] = pxDisp entries
// (the filter runs after refineTrackedPts in trackAndMatchCurrImageInit :236-242, so pxDisp keeps the filtered matches' entries).
#define LK_REFINE_THREADS 256
__global__ void __launch_bounds__(LK_REFINE_THREADS) lk_refine_kernel(const float2* __restrict__ curr, const uint8_t* __restrict__ status,
                                                                      const eorb_keypoint* __restrict__ ref, int n, int W, int H,
                                                                      int firstOctaveOnly, eorb_keypoint* __restrict__ tracked,
                                                                      uint8_t* __restrict__ matched, float* __restrict__ pxDisp,
                                                                      int* __restrict__ counts) {
    __shared__ int s_warp[LK_REFINE_THREADS / 32];
    __shared__ int s_base, s_kept;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_base = 0; s_kept = 0; }
    __syncthreads();
    const float fw = (float)W, fh = (float)H;
    for (int i0 = 0; i0 < n; i0 += LK_REFINE_THREADS) {
        const int i = i0 + tid;
        bool m = false, keep = false;
        float disp = 0.f;
        if (i < n) {
            const float2 c = curr[i];
            eorb_keypoint k = ref[i];
            const float dx = fsub(c.x, k.x), dy = fsub(c.y, k.y);
            m = status[i] == 1 && c.x >= 0.f && c.x < fw && c.y >= 0.f && c.y < fh;
            keep = m && !(firstOctaveOnly && k.octave > 0);
            disp = __fsqrt_rn(fadd(fmul(dx, dx), fmul(dy, dy)));
            k.x = c.x; k.y = c.y;
            tracked[i] = k;
            matched[i] = (uint8_t)((m ? 1 : 0) | (keep ? 2 : 0));
        }
        const unsigned bal = __ballot_sync(0xffffffffu, m);
        const unsigned kbal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; w++) before += s_warp[w];
        if (m) pxDisp[before + __popc(bal & ((1u << lane) - 1u))] = disp;
        if (lane == 0 && kbal) atomicAdd(&s_kept, __popc(kbal));
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < LK_REFINE_THREADS / 32; w++) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
    if (tid == 0) { counts[0] = s_kept; counts[1] = s_base; }
}

cudaError_t launch_lk_refine(const float2* curr, const uint8_t* status, const eorb_keypoint* ref, int n, int W, int H, int firstOctaveOnly,
                             eorb_keypoint* tracked, uint8_t* matched, float* pxDisp, int* counts, cudaStream_t st) {
    lk_refine_kernel<<<1, LK_REFINE_THREADS, 0, st>>>(curr, status, ref, n, W, H, firstOctaveOnly, tracked, matched, pxDisp, counts);
    return cudaGetLastError();
}

cudaError_t launch_lk_pyrdown(const uint8_t* src, int w, int h, int pitch, uint8_t* dst, int dw, int dh, int dpitch, cudaStream_t st) {
    dim3 blk(32, 8), grd((dw + 31) / 32, (dh + 7) / 8);
    lk_pyrdown_kernel<<<grd, blk, 0, st>>>(src, w, h, pitch, dst, dw, dh, dpitch);
    return cudaGetLastError();
}

cudaError_t launch_lk_scharr(const uint8_t* src, int w, int h, int pitch, short2* dst, cudaStream_t st) {
    dim3 blk(32, 8), grd((w + 31) / 32, (h + 7) / 8);
    lk_scharr_kernel<<<grd, blk, 0, st>>>(src, w, h, pitch, dst);
    return cudaGetLastError();
}

cudaError_t launch_lk_track(const LkLevels& L, const LkParams& p, const float2* prevPts, float2* nextPts, int n, uint8_t* status, float* err,
                            cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int S = p.win + 1 + 2 * LK_MARGIN;
    const size_t smem = (size_t)p.win * p.win * 3 * sizeof(short) + (size_t)S * S;
    lk_track_kernel<<<n, LK_THREADS, smem, st>>>(L, p, prevPts, nextPts, n, status, err);
    return cudaGetLastError();
}

}  // namespace eorb
