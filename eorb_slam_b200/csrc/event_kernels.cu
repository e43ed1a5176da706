// event_kernels.cu — event windows -> event frames (reference src/Event/EventConversion.cc):
//   ev_splat_nearest_kernel   ev2im                      :173-212
//   ev_splat_gauss_kernel     ev2im_gauss                :215-269
//                             ev2mci_gg_f(Tcw,medDepth)  :279-360   (per-event rigid warp, double geometry)
//                             ev2mci_gg_f(params2D)      :362-448   (per-event SE2(+scale) warp, float)
//   ev_minmax_kernel + ev_normalize_kernel   normalizeImage :67-72 / cv::normalize(NORM_MINMAX, CV_8UC1)
//
// Accumulation model: the frames of all windows of a batch live in HBM/L2 and every splat tap is one
// fire-and-forget fp32 reduction (RED.E.ADD.F32) — shared-memory fp32 atomics are CAS loops on this
// architecture and a 346x260 frame does not fit next to them anyway.  A group of LPE lanes owns one event:
// lane i handles column xi+i-half and walks the rows, so the lanes of a group hit consecutive addresses of
// one image row (one or two 32-byte sectors per warp-wide reduction).  Event records are read in the
// reference's 24-byte AoS layout.  Sums are order-dependent in fp32, hence toleranced parity (<= 1e-4*peak).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eorb_b200.h"
#include "event_kernels.h"

namespace eorb {

__device__ __forceinline__ void splat_taps(float* __restrict__ im, int W, int H, float X, float Y, float polSign,
                                           const EvConst& c, int sub, int lpe) {
    const float fxi = floorf(X), fyi = floorf(Y);
    const int half = c.half;
    // positions whose window cannot touch the frame — and NaN / inf (e.g. the SE2 warp of a single-timestamp window divides by
    // DT = 0; the reference's float -> int conversion sends those far outside the image) — contribute nothing
    if (!(fxi >= (float)(-half - 1) && fxi <= (float)(W + half) && fyi >= (float)(-half - 1) && fyi <= (float)(H + half))) return;
    const int xi = (int)fxi, yi = (int)fyi;
    const float xr = __fsub_rn(X, (float)xi), yr = __fsub_rn(Y, (float)yi);
    const float den = __fmul_rn(2.0f, c.sig2);
    for (int ii = sub; ii <= 2 * half; ii += lpe) {
        const int i = ii - half;
        const int xn = xi + i;
        if (xn < 0 || xn >= W) continue;
        const float dx = __fsub_rn((float)i, xr);
        const float dx2 = __fmul_rn(dx, dx);
        for (int j = -half; j <= half; j++) {
            const int yn = yi + j;
            if (yn < 0 || yn >= H) continue;
            const float dy = __fsub_rn((float)j, yr);
            const float dd = __fdiv_rn(__fadd_rn(dx2, __fmul_rn(dy, dy)), den);
            const float val = __fdiv_rn(expf(-dd), c.norm);
            atomicAdd(im + (size_t)yn * W + xn, __fmul_rn(polSign, val));
        }
    }
}

// ---- KannalaBrandt8 (src/CameraModels/KannalaBrandt8.cpp): unproject :163-190 (ten Newton steps on theta in float, precision 1e-6), project
// (cv::Point3f) :86-103, project (Eigen::Vector3d) :111-129.  Float operation order as in the reference, no contraction; atan2f / tanf / cos
// / sin are CUDA's (within 2 ulp of glibc's), which moves a warped position by < 1e-5 px -- inside the 1e-4 * peak parity bar of the frames.
__device__ __forceinline__ void kb8_unproject(const EvConst& c, float x, float y, float& X, float& Y) {
    const float pwx = __fdiv_rn(__fsub_rn(x, c.cx), c.fx), pwy = __fdiv_rn(__fsub_rn(y, c.cy), c.fy);
    float scale = 1.f;
    float theta_d = sqrtf(__fadd_rn(__fmul_rn(pwx, pwx), __fmul_rn(pwy, pwy)));
    theta_d = fminf(fmaxf(-1.57079637f, theta_d), 1.57079637f);
    if ((double)theta_d > 1e-8) {
        float theta = theta_d;
        for (int j = 0; j < 10; j++) {
            const float theta2 = __fmul_rn(theta, theta), theta4 = __fmul_rn(theta2, theta2), theta6 = __fmul_rn(theta4, theta2),
                        theta8 = __fmul_rn(theta4, theta4);
            const float k0 = __fmul_rn(c.kb[0], theta2), k1 = __fmul_rn(c.kb[1], theta4), k2 = __fmul_rn(c.kb[2], theta6), k3 = __fmul_rn(c.kb[3], theta8);
            const float num = __fsub_rn(__fmul_rn(theta, __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(1.f, k0), k1), k2), k3)), theta_d);
            const float den = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(1.f, __fmul_rn(3.f, k0)), __fmul_rn(5.f, k1)), __fmul_rn(7.f, k2)), __fmul_rn(9.f, k3));
            const float fix = __fdiv_rn(num, den);
            theta = __fsub_rn(theta, fix);
            if (fabsf(fix) < 1e-6f) break;
        }
        scale = __fdiv_rn(tanf(theta), theta_d);
    }
    X = __fmul_rn(pwx, scale); Y = __fmul_rn(pwy, scale);
}
__device__ __forceinline__ void kb8_project_d(const EvConst& c, double x, double y, double z, float& U, float& V) {
    const double x2y2 = x * x + y * y;
    const double theta = (double)atan2f(sqrtf((float)x2y2), (float)z);
    const double psi = (double)atan2f((float)y, (float)x);
    const double t2 = theta * theta, t3 = theta * t2, t5 = t3 * t2, t7 = t5 * t2, t9 = t7 * t2;
    const double r = theta + (double)c.kb[0] * t3 + (double)c.kb[1] * t5 + (double)c.kb[2] * t7 + (double)c.kb[3] * t9;
    double sp, cp;
    sincos(psi, &sp, &cp);
    U = (float)((double)c.fx * r * cp + (double)c.cx);
    V = (float)((double)c.fy * r * sp + (double)c.cy);
}
__device__ __forceinline__ void kb8_project_f(const EvConst& c, float x, float y, float z, float& U, float& V) {
    const float x2y2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
    const float theta = atan2f(sqrtf(x2y2), z);
    const float psi = atan2f(y, x);
    const float t2 = __fmul_rn(theta, theta), t3 = __fmul_rn(theta, t2), t5 = __fmul_rn(t3, t2), t7 = __fmul_rn(t5, t2), t9 = __fmul_rn(t7, t2);
    const float r = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(theta, __fmul_rn(c.kb[0], t3)), __fmul_rn(c.kb[1], t5)), __fmul_rn(c.kb[2], t7)), __fmul_rn(c.kb[3], t9));
    double sp, cp;
    sincos((double)psi, &sp, &cp);                       // the reference's cos(psi) / sin(psi) are the double functions on the float angle
    U = (float)((double)__fmul_rn(c.fx, r) * cp + (double)c.cx);
    V = (float)((double)__fmul_rn(c.fy, r) * sp + (double)c.cy);
}

// per-event warp of the SE3 (:279-360) and SE2 (:362-448) overloads; X, Y come in as the event position
__device__ __forceinline__ void ev_warp_point(const eorb_event* __restrict__ evs, const EvWindow& w, const EvConst& c, double ts, float ex,
                                              float ey, float& X, float& Y) {
    if (c.mode == EORB_EV_SE3) {
        const double t1 = evs[w.end - 1].ts, DT = t1 - evs[w.begin].ts;
        const double rate = DT > 0 ? (t1 - ts) * (1.0 / DT) : 0.0;
        float Xs, Ys;
        if (c.cam == 1) kb8_unproject(c, ex, ey, Xs, Ys);
        else { Xs = __fdiv_rn(__fsub_rn(ex, c.cx), c.fx); Ys = __fdiv_rn(__fsub_rn(ey, c.cy), c.fy); }
        const double P0 = (double)Xs, P1 = (double)Ys, P2 = 1.0;
        // Eigen::AngleAxisd(angle*rate, axis).toRotationMatrix()
        const double ang = w.angle * rate;
        double s, co;
        sincos(ang, &s, &co);
        const double a0 = w.axis[0], a1 = w.axis[1], a2 = w.axis[2];
        const double s0 = s * a0, s1 = s * a1, s2 = s * a2;
        const double c0 = (1 - co) * a0, c1 = (1 - co) * a1, c2 = (1 - co) * a2;
        double t;
        double R01, R10, R02, R20, R12, R21;
        t = c0 * a1; R01 = t - s2; R10 = t + s2;
        t = c0 * a2; R02 = t + s1; R20 = t - s1;
        t = c1 * a2; R12 = t - s0; R21 = t + s0;
        const double R00 = c0 * a0 + co, R11 = c1 * a1 + co, R22 = c2 * a2 + co;
        const double dep = (double)c.depth;
        const double n0 = (dep * R00) * P0 + (dep * R01) * P1 + (dep * R02) * P2 + w.t[0] * rate;
        const double n1 = (dep * R10) * P0 + (dep * R11) * P1 + (dep * R12) * P2 + w.t[1] * rate;
        const double n2 = (dep * R20) * P0 + (dep * R21) * P1 + (dep * R22) * P2 + w.t[2] * rate;
        if (c.cam == 1) kb8_project_d(c, n0, n1, n2, X, Y);
        else {
            X = (float)((double)c.fx * n0 / n2 + (double)c.cx);
            Y = (float)((double)c.fy * n1 / n2 + (double)c.cy);
        }
    } else if (c.mode == EORB_EV_SE2) {
        const double t1 = evs[w.end - 1].ts;
        const float DT = (float)(t1 - evs[w.begin].ts);
        const float invDT = __fdiv_rn(1.f, DT);
        const float omega0 = __fmul_rn(c.se2[0], invDT), vx0 = __fmul_rn(c.se2[1], invDT), vy0 = __fmul_rn(c.se2[2], invDT);
        const float sc = c.se2_n > 3 ? c.se2[3] : 1.f;
        const float scDiff = __fsub_rn(1.f, sc);
        const float tk = (float)(t1 - ts);
        float Xs, Ys;
        if (c.cam == 1) kb8_unproject(c, ex, ey, Xs, Ys);
        else { Xs = __fdiv_rn(__fsub_rn(ex, c.cx), c.fx); Ys = __fdiv_rn(__fsub_rn(ey, c.cy), c.fy); }
        const float th = __fmul_rn(tk, omega0);
        const float cs = __fadd_rn(__fmul_rn(scDiff, __fsub_rn(1.f, __fmul_rn(tk, invDT))), sc);
        const float ct = cosf(th), st = sinf(th);
        const float xp = __fadd_rn(__fmul_rn(cs, __fsub_rn(__fmul_rn(Xs, ct), __fmul_rn(Ys, st))), __fmul_rn(vx0, tk));
        const float yp = __fadd_rn(__fmul_rn(cs, __fadd_rn(__fmul_rn(Xs, st), __fmul_rn(Ys, ct))), __fmul_rn(vy0, tk));
        if (c.cam == 1) kb8_project_f(c, xp, yp, 1.f, X, Y);
        else {
            X = __fadd_rn(__fdiv_rn(__fmul_rn(c.fx, xp), 1.f), c.cx);
            Y = __fadd_rn(__fdiv_rn(__fmul_rn(c.fy, yp), 1.f), c.cy);
        }
    }
}

__global__ void __launch_bounds__(256) ev_splat_gauss_kernel(const eorb_event* __restrict__ evs,
                                                             const EvWindow* __restrict__ wins, EvConst c,
                                                             float* __restrict__ img, int lpe) {
    const EvWindow w = wins[blockIdx.y];
    const long long nev = w.end - w.begin;
    const int perBlock = blockDim.x / lpe;
    const long long e = (long long)blockIdx.x * perBlock + threadIdx.x / lpe;
    const int sub = threadIdx.x % lpe;
    if (e >= nev) return;
    const eorb_event* evp = evs + w.begin + e;
    const double ts = evp->ts;
    const float ex = evp->x, ey = evp->y;
    const bool p = evp->p != 0;
    const float polSign = (c.pol && !p) ? -1.0f : 1.0f;
    float X = ex, Y = ey;
    ev_warp_point(evs, w, c, ts, ex, ey, X, Y);
    float* im = img + (size_t)blockIdx.y * (size_t)c.width * c.height;
    splat_taps(im, c.width, c.height, X, Y, polSign, c, sub, lpe);
}

// ---- shared-memory event frames (sigma with ceil(3*sigma) == 3, i.e. the 7x7 splat every live caller uses) ----------
// One block per (window, row band); the band of the frame lives in shared memory as int32 fixed point, so a tap is
// a NATIVE shared-memory atomic (ATOMS.ADD; fp32 shared atomics are CAS loops) and the sum is order independent.
// Scale 2^k with k = min(24, floor(log2(2^31 / (events * peakTap)))): no pixel can overflow whatever the event
// distribution is.  One thread per event: warp geometry once, the Gaussian is evaluated separably
// (7 + 7 expf instead of 49; exp(-(dx^2+dy^2)/2s^2) = exp(-dx^2/2s^2) * exp(-dy^2/2s^2), a few ulp from the
// reference's single expf, far inside the 1e-4*peak parity tolerance).  The band is written out once as fp32
// (coalesced) and, when the frame is a single band, min/max + the u8 normalisation are fused into the same block.
#define EV_SMEM_HALF 3
#ifndef EV_SMEM_THREADS
#define EV_SMEM_THREADS 512   // measured on B200 (40-step runs): 1024 -> 15.4, 768 -> 15.7, 512 -> 16.0 Gev/s
#endif
#define EV_SMEM_PAD 8

__device__ __forceinline__ void ev_norm_coeffs(int normMode, float mn, float mx, float& alpha, float& beta) {
    if (normMode == EORB_NORM_RUNNING) {
        // running min starts at 0, running max at -1e6 (EventConversion.cc:219-220)
        mn = fminf(mn, 0.0f);
        if (!(mx > mn)) { alpha = 0.f; beta = 0.f; }
        else { alpha = __fdiv_rn(255.f, __fsub_rn(mx, mn)); beta = __fmul_rn(-mn, alpha); }
    } else {
        const double d = (double)mx - (double)mn;
        const double scale = 255.0 * (d > 2.220446049250313e-16 ? 1.0 / d : 0.0);
        alpha = (float)scale; beta = (float)(0.0 - (double)mn * scale);
    }
}

// warped position of every event of every window, once (used when a frame is split into row bands: every band block would
// otherwise repeat the warp, and the SE3 warp is ~300 double-precision instructions per event)
__global__ void __launch_bounds__(256) ev_warp_kernel(const eorb_event* __restrict__ evs, const EvWindow* __restrict__ wins, EvConst c,
                                                      float2* __restrict__ xy) {
    const EvWindow w = wins[blockIdx.y];
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= w.end - w.begin) return;
    const eorb_event* evp = evs + w.begin + e;
    const float ex = evp->x, ey = evp->y;
    float X = ex, Y = ey;
    ev_warp_point(evs, w, c, evp->ts, ex, ey, X, Y);
    xy[w.begin + e] = make_float2(X, Y);
}

__global__ void __launch_bounds__(EV_SMEM_THREADS) ev_frame_smem_kernel(const eorb_event* __restrict__ evs, const EvWindow* __restrict__ wins,
                                                                        EvConst c, int bandRows, int fuseNorm, int normMode,
                                                                        float* __restrict__ img, float* __restrict__ minmax,
                                                                        uint8_t* __restrict__ u8, const float2* __restrict__ xy) {
    // the band is preceded and followed by EV_SMEM_PAD ints: taps that fall outside the frame add 0 to a clamped row at an
    // unclamped column, which may run up to 6 ints past either end of the band (see the tap loop)
    extern __shared__ __align__(16) int s_raw[];
    int* const s_acc = s_raw + EV_SMEM_PAD;
    __shared__ float s_mn[32], s_mx[32];
    const EvWindow w = wins[blockIdx.y];
    const int W = c.width, H = c.height;
    const int r0 = blockIdx.x * bandRows, r1 = min(r0 + bandRows, H);
    const int nrow = r1 - r0, npx = nrow * W;
    const int tid = threadIdx.x;
    for (int i = tid; i < (npx + 3) >> 2; i += EV_SMEM_THREADS) reinterpret_cast<int4*>(s_acc)[i] = make_int4(0, 0, 0, 0);   // band is padded to 16 B
    const long long nev = w.end - w.begin;
    // fixed-point scale: |pixel sum| <= nev * peakTap, peakTap = 1 / (2*pi*sigma^2)
    const float peakTap = __fdiv_rn(1.0f, c.norm);
    int k = (int)floorf(log2f(2147483648.0f / ((float)nev * peakTap + 1.0f)));
    k = min(max(k, 0), 24);
    const float scale = exp2f((float)k), invScale = exp2f(-(float)k);
    const float invDen = __fdiv_rn(1.0f, __fmul_rn(2.0f, c.sig2));
    __syncthreads();

    for (long long e = tid; e < nev; e += EV_SMEM_THREADS) {
        const eorb_event* evp = evs + w.begin + e;
        const float polSign = (c.pol && evp->p == 0) ? -1.0f : 1.0f;
        float X, Y;
        if (xy) {   // multi-band warped frames: the per-event warp (double precision for SE3) was done once by ev_warp_kernel
            const float2 v = __ldg(xy + w.begin + e);
            X = v.x; Y = v.y;
        } else {
            const double ts = evp->ts;
            const float ex = evp->x, ey = evp->y;
            X = ex; Y = ey;
            ev_warp_point(evs, w, c, ts, ex, ey, X, Y);
        }
        const float fxi = floorf(X), fyi = floorf(Y);
        if (!(fxi >= -8.f && fxi <= (float)(W + 8) && fyi >= -8.f && fyi <= (float)(H + 8))) continue;   // also NaN / inf
        const int xi = (int)fxi, yi = (int)fyi;
        if (yi + EV_SMEM_HALF < r0 || yi - EV_SMEM_HALF >= r1) continue;
        if (xi + EV_SMEM_HALF < 0 || xi - EV_SMEM_HALF >= W) continue;          // no tap inside the frame
        const float xr = __fsub_rn(X, fxi), yr = __fsub_rn(Y, fyi);
        float gx[2 * EV_SMEM_HALF + 1], gy[2 * EV_SMEM_HALF + 1];
        const float amp = polSign * peakTap * scale;
        // Branch-free taps: a column / row outside the frame gets a ZERO factor instead of a test around its atomics, and
        // its row pointer is clamped into the band, so every tap is FMUL + F2I + ATOMS at a compile-time offset (adding 0
        // changes nothing; the uniform form keeps the 32 events of a warp converged).
#pragma unroll
        for (int i = -EV_SMEM_HALF; i <= EV_SMEM_HALF; i++) {
            const float dx = __fsub_rn((float)i, xr), dy = __fsub_rn((float)i, yr);
            const float ex_ = expf(-(dx * dx) * invDen) * amp, ey_ = expf(-(dy * dy) * invDen);
            gx[i + EV_SMEM_HALF] = (unsigned)(xi + i) < (unsigned)W ? ex_ : 0.0f;
            gy[i + EV_SMEM_HALF] = (yi + i >= r0 && yi + i < r1) ? ey_ : 0.0f;
        }
#pragma unroll
        for (int j = -EV_SMEM_HALF; j <= EV_SMEM_HALF; j++) {
            int* row = s_acc + (min(max(yi + j, r0), r1 - 1) - r0) * W + xi;
#pragma unroll
            for (int i = -EV_SMEM_HALF; i <= EV_SMEM_HALF; i++)
                atomicAdd(row + i, __float2int_rn(gx[i + EV_SMEM_HALF] * gy[j + EV_SMEM_HALF]));
        }
    }
    __syncthreads();

    // ---- write the band out as fp32, track min / max (four pixels per thread and step when the band is 16-byte aligned)
    float* im = img + (size_t)blockIdx.y * (size_t)W * H + (size_t)r0 * W;
    float mn = 3.4e38f, mx = -3.4e38f;
    const bool vec = ((npx & 3) == 0) && ((reinterpret_cast<uintptr_t>(im) & 15) == 0);
    if (vec) {
        for (int i = tid; i < npx >> 2; i += EV_SMEM_THREADS) {
            const int4 a = reinterpret_cast<const int4*>(s_acc)[i];
            const float4 v = make_float4((float)a.x * invScale, (float)a.y * invScale, (float)a.z * invScale, (float)a.w * invScale);
            reinterpret_cast<float4*>(im)[i] = v;
            mn = fminf(fminf(mn, fminf(v.x, v.y)), fminf(v.z, v.w)); mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
        }
    } else {
        for (int i = tid; i < npx; i += EV_SMEM_THREADS) {
            const float v = (float)s_acc[i] * invScale;
            im[i] = v;
            mn = fminf(mn, v); mx = fmaxf(mx, v);
        }
    }
    if (!fuseNorm) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((tid & 31) == 0) { s_mn[tid >> 5] = mn; s_mx[tid >> 5] = mx; }
    __syncthreads();
    {   // second stage over the block's warps (fewer than 32 when the block has fewer than 1024 threads)
        const bool has = (tid & 31) < EV_SMEM_THREADS / 32;
        mn = has ? s_mn[tid & 31] : 3.4e38f; mx = has ? s_mx[tid & 31] : -3.4e38f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (tid == 0) { minmax[2 * blockIdx.y] = mn; minmax[2 * blockIdx.y + 1] = mx; }
    if (normMode == EORB_NORM_NONE || !u8) return;
    float alpha, beta;
    ev_norm_coeffs(normMode, mn, mx, alpha, beta);
    uint8_t* o8 = u8 + (size_t)blockIdx.y * (size_t)W * H;
    if (((npx & 3) == 0) && ((reinterpret_cast<uintptr_t>(o8) & 3) == 0)) {
        for (int i = tid; i < npx >> 2; i += EV_SMEM_THREADS) {
            const int4 a = reinterpret_cast<const int4*>(s_acc)[i];
            const int q0 = min(max(__float2int_rn(__fadd_rn(__fmul_rn((float)a.x * invScale, alpha), beta)), 0), 255);
            const int q1 = min(max(__float2int_rn(__fadd_rn(__fmul_rn((float)a.y * invScale, alpha), beta)), 0), 255);
            const int q2 = min(max(__float2int_rn(__fadd_rn(__fmul_rn((float)a.z * invScale, alpha), beta)), 0), 255);
            const int q3 = min(max(__float2int_rn(__fadd_rn(__fmul_rn((float)a.w * invScale, alpha), beta)), 0), 255);
            reinterpret_cast<uint32_t*>(o8)[i] = (uint32_t)q0 | ((uint32_t)q1 << 8) | ((uint32_t)q2 << 16) | ((uint32_t)q3 << 24);
        }
    } else {
        for (int i = tid; i < npx; i += EV_SMEM_THREADS) {
            const float v = (float)s_acc[i] * invScale;
            const int r = __float2int_rn(__fadd_rn(__fmul_rn(v, alpha), beta));
            o8[i] = (uint8_t)min(max(r, 0), 255);
        }
    }
}

__global__ void __launch_bounds__(256) ev_splat_nearest_kernel(const eorb_event* __restrict__ evs,
                                                               const EvWindow* __restrict__ wins, EvConst c,
                                                               float* __restrict__ img) {
    const EvWindow w = wins[blockIdx.y];
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= w.end - w.begin) return;
    const eorb_event* evp = evs + w.begin + e;
    const float polSign = (c.pol && evp->p == 0) ? -1.0f : 1.0f;
    const int px = (int)roundf(evp->x), py = (int)roundf(evp->y);
    if (px < 0 || px >= c.width || py < 0 || py >= c.height) return;
    float* im = img + (size_t)blockIdx.y * (size_t)c.width * c.height;
    atomicAdd(im + (size_t)py * c.width + px, __fmul_rn(polSign, 0.001f));
}

// one block per window: (min, max) over the whole frame
__global__ void __launch_bounds__(1024) ev_minmax_kernel(const float* __restrict__ img, int npix, float* __restrict__ minmax) {
    __shared__ float s_mn[32], s_mx[32];
    const float* im = img + (size_t)blockIdx.x * npix;
    float mn = 3.4e38f, mx = -3.4e38f;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) { const float v = im[i]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = (blockDim.x + 31) >> 5;
        mn = threadIdx.x < nw ? s_mn[threadIdx.x] : 3.4e38f;
        mx = threadIdx.x < nw ? s_mx[threadIdx.x] : -3.4e38f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (threadIdx.x == 0) { minmax[2 * blockIdx.x] = mn; minmax[2 * blockIdx.x + 1] = mx; }
    }
}

// u8 = saturate(rint(v*alpha + beta)) with the two parameterisations the reference uses
__global__ void __launch_bounds__(256) ev_normalize_kernel(const float* __restrict__ img, int npix, int normMode,
                                                           float* __restrict__ minmax, uint8_t* __restrict__ out) {
    const int win = blockIdx.y;
    const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= npix) return;
    float alpha, beta;
    ev_norm_coeffs(normMode, minmax[2 * win], minmax[2 * win + 1], alpha, beta);
    const float* im = img + (size_t)win * npix;
    uint8_t* o = out + (size_t)win * npix;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = i0 + k;
        if (i < npix) {
            const int r = __float2int_rn(__fadd_rn(__fmul_rn(im[i], alpha), beta));
            o[i] = (uint8_t)min(max(r, 0), 255);
        }
    }
}

cudaError_t launch_ev_splat(const eorb_event* d_evs, const EvWindow* d_wins, int nwin, long long maxEventsPerWindow,
                            const EvConst& c, float* d_img, cudaStream_t st, long long* launches) {
    if (nwin <= 0 || maxEventsPerWindow <= 0) return cudaSuccess;
    if (c.mode == EORB_EV_NEAREST) {
        dim3 grd((unsigned)((maxEventsPerWindow + 255) / 256), nwin);
        ev_splat_nearest_kernel<<<grd, 256, 0, st>>>(d_evs, d_wins, c, d_img);
    } else {
        const int win = 2 * c.half + 1;
        const int lpe = win <= 8 ? 8 : (win <= 16 ? 16 : 32);
        const int perBlock = 256 / lpe;
        dim3 grd((unsigned)((maxEventsPerWindow + perBlock - 1) / perBlock), nwin);
        ev_splat_gauss_kernel<<<grd, 256, 0, st>>>(d_evs, d_wins, c, d_img, lpe);
    }
    (*launches)++;
    return cudaGetLastError();
}

// ---- contrast metric of event frames -------------------------------------------------------
// EvImConverter::measureImageFocusLocal / measureImageFocusGlobal / imageMeanLocal (EventConversion.cc:79-162): the
// image is cut into patch x patch cells (30, DEF_PATCH_SIZE_STD), cv::meanStdDev of every cell (sum and sum of
// squares in double), the per-cell value cast to float and accumulated in float IN CELL ORDER, or sorted for the
// median.  One block per frame, one warp per cell (double shuffle reduction), thread 0 does the ordered float part.
// what: 0 local std-dev, 1 global std-dev, 2 local mean;  avg: 1 average, 0 median.
#define EV_FOCUS_MAX_CELLS 1024
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// With img2 != nullptr the metric is taken of the pixel-wise float product img (fixed) x img2 (frame blockIdx.x): that is
// I.mul(I_k) followed by imageMean of ev2mci_gg_f_jac (EventConversion.cc:646-659); what == 3 is cv::mean (global mean).
__global__ void __launch_bounds__(256) ev_focus_kernel(const float* __restrict__ img, const float* __restrict__ img2, int W, int H, int patch,
                                                       int what, int avg, float* __restrict__ out) {
    __shared__ float s_val[EV_FOCUS_MAX_CELLS];
    __shared__ double s_part[8][2];
    const float* im = img2 ? img : img + (size_t)blockIdx.x * (size_t)W * H;
    const float* im2 = img2 ? img2 + (size_t)blockIdx.x * (size_t)W * H : nullptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (what == 1 || what == 3) {   // one cell = the whole frame, all warps
        double s = 0, sq = 0;
        for (int i = threadIdx.x; i < W * H; i += 256) { const double v = im2 ? (double)__fmul_rn(im[i], im2[i]) : (double)im[i]; s += v; sq += v * v; }
        s = warp_sum_d(s); sq = warp_sum_d(sq);
        if (lane == 0) { s_part[warp][0] = s; s_part[warp][1] = sq; }
        __syncthreads();
        if (threadIdx.x == 0) {
            s = 0; sq = 0;
            for (int k = 0; k < 8; k++) { s += s_part[k][0]; sq += s_part[k][1]; }
            const double n = (double)W * H, mean = s / n, var = sq / n - mean * mean;
            out[blockIdx.x] = what == 3 ? (float)mean : (float)sqrt(var > 0 ? var : 0);
        }
        return;
    }
    const int npc = (W + patch - 1) / patch, npr = (H + patch - 1) / patch, cells = npc * npr;
    for (int c = warp; c < cells; c += 8) {
        const int cr = c / npc, cc = c - cr * npc;
        const int r0 = cr * patch, r1 = min(r0 + patch, H), c0 = cc * patch, c1 = min(c0 + patch, W);
        const int pw = c1 - c0, area = pw * (r1 - r0);
        double s = 0, sq = 0;
        for (int i = lane; i < area; i += 32) {
            const int y = i / pw, x = i - y * pw;
            const size_t o = (size_t)(r0 + y) * W + c0 + x;
            const double v = im2 ? (double)__fmul_rn(im[o], im2[o]) : (double)im[o];
            s += v; sq += v * v;
        }
        s = warp_sum_d(s); sq = warp_sum_d(sq);
        if (lane == 0) {
            const double n = (double)area, mean = s / n, var = sq / n - mean * mean;
            s_val[c] = (float)(what == 2 ? mean : sqrt(var > 0 ? var : 0));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (avg) {
            float acc = 0.f;
            for (int c = 0; c < cells; c++) acc = __fadd_rn(acc, s_val[c]);
            out[blockIdx.x] = __fdiv_rn(acc, (float)cells);
        } else {
            for (int a = 1; a < cells; a++) { const float v = s_val[a]; int b = a - 1; while (b >= 0 && s_val[b] > v) { s_val[b + 1] = s_val[b]; b--; } s_val[b + 1] = v; }
            out[blockIdx.x] = s_val[cells / 2];
        }
    }
}

cudaError_t launch_ev_focus(const float* d_img, int nwin, int W, int H, int patch, int what, int avg, float* d_out, cudaStream_t st,
                            long long* launches) {
    if (nwin <= 0) return cudaSuccess;
    ev_focus_kernel<<<nwin, 256, 0, st>>>(d_img, nullptr, W, H, patch, what, avg, d_out);
    (*launches)++;
    return cudaGetLastError();
}

// ---- Jacobian of the contrast objective (second half): ev2mci_gg_f_jac, EventConversion.cc:533-662
// Seven frames (I and its derivatives w.r.t. the six components of the window motion) are splatted with fp32 L2
// reductions; a group of LPE lanes owns one event (lane = tap column), the per-event geometry and the 2x6 image
// Jacobian are evaluated in double exactly as the reference does through Eigen.
__global__ void __launch_bounds__(256) ev_jac_splat_kernel(const eorb_event* __restrict__ evs, const EvWindow* __restrict__ wins, EvConst c,
                                                           float* __restrict__ frames, int lpe) {
    const EvWindow w = wins[0];
    const long long nev = w.end - w.begin;
    const int perBlock = blockDim.x / lpe;
    const long long e = (long long)blockIdx.x * perBlock + threadIdx.x / lpe;
    const int sub = threadIdx.x % lpe;
    if (e >= nev) return;
    const eorb_event* evp = evs + w.begin + e;
    const double t1 = evs[w.end - 1].ts, DT = t1 - evs[w.begin].ts, invDT = 1.0 / DT;
    const double rate = (t1 - evp->ts) * invDT;
    const float ux = __fdiv_rn(__fsub_rn(evp->x, c.cx), c.fx), uy = __fdiv_rn(__fsub_rn(evp->y, c.cy), c.fy);
    const double dep = (double)c.depth;
    const double X0 = dep * (double)ux, X1 = dep * (double)uy, X2 = dep * 1.0;
    double s, co;
    sincos(w.angle * rate, &s, &co);
    const double a0 = w.axis[0], a1 = w.axis[1], a2 = w.axis[2];
    const double s0 = s * a0, s1 = s * a1, s2 = s * a2;
    const double c0 = (1 - co) * a0, c1 = (1 - co) * a1, c2 = (1 - co) * a2;
    double t;
    double R01, R10, R02, R20, R12, R21;
    t = c0 * a1; R01 = t - s2; R10 = t + s2;
    t = c0 * a2; R02 = t + s1; R20 = t - s1;
    t = c1 * a2; R12 = t - s0; R21 = t + s0;
    const double R00 = c0 * a0 + co, R11 = c1 * a1 + co, R22 = c2 * a2 + co;
    const double X = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(R00, X0), __dmul_rn(R01, X1)), __dmul_rn(R02, X2)), __dmul_rn(w.t[0], rate));
    const double Y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(R10, X0), __dmul_rn(R11, X1)), __dmul_rn(R12, X2)), __dmul_rn(w.t[1], rate));
    const double Z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(R20, X0), __dmul_rn(R21, X1)), __dmul_rn(R22, X2)), __dmul_rn(w.t[2], rate));
    // JP2D = -projectJac(xyz) * (SE3deriv * rate)      (Pinhole::projectJac, Pinhole.cpp:81-91)
    const double J00 = (double)c.fx / Z, J02 = -(double)c.fx * X / (Z * Z), J11 = (double)c.fy / Z, J12 = -(double)c.fy * Y / (Z * Z);
    const double S0[6] = {0.0, Z * rate, -Y * rate, rate, 0.0, 0.0};
    const double S1[6] = {-Z * rate, 0.0, X * rate, 0.0, rate, 0.0};
    const double S2[6] = {Y * rate, -X * rate, 0.0, 0.0, 0.0, rate};
    double JP0[6], JP1[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        JP0[k] = -__dadd_rn(__dmul_rn(J00, S0[k]), __dmul_rn(J02, S2[k]));
        JP1[k] = -__dadd_rn(__dmul_rn(J11, S1[k]), __dmul_rn(J12, S2[k]));
    }
    const float U = (float)((double)c.fx * X / Z + (double)c.cx), V = (float)((double)c.fy * Y / Z + (double)c.cy);
    const float fxi = floorf(U), fyi = floorf(V);
    if (!(fxi >= -64.f && fxi <= (float)(c.width + 64) && fyi >= -64.f && fyi <= (float)(c.height + 64))) return;   // also NaN / inf
    const int xi = (int)fxi, yi = (int)fyi;
    const float xr = __fsub_rn(U, fxi), yr = __fsub_rn(V, fyi);
    const float ps = (c.pol && evp->p == 0) ? -1.0f : 1.0f;
    const float invSig2 = __fdiv_rn(1.f, c.sig2), den = __fmul_rn(2.0f, c.sig2);
    const size_t npx = (size_t)c.width * c.height;
    for (int ii = sub; ii <= 2 * c.half; ii += lpe) {
        const int i = ii - c.half, xn = xi + i;
        if (xn < 0 || xn >= c.width) continue;
        const float dx = __fsub_rn((float)i, xr), dx2 = __fmul_rn(dx, dx);
        for (int j = -c.half; j <= c.half; j++) {
            const int yn = yi + j;
            if (yn < 0 || yn >= c.height) continue;
            const float dy = __fsub_rn((float)j, yr);
            const float val = __fdiv_rn(expf(-__fdiv_rn(__fadd_rn(dx2, __fmul_rn(dy, dy)), den)), c.norm);
            const float gx = __fmul_rn(__fmul_rn(invSig2, dx), val), gy = __fmul_rn(__fmul_rn(invSig2, dy), val);
            float* p = frames + (size_t)yn * c.width + xn;
            atomicAdd(p, __fmul_rn(ps, val));
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const double JI = __dadd_rn(__dmul_rn((double)gx, JP0[k]), __dmul_rn((double)gy, JP1[k]));
                atomicAdd(p + (size_t)(k + 1) * npx, (float)__dmul_rn((double)ps, JI));
            }
        }
    }
}

cudaError_t launch_ev_jac(const eorb_event* d_evs, const EvWindow* d_win, long long nev, const EvConst& c, int globalMean, float* d_frames7,
                          float* d_out6, cudaStream_t st, long long* launches) {
    const int win = 2 * c.half + 1;
    const int lpe = win <= 8 ? 8 : (win <= 16 ? 16 : 32);
    const int perBlock = 256 / lpe;
    ev_jac_splat_kernel<<<(unsigned)((nev + perBlock - 1) / perBlock), 256, 0, st>>>(d_evs, d_win, c, d_frames7, lpe);
    (*launches)++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t npx = (size_t)c.width * c.height;
    // six products I x I_k, mean over all pixels (cv::mean) or mean of the 30x30-cell means (imageMeanLocal)
    ev_focus_kernel<<<6, 256, 0, st>>>(d_frames7, d_frames7 + npx, c.width, c.height, 30, globalMean ? 3 : 2, 1, d_out6);
    (*launches)++;
    return cudaGetLastError();
}

// whole path: zero + splat + min/max + normalise.  7x7 Gaussian windows take the shared-memory kernel (which also
// zeroes and, for single-band frames, normalises); everything else the L2-reduction kernels.
// ---- ordered frames: polarity-signed sums WITH the reference's running normalisation ------------------------------------------------
// With pol = true the reference's resolveMinMaxVals (EventConversion.cc:30-38, called after every tap :201, :257, :341, :432) sees every
// INTERMEDIATE pixel value: a pixel that climbs to +3 and is pulled back to 0 by negative events still raised maxVal to 3, so the
// extremes used by normalizeImage are a function of the event ORDER, not of the final frame (with pol = false every tap is positive and
// the two coincide).  This kernel replays the window in order: one block per window, one thread per tap, one event per step, a block
// barrier between events.  Every pixel therefore receives its contributions in the reference's order (the float frame equals the
// serial sum up to expf rounding) and every thread tracks the extremes of the values it writes.  It is a latency-bound replay
// (about 0.05 us per event with the frame in shared memory, about 1 us in global memory); only pol && NORM_RUNNING takes it.
__global__ void __launch_bounds__(256) ev_ordered_kernel(const eorb_event* __restrict__ evs, const EvWindow* __restrict__ wins, EvConst c,
                                                         int frameInSmem, float* __restrict__ img, float* __restrict__ minmax,
                                                         const float2* __restrict__ xy) {
    extern __shared__ __align__(16) float s_frame[];
    __shared__ float s_mn[8], s_mx[8];
    const EvWindow w = wins[blockIdx.x];
    const int W = c.width, H = c.height, npix = W * H, tid = threadIdx.x;
    float* gim = img + (size_t)blockIdx.x * (size_t)npix;
    float* im = frameInSmem ? s_frame : gim;
    for (int i = tid; i < npix; i += blockDim.x) im[i] = 0.0f;
    __syncthreads();
    const bool nearest = c.mode == EORB_EV_NEAREST;
    const int side = nearest ? 1 : 2 * c.half + 1, ntaps = side * side;
    const float den = __fmul_rn(2.0f, c.sig2);
    float mn = 0.0f, mx = -1000000.0f;            // EventConversion.cc:177-178, 219-220
    const long long nev = w.end - w.begin;
    for (long long e = 0; e < nev; e++) {
        const eorb_event* evp = evs + w.begin + e;
        const float polSign = (c.pol && evp->p == 0) ? -1.0f : 1.0f;
        for (int t = tid; t < ntaps; t += blockDim.x) {
            if (nearest) {
                const int px = (int)roundf(evp->x), py = (int)roundf(evp->y);
                if (px >= 0 && px < W && py >= 0 && py < H) {
                    const float nv = __fadd_rn(im[(size_t)py * W + px], __fmul_rn(polSign, 0.001f));
                    im[(size_t)py * W + px] = nv;
                    mx = fmaxf(mx, nv); mn = fminf(mn, nv);
                }
            } else {
                float X, Y;
                if (xy) { const float2 v = __ldg(xy + w.begin + e); X = v.x; Y = v.y; }
                else { X = evp->x; Y = evp->y; }
                const float fxi = floorf(X), fyi = floorf(Y);
                if (fxi >= (float)(-c.half - 1) && fxi <= (float)(W + c.half) && fyi >= (float)(-c.half - 1) && fyi <= (float)(H + c.half)) {
                    const int xi = (int)fxi, yi = (int)fyi;
                    const int i = t / side - c.half, j = t % side - c.half;      // the reference's loops: i over x outside, j over y inside
                    const int xn = xi + i, yn = yi + j;
                    if (xn >= 0 && xn < W && yn >= 0 && yn < H) {
                        const float dx = __fsub_rn((float)i, __fsub_rn(X, (float)xi)), dy = __fsub_rn((float)j, __fsub_rn(Y, (float)yi));
                        const float dd = __fdiv_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), den);
                        const float val = __fdiv_rn(expf(-dd), c.norm);
                        const float nv = __fadd_rn(im[(size_t)yn * W + xn], __fmul_rn(polSign, val));
                        im[(size_t)yn * W + xn] = nv;
                        mx = fmaxf(mx, nv); mn = fminf(mn, nv);
                    }
                }
            }
        }
        __syncthreads();                          // the next event may touch the same pixels
    }
    if (frameInSmem)
        for (int i = tid; i < npix; i += blockDim.x) gim[i] = s_frame[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((tid & 31) == 0) { s_mn[tid >> 5] = mn; s_mx[tid >> 5] = mx; }
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); k++) { mn = fminf(mn, s_mn[k]); mx = fmaxf(mx, s_mx[k]); }
        minmax[2 * blockIdx.x] = mn; minmax[2 * blockIdx.x + 1] = mx;
    }
}

cudaError_t launch_ev_frames(const eorb_event* d_evs, const EvWindow* d_wins, int nwin, long long maxEventsPerWindow, const EvConst& c,
                             int normMode, float* d_img, float* d_minmax, uint8_t* d_u8, cudaStream_t st, long long* launches,
                             float2* d_xyScratch) {
    if (nwin <= 0) return cudaSuccess;
    const int npix = c.width * c.height;
    const size_t smemBudget = 200 * 1024;
    if (c.pol && normMode == EORB_NORM_RUNNING && maxEventsPerWindow > 0) {
        // order-dependent extremes: replay the window in order (see ev_ordered_kernel)
        const float2* xy = nullptr;
        if (c.mode == EORB_EV_SE3 || c.mode == EORB_EV_SE2) {
            if (!d_xyScratch) return cudaErrorInvalidValue;
            dim3 gw((unsigned)((maxEventsPerWindow + 255) / 256), nwin);
            ev_warp_kernel<<<gw, 256, 0, st>>>(d_evs, d_wins, c, d_xyScratch);
            (*launches)++;
            xy = d_xyScratch;
        }
        const size_t fb = (size_t)npix * sizeof(float);
        const int inSmem = fb <= smemBudget ? 1 : 0;
        cudaError_t ea = cudaFuncSetAttribute(ev_ordered_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBudget + 4096);
        if (ea != cudaSuccess) return ea;
        ev_ordered_kernel<<<nwin, 256, inSmem ? fb : 0, st>>>(d_evs, d_wins, c, inSmem, d_img, d_minmax, xy);
        (*launches)++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (d_u8) {
            dim3 grd((npix / 4 + 255) / 256 + 1, nwin);
            ev_normalize_kernel<<<grd, 256, 0, st>>>(d_img, npix, normMode, d_minmax, d_u8);
            (*launches)++;
        }
        return cudaGetLastError();
    }
    // fixed-point guard of the shared-memory path: its scale 2^k shrinks with the window (k = floor(log2(2^31 / (events * peakTap)))),
    // and below k = 13 (more than 1.6 M events per window at sigma = 1) the accumulated rounding approaches the 1e-4 * peak bar
    const bool fixedOk = (double)maxEventsPerWindow * (1.0 / (double)c.norm) + 1.0 <= 2147483648.0 / 8192.0;
    if (fixedOk && c.mode != EORB_EV_NEAREST && c.half == EV_SMEM_HALF && maxEventsPerWindow > 0 && (size_t)c.width * 4 * 8 <= smemBudget) {
        const int bands = (int)(((size_t)npix * 4 + smemBudget - 1) / smemBudget);
        const int bandRows = (c.height + bands - 1) / bands;
        const size_t smem = ((((size_t)bandRows * c.width + 2 * EV_SMEM_PAD) * 4) + 15) & ~(size_t)15;
        cudaError_t ea = cudaFuncSetAttribute(ev_frame_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBudget + 4096);
        if (ea != cudaSuccess) return ea;
        const int nb = (c.height + bandRows - 1) / bandRows;
        const int fuse = nb == 1 ? 1 : 0;
        dim3 grd(nb, nwin);
        const float2* xy = nullptr;
        if (nb > 1 && d_xyScratch && (c.mode == EORB_EV_SE3 || c.mode == EORB_EV_SE2)) {
            dim3 gw((unsigned)((maxEventsPerWindow + 255) / 256), nwin);
            ev_warp_kernel<<<gw, 256, 0, st>>>(d_evs, d_wins, c, d_xyScratch);
            (*launches)++;
            xy = d_xyScratch;
        }
        ev_frame_smem_kernel<<<grd, EV_SMEM_THREADS, smem, st>>>(d_evs, d_wins, c, bandRows, fuse, normMode, d_img, d_minmax, d_u8, xy);
        (*launches)++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess || fuse) return e;
        return launch_ev_normalize(d_img, nwin, npix, normMode, d_minmax, d_u8, st, launches);
    }
    cudaError_t e = cudaMemsetAsync(d_img, 0, (size_t)nwin * npix * sizeof(float), st);
    if (e != cudaSuccess) return e;
    (*launches)++;   // memset node
    e = launch_ev_splat(d_evs, d_wins, nwin, maxEventsPerWindow, c, d_img, st, launches);
    if (e != cudaSuccess) return e;
    return launch_ev_normalize(d_img, nwin, npix, normMode, d_minmax, d_u8, st, launches);
}

cudaError_t launch_ev_normalize(const float* d_img, int nwin, int npix, int normMode, float* d_minmax, uint8_t* d_u8,
                                cudaStream_t st, long long* launches) {
    if (nwin <= 0) return cudaSuccess;
    ev_minmax_kernel<<<nwin, 1024, 0, st>>>(d_img, npix, d_minmax);
    (*launches)++;
    if (normMode != EORB_NORM_NONE && d_u8) {
        dim3 grd((npix / 4 + 255) / 256 + 1, nwin);
        ev_normalize_kernel<<<grd, 256, 0, st>>>(d_img, npix, normMode, d_minmax, d_u8);
        (*launches)++;
    }
    return cudaGetLastError();
}

}  // namespace eorb
