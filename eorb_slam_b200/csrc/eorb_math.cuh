// eorb_math.cuh — scalar arithmetic shared by the CUDA kernels and (compiled as plain C++) by the host-model
// unit tests.  Every function here is the exact integer / no-FMA float arithmetic the reference gets from
// OpenCV on the CPU; see DESIGN.md §"Arithmetic pins".
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define EORB_HD __host__ __device__ __forceinline__
#else
#define EORB_HD inline
#endif

namespace eorb {

// float ops that must never be contracted into FMA (reference arithmetic is mul-then-add)
#ifdef __CUDA_ARCH__
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ int   round_rne(float v) { return __float2int_rn(v); }
#else
// host build of this header must be compiled with -ffp-contract=off
inline float fmul(float a, float b) { volatile float r = a * b; return r; }
inline float fadd(float a, float b) { volatile float r = a + b; return r; }
inline float fsub(float a, float b) { volatile float r = a - b; return r; }
inline float fdiv(float a, float b) { volatile float r = a / b; return r; }
inline int   round_rne(float v) { return (int)lrintf(v); }
#endif

EORB_HD int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) { p = p < 0 ? -p : 2 * len - 2 - p; }
    return p;
}

// ---- cv::resize INTER_LINEAR 8-bit, one output pixel from the two horizontally-interpolated sums
// (VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>; reference call ORBextractor.cc:1253)
EORB_HD int resize_hsum(int p0, int p1, int a0, int a1) { return p0 * a0 + p1 * a1; }
EORB_HD int resize_vsum(int r0, int r1, int b0, int b1) {
    return (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
}

// ---- cv::GaussianBlur 5x5 sigma 2 on 8-bit: 8.8 fixed-point separable kernel [39 57 64 57 39]
// (reference call ORBextractor.cc:1141-1142).  hsum of 5 pixels, then vsum of 5 hsums -> (s+32768)>>16.
EORB_HD int gauss5_h(int p0, int p1, int p2, int p3, int p4) { return 39 * (p0 + p4) + 57 * (p1 + p3) + 64 * p2; }
EORB_HD int gauss5_v(int h0, int h1, int h2, int h3, int h4) {
    return (39 * (h0 + h4) + 57 * (h1 + h3) + 64 * h2 + 32768) >> 16;
}

// ---- FAST-9/16 (cv::FAST called at ORBextractor.cc:832,851).  v = centre, ring[k], k = 0..15 in the ring order
// (0,3),(1,3),(2,2),(3,1),(3,0),(3,-1),(2,-2),(1,-3),(0,-3),(-1,-3),(-2,-2),(-3,-1),(-3,0),(-3,1),(-2,2),(-1,3).
// Returns m = max over the 16 arcs of 9 contiguous ring pixels of min(|v-p|) with a common sign, clamped at 0.
// corner at threshold t  <=>  m > t ;  OpenCV's NMS score (cornerScore<16>) = m - 1.
//
// Written as two MIN-only sliding-window networks over d = v-p and e = p-v (window 9 on a circular array of
// 16 by doubling: 2, 4, 8, +1).  Do NOT fold the second one into max(-max(d)): ptxas 12.9 for sm_100a fuses
// `max(best, mn, -mx)` into VIMNMX3 and drops the negation (measured on B200: returns max(d) instead), see
// DESIGN.md "Toolchain findings".
EORB_HD int fast_arc_min16(const int* d) {
    int m2[16], m4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { const int a = d[k], b = d[(k + 1) & 15]; m2[k] = a < b ? a : b; }
#pragma unroll
    for (int k = 0; k < 16; k++) { const int a = m2[k], b = m2[(k + 2) & 15]; m4[k] = a < b ? a : b; }
    int best = -100000;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int a = m4[k], b = m4[(k + 4) & 15];
        int m = a < b ? a : b;
        const int x = d[(k + 8) & 15];
        m = m < x ? m : x;
        best = best > m ? best : m;
    }
    return best;
}

EORB_HD int fast_max_arc_min(int v, const int* ring) {
    int d[16], e[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { d[k] = v - ring[k]; e[k] = ring[k] - v; }
    const int a = fast_arc_min16(d);   // ring darker than the centre
    const int b = fast_arc_min16(e);   // ring brighter than the centre
    int m = a > b ? a : b;
    return m > 0 ? m : 0;
}

// ---- cv::fastAtan2 (degrees, [0,360]); called at ORBextractor.cc:103
EORB_HD float fast_atan2_deg(float y, float x) {
    const float p1 = 0.9997878412794807f * (float)(180 / 3.14159265358979323846);
    const float p3 = -0.3258083974640975f * (float)(180 / 3.14159265358979323846);
    const float p5 = 0.1555786518463281f * (float)(180 / 3.14159265358979323846);
    const float p7 = -0.04432655554792128f * (float)(180 / 3.14159265358979323846);
    const float eps = (float)2.2204460492503131e-16;
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = fdiv(ay, fadd(ax, eps));
        c2 = fmul(c, c);
        a = fmul(fadd(fmul(fadd(fmul(fadd(fmul(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = fdiv(ax, fadd(ay, eps));
        c2 = fmul(c, c);
        a = fsub(90.f, fmul(fadd(fmul(fadd(fmul(fadd(fmul(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = fsub(180.f, a);
    if (y < 0) a = fsub(360.f, a);
    return a;
}

// ---- steered BRIEF sample offset (computeOrbDescriptor, ORBextractor.cc:118-123): a = cos, b = sin
EORB_HD void brief_offset(int px, int py, float a, float b, int& row, int& col) {
    float fx = (float)px, fy = (float)py;
    row = round_rne(fadd(fmul(fx, b), fmul(fy, a)));
    col = round_rne(fsub(fmul(fx, a), fmul(fy, b)));
}

// same with float pattern coordinates and, on the device, round-half-even through the 1.5*2^23 trick (an FADD and
// an integer subtract instead of two conversion-pipe F2I; exact for |v| < 2^22, pattern offsets are < 32)
#ifdef __CUDA_ARCH__
__device__ __forceinline__ int round_rne_small(float v) { return __float_as_int(__fadd_rn(v, 12582912.0f)) - 0x4B400000; }
#else
inline int round_rne_small(float v) { return round_rne(v); }
#endif
#ifdef __CUDACC__
// the same two roundings left as raw bits: value = 0x4B400000 + round(v); the caller removes the bias where it is cheapest
__device__ __forceinline__ void brief_offset_biased(float fx, float fy, float a, float b, unsigned& row, unsigned& col) {
    row = (unsigned)__float_as_int(__fadd_rn(fadd(fmul(fx, b), fmul(fy, a)), 12582912.0f));
    col = (unsigned)__float_as_int(__fadd_rn(fsub(fmul(fx, a), fmul(fy, b)), 12582912.0f));
}
#endif
EORB_HD void brief_offset_f(float fx, float fy, float a, float b, int& row, int& col) {
    row = round_rne_small(fadd(fmul(fx, b), fmul(fy, a)));
    col = round_rne_small(fsub(fmul(fx, a), fmul(fy, b)));
}

// ---- DescriptorDistance (ORBmatcher.cc:2360-2378): 8 x popcount(xor)
EORB_HD int hamming256(const uint32_t* a, const uint32_t* b) {
    int d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
#ifdef __CUDA_ARCH__
        d += __popc(a[i] ^ b[i]);
#else
        d += __builtin_popcount(a[i] ^ b[i]);
#endif
    }
    return d;
}

}  // namespace eorb
