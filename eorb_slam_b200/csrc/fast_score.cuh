// fast_score.cuh — device-only exact FAST-9/16 arc score on packed s16x2 lanes (used by orb_fast.cu and by the
// device-vs-host self-test kernel in orb_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace eorb {

// ---- exact arc score on packed s16x2 lanes -------------------------------------------------------------------
// D_k = (ring_k - v, v - ring_k) per 16-bit half; arcs of 9 = min3 of min3 (VIMNMX3.S16x2), then a max3 tree.
// ring[k] in the order of eorb_math.cuh::fast_max_arc_min; result identical to that function (device-vs-host
// equality is checked by eorb_selftest_math on 20 000 rings).
__device__ __forceinline__ int fast_max_arc_min_packed(int v, const int* ring) {
    // one IMAD per ring pixel builds both polarities: r * 0xFFFF0001 = (lo: r, hi: -r), plus C = (lo: 0x4000 - v,
    // hi: +v) gives D = (lo: r - v + 0x4000, hi: v - r).  The low half is biased by 0x4000 so that it never
    // carries into the high half (0 <= r - v + 0x4000 < 2^16); min/max commute with the common bias.
    const unsigned C = ((unsigned)v << 16) | (unsigned)(0x4000 - v);
    unsigned D[16], A[16];
#pragma unroll
    for (int k = 0; k < 16; k++) D[k] = (unsigned)ring[k] * 0xFFFF0001u + C;
#pragma unroll
    for (int k = 0; k < 16; k++) A[k] = __vimin3_s16x2(D[k], D[(k + 1) & 15], D[(k + 2) & 15]);          // arcs of 3
#pragma unroll
    for (int k = 0; k < 16; k++) D[k] = __vimin3_s16x2(A[k], A[(k + 3) & 15], A[(k + 6) & 15]);          // arcs of 9
    unsigned b0 = __vimax3_s16x2(D[0], D[1], D[2]);
    unsigned b1 = __vimax3_s16x2(D[3], D[4], D[5]);
    unsigned b2 = __vimax3_s16x2(D[6], D[7], D[8]);
    unsigned b3 = __vimax3_s16x2(D[9], D[10], D[11]);
    unsigned b4 = __vimax3_s16x2(D[12], D[13], D[14]);
    b0 = __vimax3_s16x2(b0, b1, b2);
    b3 = __vimax3_s16x2(b3, b4, D[15]);
    b0 = __vmaxs2(b0, b3);
    const int bright = (int)(b0 & 0xffffu) - 0x4000, dark = (int)b0 >> 16;
    return max(max(bright, dark), 0);
}

}  // namespace eorb
