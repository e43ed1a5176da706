// hamming_tc.cu — 256-bit Hamming best-2 search on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
//
// Reference: ORBmatcher::DescriptorDistance src/ORBmatcher.cc:2360-2378 and the best / second-best scan with strict '<'
// (:741-770, :318-378); brute-force shape of src/Frame.cc:1228-1235.  Same contract as hamming_best2_kernel (match_kernels.cu):
// per (database chunk, query) the two smallest keys (dist << 32 | global row) in scan order, merged by merge_best2_kernel.
//
// A 256-bit Hamming distance is an exact int8 contraction.  With a_i = 1 - 2 q_i (+-1, the queries) and b_i = -r_i (0 / -1, the database),
//     sum_i a_i b_i = -popc(r) + 2 popc(q & r)      =>      dist = popc(q) + popc(r) - 2 popc(q & r) = popc(q) - dot
// (int32 accumulation, no rounding anywhere; popc(q) is one constant per query).  The asymmetric encoding makes the database side one
// PRMT per four elements: `prmt` with the sign-replicate bit turns the top bit of each byte of (word << k) into 0x00 / 0xFF.  Element
// order along K is therefore not bit order but (word w, shift k, byte b) <-> bit 8 b + 7 - k of word w, the same for both operands.
// One CTA (one per SM) owns 128 queries (the M = 128 rows of the accumulator = the 128 TMEM lanes) and walks a chunk of the database in
// tiles of 256 rows (N = 256 accumulator columns); K = 256 bits = 8 instructions of K = 32.
//   warp 0          allocates TMEM (512 columns = two accumulators) and issues the MMAs (one elected thread)
//   warps 1-8       producers: thread = one database row of the tile, packed 32 B -> 256 int8 in the no-swizzle K-major core-matrix
//                   layout the shared-memory descriptors describe (8 rows x 16 bytes per core matrix); two 64 KB stages
//   warps 9-16      epilogue: thread = one query (TMEM lane), tcgen05.ld 32 accumulator columns at a time, the next load in flight while
//                   the current columns are examined.  The scan order makes "this row enters the best two" equivalent to
//                   dot > popc(q) - (second-best distance), so the fast path is a max over the 32 dots and one compare; the rare slow
//                   path replays the 32 columns in order with the packed-key min / max network of the POPC kernel (lowest index wins
//                   ties, as the reference's sequential scan).
// Pipelines: full / empty mbarriers per shared-memory stage (producers <-> MMA, the "empty" side arrives through tcgen05.commit) and
// per accumulator (MMA <-> epilogue).  Every wait is bounded: a broken pipeline traps instead of hanging the GPU.
// What bounds it (measured with the EORB_HT_PROBE switches, profiles/r02_hamming_tensor.md): a tile is 1030 cycles of int8 math
// (8 x 128 cycles at the B200's dense rate), but reading the 128 x 256 int32 accumulator back (tcgen05.ld, 128 KB per tile) takes about
// as long, and the accumulating MMAs use the same TMEM port (every K = 32 step reads and writes the whole accumulator), so readout and
// math do not overlap well: ~2250 cycles per tile in all, 3.9 T pairs/s = 7 x the POPC kernel at 97 % of its own pipe.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <mutex>

#include "../../include/eorb_b200.h"
#include "match_kernels.h"
#include "tma_utils.cuh"

namespace eorb {

#define HT_THREADS 544
#define HT_M 128
#define HT_N 256
#define HT_A_BYTES (HT_M * 256)
#define HT_B_BYTES (HT_N * 256)
#define HT_BAR_OFF (HT_A_BYTES + 2 * HT_B_BYTES)
#define HT_SMEM (HT_BAR_OFF + 128)
#define HT_TMEM_COLS 512

__device__ __forceinline__ void ht_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spin = 0; !done; spin++) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && spin > (1u << 28)) __trap();   // a broken pipeline must not hang the device (seconds; legitimate waits are microseconds)
    }
}
__device__ __forceinline__ void ht_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ht_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ht_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ht_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 contiguous bytes);
// LBO = distance between the two core matrices of one K = 32 step, SBO = distance between 8-row groups (both in bytes)
__device__ __forceinline__ unsigned long long ht_desc(unsigned smemAddr, unsigned lbo, unsigned sbo) {
    return (unsigned long long)((smemAddr >> 4) & 0x3FFFu) | ((unsigned long long)((lbo >> 4) & 0x3FFFu) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void ht_mma_i8(unsigned tmemD, unsigned long long descA, unsigned long long descB, unsigned idesc, unsigned accumulate) {
    const unsigned z = 0;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
        "}\n" ::"r"(tmemD),
        "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate), "r"(z), "r"(z), "r"(z), "r"(z)
        : "memory");
}
__device__ __forceinline__ void ht_tmem_ld32(unsigned addr, int* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(addr)
        : "memory");
}
// 64 accumulator columns as 32 registers: .pack::16b keeps the low 16 bits of two adjacent columns per register (low half = the even
// column); the dots lie in [-256, 256], so nothing is lost and the readout moves half the registers
__device__ __forceinline__ void ht_tmem_ld64p(unsigned addr, unsigned* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(addr)
        : "memory");
}
__device__ __forceinline__ void ht_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// every byte -> 0xFF when its top bit is set, else 0x00: prmt with the sign-replicate bit (8) in every selector nibble
// (the __byte_perm intrinsic masks the selector to three bits per nibble, hence the PTX)
__device__ __forceinline__ unsigned ht_sign_bytes(unsigned x) {
    unsigned d;
    asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(d) : "r"(x));
    return d;
}
// one packed row (8 words) -> 256 int8: element (w, k, b) = -(bit 8 b + 7 - k of word w) for the database (orMask = 0), or
// +1 / -1 for the queries (orMask = 0x01010101 turns the 0x00 of a clear bit into +1).  K chunk 2 w + (k >> 2) holds shifts k of
// word w as its four 32-bit words; dst = address of (k chunk 0, this row), chunkStride = bytes between k chunks.
__device__ __forceinline__ void ht_expand_row(const uint4& a, const uint4& b, unsigned char* dst, unsigned chunkStride, unsigned orMask) {
    const unsigned w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int kc = 0; kc < 16; kc++) {
        const unsigned x = w[kc >> 1];
        unsigned o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) o[j] = ht_sign_bytes(x << ((kc & 1) * 4 + j)) | orMask;
        *reinterpret_cast<uint4*>(dst + (size_t)kc * chunkStride) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

__global__ void __launch_bounds__(HT_THREADS, 1) hamming_tc_kernel(const uint4* __restrict__ q, int nq, const uint4* __restrict__ db, long long ndb,
                                                                   long long chunkRows, long long indexOffset, eorb_best2* __restrict__ partial, int probe) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned s_tmemBase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* sA = smem;
    unsigned char* sB = smem + HT_A_BYTES;
    const unsigned barBase = smem_u32(smem + HT_BAR_OFF);
    // barriers (8 bytes each): fullB[0..1] @0,8  emptyB[0..1] @16,24  tmemFull[0..1] @32,40  tmemEmpty[0..1] @48,56
    const unsigned fullB = barBase, emptyB = barBase + 16, tmemFull = barBase + 32, tmemEmpty = barBase + 48;

    const long long j0 = (long long)blockIdx.x * chunkRows;
    const long long j1 = (j0 + chunkRows < ndb) ? j0 + chunkRows : ndb;
    const long long rows = j1 > j0 ? j1 - j0 : 0;
    const int ntiles = (int)((rows + HT_N - 1) / HT_N);
    const int q0 = blockIdx.y * HT_M;

    // ---- setup: barriers, TMEM, the CTA's 128 queries as int8
    if (tid == 0) {
        mbar_init(fullB, 8); mbar_init(fullB + 8, 8);
        mbar_init(emptyB, 1); mbar_init(emptyB + 8, 1);
        mbar_init(tmemFull, 1); mbar_init(tmemFull + 8, 1);
        mbar_init(tmemEmpty, 8); mbar_init(tmemEmpty + 8, 8);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmemBase)), "r"(HT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 1 && warp <= 4) {
        const int m = tid - 32;   // query row of the tile
        const int qi = q0 + m;
        uint4 a = make_uint4(0, 0, 0, 0), b = a;
        if (qi < nq) { a = __ldg(&q[2 * qi]); b = __ldg(&q[2 * qi + 1]); }
        ht_expand_row(a, b, sA + (m >> 3) * 128 + (m & 7) * 16, (HT_M / 8) * 128, 0x01010101u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core's reads
    }
    ht_fence_before();
    __syncthreads();
    ht_fence_after();
    const unsigned tmemBase = s_tmemBase;

    if (warp == 0) {
        // ================================================================= MMA issuer (the warp waits together, one lane issues)
        // instruction descriptor (kind::i8): D = S32 (bits 4-5 = 2), A = B = signed 8 bit (bits 7-9, 10-12 = 1), both K-major,
        // N >> 3 at bit 17, M >> 4 at bit 24
        const unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(HT_N >> 3) << 17) | ((unsigned)(HT_M >> 4) << 24);
        const unsigned aAddr = smem_u32(sA), bAddr = smem_u32(sB);
        for (int t = 0; t < ntiles; t++) {
            const int s = t & 1;
            const unsigned ph = (unsigned)(t >> 1) & 1u;
            ht_wait(fullB + 8 * s, ph);
            ht_wait(tmemEmpty + 8 * s, ph ^ 1u);
            ht_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < ((probe & 4) ? 0 : 8); j++) {
                    const unsigned long long da = ht_desc(aAddr + (unsigned)j * 2u * (HT_M / 8) * 128u, (HT_M / 8) * 128u, 128u);
                    const unsigned long long dbd = ht_desc(bAddr + (unsigned)s * HT_B_BYTES + (unsigned)j * 2u * (HT_N / 8) * 128u, (HT_N / 8) * 128u, 128u);
                    ht_mma_i8(tmemBase + (unsigned)s * HT_N, da, dbd, idesc, j > 0 ? 1u : 0u);
                }
                ht_commit(emptyB + 8 * s);     // the stage may be refilled once these MMAs have read it
                ht_commit(tmemFull + 8 * s);   // ... and the accumulator is complete
            }
            __syncwarp();
        }
    } else if (warp <= 8) {
        // ================================================================= producers: thread <-> row p of every tile
        const int p = tid - 32;
        uint4 r0, r1;
        auto load = [&](int t) {
            const long long row = j0 + (long long)t * HT_N + p;
            if (row < j1) { r0 = __ldg(&db[2 * row]); r1 = __ldg(&db[2 * row + 1]); }
            else { r0 = make_uint4(0, 0, 0, 0); r1 = r0; }
        };
        if (ntiles > 0) load(0);
        for (int t = 0; t < ntiles; t++) {
            const int s = t & 1;
            const unsigned ph = (unsigned)(t >> 1) & 1u;
            const uint4 c0 = r0, c1 = r1;
            if (t + 1 < ntiles) load(t + 1);   // the next tile's row is in flight while this one is expanded
            ht_wait(emptyB + 8 * s, ph ^ 1u);
            if (!(probe & 2)) ht_expand_row(c0, c1, sB + (size_t)s * HT_B_BYTES + (p >> 3) * 128 + (p & 7) * 16, (HT_N / 8) * 128, 0u);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) ht_arrive(fullB + 8 * s);
        }
    } else {
        // ================================================================= epilogue: thread <-> query 32 * (warp & 3) + lane
        const int quarter = warp & 3;                 // the TMEM lanes this warp may read
        const int half = (warp - 9) >> 2;             // columns [128 * half, 128 * half + 128) of every tile
        const unsigned laneAddr = tmemBase + ((unsigned)(quarter * 32) << 16);
        const int qmine = q0 + quarter * 32 + lane;
        int pq = 0;                                   // popc of this thread's query
        if (qmine < nq) {
            const uint4 a = __ldg(&q[2 * qmine]), b = __ldg(&q[2 * qmine + 1]);
            pq = __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w);
        }
        uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
        int thrDot = -100000;                         // dot > thrDot  <=>  dist < second-best distance  <=>  the row enters the best two
        // 32 columns of one query.  Fast path: maxima of the eight groups of four columns (independent, so the warp is not waiting on one
        // long dependency chain), their maximum, one compare.  Slow path: only the groups that hold a qualifying dot are replayed, in
        // column order, through the packed-key network; the threshold tightens after every group, which keeps the scan sequential.
        auto examine = [&](const int* v, int t, int col0, int cnt) {
            if (col0 >= cnt || (probe & 1)) return;   // warp-uniform
            int g[8];
#pragma unroll
            for (int j = 0; j < 8; j++) g[j] = max(max(v[4 * j], v[4 * j + 1]), max(v[4 * j + 2], v[4 * j + 3]));
            const int m = max(max(max(g[0], g[1]), max(g[2], g[3])), max(max(g[4], g[5]), max(g[6], g[7])));
            const bool partial = col0 + 32 > cnt;     // the chunk's last tile may be partial: columns >= cnt are not rows
            if (m > thrDot || partial) {
                const uint32_t local0 = (uint32_t)(t * HT_N + col0);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (g[j] > thrDot || partial) {
#pragma unroll
                        for (int i = 4 * j; i < 4 * j + 4; i++) {
                            if (col0 + i < cnt) {
                                const uint32_t dist = (uint32_t)(pq - v[i]);
                                const uint32_t key = (dist << 22) | (local0 + (uint32_t)i);
                                const uint32_t hi = max(k1, key);
                                k1 = min(k1, key);
                                k2 = min(k2, hi);
                            }
                        }
                        thrDot = (k2 == 0xFFFFFFFFu) ? -100000 : pq - (int)(k2 >> 22);
                    }
                }
            }
        };
        // the same on 64 columns held as 32 registers of two packed 16-bit dots (register i = columns 2 i, 2 i + 1)
        auto examine64 = [&](const unsigned* v, int t, int col0, int cnt) {
            if (col0 >= cnt || (probe & 1)) return;   // warp-uniform
            unsigned g[8];
#pragma unroll
            for (int j = 0; j < 8; j++) g[j] = __vmaxs2(__vmaxs2(v[4 * j], v[4 * j + 1]), __vmaxs2(v[4 * j + 2], v[4 * j + 3]));
            const unsigned mm = __vmaxs2(__vmaxs2(__vmaxs2(g[0], g[1]), __vmaxs2(g[2], g[3])), __vmaxs2(__vmaxs2(g[4], g[5]), __vmaxs2(g[6], g[7])));
            const int m = max((int)(short)(mm & 0xffffu), (int)mm >> 16);
            const bool partial = col0 + 64 > cnt;
            if (m > thrDot || partial) {
                const uint32_t local0 = (uint32_t)(t * HT_N + col0);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int gm = max((int)(short)(g[j] & 0xffffu), (int)g[j] >> 16);
                    if (gm > thrDot || partial) {
#pragma unroll
                        for (int i = 8 * j; i < 8 * j + 8; i++) {
                            if (col0 + i < cnt) {
                                const unsigned r = v[i >> 1];
                                const int dot = (i & 1) ? ((int)r >> 16) : (int)(short)(r & 0xffffu);
                                const uint32_t dist = (uint32_t)(pq - dot);
                                const uint32_t key = (dist << 22) | (local0 + (uint32_t)i);
                                const uint32_t hi = max(k1, key);
                                k1 = min(k1, key);
                                k2 = min(k2, hi);
                            }
                        }
                        thrDot = (k2 == 0xFFFFFFFFu) ? -100000 : pq - (int)(k2 >> 22);
                    }
                }
            }
        };
        if (!(probe & 8)) {
            for (int t = 0; t < ntiles; t++) {
                const int s = t & 1;
                const unsigned ph = (unsigned)(t >> 1) & 1u;
                const long long left = rows - (long long)t * HT_N;
                const int cnt = left < HT_N ? (int)left : HT_N;
                const unsigned colAddr = laneAddr + (unsigned)(s * HT_N + half * 128);
                const int col0 = half * 128;
                unsigned pa[32], pb[32];
                ht_wait(tmemFull + 8 * s, ph);
                ht_fence_after();
                __syncwarp();
                ht_tmem_ld64p(colAddr, pa);
                ht_tmem_ld64p(colAddr + 64, pb);
                ht_tmem_wait_ld();
                ht_fence_before();
                if (lane == 0) ht_arrive(tmemEmpty + 8 * s);   // this warp's 128 columns are in registers: the accumulator may be overwritten
                examine64(pa, t, col0, cnt);
                examine64(pb, t, col0 + 64, cnt);
                __syncwarp();
            }
        } else
        for (int t = 0; t < ntiles; t++) {
            const int s = t & 1;
            const unsigned ph = (unsigned)(t >> 1) & 1u;
            const long long left = rows - (long long)t * HT_N;
            const int cnt = left < HT_N ? (int)left : HT_N;
            const unsigned colAddr = laneAddr + (unsigned)(s * HT_N + half * 128);
            const int col0 = half * 128;
            int va[32], vb[32];
            ht_wait(tmemFull + 8 * s, ph);
            ht_fence_after();
            __syncwarp();
            ht_tmem_ld32(colAddr, va);
            ht_tmem_wait_ld();
            ht_tmem_ld32(colAddr + 32, vb);
            examine(va, t, col0, cnt);
            __syncwarp();
            ht_tmem_wait_ld();
            ht_tmem_ld32(colAddr + 64, va);
            examine(vb, t, col0 + 32, cnt);
            __syncwarp();
            ht_tmem_wait_ld();
            ht_tmem_ld32(colAddr + 96, vb);
            examine(va, t, col0 + 64, cnt);
            __syncwarp();
            ht_tmem_wait_ld();
            ht_fence_before();
            if (lane == 0) ht_arrive(tmemEmpty + 8 * s);   // every column of this warp's half is in registers: the accumulator may be overwritten
            examine(vb, t, col0 + 96, cnt);
        }
        const int qi = q0 + quarter * 32 + lane;
        if (qi < nq) {
            eorb_best2 o;
            const unsigned long long base = (unsigned long long)(indexOffset + j0);
            o.key1 = (k1 == 0xFFFFFFFFu) ? ~0ull : (((unsigned long long)(k1 >> 22)) << 32) | (base + (k1 & 0x3FFFFFu));
            o.key2 = (k2 == 0xFFFFFFFFu) ? ~0ull : (((unsigned long long)(k2 >> 22)) << 32) | (base + (k2 & 0x3FFFFFu));
            partial[((size_t)blockIdx.x * 2 + half) * nq + qi] = o;
        }
    }
    ht_fence_before();
    __syncthreads();
    if (warp == 0) {
        ht_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(HT_TMEM_COLS) : "memory");
    }
}

// chunks of the tensor-core search: multiples of 256 rows, <= 2^22 rows (22-bit local index), and a CTA count (query tiles x chunks) that
// fills whole waves of one CTA per SM: with 16 query tiles, 19 chunks are 304 CTAs on 148 SMs, i.e. a third wave of 8 CTAs that costs as
// much as a full one (measured: 2.09 ms against 1.45 ms with 9 chunks = 144 CTAs in one wave)
int hamming_tc_chunks(long long ndb, int nq, int sms, long long* chunkRows) {
    const int qtiles = (nq + HT_M - 1) / HT_M;
    const long long TN = HT_N;
    const long long maxRows = (1ll << 22) / TN * TN;
    long long rows = TN;
    for (int waves = 1; waves <= 4096; waves++) {
        long long want = (long long)sms * waves / qtiles;
        if (want < 1) want = 1;
        rows = (ndb + want - 1) / want;
        if (rows < TN) rows = TN;
        rows = (rows + TN - 1) / TN * TN;
        if (rows <= maxRows) break;
        rows = maxRows;
    }
    *chunkRows = rows;
    return (int)((ndb + rows - 1) / rows);
}
int hamming_tc_parts_per_chunk() { return 2; }

cudaError_t launch_hamming_best2_tc(const uint8_t* d_q, int nq, const uint8_t* d_db, long long ndb, long long indexOffset, long long chunkRows,
                                    int nchunks, eorb_best2* d_partial, cudaStream_t st) {
    if (nq <= 0 || nchunks <= 0) return cudaSuccess;
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!(dev >= 0 && dev < 64 && done[dev])) {
            e = cudaFuncSetAttribute(hamming_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_SMEM);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) done[dev] = true;
        }
    }
    static const int probe = getenv("EORB_HT_PROBE") ? atoi(getenv("EORB_HT_PROBE")) : 0;   // probes: 1 no epilogue work, 2 no expansion, 4 no MMA (timing only), 8 unpacked 32-bit readout (A/B)
    dim3 grd(nchunks, (nq + HT_M - 1) / HT_M);
    hamming_tc_kernel<<<grd, HT_THREADS, HT_SMEM, st>>>(reinterpret_cast<const uint4*>(d_q), nq, reinterpret_cast<const uint4*>(d_db), ndb, chunkRows,
                                                        indexOffset, d_partial, probe);
    return cudaGetLastError();
}

}  // namespace eorb
