// capi.cu — the C ABI of libeorb_b200.so (include/eorb_b200.h): handles, geometry plans, HBM slabs, streams,
// launch sequences.  Host code only; every compute step is a kernel in orb_kernels.cu / match_kernels.cu /
// event_kernels.cu.  No CPU fallback: without a CUDA device the entry points return EORB_ERR_CUDA.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/eorb_b200.h"
#include "event_kernels.h"
#include "match_kernels.h"
#include "octree_core.cuh"
#include "orb_kernels.h"
#include "orb_plan.h"

using namespace eorb;

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
// the same for the other translation units of the library (capi_lk.cu); not part of the public header
extern "C" int eorb_internal_fail(int code, const char* msg) { return fail(code, "%s", msg); }

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(EORB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

template <class T>
static cudaError_t devAlloc(T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    return cudaMalloc((void**)p, count * sizeof(T));
}

static inline int roundUp(int v, int a) { return (v + a - 1) / a * a; }

// grow-only device buffer: the capacity is zeroed before the old block is freed and set only after the new one exists, so a
// failed allocation never leaves a stale capacity beside a null pointer
template <typename T>
static int growBuf(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return EORB_OK;
    *cap = 0;
    cudaFree(*p); *p = nullptr;
    CU(devAlloc(p, need));
    *cap = need;
    return EORB_OK;
}


// ---- TMA tensor maps (driver entry point resolved through the runtime: no link-time dependency on libcuda)
static PFN_cuTensorMapEncodeTiled_v12000 g_tmaEncode = nullptr;
static std::once_flag g_tmaOnce;
static int tmaEncoder() {
    std::call_once(g_tmaOnce, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            g_tmaEncode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
        else
            cudaGetLastError();
    });
    return g_tmaEncode ? EORB_OK : fail(EORB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
}

// u8 tensor {x = w, y = h, z = frames} with byte strides {pitch, frameStride}; box {boxW, boxH, 1}; zero fill outside
static int tmaEncodeFrames(CUtensorMap* out, const void* base, int w, int hgt, int frames, size_t pitch, size_t frameStride, int boxW, int boxH) {
    int rc = tmaEncoder();
    if (rc != EORB_OK) return rc;
    if (((uintptr_t)base & 15) || (pitch & 15) || (frameStride & 15)) return fail(EORB_ERR_ARG, "TMA source must be 16-byte aligned with 16-byte strides");
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)hgt, (cuuint64_t)std::max(frames, 1)};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)std::max(frameStride, (size_t)16)};
    cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = g_tmaEncode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(EORB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for %dx%dx%d pitch %zu", (int)r, w, hgt, frames, pitch);
    return EORB_OK;
}
static inline int rne(float v) { return (int)lrintf(v); }
static inline short satShort(float v) { int i = rne(v); return (short)(i < -32768 ? -32768 : (i > 32767 ? 32767 : i)); }

extern "C" int eorb_version(void) { return 100; }
extern "C" const char* eorb_last_error(void) { return g_err; }
extern "C" int eorb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ------------------------------------------------------------------------------------------------ timers
struct EorbTimer { cudaEvent_t a, b; };
extern "C" int eorb_timer_create(void** t) {
    if (!t) return fail(EORB_ERR_ARG, "null timer");
    std::unique_ptr<EorbTimer> x(new EorbTimer());
    CU(cudaEventCreate(&x->a));
    if (cudaEventCreate(&x->b) != cudaSuccess) { cudaEventDestroy(x->a); return fail(EORB_ERR_CUDA, "cudaEventCreate failed"); }
    *t = x.release();
    return EORB_OK;
}
extern "C" int eorb_timer_destroy(void* t) {
    if (!t) return EORB_OK;
    EorbTimer* x = (EorbTimer*)t;
    cudaEventDestroy(x->a); cudaEventDestroy(x->b);
    delete x;
    return EORB_OK;
}
extern "C" int eorb_timer_start(void* t, void* s) { CU(cudaEventRecord(((EorbTimer*)t)->a, (cudaStream_t)s)); return EORB_OK; }
extern "C" int eorb_timer_stop(void* t, void* s) { CU(cudaEventRecord(((EorbTimer*)t)->b, (cudaStream_t)s)); return EORB_OK; }
extern "C" int eorb_timer_elapsed_ms(void* t, float* ms) {
    EorbTimer* x = (EorbTimer*)t;
    CU(cudaEventSynchronize(x->b));
    CU(cudaEventElapsedTime(ms, x->a, x->b));
    return EORB_OK;
}

extern "C" int eorb_probe_popc_rate(int device, double* rate) {
    if (!rate) return fail(EORB_ERR_ARG, "null rate");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    unsigned* d = nullptr;
    CU(devAlloc(&d, 4));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    CU(launch_popc_probe(d, blocks, 64, 0));   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CU(cudaEventRecord(a, 0));
        CU(launch_popc_probe(d, blocks, iters, 0));
        CU(cudaEventRecord(b, 0));
        CU(cudaEventSynchronize(b));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, a, b));
        best = std::min(best, ms);
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    *rate = (double)blocks * 256.0 * 8.0 * iters / (best * 1e-3);
    return EORB_OK;
}

// Runs the shared scalar arithmetic (eorb_math.cuh) on the device and compares it with the host compilation of
// the same header on pseudo-random inputs.  *mismatches = number of differing results (0 expected).
#include "eorb_math.cuh"
extern "C" int eorb_selftest_math(int device, int* mismatches) {
    if (!mismatches) return fail(EORB_ERR_ARG, "null argument");
    CU(cudaSetDevice(device));
    const int nF = 20000, nA = 20000;
    std::vector<int> fin((size_t)nF * 17), fout(nF), bout((size_t)nA * 2);
    std::vector<float> ain((size_t)nA * 2), aout(nA);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 20); };
    for (int i = 0; i < nF; i++) {
        const int base = rnd() % 256, spread = (i % 4 == 0) ? 255 : (int)(rnd() % 40) + 1;
        for (int k = 0; k < 17; k++) {
            int v = base + (int)(rnd() % (2 * spread + 1)) - spread;
            fin[(size_t)17 * i + k] = v < 0 ? 0 : (v > 255 ? 255 : v);
        }
        if (i % 3 == 0) {   // plant a contiguous dark/bright arc so that real corners are exercised
            const int start = rnd() % 16, len = 8 + rnd() % 4, delta = (rnd() & 1) ? 60 : -60;
            for (int k = 0; k < len; k++) {
                int v = fin[(size_t)17 * i] + delta + (int)(rnd() % 9) - 4;
                fin[(size_t)17 * i + 1 + (start + k) % 16] = v < 0 ? 0 : (v > 255 ? 255 : v);
            }
        }
    }
    for (int i = 0; i < nA; i++) {
        ain[2 * i] = (float)((int)(rnd() % 6000001) - 3000000);
        ain[2 * i + 1] = (float)((int)(rnd() % 6000001) - 3000000);
        if (i < 16) { ain[2 * i] = (float)((i & 3) - 1); ain[2 * i + 1] = (float)(((i >> 2) & 3) - 1); }
    }
    int *d_fin = nullptr, *d_fout = nullptr, *d_bout = nullptr; float *d_ain = nullptr, *d_aout = nullptr;
    CU(devAlloc(&d_fin, fin.size())); CU(devAlloc(&d_fout, fout.size())); CU(devAlloc(&d_bout, bout.size()));
    CU(devAlloc(&d_ain, ain.size())); CU(devAlloc(&d_aout, aout.size()));
    CU(cudaMemcpy(d_fin, fin.data(), fin.size() * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_ain, ain.data(), ain.size() * 4, cudaMemcpyHostToDevice));
    CU(launch_selftest_math(d_fin, nF, d_fout, d_ain, nA, d_aout, d_bout, 0));
    CU(cudaMemcpy(fout.data(), d_fout, fout.size() * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(aout.data(), d_aout, aout.size() * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(bout.data(), d_bout, bout.size() * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_fin); cudaFree(d_fout); cudaFree(d_bout); cudaFree(d_ain); cudaFree(d_aout);
    int bad = 0;
    for (int i = 0; i < nF; i++) bad += (fout[i] != eorb::fast_max_arc_min(fin[(size_t)17 * i], &fin[(size_t)17 * i + 1]));
    for (int i = 0; i < nA; i++) {
        const float ang = eorb::fast_atan2_deg(ain[2 * i], ain[2 * i + 1]);
        uint32_t a, b;
        memcpy(&a, &ang, 4); memcpy(&b, &aout[i], 4);
        bad += (a != b);
        const float rad = eorb::fmul(ang, (float)(M_PI / 180.f));
        const float ca = (float)std::cos((double)rad), sa = (float)std::sin((double)rad);
        int r, c;
        eorb::brief_offset((i % 27) - 13, ((i / 27) % 27) - 13, ca, sa, r, c);
        bad += (r != bout[2 * i] || c != bout[2 * i + 1]);
    }
    *mismatches = bad;
    return EORB_OK;
}

// ================================================================================================ ORB
struct eorb_orb {
    eorb_orb_params par{};
    int device = 0, maxBatch = 1, pipeBatch = 1, sms = 148;   // pipeBatch: frames per slot of the host pipeline
    cudaStream_t ownStream = nullptr, stream = nullptr;
    int nlevels = 0, edge = 19;
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> quota;
    int umax[16];
    // plan
    int planW = 0, planH = 0;
    OrbPlan hp{};
    std::vector<CellPlan> cells;
    OrbPlan* d_plan = nullptr; CellPlan* d_cells = nullptr; short4* d_xtab = nullptr; int4* d_ytab = nullptr; int4* d_ytabT = nullptr; int2* d_icTab = nullptr;
    float* d_invScale = nullptr;
    // slabs (maxBatch frames): `main` serves the device entry points and single calls; `pipe` holds the extra
    // slots (own stream + slabs + pinned staging) that eorb_orb_extract_batch cycles through so that the H2D copy
    // of chunk i+1, the kernels of chunk i and the D2H copy of chunk i-1 overlap
    int pitch0 = 0;
    int cap = 0;
    struct Bufs {
        uint8_t* d_img0 = nullptr; uint8_t* d_pyr = nullptr; uint8_t* d_blur = nullptr;
        uint16_t* d_cellCount = nullptr; uint32_t* d_cand = nullptr; uint32_t* d_okeys = nullptr; uint16_t* d_knode = nullptr;
        uint32_t* d_sel = nullptr; int* d_selCount = nullptr; int* d_candCount = nullptr; int* d_dstIdx = nullptr;
        uint32_t* d_kpList = nullptr;
        float* d_levelAngle = nullptr;
        CUtensorMap* d_tmaps = nullptr;   // [nlevels] maps of this slab's pyramid levels >= 1
        int* d_pyrDone = nullptr;            // [8][EORB_MAX_LEVELS][64 block rows] (pyr_chain_kernel)
        CUtensorMap* d_icMaps = nullptr;     // [nlevels] maps of the levels >= 1, box = one keypoint's orientation patch (orient_desc_kernel)
        CUtensorMap* d_briefMaps = nullptr;  // [nlevels] maps of the blurred levels, box = one keypoint's BRIEF patch (orient_desc_kernel)
        CUtensorMap* d_blurMaps = nullptr;   // [nlevels] the same levels with blur_tma_kernel's box (EORB_BLUR_TMA=1)
        CUtensorMap pyrMaps[EORB_MAX_LEVELS];   // host: source map of the TMA-staged resize INTO level l (l >= 2; level 1's source is the caller's frame)
        eorb_keypoint* d_outKps = nullptr; uint8_t* d_outDesc = nullptr; int* d_outN = nullptr; int* d_outMono = nullptr;
        eorb_keypoint* h_kps = nullptr; uint8_t* h_desc = nullptr; int* h_n = nullptr; int* h_mono = nullptr;   // pinned
        cudaStream_t stream = nullptr; cudaEvent_t done = nullptr;   // pipeline slots only
        int f0 = 0, nb = 0; bool pending = false;
    };
    Bufs main;
    std::vector<Bufs> pipe;
    Bufs* last = nullptr;          // buffers of the most recent launch set (stage taps)
    // last call (for the stage taps)
    const uint8_t* lastLvl0 = nullptr; long long lastPitch0 = 0, lastFrameStride0 = 0; int lastFrames = 0;
    long long launches = 0;
    // optional per-stage timing (bench.py): event sets recorded around every stage of every call
    bool stageTiming = false;
    std::vector<cudaEvent_t> evPool;   // (EORB_ORB_STAGES+1) events per recorded call
    size_t evCalls = 0;
    long long stageLaunches[EORB_ORB_STAGES] = {0, 0, 0, 0, 0, 0};
    // CUDA graph of one launch set on the main slab + its result copies (the single-frame call is launch bound: 12 kernels and
    // 4 copies per frame); replayed while (frames, lapping area, descriptors wanted) stay the same, rebuilt otherwise
    cudaGraphExec_t graphExec = nullptr;
    int graphKey[4] = {-1, 0, 0, 0};
    long long graphLaunches = 0;   // kernel launches inside the captured graph
    bool useGraph = true;          // EORB_ORB_GRAPH=0 disables
    OrbFork graphFork;             // side stream + events of the captured single-call graph (blur beside FAST / octree)
    bool fastPadTile = true;       // EORB_FAST_PAD=0: FAST tile pitch left at the next multiple of 16 (for A/B)
    int useBlurTma = 3;           // blur_tma_kernel variant (TMA-staged strips; 3 = 64-row bands, neighbour words read from the tile: 0.705 -> 0.609 us/frame); EORB_BLUR_TMA=0: blur_kernel (direct global loads), for A/B
    bool useBriefTma = true;       // EORB_BRIEF_TMA=0: orient_desc_kernel gathers the BRIEF samples from global memory (for A/B)
    bool usePyrChain = true;       // EORB_PYR_CHAIN=0: small batches launch the pyramid level by level like launch sets do
    bool usePyrTma = true;         // EORB_PYR_TMA=0: every pyramid level through pyr_resize_kernel (direct global loads), for A/B
    int pyrTileRows = 112;         // EORB_PYR_TH: destination rows per TMA-staged tile (sweep on B200, us/frame of the pyramid: 32 -> 0.657, 48 -> 0.609, 64 -> 0.591, 80 -> 0.575, 96 -> 0.570, 112 -> 0.568, 128 -> 0.570, 144 -> 0.590, 192 -> 0.606)
};

static cudaEvent_t* orbStageEvents(eorb_orb* h) {
    if (!h->stageTiming) return nullptr;
    const size_t per = EORB_ORB_STAGES + 1;
    if ((h->evCalls + 1) * per > h->evPool.size()) {
        const size_t old = h->evPool.size();
        h->evPool.resize(old + 64 * per);
        for (size_t i = old; i < h->evPool.size(); i++) cudaEventCreate(&h->evPool[i]);
    }
    cudaEvent_t* e = &h->evPool[h->evCalls * per];
    h->evCalls++;
    const int lv = h->hp.nlevels;
    h->stageLaunches[0] += lv - 1; h->stageLaunches[1]++; h->stageLaunches[2]++; h->stageLaunches[3]++;
    h->stageLaunches[4]++; h->stageLaunches[5]++;
    return e;
}

static void orbFreeBufs(eorb_orb::Bufs& b) {
    cudaFree(b.d_img0); cudaFree(b.d_pyr); cudaFree(b.d_blur); cudaFree(b.d_cellCount); cudaFree(b.d_cand);
    cudaFree(b.d_okeys); cudaFree(b.d_knode); cudaFree(b.d_sel); cudaFree(b.d_selCount); cudaFree(b.d_candCount);
    cudaFree(b.d_dstIdx); cudaFree(b.d_kpList); cudaFree(b.d_levelAngle); cudaFree(b.d_tmaps); cudaFree(b.d_pyrDone); cudaFree(b.d_blurMaps); cudaFree(b.d_briefMaps); cudaFree(b.d_icMaps); cudaFree(b.d_outKps); cudaFree(b.d_outDesc);
    cudaFree(b.d_outN); cudaFree(b.d_outMono);
    cudaFreeHost(b.h_kps); cudaFreeHost(b.h_desc); cudaFreeHost(b.h_n); cudaFreeHost(b.h_mono);
    if (b.done) cudaEventDestroy(b.done);
    if (b.stream) cudaStreamDestroy(b.stream);
    b = eorb_orb::Bufs();
}

static int orbAllocBufs(eorb_orb* h, eorb_orb::Bufs& b, bool pipeline) {
    const OrbPlan& P = h->hp;
    const size_t B = (size_t)(pipeline ? h->pipeBatch : h->maxBatch);
    const int nl = h->nlevels;
    CU(devAlloc(&b.d_img0, B * (size_t)h->pitch0 * P.H));
    CU(devAlloc(&b.d_pyr, B * (size_t)P.pyrBytesPerFrame));
    CU(devAlloc(&b.d_blur, B * (size_t)P.blurBytesPerFrame));
    CU(devAlloc(&b.d_cellCount, B * (size_t)std::max(P.nCells, 1)));
    CU(devAlloc(&b.d_cand, B * (size_t)P.slotsPerFrame));
    CU(devAlloc(&b.d_okeys, B * (size_t)P.slotsPerFrame));
    CU(devAlloc(&b.d_knode, B * (size_t)P.slotsPerFrame));
    CU(devAlloc(&b.d_sel, B * (size_t)P.selPerFrame));
    CU(devAlloc(&b.d_selCount, B * (size_t)nl));
    CU(devAlloc(&b.d_candCount, B * (size_t)nl));
    CU(devAlloc(&b.d_dstIdx, B * (size_t)P.selPerFrame));
    CU(devAlloc(&b.d_kpList, B * (size_t)P.selPerFrame));
    CU(devAlloc(&b.d_levelAngle, B * (size_t)P.selPerFrame));
    {   // TMA maps of the pyramid levels held by this slab (FAST stages its cell tiles with them)
        std::vector<CUtensorMap> maps((size_t)nl);
        memset(maps.data(), 0, maps.size() * sizeof(CUtensorMap));
        for (int l = 1; l < nl; l++) {
            int rct = tmaEncodeFrames(&maps[l], b.d_pyr + P.lv[l].off, P.lv[l].w, P.lv[l].h, (int)B, (size_t)P.lv[l].pitch,
                                      (size_t)P.pyrBytesPerFrame, P.cellTileStride, P.cellTileRows);
            if (rct != EORB_OK) return rct;
        }
        CU(devAlloc(&b.d_tmaps, (size_t)nl));
        CU(cudaMemcpy(b.d_tmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
        CU(devAlloc(&b.d_pyrDone, (size_t)8 * EORB_MAX_LEVELS * 64));   // [8][EORB_MAX_LEVELS][EORB_PYR_CHAIN_ROWS]
        bool blurTmaOk = h->useBlurTma != 0;
        for (int l = 0; l < nl; l++)   // degenerate levels (rows bounce more than once) stay with blur_kernel's per-pixel path
            if (P.lv[l].w > 0 && P.lv[l].h > 0 && (P.lv[l].h < 3 || P.lv[l].w < 8)) blurTmaOk = false;
        if (blurTmaOk) {
            for (int l = 1; l < nl; l++) {
                int rct = tmaEncodeFrames(&maps[l], b.d_pyr + P.lv[l].off, P.lv[l].w, P.lv[l].h, (int)B, (size_t)P.lv[l].pitch,
                                          (size_t)P.pyrBytesPerFrame, blur_tma_box_w(), blur_tma_box_h(h->useBlurTma));
                if (rct != EORB_OK) return rct;
            }
            CU(devAlloc(&b.d_blurMaps, (size_t)nl));
            CU(cudaMemcpy(b.d_blurMaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
        }
        if (h->useBriefTma) {   // the blurred levels with the box of one keypoint's BRIEF patch
            bool ok = true;
            for (int l = 0; l < nl && ok; l++) {
                if (P.lv[l].w <= 0 || P.lv[l].h <= 0) { memset(&maps[l], 0, sizeof(CUtensorMap)); continue; }
                ok = tmaEncodeFrames(&maps[l], b.d_blur + P.lv[l].blurOff, P.lv[l].w, P.lv[l].h, (int)B, (size_t)P.lv[l].bpitch,
                                     (size_t)P.blurBytesPerFrame, brief_tma_box_w(), brief_tma_box_h()) == EORB_OK;
            }
            if (ok) {
                CU(devAlloc(&b.d_briefMaps, (size_t)nl));
                CU(cudaMemcpy(b.d_briefMaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
            }
            memset(&maps[0], 0, sizeof(CUtensorMap));   // level 0 = the caller's frames: its map is made per call (orbIcMap0)
            for (int l = 1; l < nl && ok; l++) {
                if (P.lv[l].w <= 0 || P.lv[l].h <= 0) { memset(&maps[l], 0, sizeof(CUtensorMap)); continue; }
                ok = tmaEncodeFrames(&maps[l], b.d_pyr + P.lv[l].off, P.lv[l].w, P.lv[l].h, (int)B, (size_t)P.lv[l].pitch,
                                     (size_t)P.pyrBytesPerFrame, ic_tma_box_w(), ic_tma_box_h()) == EORB_OK;
            }
            if (ok) {
                CU(devAlloc(&b.d_icMaps, (size_t)nl));
                CU(cudaMemcpy(b.d_icMaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
            }
        }
        memset(b.pyrMaps, 0, sizeof(b.pyrMaps));
        for (int l = 2; l < nl; l++) {
            if (P.lv[l].pyrTW <= 0) continue;
            int rct = tmaEncodeFrames(&b.pyrMaps[l], b.d_pyr + P.lv[l - 1].off, P.lv[l - 1].w, P.lv[l - 1].h, (int)B, (size_t)P.lv[l - 1].pitch,
                                      (size_t)P.pyrBytesPerFrame, P.lv[l].pyrBW, P.lv[l].pyrBH);
            if (rct != EORB_OK) return rct;
        }
    }
    CU(devAlloc(&b.d_outKps, B * (size_t)h->cap));
    CU(devAlloc(&b.d_outDesc, B * (size_t)h->cap * 32));
    CU(devAlloc(&b.d_outN, B));
    CU(devAlloc(&b.d_outMono, B));
    CU(cudaMallocHost((void**)&b.h_kps, B * (size_t)h->cap * sizeof(eorb_keypoint)));
    CU(cudaMallocHost((void**)&b.h_desc, B * (size_t)h->cap * 32));
    CU(cudaMallocHost((void**)&b.h_n, B * sizeof(int)));
    CU(cudaMallocHost((void**)&b.h_mono, B * sizeof(int)));
    if (pipeline) {
        CU(cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
    }
    return EORB_OK;
}

static void orbDropGraph(eorb_orb* h) {
    if (h->graphExec) cudaGraphExecDestroy(h->graphExec);
    h->graphExec = nullptr; h->graphKey[0] = -1;
}

static void orbFreePlan(eorb_orb* h) {
    orbDropGraph(h);
    cudaFree(h->d_plan); cudaFree(h->d_cells); cudaFree(h->d_xtab); cudaFree(h->d_ytab); cudaFree(h->d_ytabT); cudaFree(h->d_invScale); cudaFree(h->d_icTab);
    h->d_plan = nullptr; h->d_cells = nullptr; h->d_xtab = nullptr; h->d_ytab = nullptr; h->d_ytabT = nullptr; h->d_invScale = nullptr; h->d_icTab = nullptr;
    orbFreeBufs(h->main);
    for (auto& b : h->pipe) orbFreeBufs(b);
    h->pipe.clear();
    h->last = nullptr;
    h->planW = h->planH = 0;
}

// ORBextractor::ORBextractor (ORBextractor.cc:420-489): scale tables, per-level quotas, umax, edge threshold
static void orbTables(eorb_orb* h) {
    const eorb_orb_params& p = h->par;
    const int nl = p.nlevels;
    h->nlevels = nl;
    const double sf = (double)p.scaleFactor;   // member is a double holding the float value (ORBextractor.h:123)
    h->scale.assign(nl, 1.f); h->sigma2.assign(nl, 1.f);
    for (int i = 1; i < nl; i++) {
        h->scale[i] = (float)((double)h->scale[i - 1] * sf);
        h->sigma2[i] = h->scale[i] * h->scale[i];
    }
    h->invScale.resize(nl); h->invSigma2.resize(nl);
    for (int i = 0; i < nl; i++) { h->invScale[i] = 1.0f / h->scale[i]; h->invSigma2[i] = 1.0f / h->sigma2[i]; }
    h->quota.assign(nl, 0);
    const float factor = (float)(1.0 / sf);
    float desired = (float)p.nfeatures * (1.f - factor) / (1.f - (float)std::pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
        h->quota[l] = rne(desired);
        sum += h->quota[l];
        desired *= factor;
    }
    h->quota[nl - 1] = std::max(p.nfeatures - sum, 0);
    // circular patch row extents (:463-478)
    for (int v = 0; v < 16; v++) h->umax[v] = 0;
    const int vmax = (int)std::floor(15 * std::sqrt(2.f) / 2 + 1), vmin = (int)std::ceil(15 * std::sqrt(2.f) / 2);
    for (int v = 0; v <= vmax; ++v) h->umax[v] = (int)lrint(std::sqrt(225.0 - (double)v * v));
    for (int v = 15, v0 = 0; v >= vmin; --v) {
        while (h->umax[v0] == h->umax[v0 + 1]) ++v0;
        h->umax[v] = v0;
        ++v0;
    }
    if (p.edgeTh < 0) {
        int e = (int)(19.f * ((float)p.imW / 752.f));
        e += (e % 2 - 1);
        h->edge = e;
    } else {
        h->edge = p.edgeTh;
    }
}

static int orbBuildPlan(eorb_orb* h, int W, int H) {
    if (h->planW == W && h->planH == H) return EORB_OK;
    if (W < 1 || H < 1 || W > EORB_MAX_DIM || H > EORB_MAX_DIM) return fail(EORB_ERR_ARG, "image size %dx%d unsupported (1..%d)", W, H, EORB_MAX_DIM);
    CU(cudaStreamSynchronize(h->stream));
    orbFreePlan(h);
    OrbPlan& P = h->hp;
    memset(&P, 0, sizeof(P));
    const int nl = h->nlevels, E = h->edge;
    // the FAST region starts at column / row E - 3 of the level (:792-795): below 3 the reference reads outside its bordered image
    if (E < 3) return fail(EORB_ERR_ARG, "edge threshold %d < 3 is unsupported (the FAST region would start outside the image)", E);
    P.nlevels = nl; P.edge = E; P.iniTh = h->par.iniThFAST; P.minTh = h->par.minThFAST; P.W = W; P.H = H;
    for (int i = 0; i < 16; i++) P.umax[i] = h->umax[i];
    h->cells.clear();
    std::vector<short4> xtab;
    std::vector<int4> ytab;   // {sy0, sy1, b0 << 16, b1 << 16}
    long long pyrOff = 0, blurOff = 0;
    int slot = 0, sel = 0, rowBlocks = 0, maxCW = 7, maxCH = 7, octSmem = 0, maxTasks = 1;
    for (int l = 0; l < nl; l++) {
        LevelPlan& lp = P.lv[l];
        lp.w = rne((float)W * h->invScale[l]);
        lp.h = rne((float)H * h->invScale[l]);
        if (lp.w < 1 || lp.h < 1) return fail(EORB_ERR_ARG, "pyramid level %d is empty (%dx%d)", l, lp.w, lp.h);
        lp.pitch = roundUp(lp.w, 16);
        lp.bpitch = roundUp(lp.w, 16);
        lp.off = 0;
        if (l > 0) { lp.off = pyrOff; pyrOff += (long long)lp.pitch * lp.h; }
        lp.blurOff = blurOff; blurOff += (long long)lp.bpitch * lp.h;
        lp.scale = h->scale[l];
        lp.sizeF = (float)(int)(31 * h->scale[l]);
        lp.quota = h->quota[l];
        lp.blurTaskBase = rowBlocks; rowBlocks += ((lp.h + EORB_BLUR_BAND - 1) / EORB_BLUR_BAND) * ((lp.w + 127) / 128);   // (EORB_BLUR_BAND rows) x (128 columns)
        // FAST grid (:792-828)
        lp.minBX = E - 3; lp.minBY = E - 3; lp.maxBX = lp.w - E + 3; lp.maxBY = lp.h - E + 3;
        const float width = (float)(lp.maxBX - lp.minBX), height = (float)(lp.maxBY - lp.minBY);
        lp.nCols = (int)(width / 30.f); lp.nRows = (int)(height / 30.f);
        lp.cellBase = (int)h->cells.size(); lp.slotBase = slot;
        if (lp.nCols > 0 && lp.nRows > 0) {
            lp.wCell = (int)std::ceil(width / lp.nCols); lp.hCell = (int)std::ceil(height / lp.nRows);
            for (int i = 0; i < lp.nRows; i++) {
                const int iniY = lp.minBY + i * lp.hCell;
                int maxY = iniY + lp.hCell + 6;
                if (iniY >= lp.maxBY - 3) continue;
                if (maxY > lp.maxBY) maxY = lp.maxBY;
                for (int j = 0; j < lp.nCols; j++) {
                    const int iniX = lp.minBX + j * lp.wCell;
                    int maxX = iniX + lp.wCell + 6;
                    if (iniX >= lp.maxBX - 3) continue;
                    if (maxX > lp.maxBX) maxX = lp.maxBX;
                    CellPlan c{};
                    c.x0 = (short)iniX; c.y0 = (short)iniY; c.w = (short)(maxX - iniX); c.h = (short)(maxY - iniY);
                    c.level = (short)l; c.ox = (short)(j * lp.wCell); c.oy = (short)(i * lp.hCell);
                    const int cw = c.w - 6, ch = c.h - 6;
                    c.slotCap = (cw > 0 && ch > 0) ? ((cw + 1) / 2) * ((ch + 1) / 2) : 0;   // strict 3x3 maxima cannot touch
                    c.slotOff = slot; slot += c.slotCap;
                    {
                        const int aoff = iniX & 15, p0 = (aoff + 3) >> 3;
                        const int np = std::max(((aoff + c.w - 4) >> 3) - p0 + 1, 1);
                        const int lastBits = aoff + c.w - 3 - 8 * (p0 + np - 1);
                        c.aoff = (unsigned char)aoff; c.p0 = (unsigned char)p0; c.np = (unsigned char)np; c.rps = (unsigned char)(32 / np);
                        c.firstMask = (unsigned char)((0xffu << ((aoff + 3) & 7)) & 0xffu);
                        c.lastMask = (unsigned char)(lastBits >= 8 ? 0xffu : (lastBits <= 0 ? 0u : ((1u << lastBits) - 1u)));
                        c.rcpNpM1 = (unsigned short)((65536 + np - 1) / np - 1);
                        maxTasks = std::max(maxTasks, np * std::max(ch, 0));
                    }
                    maxCW = std::max(maxCW, (int)c.w); maxCH = std::max(maxCH, (int)c.h);
                    h->cells.push_back(c);
                }
            }
        }
        lp.nCells = (int)h->cells.size() - lp.cellBase;
        lp.slotCount = slot - lp.slotBase;
        if (lp.slotCount > (1 << 20)) return fail(EORB_ERR_ARG, "level %d too large for the 20-bit candidate index", l);
        // octree roots (:562-563)
        lp.nIni = 0; lp.hX = 1.f;
        if (lp.maxBX > lp.minBX && lp.maxBY > lp.minBY) {
            lp.nIni = (int)std::round((float)(lp.maxBX - lp.minBX) / (float)(lp.maxBY - lp.minBY));
            if (lp.nIni > 0) lp.hX = (float)(lp.maxBX - lp.minBX) / (float)lp.nIni;
        }
        lp.nodeCap = std::max(lp.quota + 3, 4 * lp.nIni) + 1;
        if (lp.nodeCap > 65535) return fail(EORB_ERR_ARG, "nfeatures too large (node capacity %d)", lp.nodeCap);
        lp.selBase = sel; sel += lp.nodeCap;
        octSmem = std::max(octSmem, (int)oct_smem_bytes(lp.nodeCap));
        // resize taps (cv::resize INTER_LINEAR 8-bit, reference call :1253)
        lp.xtabOff = (int)xtab.size(); lp.ytabOff = (int)ytab.size();
        if (l > 0) {
            const int sw = P.lv[l - 1].w, sh = P.lv[l - 1].h;
            const double scale_x = 1.0 / ((double)lp.w / sw), scale_y = 1.0 / ((double)lp.h / sh);
            for (int dx = 0; dx < lp.w; dx++) {
                float fx = (float)((dx + 0.5) * scale_x - 0.5);
                int sx = (int)std::floor(fx);
                fx -= sx;
                if (sx < 0) { fx = 0; sx = 0; }
                if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
                short4 t;
                t.x = (short)sx; t.y = (short)std::min(sx + 1, sw - 1);
                t.z = satShort((1.f - fx) * 2048.f); t.w = satShort(fx * 2048.f);
                xtab.push_back(t);
            }
            for (int dy = 0; dy < lp.h; dy++) {
                float fy = (float)((dy + 0.5) * scale_y - 0.5);
                int sy = (int)std::floor(fy);
                fy -= sy;
                int4 t;
                t.x = std::min(std::max(sy, 0), sh - 1); t.y = std::min(std::max(sy + 1, 0), sh - 1);
                t.z = (int)satShort((1.f - fy) * 2048.f) << 16; t.w = (int)satShort(fy * 2048.f) << 16;
                ytab.push_back(t);
            }
            // TMA-staged resize (orb_tiles.cu): even partition of the level into destination tiles of at most 128 x 64 pixels; the
            // source box is the largest footprint over the tiles (first tap column rounded down to 16 bytes .. last 8-byte window)
            lp.pyrTW = lp.pyrTH = lp.pyrBW = lp.pyrBH = 0;
            if (h->usePyrTma && sw >= 16 && sh >= 2 && lp.w >= 4 && scale_x <= 2.0 && scale_y <= 4.0) {
                const int ntx = (lp.w + 127) / 128;
                const int TW = roundUp((lp.w + ntx - 1) / ntx, 4);
                const int thMax = h->pyrTileRows;
                int TH = roundUp((lp.h + (lp.h + thMax - 1) / thMax - 1) / ((lp.h + thMax - 1) / thMax), 4);
                for (;;) {
                    int BW = 16, BH = 2;
                    for (int x0 = 0; x0 < lp.w; x0 += TW) {
                        const int first = xtab[lp.xtabOff + x0].x & ~15, last = xtab[lp.xtabOff + std::min(x0 + TW, lp.w) - 1].x;
                        BW = std::max(BW, roundUp((last - first) + 8, 16));
                    }
                    for (int y0 = 0; y0 < lp.h; y0 += TH)
                        BH = std::max(BH, ytab[lp.ytabOff + std::min(y0 + TH, lp.h) - 1].y - ytab[lp.ytabOff + y0].x + 1);
                    if (BW <= 256 && BH <= 256 && BW * BH + 32 <= 40 * 1024) { lp.pyrTW = TW; lp.pyrTH = TH; lp.pyrBW = BW; lp.pyrBH = BH; break; }
                    if (TH <= 8 || BW > 256) break;       // cannot be tiled within the box limits: direct-load kernel
                    TH = roundUp(TH / 2, 4);
                }
            }
        }
    }
    P.nCells = (int)h->cells.size();
    P.slotsPerFrame = std::max(slot, 1);
    P.selPerFrame = sel;
    P.pyrBytesPerFrame = std::max(pyrOff, 16ll);
    P.blurBytesPerFrame = blurOff;
    P.blurTasksTotal = rowBlocks;
    // FAST smem region of one warp: [TMA tile BW x BH][score map (ch+2) x MS][survivor list u16][flagged groups u32][mbarrier]
    P.cellTileStride = roundUp(maxCW + 15, 16);   // the TMA box starts at x0 & ~15 (16-byte inner-coordinate rule)
    // A row pitch that is a multiple of 32 words maps rows two apart onto the same banks: the byte gathers of the exact-score phase
    // (lanes on different rows of one cell) then serialise 2.5-2.7 deep (ncu r01k: 46 % of the kernel's shared-memory wavefronts were
    // bank conflicts, and the shared-memory pipe was its busiest unit at 82 %).  16 bytes more spread eight rows over all banks.
    if (h->fastPadTile && P.cellTileStride % 32 == 0 && P.cellTileStride + 16 <= 256) P.cellTileStride += 16;
    P.cellTileRows = maxCH;
    P.cellMapStride = roundUp(maxCW - 6 + 2, 4);
    P.cellMapOff = roundUp(P.cellTileRows * P.cellTileStride, 16);
    P.cellListOff = P.cellMapOff + roundUp((maxCH - 6 + 2) * P.cellMapStride, 16);
    P.cellTaskOff = roundUp(P.cellListOff + (maxCW - 6) * (maxCH - 6) * 2, 16);
    P.cellBarOff = roundUp(P.cellTaskOff + maxTasks * 4, 16);
    P.cellSmemPerWarp = roundUp(P.cellBarOff + 16, 128);
    if (P.cellTileStride > 256 || P.cellTileRows > 256) return fail(EORB_ERR_ARG, "FAST cell %dx%d exceeds the TMA box limit", maxCW, maxCH);
    P.octSmemBytes = octSmem;
    if (P.cellSmemPerWarp * EORB_FAST_WARPS > 200 * 1024 || octSmem > 200 * 1024)
        return fail(EORB_ERR_ARG, "shared-memory budget exceeded (fast %d B, octree %d B)", P.cellSmemPerWarp * EORB_FAST_WARPS, octSmem);

    h->pitch0 = roundUp(W, 16);
    h->cap = eorb_orb_max_keypoints_for_size(h, W, H);
    CU(devAlloc(&h->d_plan, 1));
    CU(devAlloc(&h->d_cells, h->cells.size()));
    CU(devAlloc(&h->d_xtab, xtab.size()));
    CU(devAlloc(&h->d_ytab, ytab.size()));
    CU(devAlloc(&h->d_ytabT, ytab.size()));
    CU(devAlloc(&h->d_invScale, (size_t)nl));
    {   // IC_Angle weight table (:77-104): for alignment a = (x-15)&3, row r = v+15, aligned word k the four columns are
        // u = 4k + j - a - 15; weight u (m10) / v (m01) inside the circle |u| <= umax[|v|], 0 outside; row 31 is padding
        std::vector<int2> tab(4 * 288);
        for (int al = 0; al < 4; al++)
            for (int i = 0; i < 288; i++) {
                const int r = i / 9, k = i % 9, v = r - 15;
                uint32_t w10 = 0, w01 = 0;
                for (int j = 0; j < 4 && r < 31; j++) {
                    const int u = 4 * k + j - al - 15;
                    if (u < -15 || u > 15 || std::abs(u) > h->umax[std::abs(v)]) continue;
                    w10 |= (uint32_t)(uint8_t)(int8_t)u << (8 * j);
                    w01 |= (uint32_t)(uint8_t)(int8_t)v << (8 * j);
                }
                tab[(size_t)al * 288 + i] = make_int2((int)w10, (int)w01);
            }
        CU(devAlloc(&h->d_icTab, tab.size()));
        CU(cudaMemcpy(h->d_icTab, tab.data(), tab.size() * sizeof(int2), cudaMemcpyHostToDevice));
    }
    {
        int rcb = orbAllocBufs(h, h->main, false);
        if (rcb != EORB_OK) return rcb;
    }
    CU(cudaMemcpy(h->d_plan, &P, sizeof(P), cudaMemcpyHostToDevice));
    if (!h->cells.empty()) CU(cudaMemcpy(h->d_cells, h->cells.data(), h->cells.size() * sizeof(CellPlan), cudaMemcpyHostToDevice));
    if (!xtab.empty()) CU(cudaMemcpy(h->d_xtab, xtab.data(), xtab.size() * sizeof(short4), cudaMemcpyHostToDevice));
    if (!ytab.empty()) CU(cudaMemcpy(h->d_ytab, ytab.data(), ytab.size() * sizeof(int4), cudaMemcpyHostToDevice));
    {   // the TMA path's y taps carry tile-row byte offsets (row index x box width of the level)
        std::vector<int4> yt2(ytab);
        for (int l = 1; l < nl; l++)
            for (int dy = 0; dy < P.lv[l].h && P.lv[l].pyrTW > 0; dy++) {
                int4& t = yt2[(size_t)P.lv[l].ytabOff + dy];
                t.x *= P.lv[l].pyrBW; t.y *= P.lv[l].pyrBW;
            }
        if (!yt2.empty()) CU(cudaMemcpy(h->d_ytabT, yt2.data(), yt2.size() * sizeof(int4), cudaMemcpyHostToDevice));
    }
    CU(cudaMemcpy(h->d_invScale, h->invScale.data(), nl * sizeof(float), cudaMemcpyHostToDevice));
    CU(orb_kernels_configure(P));
    h->planW = W; h->planH = H;
    return EORB_OK;
}

static OrbArgs orbArgs(eorb_orb* h, eorb_orb::Bufs& b, const uint8_t* lvl0, long long pitch0, long long frameStride0, int lap0, int lap1,
                       int wantDesc, eorb_keypoint* kps, uint8_t* desc, int cap, int* nOut, int* monoOut) {
    OrbArgs a{};
    a.plan = h->d_plan; a.cells = h->d_cells; a.xtab = h->d_xtab; a.ytab = h->d_ytab; a.ytabT = h->d_ytabT;
    a.lvl0 = lvl0; a.lvl0Pitch = pitch0; a.lvl0FrameStride = frameStride0;
    a.tmaps = b.d_tmaps;
    a.blurMaps = b.d_blurMaps;
    a.briefMaps = b.d_briefMaps;
    a.icMaps = b.d_icMaps;
    a.pyrDone = h->usePyrChain ? b.d_pyrDone : nullptr;
    a.blurVariant = h->useBlurTma;
    a.pyr = b.d_pyr; a.blur = b.d_blur; a.cellCount = b.d_cellCount; a.cand = b.d_cand; a.okeys = b.d_okeys;
    a.knode = b.d_knode; a.sel = b.d_sel; a.selCount = b.d_selCount; a.candCount = b.d_candCount;
    a.dstIdx = b.d_dstIdx; a.kpList = b.d_kpList; a.icTab = h->d_icTab; a.levelAngle = b.d_levelAngle;
    a.outKps = kps; a.outDesc = desc; a.outN = nOut; a.outMono = monoOut; a.cap = cap;
    a.lap0 = lap0; a.lap1 = lap1; a.wantDesc = wantDesc;
    h->last = &b;
    return a;
}

extern "C" int eorb_orb_create(const eorb_orb_params* params, int device, int max_batch, eorb_orb** out) {
    if (!params || !out) return fail(EORB_ERR_ARG, "null argument");
    if (params->nlevels < 1 || params->nlevels > EORB_MAX_LEVELS) return fail(EORB_ERR_ARG, "nlevels must be 1..%d", EORB_MAX_LEVELS);
    if (params->nfeatures < 0 || !(params->scaleFactor >= 1.0f)) return fail(EORB_ERR_ARG, "bad nfeatures/scaleFactor");
    if (max_batch < 1) return fail(EORB_ERR_ARG, "max_batch must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(EORB_ERR_CUDA, "no CUDA device: eorb_b200 has no CPU fallback"); }
    if (device < 0 || device >= ndev) return fail(EORB_ERR_ARG, "device %d out of range", device);
    CU(cudaSetDevice(device));
    std::unique_ptr<eorb_orb> guard(new eorb_orb());      // released to *out only when every CUDA call below succeeded
    eorb_orb* h = guard.get();
    h->par = *params; h->device = device; h->maxBatch = max_batch;
    if (const char* e = getenv("EORB_ORB_GRAPH")) h->useGraph = atoi(e) != 0;
    if (const char* e = getenv("EORB_PYR_TMA")) h->usePyrTma = atoi(e) != 0;
    if (const char* e = getenv("EORB_BLUR_TMA")) h->useBlurTma = atoi(e);
    if (const char* e = getenv("EORB_PYR_CHAIN")) h->usePyrChain = atoi(e) != 0;
    if (const char* e = getenv("EORB_BRIEF_TMA")) h->useBriefTma = atoi(e) != 0;
    if (const char* e = getenv("EORB_FAST_PAD")) h->fastPadTile = atoi(e) != 0;
    if (const char* e = getenv("EORB_PYR_TH")) h->pyrTileRows = std::min(std::max(atoi(e), 8), 200);
    h->pipeBatch = std::min(max_batch, 128);   // measured on B200: H2D/compute/D2H overlap is best with 128-frame slots
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    h->sms = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&h->ownStream, cudaStreamNonBlocking));
    h->stream = h->ownStream;
    orbTables(h);
    *out = guard.release();
    return EORB_OK;
}

static void orbFreeFork(OrbFork& f) {
    if (f.forked) cudaEventDestroy(f.forked);
    if (f.joined) cudaEventDestroy(f.joined);
    if (f.side) cudaStreamDestroy(f.side);
    f = OrbFork();
}

extern "C" int eorb_orb_destroy(eorb_orb* h) {
    if (!h) return EORB_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    orbFreePlan(h);
    for (cudaEvent_t e : h->evPool) cudaEventDestroy(e);
    orbDropGraph(h);
    orbFreeFork(h->graphFork);
    cudaStreamDestroy(h->ownStream);
    delete h;
    return EORB_OK;
}

extern "C" int eorb_orb_set_stream(eorb_orb* h, void* s) {
    if (!h) return fail(EORB_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    orbDropGraph(h);
    h->stream = (cudaStream_t)s;   // NULL is the CUDA legacy default stream, a legitimate choice
    return EORB_OK;
}
extern "C" int eorb_orb_reset_stream(eorb_orb* h) {
    if (!h) return fail(EORB_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    orbDropGraph(h);               // as set_stream: a graph captured for another stream configuration is not reused
    h->stream = h->ownStream;
    return EORB_OK;
}
extern "C" void* eorb_orb_get_stream(eorb_orb* h) { return h ? (void*)h->stream : nullptr; }
extern "C" int eorb_orb_synchronize(eorb_orb* h) {
    if (!h) return fail(EORB_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    return EORB_OK;
}
extern "C" long long eorb_orb_launch_count(const eorb_orb* h) { return h ? h->launches : 0; }

extern "C" int eorb_orb_stage_timing(eorb_orb* h, int enable) {
    if (!h) return fail(EORB_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    h->stageTiming = enable != 0;
    h->evCalls = 0;
    for (int i = 0; i < EORB_ORB_STAGES; i++) h->stageLaunches[i] = 0;
    return EORB_OK;
}

extern "C" int eorb_orb_stage_times(eorb_orb* h, float* ms6, long long* launches6) {
    if (!h || !ms6) return fail(EORB_ERR_ARG, "null argument");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    const size_t per = EORB_ORB_STAGES + 1;
    for (int i = 0; i < EORB_ORB_STAGES; i++) ms6[i] = 0.f;
    for (size_t c = 0; c < h->evCalls; c++)
        for (int i = 0; i < EORB_ORB_STAGES; i++) {
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, h->evPool[c * per + i], h->evPool[c * per + i + 1]));
            ms6[i] += ms;
        }
    if (launches6) for (int i = 0; i < EORB_ORB_STAGES; i++) launches6[i] = h->stageLaunches[i];
    h->evCalls = 0;
    for (int i = 0; i < EORB_ORB_STAGES; i++) h->stageLaunches[i] = 0;
    return EORB_OK;
}

extern "C" int eorb_orb_tables(const eorb_orb* h, int* nlevels, int* edge, float* scale, float* inv_scale, float* sigma2,
                               float* inv_sigma2, int* fpl) {
    if (!h) return fail(EORB_ERR_ARG, "null handle");
    if (nlevels) *nlevels = h->nlevels;
    if (edge) *edge = h->edge;
    for (int i = 0; i < h->nlevels; i++) {
        if (scale) scale[i] = h->scale[i];
        if (inv_scale) inv_scale[i] = h->invScale[i];
        if (sigma2) sigma2[i] = h->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = h->invSigma2[i];
        if (fpl) fpl[i] = h->quota[i];
    }
    return EORB_OK;
}

extern "C" int eorb_orb_params_tables(const eorb_orb_params* params, int* nlevels, int* edge, float* scale, float* inv_scale, float* sigma2,
                                      float* inv_sigma2, int* fpl) {
    if (!params) return fail(EORB_ERR_ARG, "null argument");
    if (params->nlevels < 1 || params->nlevels > EORB_MAX_LEVELS) return fail(EORB_ERR_ARG, "nlevels must be 1..%d", EORB_MAX_LEVELS);
    if (params->nfeatures < 0 || !(params->scaleFactor >= 1.0f)) return fail(EORB_ERR_ARG, "bad nfeatures/scaleFactor");
    eorb_orb tmp;                      // host fields only; no CUDA call is made
    tmp.par = *params;
    orbTables(&tmp);
    return eorb_orb_tables(&tmp, nlevels, edge, scale, inv_scale, sigma2, inv_sigma2, fpl);
}

// Upper bound of the keypoints one frame of size w x hgt can produce: per level the octree ends with at most quota + 2 nodes
// (its last splits add up to 3 to a list below the quota) or, when the quota is tiny, with the 4 * nIni children of its first
// pass (:562-563, 620-683; nIni = round(width / height) of the level's FAST region, so elongated images have many roots).
extern "C" int eorb_orb_max_keypoints_for_size(const eorb_orb* h, int w, int hgt) {
    if (!h) return 0;
    int cap = 0;
    for (int l = 0; l < h->nlevels; l++) {
        int nIni = 4;
        if (w > 0 && hgt > 0) {
            const int lw = rne((float)w * h->invScale[l]), lh = rne((float)hgt * h->invScale[l]);
            const int bw = lw - 2 * h->edge + 6, bh = lh - 2 * h->edge + 6;   // maxBorder - minBorder
            nIni = (bw > 0 && bh > 0) ? (int)std::round((float)bw / (float)bh) : 0;
        }
        cap += std::max(h->quota[l] + 3, std::max(4 * nIni, 16));
    }
    return cap;
}

extern "C" int eorb_orb_max_keypoints(const eorb_orb* h) {
    if (!h) return 0;
    if (h->planW > 0) return eorb_orb_max_keypoints_for_size(h, h->planW, h->planH);   // the size of the last frames
    return eorb_orb_max_keypoints_for_size(h, h->par.imW, h->par.imH);                 // the size announced at construction
}

// the pyramid's TMA source maps for one launch set: the slab's maps plus level 1's source = the caller's frames
static int orbPyrMaps(eorb_orb* h, eorb_orb::Bufs& b, const uint8_t* lvl0, int w, int hgt, int nframes, long long p0, long long fs0, CUtensorMap* out) {
    memcpy(out, b.pyrMaps, sizeof(b.pyrMaps));
    if (b.d_blurMaps) {   // slot 0 is free (level 0 has no source): the blur's map of level 0 = the caller's frames
        int rcb = tmaEncodeFrames(&out[0], lvl0, w, hgt, nframes, (size_t)p0, (size_t)fs0, blur_tma_box_w(), blur_tma_box_h(h->useBlurTma));
        if (rcb != EORB_OK) return rcb;
    }
    if (h->nlevels > 1 && h->hp.lv[1].pyrTW > 0)
        return tmaEncodeFrames(&out[1], lvl0, w, hgt, nframes, (size_t)p0, (size_t)fs0, h->hp.lv[1].pyrBW, h->hp.lv[1].pyrBH);
    return EORB_OK;
}

// orient_desc_kernel's map of level 0 (the caller's frames) with the orientation patch's box; *use = false when there is none
static int orbIcMap0(eorb_orb::Bufs& b, const uint8_t* lvl0, int w, int hgt, int nframes, long long p0, long long fs0, CUtensorMap* out, bool* use) {
    *use = false;
    if (!b.d_icMaps) return EORB_OK;
    // a frame the box cannot be encoded for is not an error: the kernel then reads the orientation patches from global memory
    *use = tmaEncodeFrames(out, lvl0, w, hgt, nframes, (size_t)p0, (size_t)fs0, ic_tma_box_w(), ic_tma_box_h()) == EORB_OK;
    return EORB_OK;
}

static bool lvl0ZeroCopyOk(const uint8_t* p, int w, size_t rowStride, size_t frameStride) {
    // TMA (FAST cell tiles) needs a 16-byte aligned base and 16-byte strides; the vectorised kernels read whole words
    return ((uintptr_t)p % 16 == 0) && (rowStride % 16 == 0) && (frameStride % 16 == 0) && rowStride >= (size_t)roundUp(w, 4);
}

extern "C" int eorb_orb_extract_batch_device(eorb_orb* h, const uint8_t* d_imgs, int nframes, int w, int hgt, size_t row_stride,
                                             size_t frame_stride, int lap0, int lap1, int want_desc, eorb_keypoint* d_kps,
                                             uint8_t* d_desc, int cap, int* d_n_out, int* d_mono_out) {
    if (!h) return fail(EORB_ERR_ARG, "null handle");
    if (!d_imgs || w <= 0 || hgt <= 0 || nframes <= 0) return EORB_EMPTY;
    if (nframes > h->maxBatch) return fail(EORB_ERR_CAPACITY, "nframes %d > max_batch %d", nframes, h->maxBatch);
    if (!d_kps || !d_n_out || !d_mono_out || (want_desc && !d_desc) || cap < 1) return fail(EORB_ERR_ARG, "null output / cap");
    CU(cudaSetDevice(h->device));
    int rc = orbBuildPlan(h, w, hgt);
    if (rc != EORB_OK) return rc;
    const uint8_t* lvl0 = d_imgs; long long p0 = (long long)row_stride, fs0 = (long long)frame_stride;
    if (!lvl0ZeroCopyOk(d_imgs, w, row_stride, frame_stride)) {
        for (int f = 0; f < nframes; f++)
            CU(cudaMemcpy2DAsync(h->main.d_img0 + (size_t)f * h->pitch0 * hgt, h->pitch0, d_imgs + (size_t)f * frame_stride, row_stride,
                                 w, hgt, cudaMemcpyDeviceToDevice, h->stream));
        lvl0 = h->main.d_img0; p0 = h->pitch0; fs0 = (long long)h->pitch0 * hgt;
    }
    OrbArgs a = orbArgs(h, h->main, lvl0, p0, fs0, lap0, lap1, want_desc, d_kps, d_desc, cap, d_n_out, d_mono_out);
    CUtensorMap tm0;
    rc = tmaEncodeFrames(&tm0, lvl0, w, hgt, nframes, (size_t)p0, (size_t)fs0, h->hp.cellTileStride, h->hp.cellTileRows);
    if (rc != EORB_OK) return rc;
    CUtensorMap pm[EORB_MAX_LEVELS];
    rc = orbPyrMaps(h, h->main, lvl0, w, hgt, nframes, p0, fs0, pm);
    if (rc != EORB_OK) return rc;
    CUtensorMap ic0; bool useIc0 = false;
    rc = orbIcMap0(h->main, lvl0, w, hgt, nframes, p0, fs0, &ic0, &useIc0);
    if (rc != EORB_OK) return rc;
    if (!useIc0) a.icMaps = nullptr;
    CU(launch_orb_pipeline(a, h->hp, nframes, tm0, h->stream, &h->launches, orbStageEvents(h), pm, nullptr, useIc0 ? &ic0 : nullptr));
    h->lastLvl0 = lvl0; h->lastPitch0 = p0; h->lastFrameStride0 = fs0; h->lastFrames = nframes;
    return EORB_OK;
}

// drains one pipeline slot: waits for its D2H copies, then hands the results to the caller's arrays
static int orbCollect(eorb_orb* h, eorb_orb::Bufs& b, int want_desc, eorb_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* mono_out,
                      bool direct) {
    if (!b.pending) return EORB_OK;
    b.pending = false;
    CU(b.done ? cudaEventSynchronize(b.done) : cudaStreamSynchronize(h->stream));
    int status = EORB_OK;
    const int icap = h->cap;
    for (int f = 0; f < b.nb; f++) {
        const int n = b.h_n[f];
        n_out[b.f0 + f] = n;
        if (mono_out) mono_out[b.f0 + f] = b.h_mono[f];
        if (n > cap || n > icap) { status = fail(EORB_ERR_CAPACITY, "frame %d produced %d keypoints > cap %d", b.f0 + f, n, std::min(cap, icap)); continue; }
        if (direct) continue;   // the device wrote straight into the caller's pinned arrays
        memcpy(kps + (size_t)(b.f0 + f) * cap, b.h_kps + (size_t)f * icap, (size_t)n * sizeof(eorb_keypoint));
        if (want_desc) memcpy(desc + (size_t)(b.f0 + f) * cap * 32, b.h_desc + (size_t)f * icap * 32, (size_t)n * 32);
    }
    return status;
}

static bool isPinnedHost(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

extern "C" int eorb_orb_extract_batch(eorb_orb* h, const uint8_t* imgs, int nframes, int w, int hgt, size_t row_stride,
                                      size_t frame_stride, int lap0, int lap1, int want_desc, eorb_keypoint* kps, uint8_t* desc,
                                      int cap, int* n_out, int* mono_out) {
    if (!h) return fail(EORB_ERR_ARG, "null handle");
    if (!imgs || w <= 0 || hgt <= 0 || nframes <= 0) return EORB_EMPTY;
    if (!kps || !n_out || (want_desc && !desc) || cap < 1) return fail(EORB_ERR_ARG, "null output / cap");
    CU(cudaSetDevice(h->device));
    int rc = orbBuildPlan(h, w, hgt);
    if (rc != EORB_OK) return rc;
    // up to maxBatch frames: one launch set on the main slab; more: pipeline slots of pipeBatch frames
    const int B = nframes <= h->maxBatch && nframes <= 2 * h->pipeBatch ? h->maxBatch : h->pipeBatch, icap = h->cap;
    const int nchunks = (nframes + B - 1) / B;
    // more than one chunk: cycle through up to 3 pipeline slots (own streams) so copies and kernels overlap
    const int nslots = nchunks > 1 ? std::min(nchunks, 3) : 0;
    while ((int)h->pipe.size() < nslots) {
        h->pipe.emplace_back();
        int rcb = orbAllocBufs(h, h->pipe.back(), true);
        if (rcb != EORB_OK) { orbFreeBufs(h->pipe.back()); h->pipe.pop_back(); return rcb; }
    }
    // A slot marked pending belongs to THIS call only: flags left over from a call that failed half-way are dropped here, and the
    // guard below drains every slot on any early return, so no D2H copy into the caller's (possibly pinned) arrays is still in
    // flight and no later call collects a stale (f0, nb) range.
    h->main.pending = false;
    for (auto& pb : h->pipe) pb.pending = false;
    struct PendingGuard {
        eorb_orb* h; bool armed = true;
        ~PendingGuard() {
            if (!armed) return;
            for (auto& pb : h->pipe) { if (pb.stream) cudaStreamSynchronize(pb.stream); pb.pending = false; }
            cudaStreamSynchronize(h->stream); h->main.pending = false;
        }
    } pendingGuard{h};
    // when the caller's arrays are pinned and laid out like ours, D2H goes straight into them (no staging copy)
    const bool direct = nslots > 0 && cap == icap && isPinnedHost(kps) && (!want_desc || isPinnedHost(desc));
    if (nslots > 0) CU(cudaStreamSynchronize(h->stream));   // order after earlier work on the handle's stream
    int status = EORB_OK;
    for (int c = 0; c < nchunks; c++) {
        const int f0 = c * B, nb = std::min(B, nframes - f0);
        eorb_orb::Bufs& b = nslots > 0 ? h->pipe[c % nslots] : h->main;
        cudaStream_t st = nslots > 0 ? b.stream : h->stream;
        int rcc = orbCollect(h, b, want_desc, kps, desc, cap, n_out, mono_out, direct);
        if (rcc != EORB_OK) status = rcc;
        if (row_stride == (size_t)w && frame_stride == (size_t)w * hgt && h->pitch0 == w) {
            CU(cudaMemcpyAsync(b.d_img0, imgs + (size_t)f0 * frame_stride, (size_t)nb * frame_stride, cudaMemcpyHostToDevice, st));
        } else {
            for (int f = 0; f < nb; f++)
                CU(cudaMemcpy2DAsync(b.d_img0 + (size_t)f * h->pitch0 * hgt, h->pitch0, imgs + (size_t)(f0 + f) * frame_stride,
                                     row_stride, w, hgt, cudaMemcpyHostToDevice, st));
        }
        OrbArgs a = orbArgs(h, b, b.d_img0, h->pitch0, (long long)h->pitch0 * hgt, lap0, lap1, want_desc, b.d_outKps, b.d_outDesc,
                            icap, b.d_outN, b.d_outMono);
        eorb_keypoint* kdst = direct ? kps + (size_t)f0 * cap : b.h_kps;
        uint8_t* ddst = direct ? desc + (size_t)f0 * cap * 32 : b.h_desc;
        // the launch set and its result copies; with `capturing` they are recorded into a graph instead of being executed
        bool capturing = false;
        auto issue = [&](long long* launchCounter, cudaEvent_t* stageEv) -> int {
            CUtensorMap tm0;
            int rct = tmaEncodeFrames(&tm0, b.d_img0, w, hgt, nb, (size_t)h->pitch0, (size_t)h->pitch0 * hgt, h->hp.cellTileStride, h->hp.cellTileRows);
            if (rct != EORB_OK) return rct;
            CUtensorMap pm[EORB_MAX_LEVELS];
            rct = orbPyrMaps(h, b, b.d_img0, w, hgt, nb, h->pitch0, (long long)h->pitch0 * hgt, pm);
            if (rct != EORB_OK) return rct;
            CUtensorMap ic0; bool useIc0 = false;
            rct = orbIcMap0(b, b.d_img0, w, hgt, nb, h->pitch0, (long long)h->pitch0 * hgt, &ic0, &useIc0);
            if (rct != EORB_OK) return rct;
            OrbArgs aa = a;
            if (!useIc0) aa.icMaps = nullptr;
            CU(launch_orb_pipeline(aa, h->hp, nb, tm0, st, launchCounter, stageEv, pm, capturing ? &h->graphFork : nullptr, useIc0 ? &ic0 : nullptr));
            CU(cudaMemcpyAsync(b.h_n, b.d_outN, nb * sizeof(int), cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(b.h_mono, b.d_outMono, nb * sizeof(int), cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(kdst, b.d_outKps, (size_t)nb * icap * sizeof(eorb_keypoint), cudaMemcpyDeviceToHost, st));
            if (want_desc) CU(cudaMemcpyAsync(ddst, b.d_outDesc, (size_t)nb * icap * 32, cudaMemcpyDeviceToHost, st));
            return EORB_OK;
        };
        // single launch set on the handle's own stream, small batch: replay a CUDA graph (the call is launch bound)
        const bool graphable = h->useGraph && nslots == 0 && !h->stageTiming && st == h->ownStream && nb <= 8;
        if (graphable && !h->graphFork.side) {
            CU(cudaStreamCreateWithFlags(&h->graphFork.side, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&h->graphFork.forked, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&h->graphFork.joined, cudaEventDisableTiming));
        }
        if (graphable) {
            const int key[4] = {nb, lap0, lap1, want_desc};
            if (!h->graphExec || memcmp(key, h->graphKey, sizeof(key)) != 0) {
                orbDropGraph(h);
                cudaGraph_t g = nullptr;
                long long cnt = 0;
                CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                capturing = true;
                const int rci = issue(&cnt, nullptr);
                capturing = false;
                const cudaError_t ce = cudaStreamEndCapture(st, &g);
                if (rci != EORB_OK) { if (g) cudaGraphDestroy(g); return rci; }
                if (ce != cudaSuccess) return fail(EORB_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
                const cudaError_t ie = cudaGraphInstantiate(&h->graphExec, g, 0);
                cudaGraphDestroy(g);
                if (ie != cudaSuccess) { h->graphExec = nullptr; return fail(EORB_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); }
                memcpy(h->graphKey, key, sizeof(key));
                h->graphLaunches = cnt;
            }
            CU(cudaGraphLaunch(h->graphExec, st));
            h->launches += h->graphLaunches;
        } else {
            rc = issue(&h->launches, nslots > 0 ? nullptr : orbStageEvents(h));
            if (rc != EORB_OK) return rc;
        }
        if (b.done) CU(cudaEventRecord(b.done, st));
        b.f0 = f0; b.nb = nb; b.pending = true;
        h->lastLvl0 = b.d_img0; h->lastPitch0 = h->pitch0; h->lastFrameStride0 = (long long)h->pitch0 * hgt; h->lastFrames = nb;
    }
    if (nslots > 0) {
        for (int c = nchunks; c < nchunks + nslots; c++) {   // drain in issue order
            int rcc = orbCollect(h, h->pipe[c % nslots], want_desc, kps, desc, cap, n_out, mono_out, direct);
            if (rcc != EORB_OK) status = rcc;
        }
    } else {
        int rcc = orbCollect(h, h->main, want_desc, kps, desc, cap, n_out, mono_out, false);
        if (rcc != EORB_OK) status = rcc;
    }
    pendingGuard.armed = false;     // every slot was collected above
    return status;
}

extern "C" int eorb_orb_extract(eorb_orb* h, const uint8_t* img, int w, int hgt, size_t stride, int lap0, int lap1, int want_desc,
                                eorb_keypoint* kps, uint8_t* desc, int cap, int* n_out) {
    if (n_out) *n_out = 0;
    if (!h) return fail(EORB_ERR_ARG, "null handle");
    if (!img || w <= 0 || hgt <= 0) return EORB_EMPTY;   // _image.empty() -> -1 (ORBextractor.cc:1096)
    int n = 0, mono = 0;
    int rc = eorb_orb_extract_batch(h, img, 1, w, hgt, stride, stride * (size_t)hgt, lap0, lap1, want_desc, kps, desc, cap, &n, &mono);
    if (n_out) *n_out = n;
    if (rc != EORB_OK) return rc;
    return mono;
}

extern "C" int eorb_orb_level_size(const eorb_orb* h, int level, int* w, int* hgt) {
    if (!h || h->planW == 0) return fail(EORB_ERR_STATE, "no image processed yet");
    if (level < 0 || level >= h->nlevels) return fail(EORB_ERR_ARG, "level out of range");
    if (w) *w = h->hp.lv[level].w;
    if (hgt) *hgt = h->hp.lv[level].h;
    return EORB_OK;
}

extern "C" int eorb_orb_pyramid_level(eorb_orb* h, int frame, int level, uint8_t* dst, size_t dst_stride) {
    if (!h || h->planW == 0 || !h->lastLvl0) return fail(EORB_ERR_STATE, "no image processed yet");
    if (level < 0 || level >= h->nlevels || frame < 0 || frame >= h->lastFrames || !dst) return fail(EORB_ERR_ARG, "bad frame/level");
    CU(cudaSetDevice(h->device));
    const LevelPlan& lp = h->hp.lv[level];
    const uint8_t* src; size_t sp;
    if (level == 0) { src = h->lastLvl0 + (size_t)frame * h->lastFrameStride0; sp = (size_t)h->lastPitch0; }
    else { src = h->last->d_pyr + (size_t)frame * h->hp.pyrBytesPerFrame + lp.off; sp = lp.pitch; }
    CU(cudaMemcpy2DAsync(dst, dst_stride, src, sp, lp.w, lp.h, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return EORB_OK;
}

extern "C" int eorb_orb_debug_blurred(eorb_orb* h, int frame, int level, uint8_t* dst, size_t dst_stride) {
    if (!h || h->planW == 0) return fail(EORB_ERR_STATE, "no image processed yet");
    if (level < 0 || level >= h->nlevels || frame < 0 || frame >= h->lastFrames || !dst) return fail(EORB_ERR_ARG, "bad frame/level");
    CU(cudaSetDevice(h->device));
    const LevelPlan& lp = h->hp.lv[level];
    CU(cudaMemcpy2DAsync(dst, dst_stride, h->last->d_blur + (size_t)frame * h->hp.blurBytesPerFrame + lp.blurOff, lp.bpitch, lp.w, lp.h,
                         cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return EORB_OK;
}

extern "C" int eorb_orb_debug_candidates(eorb_orb* h, int frame, int level, int* xs, int* ys, int* scores, int cap) {
    if (!h || h->planW == 0) return fail(EORB_ERR_STATE, "no image processed yet");
    if (level < 0 || level >= h->nlevels || frame < 0 || frame >= h->lastFrames) return fail(EORB_ERR_ARG, "bad frame/level");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    int n = 0;
    CU(cudaMemcpy(&n, h->last->d_candCount + (size_t)frame * h->nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<uint32_t> k((size_t)std::max(n, 1));
    if (n > 0) CU(cudaMemcpy(k.data(), h->last->d_okeys + (size_t)frame * h->hp.slotsPerFrame + h->hp.lv[level].slotBase, (size_t)n * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n && i < cap; i++) { xs[i] = k[i] & 0xFFF; ys[i] = (k[i] >> 12) & 0xFFF; scores[i] = k[i] >> 24; }
    return n;
}

extern "C" int eorb_orb_debug_level_kps(eorb_orb* h, int frame, int level, int* xs, int* ys, int* scores, float* angles, int cap) {
    if (!h || h->planW == 0) return fail(EORB_ERR_STATE, "no image processed yet");
    if (level < 0 || level >= h->nlevels || frame < 0 || frame >= h->lastFrames) return fail(EORB_ERR_ARG, "bad frame/level");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    int n = 0;
    CU(cudaMemcpy(&n, h->last->d_selCount + (size_t)frame * h->nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    const LevelPlan& lp = h->hp.lv[level];
    std::vector<uint32_t> k((size_t)std::max(n, 1));
    std::vector<float> an((size_t)std::max(n, 1));
    if (n > 0) {
        CU(cudaMemcpy(k.data(), h->last->d_sel + (size_t)frame * h->hp.selPerFrame + lp.selBase, (size_t)n * 4, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(an.data(), h->last->d_levelAngle + (size_t)frame * h->hp.selPerFrame + lp.selBase, (size_t)n * 4, cudaMemcpyDeviceToHost));
    }
    for (int i = 0; i < n && i < cap; i++) {
        xs[i] = (int)(k[i] & 0xFFF) + lp.minBX; ys[i] = (int)((k[i] >> 12) & 0xFFF) + lp.minBY; scores[i] = k[i] >> 24;
        if (angles) angles[i] = an[i];
    }
    return n;
}

// ComputeTrackedKPtsDesc (:1316-1363) / AssignKPtLevelByBestDesc (:1267-1314): pyramid + blur of every level,
// then one warp per (keypoint, level)
static int orbTracked(eorb_orb* h, const uint8_t* img, int w, int hgt, size_t stride, eorb_keypoint* kps, int n, int mode,
                      const uint8_t* ref_desc, uint8_t* desc) {
    if (!h) return fail(EORB_ERR_ARG, "null handle");
    if (!img || w <= 0 || hgt <= 0) return EORB_EMPTY;   // trackedImage.empty() -> return
    if (n <= 0) return EORB_OK;
    if (!kps || (mode == 0 && !desc) || (mode == 1 && !ref_desc)) return fail(EORB_ERR_ARG, "null argument");
    CU(cudaSetDevice(h->device));
    int rc = orbBuildPlan(h, w, hgt);
    if (rc != EORB_OK) return rc;
    CU(cudaMemcpy2DAsync(h->main.d_img0, h->pitch0, img, stride, w, hgt, cudaMemcpyHostToDevice, h->stream));
    OrbArgs a = orbArgs(h, h->main, h->main.d_img0, h->pitch0, (long long)h->pitch0 * hgt, 0, 0, 1, h->main.d_outKps, h->main.d_outDesc, h->cap, h->main.d_outN, h->main.d_outMono);
    CUtensorMap pm[EORB_MAX_LEVELS];
    rc = orbPyrMaps(h, h->main, h->main.d_img0, w, hgt, 1, h->pitch0, (long long)h->pitch0 * hgt, pm);
    if (rc != EORB_OK) return rc;
    CU(launch_pyramid_and_blur(a, h->hp, h->stream, &h->launches, pm));
    eorb_keypoint* d_k = nullptr; uint8_t* d_ref = nullptr; uint8_t* d_desc = nullptr; int* d_dist = nullptr;
    CU(devAlloc(&d_k, (size_t)n));
    CU(cudaMemcpyAsync(d_k, kps, (size_t)n * sizeof(eorb_keypoint), cudaMemcpyHostToDevice, h->stream));
    if (mode == 0) { CU(devAlloc(&d_desc, (size_t)n * 32)); CU(cudaMemsetAsync(d_desc, 0, (size_t)n * 32, h->stream)); }
    else {
        CU(devAlloc(&d_ref, (size_t)n * 32)); CU(devAlloc(&d_dist, (size_t)n * h->nlevels));
        CU(cudaMemcpyAsync(d_ref, ref_desc, (size_t)n * 32, cudaMemcpyHostToDevice, h->stream));
    }
    CU(launch_tracked_desc(a, h->hp, d_k, n, mode, h->d_invScale, d_ref, d_desc, d_dist, h->stream, &h->launches));
    if (mode == 0) {
        CU(cudaMemcpyAsync(desc, d_desc, (size_t)n * 32, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    } else {
        std::vector<int> dist((size_t)n * h->nlevels);
        CU(cudaMemcpyAsync(dist.data(), d_dist, dist.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for (int i = 0; i < n; i++) {
            int best = INT32_MAX;
            for (int l = 0; l < h->nlevels; l++) {
                const int d = dist[(size_t)l * n + i];
                if (d < best) { best = d; kps[i].octave = l; }   // strict '<': lowest level wins ties (:1308-1311)
            }
        }
    }
    cudaFree(d_k); cudaFree(d_ref); cudaFree(d_desc); cudaFree(d_dist);
    h->lastLvl0 = h->main.d_img0; h->lastPitch0 = h->pitch0; h->lastFrameStride0 = (long long)h->pitch0 * hgt; h->lastFrames = 1;
    return EORB_OK;
}

extern "C" int eorb_orb_tracked_desc(eorb_orb* h, const uint8_t* img, int w, int hgt, size_t stride, const eorb_keypoint* kps,
                                     int n, uint8_t* desc) {
    return orbTracked(h, img, w, hgt, stride, const_cast<eorb_keypoint*>(kps), n, 0, nullptr, desc);
}
extern "C" int eorb_orb_assign_level_by_best_desc(eorb_orb* h, const uint8_t* ref_desc, const uint8_t* img, int w, int hgt,
                                                  size_t stride, eorb_keypoint* kps, int n) {
    return orbTracked(h, img, w, hgt, stride, kps, n, 1, ref_desc, nullptr);
}

// ================================================================================================ matcher
struct eorb_matcher {
    int device = 0, sms = 148;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    const uint8_t* d_db = nullptr; uint8_t* ownedDb = nullptr;
    long long ndb = 0, indexOffset = 0, chunkRows = 0;
    int nchunks = 0;
    eorb_best2* d_partial = nullptr; size_t partialCap = 0;
    uint8_t* d_q = nullptr; size_t qCap = 0;
    eorb_match* d_out = nullptr; size_t outCap = 0;
    eorb_best2* d_mine = nullptr; size_t mineCap = 0;           // sharded search: this shard's best-2 per query (nq entries)
    eorb_best2* d_gathered = nullptr; size_t gatherCap = 0;     //                 every shard's best-2 (nshards * nq entries)
    int engine = EORB_HAMMING_AUTO;                               // eorb_matcher_set_engine
    int lastEngine = EORB_HAMMING_POPC;                           // what the last search ran on
    bool tensorOk = false;                                        // compute capability 10.x (tcgen05)
    long long launches = 0;
};

extern "C" int eorb_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    if (!a || !b) return fail(EORB_ERR_ARG, "null descriptor");
    uint32_t x[8], y[8];
    memcpy(x, a, 32); memcpy(y, b, 32);
    int d = 0;
    for (int i = 0; i < 8; i++) d += __builtin_popcount(x[i] ^ y[i]);
    return d;
}

extern "C" int eorb_matcher_create(int device, eorb_matcher** out) {
    if (!out) return fail(EORB_ERR_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(EORB_ERR_CUDA, "no CUDA device: eorb_b200 has no CPU fallback"); }
    if (device < 0 || device >= ndev) return fail(EORB_ERR_ARG, "device %d out of range", device);
    CU(cudaSetDevice(device));
    eorb_matcher* m = new eorb_matcher();
    m->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    m->sms = prop.multiProcessorCount;
    m->tensorOk = prop.major == 10;
    if (const char* e = getenv("EORB_HAMMING_ENGINE")) m->engine = std::min(std::max(atoi(e), 0), 2);
    CU(cudaStreamCreateWithFlags(&m->ownStream, cudaStreamNonBlocking));
    m->stream = m->ownStream;
    *out = m;
    return EORB_OK;
}
extern "C" int eorb_matcher_set_engine(eorb_matcher* m, int engine) {
    if (!m || engine < EORB_HAMMING_POPC || engine > EORB_HAMMING_AUTO) return fail(EORB_ERR_ARG, "engine must be 0 (POPC), 1 (tensor) or 2 (auto)");
    if (engine == EORB_HAMMING_TENSOR && !m->tensorOk) return fail(EORB_ERR_CUDA, "the tensor engine needs tcgen05 (compute capability 10.x)");
    m->engine = engine;
    return EORB_OK;
}
extern "C" int eorb_matcher_last_engine(const eorb_matcher* m) { return m ? m->lastEngine : -1; }
extern "C" int eorb_matcher_destroy(eorb_matcher* m) {
    if (!m) return EORB_OK;
    cudaSetDevice(m->device);
    cudaStreamSynchronize(m->stream);
    cudaFree(m->ownedDb); cudaFree(m->d_partial); cudaFree(m->d_q); cudaFree(m->d_out); cudaFree(m->d_mine); cudaFree(m->d_gathered);
    cudaStreamDestroy(m->ownStream);
    delete m;
    return EORB_OK;
}
extern "C" int eorb_matcher_set_stream(eorb_matcher* m, void* s) {
    if (!m) return fail(EORB_ERR_ARG, "null handle");
    CU(cudaStreamSynchronize(m->stream));
    m->stream = (cudaStream_t)s;   // NULL is the CUDA legacy default stream, a legitimate choice
    return EORB_OK;
}
extern "C" int eorb_matcher_reset_stream(eorb_matcher* m) {
    if (!m) return fail(EORB_ERR_ARG, "null handle");
    CU(cudaStreamSynchronize(m->stream));
    m->stream = m->ownStream;
    return EORB_OK;
}
extern "C" int eorb_matcher_synchronize(eorb_matcher* m) {
    if (!m) return fail(EORB_ERR_ARG, "null handle");
    CU(cudaSetDevice(m->device));
    CU(cudaStreamSynchronize(m->stream));
    return EORB_OK;
}
extern "C" long long eorb_matcher_launch_count(const eorb_matcher* m) { return m ? m->launches : 0; }

static int matcherAdopt(eorb_matcher* m, const uint8_t* d_db, long long ndb, long long indexOffset) {
    if (ndb < 0 || indexOffset < 0 || indexOffset + ndb > 0x7fffffffLL) return fail(EORB_ERR_ARG, "database rows must fit int32 indices");
    if ((uintptr_t)d_db % 16 != 0) return fail(EORB_ERR_ARG, "database pointer must be 16-byte aligned");
    m->d_db = d_db; m->ndb = ndb; m->indexOffset = indexOffset;
    m->nchunks = ndb > 0 ? hamming_chunks(ndb, m->sms, &m->chunkRows) : 0;
    return EORB_OK;
}
extern "C" int eorb_matcher_set_db_device(eorb_matcher* m, const uint8_t* d_db, int64_t ndb, int64_t index_offset) {
    if (!m || (!d_db && ndb > 0)) return fail(EORB_ERR_ARG, "null argument");
    CU(cudaSetDevice(m->device));
    CU(cudaStreamSynchronize(m->stream));
    cudaFree(m->ownedDb); m->ownedDb = nullptr;
    return matcherAdopt(m, d_db, ndb, index_offset);
}
extern "C" int eorb_matcher_set_db_host(eorb_matcher* m, const uint8_t* db, int64_t ndb, int64_t index_offset) {
    if (!m || (!db && ndb > 0)) return fail(EORB_ERR_ARG, "null argument");
    CU(cudaSetDevice(m->device));
    CU(cudaStreamSynchronize(m->stream));
    cudaFree(m->ownedDb); m->ownedDb = nullptr;
    CU(devAlloc(&m->ownedDb, (size_t)std::max<int64_t>(ndb, 1) * 32));
    if (ndb > 0) CU(cudaMemcpy(m->ownedDb, db, (size_t)ndb * 32, cudaMemcpyHostToDevice));
    return matcherAdopt(m, m->ownedDb, ndb, index_offset);
}

// scan of the whole database for nq queries: per-chunk partial best-2 arrays in m->d_partial, *nparts of them.
// Engine: the tensor cores (hamming_tc.cu) pay off when the accumulator's 128 query rows are mostly real queries and the database is
// large enough to amortise the per-CTA setup; small searches stay on the POPC kernel.
static int matcherScan(eorb_matcher* m, const uint8_t* d_q, int nq, int* nparts) {
    *nparts = 0;
    if (m->ndb <= 0) return growBuf(&m->d_partial, &m->partialCap, (size_t)nq);
    const bool tensor = m->tensorOk && (m->engine == EORB_HAMMING_TENSOR || (m->engine == EORB_HAMMING_AUTO && nq >= 96 && m->ndb >= 65536));
    if (tensor) {
        long long rows = 0;
        const int nch = hamming_tc_chunks(m->ndb, nq, m->sms, &rows);
        const int parts = nch * hamming_tc_parts_per_chunk();
        int rc = growBuf(&m->d_partial, &m->partialCap, (size_t)parts * nq);
        if (rc != EORB_OK) return rc;
        CU(launch_hamming_best2_tc(d_q, nq, m->d_db, m->ndb, m->indexOffset, rows, nch, m->d_partial, m->stream));
        m->launches++;
        m->lastEngine = EORB_HAMMING_TENSOR;
        *nparts = parts;
        return EORB_OK;
    }
    int rc = growBuf(&m->d_partial, &m->partialCap, (size_t)std::max(m->nchunks, 1) * nq);
    if (rc != EORB_OK) return rc;
    CU(launch_hamming_best2(d_q, nq, m->d_db, m->ndb, m->indexOffset, m->chunkRows, m->nchunks, m->d_partial, m->stream));
    m->launches++;
    m->lastEngine = EORB_HAMMING_POPC;
    *nparts = m->nchunks;
    return EORB_OK;
}

extern "C" int eorb_matcher_search_device(eorb_matcher* m, const uint8_t* d_q, int nq, eorb_best2* d_partial) {
    if (!m || !d_partial) return fail(EORB_ERR_ARG, "null argument");
    if (nq <= 0) return EORB_OK;
    if ((uintptr_t)d_q % 16 != 0) return fail(EORB_ERR_ARG, "query pointer must be 16-byte aligned");
    CU(cudaSetDevice(m->device));
    int nparts = 0;
    int rc = matcherScan(m, d_q, nq, &nparts);
    if (rc != EORB_OK) return rc;
    CU(launch_merge_best2(m->d_partial, nparts, nq, d_partial, nullptr, 0, 0.f, m->stream));
    m->launches++;
    return EORB_OK;
}

extern "C" int eorb_matcher_merge_device(eorb_matcher* m, const eorb_best2* d_gathered, int nshards, int nq, int th, float ratio,
                                         eorb_match* d_out) {
    if (!m || !d_gathered || !d_out || nshards < 1) return fail(EORB_ERR_ARG, "null argument");
    if (nq <= 0) return EORB_OK;
    CU(cudaSetDevice(m->device));
    CU(launch_merge_best2(d_gathered, nshards, nq, nullptr, d_out, th, ratio, m->stream));
    m->launches++;
    return EORB_OK;
}

// ---- NCCL, resolved at run time (the process may already hold torch's bundled libnccl.so.2; a C++ host loads the
//      system one).  Only the four entry points the sharded search needs.
struct EorbNcclId { char internal[128]; };
struct EorbNccl {
    void* lib = nullptr;
    int (*getUniqueId)(EorbNcclId*) = nullptr;
    int (*commInitRank)(void**, int, EorbNcclId, int) = nullptr;
    int (*commDestroy)(void*) = nullptr;
    int (*allGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*getErrorString)(int) = nullptr;
};
static EorbNccl g_nccl;
static std::once_flag g_ncclOnce;
static int ncclApi() {
    std::call_once(g_ncclOnce, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.lib) break;
        }
        if (!g_nccl.lib) return;
        g_nccl.getUniqueId = (int (*)(EorbNcclId*))dlsym(g_nccl.lib, "ncclGetUniqueId");
        g_nccl.commInitRank = (int (*)(void**, int, EorbNcclId, int))dlsym(g_nccl.lib, "ncclCommInitRank");
        g_nccl.commDestroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
        g_nccl.allGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllGather");
        g_nccl.getErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
    });
    if (!g_nccl.lib || !g_nccl.getUniqueId || !g_nccl.commInitRank || !g_nccl.commDestroy || !g_nccl.allGather)
        return fail(EORB_ERR_STATE, "libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "missing symbols");
    return EORB_OK;
}
#define NCCL_CALL(call)                                                                                              \
    do {                                                                                                             \
        int r__ = (call);                                                                                            \
        if (r__ != 0) return fail(EORB_ERR_CUDA, "%s failed: %s", #call, g_nccl.getErrorString ? g_nccl.getErrorString(r__) : "nccl error"); \
    } while (0)

extern "C" int eorb_nccl_unique_id(uint8_t* id128) {
    if (!id128) return fail(EORB_ERR_ARG, "null argument");
    int rc = ncclApi();
    if (rc != EORB_OK) return rc;
    EorbNcclId id;
    NCCL_CALL(g_nccl.getUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return EORB_OK;
}

extern "C" int eorb_nccl_comm_init_rank(void** comm, int nranks, const uint8_t* id128, int rank, int device) {
    if (!comm || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(EORB_ERR_ARG, "bad communicator arguments");
    int rc = ncclApi();
    if (rc != EORB_OK) return rc;
    CU(cudaSetDevice(device));
    EorbNcclId id;
    memcpy(id.internal, id128, 128);
    NCCL_CALL(g_nccl.commInitRank(comm, nranks, id, rank));
    return EORB_OK;
}

extern "C" int eorb_nccl_comm_destroy(void* comm) {
    if (!comm) return EORB_OK;
    int rc = ncclApi();
    if (rc != EORB_OK) return rc;
    NCCL_CALL(g_nccl.commDestroy(comm));
    return EORB_OK;
}

// database sharded by rows over the ranks; each rank scans its shard, ONE ncclAllGather moves nq x 16 B per
// rank over NVLink, every rank merges with the (dist, global index) ordering and applies threshold + ratio.
extern "C" int eorb_matcher_search_sharded(eorb_matcher* m, const uint8_t* d_q, int nq, int th, float ratio, void* nccl_comm,
                                           int nshards, eorb_match* d_out) {
    if (!m || !d_out || !nccl_comm || nshards < 1) return fail(EORB_ERR_ARG, "null argument");
    if (nq <= 0) return EORB_OK;
    int rc = ncclApi();
    if (rc != EORB_OK) return rc;
    CU(cudaSetDevice(m->device));
    // the two buffers grow independently: 8 shards x 100 queries followed by 2 shards x 400 needs a larger d_mine, not a larger gather
    if (((size_t)nq > m->mineCap) || ((size_t)nshards * nq > m->gatherCap)) CU(cudaStreamSynchronize(m->stream));
    rc = growBuf(&m->d_mine, &m->mineCap, (size_t)nq);
    if (rc != EORB_OK) return rc;
    rc = growBuf(&m->d_gathered, &m->gatherCap, (size_t)nshards * nq);
    if (rc != EORB_OK) return rc;
    rc = eorb_matcher_search_device(m, d_q, nq, m->d_mine);
    if (rc != EORB_OK) return rc;
    NCCL_CALL(g_nccl.allGather(m->d_mine, m->d_gathered, (size_t)nq * sizeof(eorb_best2), /*ncclChar*/ 0, nccl_comm, m->stream));
    return eorb_matcher_merge_device(m, m->d_gathered, nshards, nq, th, ratio, d_out);
}

extern "C" int eorb_matcher_search(eorb_matcher* m, const uint8_t* q, int nq, int th, float ratio, eorb_match* out) {
    if (!m || !out || (!q && nq > 0)) return fail(EORB_ERR_ARG, "null argument");
    if (nq <= 0) return EORB_OK;
    CU(cudaSetDevice(m->device));
    int rcg = growBuf(&m->d_q, &m->qCap, (size_t)nq * 32);
    if (rcg != EORB_OK) return rcg;
    rcg = growBuf(&m->d_out, &m->outCap, (size_t)nq);
    if (rcg != EORB_OK) return rcg;
    CU(cudaMemcpyAsync(m->d_q, q, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream));
    int nparts = 0;
    int rc = matcherScan(m, m->d_q, nq, &nparts);
    if (rc != EORB_OK) return rc;
    CU(launch_merge_best2(m->d_partial, nparts, nq, nullptr, m->d_out, th, ratio, m->stream));
    m->launches++;
    CU(cudaMemcpyAsync(out, m->d_out, (size_t)nq * sizeof(eorb_match), cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    return EORB_OK;
}

extern "C" int eorb_hamming_best2(const uint8_t* q, int nq, const uint8_t* db, int64_t ndb, int th, float ratio, eorb_match* out) {
    eorb_matcher* m = nullptr;
    int rc = eorb_matcher_create(0, &m);
    if (rc != EORB_OK) return rc;
    rc = eorb_matcher_set_db_host(m, db, ndb, 0);
    if (rc == EORB_OK) rc = eorb_matcher_search(m, q, nq, th, ratio, out);
    eorb_matcher_destroy(m);
    return rc;
}

// ORBmatcher.cc:784-794 (histogram fill), :800-823 (filter), ComputeThreeMaxima :2314-2355.  Host code like the
// reference: O(matches), sequential list semantics.  factor = 1/HISTO_LENGTH is the reference's (quirky) bin width.
extern "C" int eorb_rotation_filter(const float* angle1, const float* angle2, int32_t* match12, int n1) {
    if ((!angle1 || !angle2 || !match12) && n1 > 0) return fail(EORB_ERR_ARG, "null argument");
    const int L = 30;
    const float factor = 1.0f / L;
    std::vector<std::vector<int>> hist(L);
    int nmatches = 0;
    for (int i = 0; i < n1; i++) {
        const int j = match12[i];
        if (j < 0) continue;
        nmatches++;
        float rot = angle1[i] - angle2[j];
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)std::round(rot * factor);
        if (bin == L) bin = 0;
        if (bin >= 0 && bin < L) hist[bin].push_back(i);
    }
    int top[3] = {-1, -1, -1}, cnt[3] = {0, 0, 0};
    for (int b = 0; b < L; b++) {
        const int s = (int)hist[b].size();
        if (s > cnt[0]) { cnt[2] = cnt[1]; top[2] = top[1]; cnt[1] = cnt[0]; top[1] = top[0]; cnt[0] = s; top[0] = b; }
        else if (s > cnt[1]) { cnt[2] = cnt[1]; top[2] = top[1]; cnt[1] = s; top[1] = b; }
        else if (s > cnt[2]) { cnt[2] = s; top[2] = b; }
    }
    if ((float)cnt[1] < 0.1f * (float)cnt[0]) { top[1] = -1; top[2] = -1; }
    else if ((float)cnt[2] < 0.1f * (float)cnt[0]) { top[2] = -1; }
    for (int b = 0; b < L; b++) {
        if (b == top[0] || b == top[1] || b == top[2]) continue;
        for (int i : hist[b])
            if (match12[i] >= 0) { match12[i] = -1; nmatches--; }
    }
    return nmatches;
}

// ================================================================================================ events
struct eorb_evconv {
    int device = 0;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    int maxWindows = 0, maxW = 0, maxH = 0;
    long long maxEvents = 0;
    eorb_event* d_evs = nullptr; float* d_img = nullptr; uint8_t* d_u8 = nullptr; float* d_minmax = nullptr;
    float2* d_xy = nullptr; long long xyCap = 0;   // warped event positions (multi-band motion-compensated frames), grown on demand
    float* d_jac = nullptr;   // 7 frames (I, dI/d[wx wy wz vx vy vz]) of ev2mci_gg_f_jac, allocated on first use
    float* d_bimg = nullptr; uint8_t* d_bu8 = nullptr; size_t bCap = 0;   // [nwin][h][w] frames of the host batch call, allocated on first use
    cudaStream_t pipeS[2] = {nullptr, nullptr};                            // the host batch call's two pipeline streams
    EvWindow* d_wins = nullptr;
    std::vector<EvWindow> h_wins;
    long long launches = 0;
};

extern "C" int eorb_ev_create(int device, int max_windows, int64_t max_events, int max_width, int max_height, eorb_evconv** out) {
    if (!out || max_windows < 1 || max_events < 1 || max_width < 1 || max_height < 1) return fail(EORB_ERR_ARG, "bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(EORB_ERR_CUDA, "no CUDA device: eorb_b200 has no CPU fallback"); }
    if (device < 0 || device >= ndev) return fail(EORB_ERR_ARG, "device %d out of range", device);
    CU(cudaSetDevice(device));
    eorb_evconv* c = new eorb_evconv();
    c->device = device; c->maxWindows = max_windows; c->maxEvents = max_events; c->maxW = max_width; c->maxH = max_height;
    CU(cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking));
    c->stream = c->ownStream;
    CU(devAlloc(&c->d_evs, (size_t)max_events));
    CU(devAlloc(&c->d_img, (size_t)max_width * max_height));      // single-window host path
    CU(devAlloc(&c->d_u8, (size_t)max_width * max_height));
    CU(devAlloc(&c->d_minmax, (size_t)std::max(max_windows * 2, 8)));   // also the 6-value scratch of the Jacobian / focus calls
    CU(devAlloc(&c->d_wins, (size_t)max_windows));
    *out = c;
    return EORB_OK;
}
extern "C" int eorb_ev_destroy(eorb_evconv* c) {
    if (!c) return EORB_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->d_evs); cudaFree(c->d_img); cudaFree(c->d_u8); cudaFree(c->d_minmax); cudaFree(c->d_bimg); cudaFree(c->d_bu8); cudaFree(c->d_jac); cudaFree(c->d_wins); cudaFree(c->d_xy);
    for (int i = 0; i < 2; i++) if (c->pipeS[i]) cudaStreamDestroy(c->pipeS[i]);
    cudaStreamDestroy(c->ownStream);
    delete c;
    return EORB_OK;
}
extern "C" int eorb_ev_set_stream(eorb_evconv* c, void* s) {
    if (!c) return fail(EORB_ERR_ARG, "null handle");
    CU(cudaStreamSynchronize(c->stream));
    c->stream = (cudaStream_t)s;   // NULL is the CUDA legacy default stream, a legitimate choice
    return EORB_OK;
}
extern "C" int eorb_ev_reset_stream(eorb_evconv* c) {
    if (!c) return fail(EORB_ERR_ARG, "null handle");
    CU(cudaStreamSynchronize(c->stream));
    c->stream = c->ownStream;
    return EORB_OK;
}
extern "C" int eorb_ev_synchronize(eorb_evconv* c) {
    if (!c) return fail(EORB_ERR_ARG, "null handle");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return EORB_OK;
}
extern "C" long long eorb_ev_launch_count(const eorb_evconv* c) { return c ? c->launches : 0; }

// Eigen::AngleAxisd(const Matrix3d&) (matrix -> quaternion -> angle/axis), used at EventConversion.cc:303
static void angleAxisFromR(const double R[3][3], EvWindow& w);
static void angleAxisFromPose(const float* T, EvWindow& w) {
    double R[3][3];
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) R[r][c] = (double)T[4 * r + c]; w.t[r] = (double)T[4 * r + 3]; }
    angleAxisFromR(R, w);
}
// Eigen::AngleAxisd(const Matrix3d&): through the quaternion, as Eigen does
static void angleAxisFromR(const double R[3][3], EvWindow& w) {
    double q[4];
    double tr = R[0][0] + R[1][1] + R[2][2];
    if (tr > 0) {
        double s = std::sqrt(tr + 1.0);
        q[3] = 0.5 * s; s = 0.5 / s;
        q[0] = (R[2][1] - R[1][2]) * s; q[1] = (R[0][2] - R[2][0]) * s; q[2] = (R[1][0] - R[0][1]) * s;
    } else {
        int i = 0;
        if (R[1][1] > R[0][0]) i = 1;
        if (R[2][2] > R[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        double s = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0);
        q[i] = 0.5 * s; s = 0.5 / s;
        q[3] = (R[k][j] - R[j][k]) * s; q[j] = (R[j][i] + R[i][j]) * s; q[k] = (R[k][i] + R[i][k]) * s;
    }
    double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    if (n != 0.0) {
        w.angle = 2.0 * std::atan2(n, std::fabs(q[3]));
        if (q[3] < 0) n = -n;
        w.axis[0] = q[0] / n; w.axis[1] = q[1] / n; w.axis[2] = q[2] / n;
    } else {
        w.angle = 0; w.axis[0] = 1; w.axis[1] = 0; w.axis[2] = 0;
    }
}

static int evConst(const eorb_ev_params* p, EvConst& c) {
    if (!p) return fail(EORB_ERR_ARG, "null params");
    if (p->mode < EORB_EV_NEAREST || p->mode > EORB_EV_SE2) return fail(EORB_ERR_ARG, "bad mode %d", p->mode);
    if (p->width < 1 || p->height < 1) return fail(EORB_ERR_ARG, "bad image size");
    if (p->mode != EORB_EV_NEAREST && !(p->sigma > 0.f)) return fail(EORB_ERR_ARG, "sigma must be > 0");
    c.mode = p->mode; c.width = p->width; c.height = p->height; c.pol = p->pol;
    c.sigma = p->sigma; c.sig2 = p->sigma * p->sigma;
    c.norm = 2.0f * (float)M_PI * c.sig2;
    c.half = (int)std::ceil(p->sigma * 3.0);
    c.depth = p->med_depth;
    c.fx = p->K[0]; c.fy = p->K[1]; c.cx = p->K[2]; c.cy = p->K[3];
    for (int i = 0; i < 4; i++) c.se2[i] = p->se2[i];
    c.se2_n = p->se2_n;
    if (p->cam_model < 0 || p->cam_model > 1) return fail(EORB_ERR_ARG, "bad camera model %d (0 Pinhole, 1 KannalaBrandt8)", p->cam_model);
    c.cam = p->cam_model;
    for (int i = 0; i < 4; i++) c.kb[i] = p->kb[i];
    return EORB_OK;
}

// windows [i0, i0 + n) of win_offsets on stream `st`; their EvWindow records and extremes live in slots [i0, i0 + n) of d_wins / d_minmax,
// so that disjoint window ranges can be in flight on different streams (the pipelined host call below)
static int evBatchRun(eorb_evconv* c, const eorb_event* d_evs, const int64_t* win_offsets, int i0, int n, long long totalEvents, const eorb_ev_params* p,
                      const EvConst& k, const float* poses, float* d_img_f32, uint8_t* d_img_u8, cudaStream_t st) {
    long long maxEv = 0;
    for (int i = i0; i < i0 + n; i++) {
        EvWindow& w = c->h_wins[i];
        w.begin = win_offsets[i]; w.end = win_offsets[i + 1];
        if (w.end < w.begin) return fail(EORB_ERR_ARG, "window offsets must be non-decreasing");
        maxEv = std::max(maxEv, w.end - w.begin);
        w.angle = 0; w.axis[0] = 1; w.axis[1] = w.axis[2] = 0; w.t[0] = w.t[1] = w.t[2] = 0;
        if (p->mode == EORB_EV_SE3) angleAxisFromPose(poses ? poses + 16 * (size_t)i : p->Tcw, w);
    }
    CU(cudaMemcpyAsync(c->d_wins + i0, c->h_wins.data() + i0, (size_t)n * sizeof(EvWindow), cudaMemcpyHostToDevice, st));
    float2* xy = nullptr;
    const bool ordered = p->pol && p->normalize == EORB_NORM_RUNNING;   // order-dependent extremes: ev_ordered_kernel reads warped positions
    if ((p->mode == EORB_EV_SE3 || p->mode == EORB_EV_SE2) && ((size_t)k.width * k.height * 4 > 200 * 1024 || ordered)) {
        if (totalEvents > c->xyCap) {
            CU(cudaDeviceSynchronize());
            cudaFree(c->d_xy); c->d_xy = nullptr; c->xyCap = 0;
            CU(devAlloc(&c->d_xy, (size_t)totalEvents));
            c->xyCap = totalEvents;
        }
        xy = c->d_xy;
    }
    CU(launch_ev_frames(d_evs, c->d_wins + i0, n, maxEv, k, p->normalize, d_img_f32, c->d_minmax + 2 * (size_t)i0, d_img_u8, st, &c->launches, xy));
    return EORB_OK;
}

extern "C" int eorb_ev_accumulate_batch_device(eorb_evconv* c, const eorb_event* d_evs, const int64_t* win_offsets, int nwin,
                                               const eorb_ev_params* p, const float* poses, float* d_img_f32, uint8_t* d_img_u8) {
    if (!c || !d_evs || !win_offsets || !d_img_f32) return fail(EORB_ERR_ARG, "null argument");
    if (nwin < 1) return EORB_OK;
    if (nwin > c->maxWindows) return fail(EORB_ERR_CAPACITY, "nwin %d > max_windows %d", nwin, c->maxWindows);
    EvConst k;
    int rc = evConst(p, k);
    if (rc != EORB_OK) return rc;
    if (p->normalize != EORB_NORM_NONE && !d_img_u8) return fail(EORB_ERR_ARG, "normalize requested without a u8 output");
    CU(cudaSetDevice(c->device));
    c->h_wins.resize(nwin);
    return evBatchRun(c, d_evs, win_offsets, 0, nwin, win_offsets[nwin], p, k, poses, d_img_f32, d_img_u8, c->stream);
}

// like CU(), but records the failure in `status` instead of returning: the caller drains its streams first
#define CUS(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) status = fail(EORB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
// nwin windows, HOST buffers: the events go in, the batch kernels run, the requested outputs come back -- pipelined over window ranges
extern "C" int eorb_ev_accumulate_batch(eorb_evconv* c, const eorb_event* evs, const int64_t* win_offsets, int nwin, const eorb_ev_params* p,
                                        const float* poses, float* img_f32, uint8_t* img_u8) {
    if (!c || !win_offsets || !p || (!img_f32 && !img_u8)) return fail(EORB_ERR_ARG, "null argument");
    if (nwin < 1) return EORB_OK;
    if (nwin > c->maxWindows) return fail(EORB_ERR_CAPACITY, "nwin %d > max_windows %d", nwin, c->maxWindows);
    const long long nev = win_offsets[nwin];
    if (win_offsets[0] != 0 || nev < 0) return fail(EORB_ERR_ARG, "window offsets must start at 0");
    if (nev > c->maxEvents) return fail(EORB_ERR_CAPACITY, "%lld events > max_events %lld", nev, c->maxEvents);
    if (nev > 0 && !evs) return fail(EORB_ERR_ARG, "null events");
    if (p->width < 1 || p->height < 1 || (long long)p->width * p->height > (long long)c->maxW * c->maxH)
        return fail(EORB_ERR_CAPACITY, "image %dx%d exceeds the converter's %dx%d", p->width, p->height, c->maxW, c->maxH);
    if (img_u8 && p->normalize == EORB_NORM_NONE) return fail(EORB_ERR_ARG, "a u8 output needs a normalisation mode");
    CU(cudaSetDevice(c->device));
    const size_t npix = (size_t)p->width * p->height, need = npix * (size_t)nwin;
    if (need > c->bCap) {
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_bimg); cudaFree(c->d_bu8); c->d_bimg = nullptr; c->d_bu8 = nullptr; c->bCap = 0;
        CU(devAlloc(&c->d_bimg, need));
        CU(devAlloc(&c->d_bu8, need));
        c->bCap = need;
    }
    EvConst k;
    int rc = evConst(p, k);
    if (rc != EORB_OK) return rc;
    c->h_wins.resize(nwin);
    uint8_t* d_u8 = p->normalize != EORB_NORM_NONE ? c->d_bu8 : nullptr;
    // Large packets run as up to four ranges of windows on two streams, so that the events of one range go in while the frames of the
    // range before come out (PCIe is full duplex); the windows' records and extremes sit in disjoint slots.  (The per-event scratch of the
    // multi-band / ordered paths is indexed by event position, so disjoint ranges do not collide there either.)
    const int nparts = (nwin >= 16 && nev >= 65536) ? 4 : 1;
    if (nparts > 1 && !c->pipeS[0]) {
        for (int i = 0; i < 2; i++) CU(cudaStreamCreateWithFlags(&c->pipeS[i], cudaStreamNonBlocking));
    }
    if (nparts > 1) CU(cudaStreamSynchronize(c->stream));      // order after earlier work on the converter's stream
    int status = EORB_OK;
    for (int part = 0; part < nparts && status == EORB_OK; part++) {
        const int i0 = (int)((long long)nwin * part / nparts), i1 = (int)((long long)nwin * (part + 1) / nparts);
        if (i1 <= i0) continue;
        cudaStream_t st = nparts > 1 ? c->pipeS[part & 1] : c->stream;
        const long long e0 = win_offsets[i0], e1 = win_offsets[i1];
        if (e1 < e0) { status = fail(EORB_ERR_ARG, "window offsets must be non-decreasing"); break; }
        if (e1 > e0) CUS(cudaMemcpyAsync(c->d_evs + e0, evs + e0, (size_t)(e1 - e0) * sizeof(eorb_event), cudaMemcpyHostToDevice, st));
        if (status != EORB_OK) break;
        status = evBatchRun(c, c->d_evs, win_offsets, i0, i1 - i0, nev, p, k, poses, c->d_bimg + npix * (size_t)i0, d_u8 ? d_u8 + npix * (size_t)i0 : nullptr, st);
        if (status != EORB_OK) break;
        const size_t cnt = npix * (size_t)(i1 - i0);
        if (img_f32) CUS(cudaMemcpyAsync(img_f32 + npix * (size_t)i0, c->d_bimg + npix * (size_t)i0, cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (img_u8 && status == EORB_OK) CUS(cudaMemcpyAsync(img_u8 + npix * (size_t)i0, c->d_bu8 + npix * (size_t)i0, cnt, cudaMemcpyDeviceToHost, st));
    }
    // every stream is drained before returning, also on an error path: no copy into the caller's arrays stays in flight
    cudaError_t es = cudaStreamSynchronize(c->stream);
    if (nparts > 1) for (int i = 0; i < 2; i++) { const cudaError_t e2 = cudaStreamSynchronize(c->pipeS[i]); if (es == cudaSuccess) es = e2; }
    if (status != EORB_OK) return status;
    CU(es);
    return EORB_OK;
}

// EvImConverter::ev2mci_gg_f_jac (EventConversion.cc:533-662; one call per optimiser iteration, MyOptimTypes.cpp:16)
extern "C" int eorb_ev_mci_jac(eorb_evconv* c, const eorb_event* evs, int64_t n, int w, int hgt, float sigma, const double* Rt12, float med_depth,
                               const float* K4, int pol, int global_mean, double* jac6) {
    if (!c || !jac6 || !Rt12 || !K4) return fail(EORB_ERR_ARG, "null argument");
    for (int k = 0; k < 6; k++) jac6[k] = 0.0;
    if (n <= 0 || !evs) return EORB_EMPTY;   // "no events": the reference returns a zero Jacobian (:543-546)
    if (w < 1 || hgt < 1 || (long long)w * hgt > (long long)c->maxW * c->maxH) return fail(EORB_ERR_CAPACITY, "image %dx%d exceeds the converter's %dx%d", w, hgt, c->maxW, c->maxH);
    if (n > c->maxEvents) return fail(EORB_ERR_CAPACITY, "%lld events > max_events %lld", (long long)n, c->maxEvents);
    if (!(sigma > 0.f)) return fail(EORB_ERR_ARG, "sigma must be > 0");
    const int patch = 30;
    if (((w + patch - 1) / patch) * ((hgt + patch - 1) / patch) > EORB_EV_FOCUS_MAX_CELLS) return fail(EORB_ERR_CAPACITY, "image too large for the local mean");
    CU(cudaSetDevice(c->device));
    if (!c->d_jac) CU(devAlloc(&c->d_jac, (size_t)7 * c->maxW * c->maxH));
    EvConst k{};
    k.mode = EORB_EV_SE3; k.width = w; k.height = hgt; k.pol = pol; k.sigma = sigma; k.sig2 = sigma * sigma;
    k.norm = 2.0f * (float)M_PI * k.sig2; k.half = (int)std::ceil(sigma * 3.0); k.depth = med_depth;
    k.fx = K4[0]; k.fy = K4[1]; k.cx = K4[2]; k.cy = K4[3];
    EvWindow win{};
    win.begin = 0; win.end = n;
    double R[3][3];
    for (int r = 0; r < 3; r++) { for (int cc = 0; cc < 3; cc++) R[r][cc] = Rt12[3 * r + cc]; win.t[r] = Rt12[9 + r]; }
    angleAxisFromR(R, win);
    const size_t npx = (size_t)w * hgt;
    CU(cudaMemcpyAsync(c->d_evs, evs, (size_t)n * sizeof(eorb_event), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_wins, &win, sizeof(EvWindow), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));   // `win` is a stack object
    CU(cudaMemsetAsync(c->d_jac, 0, 7 * npx * sizeof(float), c->stream));
    c->launches++;
    CU(launch_ev_jac(c->d_evs, c->d_wins, n, k, global_mean, c->d_jac, c->d_minmax, c->stream, &c->launches));
    float m[6];
    CU(cudaMemcpyAsync(m, c->d_minmax, sizeof(m), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 6; i++) jac6[i] = (double)(-m[i] * 2);
    return EORB_OK;
}

// measureImageFocusLocal / measureImageFocusGlobal / imageMeanLocal (EventConversion.cc:79-162) of nwin device frames
extern "C" int eorb_ev_image_focus_device(eorb_evconv* c, const float* d_img_f32, int nwin, int w, int hgt, int what, int avg, float* focus_out) {
    if (!c || !d_img_f32 || !focus_out) return fail(EORB_ERR_ARG, "null argument");
    if (nwin < 1) return EORB_OK;
    if (w < 1 || hgt < 1 || what < 0 || what > 2) return fail(EORB_ERR_ARG, "bad image size / metric");
    if (nwin > c->maxWindows) return fail(EORB_ERR_CAPACITY, "nwin %d > max_windows %d", nwin, c->maxWindows);
    const int patch = 30;   // DEF_PATCH_SIZE_STD (include/Event/EventConversion.h:28)
    if (((w + patch - 1) / patch) * ((hgt + patch - 1) / patch) > EORB_EV_FOCUS_MAX_CELLS) return fail(EORB_ERR_CAPACITY, "image too large for the focus metric");
    CU(cudaSetDevice(c->device));
    // the per-window min/max slots double as the output scratch (2 floats per window)
    CU(launch_ev_focus(d_img_f32, nwin, w, hgt, patch, what, avg ? 1 : 0, c->d_minmax, c->stream, &c->launches));
    CU(cudaMemcpyAsync(focus_out, c->d_minmax, (size_t)nwin * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return EORB_OK;
}

// the same for one host image (the shape of the reference's static call): H2D, metric, result
extern "C" int eorb_ev_image_focus(eorb_evconv* c, const float* img_f32, int w, int hgt, size_t stride_bytes, int what, int avg, float* focus_out) {
    if (!c || !focus_out) return fail(EORB_ERR_ARG, "null argument");
    if (!img_f32 || w < 1 || hgt < 1) return EORB_EMPTY;
    if ((long long)w * hgt > (long long)c->maxW * c->maxH) return fail(EORB_ERR_CAPACITY, "image %dx%d exceeds the converter's %dx%d", w, hgt, c->maxW, c->maxH);
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpy2DAsync(c->d_img, (size_t)w * 4, img_f32, stride_bytes, (size_t)w * 4, hgt, cudaMemcpyHostToDevice, c->stream));
    return eorb_ev_image_focus_device(c, c->d_img, 1, w, hgt, what, avg, focus_out);
}

extern "C" int eorb_ev_accumulate(eorb_evconv* c, const eorb_event* evs, int64_t n, const eorb_ev_params* p, float* img_f32,
                                  uint8_t* img_u8, float* minmax) {
    if (!c || !p) return fail(EORB_ERR_ARG, "null argument");
    if (p->width > c->maxW || p->height > c->maxH || (long long)p->width * p->height > (long long)c->maxW * c->maxH)
        return fail(EORB_ERR_CAPACITY, "image %dx%d exceeds the converter's %dx%d", p->width, p->height, c->maxW, c->maxH);
    if (n > c->maxEvents) return fail(EORB_ERR_CAPACITY, "%lld events > max_events %lld", (long long)n, c->maxEvents);
    const int npix = p->width * p->height;
    if (n <= 0 || !evs) {
        // reference: zero image; the SE3/SE2 overloads log "no events" and return it un-normalised (:292-295)
        if (img_f32) memset(img_f32, 0, (size_t)npix * sizeof(float));
        if (img_u8) memset(img_u8, 0, (size_t)npix);
        if (minmax) { minmax[0] = 0.f; minmax[1] = 0.f; }
        return EORB_EMPTY;
    }
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(c->d_evs, evs, (size_t)n * sizeof(eorb_event), cudaMemcpyHostToDevice, c->stream));
    const int64_t offs[2] = {0, n};
    eorb_ev_params q = *p;
    if (!img_u8) q.normalize = EORB_NORM_NONE;
    int rc = eorb_ev_accumulate_batch_device(c, c->d_evs, offs, 1, &q, nullptr, c->d_img, c->d_u8);
    if (rc != EORB_OK) return rc;
    if (img_f32) CU(cudaMemcpyAsync(img_f32, c->d_img, (size_t)npix * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (img_u8 && q.normalize != EORB_NORM_NONE) CU(cudaMemcpyAsync(img_u8, c->d_u8, (size_t)npix, cudaMemcpyDeviceToHost, c->stream));
    else if (img_u8) memset(img_u8, 0, (size_t)npix);
    if (minmax) CU(cudaMemcpyAsync(minmax, c->d_minmax, 2 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return EORB_OK;
}
