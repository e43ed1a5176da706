// oracle/ref_api.cc — C entry points of oracle/_ref/libref.so.  TEST INFRASTRUCTURE ONLY.
//
// libref.so = the reference's OWN translation units, compiled unmodified from /root/reference by `make -C oracle ref`:
//     src/ORBextractor.cc                     (whole file)
//     src/Event/EventConversion.cc            (whole file)
//     src/ORBmatcher.cc:2360-2378             (DescriptorDistance, cut out of the source at build time)
// against the header-only stand-ins in oracle/ref_mock/ (cvmini.hpp, eigenmini.hpp, ...), plus this file, which only
// marshals plain C arrays in and out.  It pins the oracle (oracle/*.cc restatement) to what the reference's code does:
// tests/test_ref_pin.py asserts oracle == libref byte for byte, tests/golden/make_ref_golden.py writes libref's outputs
// to tests/golden/ref_*.npz, and the GPU suite checks the CUDA path against those files.
//
// Heap policy (the octree's size-tie rule).  DistributeOctTree sorts pair<int, ExtractorNode*> (ORBextractor.cc:703), so
// equal-sized nodes are ordered by the ADDRESS of their std::list node.  alloc_mode selects what `operator new` inside
// this library (and only inside it: the version script keeps it local) hands out:
//     REF_ALLOC_BUMP   a monotone bump arena: address order == creation order, never reused.  This is the pin the oracle
//                      states ("ascending creation sequence plays the role of ascending pointer value").
//     REF_ALLOC_MALLOC glibc malloc as a stock build would use; address order then depends on tcache/bin reuse and on the
//                      heap's history.  ref_orb_extract reports results under both so the difference can be counted.
#include <pthread.h>
#include <sys/mman.h>

#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "ORBextractor.h"   // the reference's include/ORBextractor.h
#include "ORBmatcher.h"     // ref_mock stand-in (one static member)
#include "oracle.h"

// ------------------------------------------------------------------------------------------------ heap policy
namespace {
const size_t kArenaBytes = (size_t)3 << 30;   // virtual; pages are touched only as used
char* g_arena = nullptr;
std::atomic<size_t> g_off{0};
std::atomic<long> g_overflow{0};
thread_local bool t_bump = false;
std::mutex g_bumpSession;

void arenaInit() {
    if (g_arena) return;
    void* p = mmap(nullptr, kArenaBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { std::fprintf(stderr, "libref: cannot map the bump arena\n"); std::abort(); }
    g_arena = (char*)p;
}
inline bool inArena(const void* p) { return g_arena && (const char*)p >= g_arena && (const char*)p < g_arena + kArenaBytes; }

inline void* refAlloc(size_t n) {
    if (t_bump) {
        n = (n + 15) & ~(size_t)15;
        size_t o = g_off.fetch_add(n);
        if (o + n <= kArenaBytes) return g_arena + o;
        g_overflow++;
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
inline void refFree(void* p) { if (p && !inArena(p)) std::free(p); }

struct BumpSession {   // one thread at a time; everything allocated inside dies inside (stateless entry points)
    bool on;
    explicit BumpSession(bool enable) : on(enable) {
        if (!on) return;
        g_bumpSession.lock();
        arenaInit();
        g_off = 0; g_overflow = 0;
        t_bump = true;
    }
    ~BumpSession() { if (on) { t_bump = false; g_bumpSession.unlock(); } }
};
}  // namespace

void* operator new(size_t n) { return refAlloc(n); }
void* operator new[](size_t n) { return refAlloc(n); }
void operator delete(void* p) noexcept { refFree(p); }
void operator delete[](void* p) noexcept { refFree(p); }
void operator delete(void* p, size_t) noexcept { refFree(p); }
void operator delete[](void* p, size_t) noexcept { refFree(p); }

// ------------------------------------------------------------------------------------------------ the reference's text
namespace ORB_SLAM3 {
extern int EDGE_THRESHOLD;   // ORBextractor.cc:73 (a mutable global the constructor overwrites, :481-488)
#include "gen_descriptor_distance.inc"   // ORBmatcher.cc:2360-2378, cut out by the Makefile
}  // namespace ORB_SLAM3

namespace {
// protected members are reached through a derived class; nothing is re-implemented
// EDGE_THRESHOLD is ONE mutable global that every constructor overwrites, and the adaptive rule (edgeTh < 0) scales its
// CURRENT value (ORBextractor.cc:481-485): in the reference the margin of an adaptive extractor depends on which extractors
// were built before it in the process.  The pin (oracle: "per extractor instance") is the first-extractor-in-a-fresh-process
// value, so the global is put back to its initialiser (:73) before every construction here.
struct FreshEdge { FreshEdge() { ORB_SLAM3::EDGE_THRESHOLD = 19; } };
struct RefExtractor : FreshEdge, ORB_SLAM3::ORBextractor {
    explicit RefExtractor(const ORB_SLAM3::ORBxParams& p) : FreshEdge(), ORB_SLAM3::ORBextractor(p) {}
    const std::vector<int>& featuresPerLevel() const { return mnFeaturesPerLevel; }
    const std::vector<int>& umaxTable() const { return umax; }
    static std::vector<cv::KeyPoint> distribute(const std::vector<cv::KeyPoint>& keys, int minX, int maxX, int minY, int maxY, int N) {
        return DistributeOctTree(keys, minX, maxX, minY, maxY, N, 0, N);
    }
};

ORB_SLAM3::ORBxParams toParams(const orc_orb_params* p) {
    return ORB_SLAM3::ORBxParams(p->nfeatures, p->scaleFactor, p->nlevels, p->iniThFAST, p->minThFAST, p->edgeTh, cv::Size(p->imW, p->imH));
}
void copyKps(const std::vector<cv::KeyPoint>& v, orc_keypoint* out, int cap) {
    static_assert(sizeof(orc_keypoint) == sizeof(cv::KeyPoint), "keypoint layout");
    const int n = (int)v.size() < cap ? (int)v.size() : cap;
    if (n > 0) std::memcpy(out, v.data(), (size_t)n * sizeof(orc_keypoint));
}
}  // namespace

extern "C" {

enum { REF_ALLOC_MALLOC = 0, REF_ALLOC_BUMP = 1 };

int ref_version(void) { return 2; }

/* ORBextractor::ORBextractor (ORBextractor.cc:420-489) and the level sizes ComputePyramid derives (:1244-1245) */
int ref_orb_tables(const orc_orb_params* p, int w, int h, int* feats_per_level, float* scale, float* inv_scale, float* sigma2,
                   float* inv_sigma2, int* umax16, int* edge, int* level_w, int* level_h) {
    RefExtractor ex(toParams(p));
    const int nl = ex.GetLevels();
    std::vector<float> s = ex.GetScaleFactors(), is = ex.GetInverseScaleFactors(), g = ex.GetScaleSigmaSquares(), ig = ex.GetInverseScaleSigmaSquares();
    for (int l = 0; l < nl; l++) {
        if (feats_per_level) feats_per_level[l] = ex.featuresPerLevel()[l];
        if (scale) scale[l] = s[l];
        if (inv_scale) inv_scale[l] = is[l];
        if (sigma2) sigma2[l] = g[l];
        if (inv_sigma2) inv_sigma2[l] = ig[l];
        if (level_w) level_w[l] = cvRound((float)w * is[l]);
        if (level_h) level_h[l] = cvRound((float)h * is[l]);
    }
    if (umax16) for (int i = 0; i < 16; i++) umax16[i] = ex.umaxTable()[i];
    if (edge) *edge = ORB_SLAM3::EDGE_THRESHOLD;
    return nl;
}

/* ORBextractor::operator() (ORBextractor.cc:1092-1238).  Stateless: builds the extractor, runs it, copies everything out.
   pyr_out / blur_out (may be NULL): the unbordered levels / the blurred levels that had keypoints, concatenated level by
   level at their natural size (levels without keypoints are skipped in blur_out and flagged 0 in blur_present[level]).
   taps[8]: FAST calls, FAST calls that returned keypoints, total candidates, bump overflow count, arena bytes used, 0, 0, 0 */
int ref_orb_extract(const orc_orb_params* p, const uint8_t* img, int w, int h, size_t stride, int lap0, int lap1, int want_desc,
                    int alloc_mode, orc_keypoint* kps, uint8_t* desc, int cap, int* n_out, uint8_t* pyr_out, uint8_t* blur_out,
                    int* blur_present, int64_t* taps) {
    BumpSession session(alloc_mode == REF_ALLOC_BUMP);
    cv::fastTap() = cv::FastTap();
    std::vector<cv::Mat> blurred;
    cv::gaussTap().sink = blur_out ? &blurred : nullptr;
    int ret;
    {
        RefExtractor ex(toParams(p));
        cv::Mat image = (w > 0 && h > 0) ? cv::Mat(h, w, CV_8UC1, (void*)img, stride) : cv::Mat();
        std::vector<cv::KeyPoint> keys;
        std::vector<int> lap = {lap0, lap1};
        cv::Mat descriptors;
        if (want_desc) ret = ex(image, cv::Mat(), keys, descriptors, lap);
        else ret = ex(image, cv::Mat(), keys, lap);
        if (n_out) *n_out = (int)keys.size();
        if (kps) copyKps(keys, kps, cap);
        if (want_desc && desc)
            for (int i = 0; i < (int)keys.size() && i < cap; i++) std::memcpy(desc + (size_t)i * 32, descriptors.ptr(i), 32);
        if (ret >= 0) {
            const int nl = ex.GetLevels();
            if (pyr_out) {
                size_t o = 0;
                for (int l = 0; l < nl; l++) {
                    const cv::Mat& m = ex.mvImagePyramid[l];
                    for (int y = 0; y < m.rows; y++) { std::memcpy(pyr_out + o, m.ptr(y), m.cols); o += m.cols; }
                }
            }
            if (blur_out) {
                std::vector<int> perLevel(nl, 0);
                for (const cv::KeyPoint& k : keys) perLevel[k.octave]++;
                size_t o = 0, b = 0;
                for (int l = 0; l < nl; l++) {
                    if (blur_present) blur_present[l] = perLevel[l] > 0;
                    if (!perLevel[l]) continue;
                    const cv::Mat& m = blurred[b++];
                    for (int y = 0; y < m.rows; y++) { std::memcpy(blur_out + o, m.ptr(y), m.cols); o += m.cols; }
                }
            }
        }
    }
    cv::gaussTap().sink = nullptr;
    if (taps) {
        const cv::FastTap& t = cv::fastTap();
        taps[0] = t.calls; taps[1] = t.calls_nonempty; taps[2] = t.candidates; taps[3] = g_overflow.load();
        taps[4] = session.on ? (int64_t)g_off.load() : 0; taps[5] = taps[6] = taps[7] = 0;
    }
    return ret;
}

/* ORBextractor::DistributeOctTree (ORBextractor.cc:558-782) on its own; keys relative to (minX, minY) as the caller at
   :877 passes them; class_id carries the input index so that the selection can be read back; returns the count */
int ref_distribute_octtree(const float* kx, const float* ky, const float* kresp, int n, int minX, int maxX, int minY, int maxY, int N,
                           int alloc_mode, int* out_idx, int cap) {
    BumpSession session(alloc_mode == REF_ALLOC_BUMP);
    int m;
    {
        std::vector<cv::KeyPoint> in(n);
        for (int i = 0; i < n; i++) { in[i] = cv::KeyPoint(kx[i], ky[i], 7.f, -1, kresp[i]); in[i].class_id = i; }
        std::vector<cv::KeyPoint> r = RefExtractor::distribute(in, minX, maxX, minY, maxY, N);
        m = (int)r.size();
        for (int i = 0; i < m && i < cap; i++) out_idx[i] = r[i].class_id;
    }
    return m;
}

/* ORBextractor::ComputeTrackedKPtsDesc (ORBextractor.cc:1316-1363) */
int ref_orb_tracked_desc(const orc_orb_params* p, const uint8_t* img, int w, int h, size_t stride, const orc_keypoint* kps, int n, uint8_t* desc) {
    RefExtractor ex(toParams(p));
    cv::Mat image(h, w, CV_8UC1, (void*)img, stride);
    std::vector<cv::KeyPoint> keys(n);
    if (n) std::memcpy(keys.data(), kps, (size_t)n * sizeof(orc_keypoint));
    cv::Mat d;
    ex.ComputeTrackedKPtsDesc(image, keys, d);
    for (int i = 0; i < n; i++) std::memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
    return n;
}

/* ORBextractor::AssignKPtLevelByBestDesc (ORBextractor.cc:1267-1314); kps[].octave is updated in place */
int ref_orb_assign_level_by_best_desc(const orc_orb_params* p, const uint8_t* ref_desc, const uint8_t* img, int w, int h, size_t stride,
                                      orc_keypoint* kps, int n) {
    RefExtractor ex(toParams(p));
    cv::Mat image(h, w, CV_8UC1, (void*)img, stride);
    cv::Mat rd(n, 32, CV_8UC1, (void*)ref_desc, 32);
    std::vector<cv::KeyPoint> keys(n);
    if (n) std::memcpy(keys.data(), kps, (size_t)n * sizeof(orc_keypoint));
    ex.AssignKPtLevelByBestDesc(rd, image, keys);
    if (n) std::memcpy(kps, keys.data(), (size_t)n * sizeof(orc_keypoint));
    return n;
}

/* ORBmatcher::DescriptorDistance (ORBmatcher.cc:2360-2378) */
int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    cv::Mat ma(1, 32, CV_8UC1, (void*)a, 32), mb(1, 32, CV_8UC1, (void*)b, 32);
    return ORB_SLAM3::ORBmatcher::DescriptorDistance(ma, mb);
}

/* nq x ndb distance matrix through the reference's DescriptorDistance (row-major int32) */
void ref_descriptor_distance_matrix(const uint8_t* q, int nq, const uint8_t* db, int ndb, int32_t* out) {
    for (int i = 0; i < nq; i++) {
        cv::Mat ma(1, 32, CV_8UC1, (void*)(q + (size_t)i * 32), 32);
        for (int j = 0; j < ndb; j++) {
            cv::Mat mb(1, 32, CV_8UC1, (void*)(db + (size_t)j * 32), 32);
            out[(size_t)i * ndb + j] = ORB_SLAM3::ORBmatcher::DescriptorDistance(ma, mb);
        }
    }
}

/* thread pool over frames, one extractor per thread, glibc malloc (the CPU baseline of bench.py --impl reference);
   frames are w*h contiguous; returns the total keypoint count */
long ref_orb_extract_batch_mt(const orc_orb_params* p, const uint8_t* imgs, int nframes, int w, int h, int nthreads, int want_desc,
                              int* n_per_frame) {
    if (nthreads < 1) nthreads = 1;
    // extractors are built one after the other on this thread (the constructor writes the EDGE_THRESHOLD global)
    std::vector<std::unique_ptr<RefExtractor>> exs;
    for (int t = 0; t < nthreads; t++) exs.emplace_back(new RefExtractor(toParams(p)));
    std::atomic<int> next{0};
    std::atomic<long> total{0};
    auto work = [&](int tid) {
        RefExtractor& ex = *exs[(size_t)tid];
        std::vector<int> lap = {0, 0};
        for (;;) {
            int f = next.fetch_add(1);
            if (f >= nframes) break;
            cv::Mat image(h, w, CV_8UC1, (void*)(imgs + (size_t)f * w * h), (size_t)w);
            std::vector<cv::KeyPoint> keys;
            cv::Mat d;
            if (want_desc) ex(image, cv::Mat(), keys, d, lap);
            else ex(image, cv::Mat(), keys, lap);
            if (n_per_frame) n_per_frame[f] = (int)keys.size();
            total += (long)keys.size();
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    return total.load();
}

}  // extern "C"
