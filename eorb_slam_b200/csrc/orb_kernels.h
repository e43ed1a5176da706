// orb_kernels.h — launch interface between the C-ABI host code (capi.cu) and orb_kernels.cu
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eorb_b200.h"
#include "orb_plan.h"

namespace eorb {

// Everything a kernel needs, passed by value.  Per-frame slabs are indexed by the frame's position in the
// current batch; offsets inside a slab come from the plan.
struct OrbArgs {
    const OrbPlan* plan;         // device copy
    const CellPlan* cells;       // device
    const short4* xtab;          // device: resize taps per destination column  {sx, sx+1, a0, a1}
    const int4* ytab;            // device: resize taps per destination row     {sy0, sy1, b0 << 16, b1 << 16}
    const int4* ytabT;           // device: the same for the TMA-staged resize  {sy0 * BW, sy1 * BW, b0 << 16, b1 << 16} (BW = the level's box width)
    const uint8_t* lvl0;         // level 0 = the input frames (zero-copy when aligned, else staged)
    long long lvl0Pitch, lvl0FrameStride;
    const CUtensorMap* tmaps;    // device: [nlevels] TMA maps {x, y, frame} of the pyramid levels >= 1 (entry 0 unused:
                                 //         level 0 is the caller's buffer, its map travels as a kernel parameter)
    const CUtensorMap* blurMaps; // device: [nlevels] maps of the levels >= 1 with the box of blur_tma_kernel, or nullptr (blur_kernel is used);
                                 //         level 0's travels as pyrMaps[0] of launch_orb_pipeline
    const CUtensorMap* briefMaps;// device: [nlevels] maps of the BLURRED levels with the box orient_desc_kernel stages a keypoint's BRIEF patch with, or nullptr
    const CUtensorMap* icMaps;   // device: [nlevels] maps of the levels >= 1 with the box of the orientation patch (level 0's travels as a kernel
                                 //         parameter: icMap0 of launch_orb_pipeline), or nullptr
    int* pyrDone;                // device: [8][EORB_MAX_LEVELS] tile counters of pyr_chain_kernel (small batches), or nullptr
    int blurVariant;             // EORB_BLUR_TMA value (1..4: band rows 32 / 64, neighbour words by shuffle / from the tile)
    uint8_t* pyr;                // [B][pyrBytesPerFrame]   levels >= 1
    uint8_t* blur;               // [B][blurBytesPerFrame]  all levels
    uint16_t* cellCount;         // [B][nCells]
    uint32_t* cand;              // [B][slotsPerFrame]      per-cell candidate lists (packed x|y<<12|score<<24)
    uint32_t* okeys;             // [B][slotsPerFrame]      per-level ordered candidates
    uint16_t* knode;             // [B][slotsPerFrame]      octree: list position of each candidate's node
    uint32_t* sel;               // [B][selPerFrame]        selected candidates per level, list order
    int* selCount;               // [B][nlevels]
    int* candCount;              // [B][nlevels]
    int* dstIdx;                 // [B][selPerFrame]        final output position
    uint32_t* kpList;            // [B][selPerFrame]        compact list of selected keypoints: slot | level << 24
    const int2* icTab;           // device: [4][288] IC_Angle weights (u, v) as packed s8x4 per (alignment, row, word)
    float* levelAngle;           // [B][selPerFrame]        (debug tap) or nullptr
    eorb_keypoint* outKps;       // [B][cap]
    uint8_t* outDesc;            // [B][cap][32]
    int* outN;                   // [B]
    int* outMono;                // [B]
    int cap;
    int lap0, lap1;
    int wantDesc;
};

// per-level constants of orient_desc_kernel, passed with the kernel parameters (constant bank)
struct OdLevels {
    int4 geo[EORB_MAX_LEVELS];     // minBX, minBY, w, h
    float2 sc[EORB_MAX_LEVELS];    // mvScaleFactor[level], (float)(int)(31 * scale)
    int selPerFrame;
};


// constants of one TMA-staged pyramid launch (orb_tiles.cu), passed by value
struct PyrTileConst {
    int TW, TH, BW, BH, barOff;
    int dw, dh, dpitch;
    int xtabOff, ytabOff;
    long long doff, pyrBytes;
};
cudaError_t launch_pyr_tma(const OrbArgs& a, const OrbPlan& hp, int level, int nframes, const CUtensorMap& tmSrc, cudaStream_t st);

cudaError_t orb_kernels_configure(const OrbPlan& hp);
int blur_tma_box_w();
int brief_tma_box_w();   // box of one keypoint's BRIEF patch on the blurred level (orient_desc_kernel)
int brief_tma_box_h();
int ic_tma_box_w();      // box of one keypoint's orientation patch on the level (orient_desc_kernel)
int ic_tma_box_h();
int blur_tma_box_h(int variant);
#define EORB_ORB_STAGES 6   // pyramid, fast, octree, index, blur, orient+desc
// side-stream fork of one launch set (small batches inside the captured graph only: a single frame leaves the GPU mostly idle, so the
// blur, which only depends on the pyramid, runs beside FAST / octree / index and joins before orientation + BRIEF; at 1024-frame
// launch sets the same fork was measured neutral, 17.69 vs 17.73 ms per 4096 frames: every kernel fills the machine on its own)
struct OrbFork { cudaStream_t side = nullptr; cudaEvent_t forked = nullptr, joined = nullptr; };
// pyrMaps (host array, [nlevels], may be null): TMA map of the SOURCE of level l (= level l-1) with that level's box, for the
// levels whose plan says pyrTW > 0; null or pyrTW == 0 -> pyr_resize_kernel
cudaError_t launch_orb_pipeline(const OrbArgs& a, const OrbPlan& hp, int nframes, const CUtensorMap& tm0, cudaStream_t st,
                                long long* launches, cudaEvent_t* ev, const CUtensorMap* pyrMaps = nullptr, const OrbFork* fork = nullptr,
                                const CUtensorMap* icMap0 = nullptr);
cudaError_t launch_fast_cells(const OrbArgs& a, const OrbPlan& hp, int nframes, const CUtensorMap& tm0, cudaStream_t st);
cudaError_t fast_cells_configure(const OrbPlan& hp);
cudaError_t launch_pyramid_and_blur(const OrbArgs& a, const OrbPlan& hp, cudaStream_t st, long long* launches, const CUtensorMap* pyrMaps = nullptr);
cudaError_t launch_tracked_desc(const OrbArgs& a, const OrbPlan& hp, const eorb_keypoint* d_kps, int n, int mode,
                                const float* d_invScale, const uint8_t* d_refDesc, uint8_t* d_desc, int* d_dist,
                                cudaStream_t st, long long* launches);

cudaError_t launch_selftest_math(const int* fastIn, int nFast, int* fastOut, const float* atanIn, int nAtan, float* atanOut,
                                 int* briefOut, cudaStream_t st);

}  // namespace eorb
