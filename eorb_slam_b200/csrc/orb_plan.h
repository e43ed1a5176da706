// orb_plan.h — per-extractor geometry tables shared by host code and kernels (plain POD, lives in HBM).
// Everything here is derived once per (params, image size) on the host exactly as the reference derives it:
// ctor tables ORBextractor.cc:420-489, level sizes :1244-1245, FAST grid :792-828, octree roots :562-563.
#pragma once
#include <stdint.h>

#define EORB_MAX_LEVELS 32
#define EORB_MAX_DIM 4095          // candidate coordinates are packed in 12 bits
#ifndef EORB_BLUR_BAND
#define EORB_BLUR_BAND 64          // rows per blur task; multiple of 8 (eight rows are loaded per group, the row-pair ring has period 4)
#endif
#ifndef EORB_FAST_WARPS
#define EORB_FAST_WARPS 1          // warps (= grid cells) per FAST thread block
#endif

namespace eorb {

struct LevelPlan {
    int w, h, pitch;               // level size, row pitch in bytes (multiple of 16); level 0 pitch is the input's
    long long off;                 // byte offset inside one frame's pyramid slab (levels >= 1)
    long long blurOff;             // byte offset inside one frame's blurred slab (all levels, own pitch = bpitch)
    int bpitch;
    int minBX, minBY, maxBX, maxBY;
    int nCols, nRows, wCell, hCell;
    int cellBase, nCells;          // cells of this level are [cellBase, cellBase + nCells), row-major
    int slotBase, slotCount;       // candidate slots of this level inside one frame's candidate slab
    int quota;                     // mnFeaturesPerLevel[level]
    int nIni;                      // octree roots: round(width/height)
    float hX;                      // width / nIni
    int nodeCap;                   // octree node capacity = max(quota+3, 4*nIni)+1
    int selBase;                   // offset of this level's selected-keypoint slots inside one frame's slab
    float scale;                   // mvScaleFactor[level]
    float sizeF;                   // (float)(int)(31*scale)
    int xtabOff, ytabOff;          // offsets (in entries) into the resize tables (levels >= 1)
    int blurTaskBase;              // first flattened (32-row band, 128-column strip) task of this level (blur grid)
    // TMA-staged resize of level l-1 into this level (orb_tiles.cu): destination tile TW x TH (TW multiple of 4, <= 128; TH multiple of
    // 4), source box BW x BH bytes (BW multiple of 16, <= 256); pyrTW == 0: the level uses pyr_resize_kernel (direct global loads)
    int pyrTW, pyrTH, pyrBW, pyrBH;
};

struct CellPlan {
    short x0, y0;                  // ROI origin in level coordinates (iniX, iniY)
    short w, h;                    // ROI size (maxX-iniX, maxY-iniY), includes the 3-px FAST apron
    short level, _pad;
    int slotOff;                   // first candidate slot of this cell (relative to the frame's slab)
    short ox, oy;                  // j*wCell, i*hCell: added to ROI coordinates (ORBextractor.cc:871-872)
    int slotCap;
    // FAST phase-1 walk, precomputed on the host: the TMA tile starts at column x0 & ~15, the interior is tile
    // columns [aoff+3, aoff+w-3), walked in np aligned groups of 8 columns starting at group p0, rps rows per step
    unsigned char aoff, p0, np, rps;
    unsigned char firstMask, lastMask;   // interior columns inside the first / last group
    unsigned short rcpNpM1;        // ceil(65536 / np) - 1: lane / np == (lane * (rcpNpM1 + 1)) >> 16 for lane < 32
};

struct OrbPlan {
    int nlevels, edge, iniTh, minTh;
    int W, H;
    int nCells;                    // total cells over all levels
    int slotsPerFrame;             // candidate slots per frame
    int selPerFrame;               // selected-keypoint slots per frame (sum of nodeCap)
    long long pyrBytesPerFrame;    // levels >= 1
    long long blurBytesPerFrame;
    int cellTileStride;            // FAST: TMA box width = smem bytes per tile row (multiple of 16)
    int cellTileRows;              // FAST: TMA box height (largest cell ROI height)
    int cellMapStride;             // FAST: score-map row stride (multiple of 4)
    int cellMapOff;                // FAST: byte offset of the score map inside a warp's smem region
    int cellListOff;               // FAST: byte offset of the survivor (pixel) list
    int cellTaskOff;               // FAST: byte offset of the flagged-group list of phase 1
    int cellBarOff;                // FAST: byte offset of the warp's mbarrier
    int cellSmemPerWarp;           // bytes (multiple of 128: TMA destinations are 128-byte aligned)
    int octSmemBytes;              // max over levels
    int blurTasksTotal;            // blur grid: total (band, strip) tasks over all levels
    int umax[16];
    LevelPlan lv[EORB_MAX_LEVELS];
};

}  // namespace eorb
