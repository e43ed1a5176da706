/*
 * eorb_b200.h — C ABI of libeorb_b200.so: the B200 (sm_100a) front-end hot path of EORB-SLAM.
 *
 * The reference (m-dayani/EORB_SLAM) has no FFI layer: the hot path sits behind three C++ class
 * surfaces linked into libORB_SLAM3.so.  This header is the thin C boundary those classes are
 * re-implemented on (shims with the reference signatures live in eorb_slam_b200/shim/ and
 * INTEGRATION.md shows the binding a maintainer adds).  Each entry point names the reference
 * interface it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary; no exceptions.
 *   - return value: >= 0 success (0, or the value the reference function returns),
 *       EORB_EMPTY (-1)  empty input, mirroring `return -1` at ORBextractor.cc:1096 /
 *                         the zero image returned at EventConversion.cc:292-295,
 *       <= -2            error (see codes); eorb_last_error() gives a thread-local message.
 *   - "host" entry points take host buffers and do H2D/D2H on the handle's stream, then synchronise.
 *     "_device" entry points take device pointers, enqueue on the handle's stream and do NOT synchronise.
 *   - a handle is not re-entrant (like the reference extractor, ORBextractor.h:105); distinct handles may
 *     be used concurrently from distinct threads (Frame.cc:122-125, EvImBuilder.cpp:1165-1193).
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns EORB_ERR_CUDA.
 */
#ifndef EORB_B200_H
#define EORB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EORB_OK            0
#define EORB_EMPTY        (-1)
#define EORB_ERR_ARG      (-2)
#define EORB_ERR_CAPACITY (-3)
#define EORB_ERR_STATE    (-4)
#define EORB_ERR_CUDA     (-10)

#define EORB_DESC_BYTES 32

/* mirrors ORB_SLAM3::ORBxParams (include/ORBextractor.h:33-47); patchSize is fixed at 31 */
typedef struct eorb_orb_params {
    int   nfeatures;
    float scaleFactor;
    int   nlevels;
    int   iniThFAST;
    int   minThFAST;
    int   edgeTh;       /* < 0 : adaptive 19*imW/752 forced odd (ORBextractor.cc:481-485) */
    int   imW, imH;     /* ORBxParams::imSize; only used for the adaptive edge threshold */
} eorb_orb_params;

/* bit-compatible with cv::KeyPoint (28 bytes) */
typedef struct eorb_keypoint {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} eorb_keypoint;

/* bit-compatible with EORB_SLAM::EventData (include/Event/EventData.h:36-58): 24 bytes */
typedef struct eorb_event {
    double  ts;
    float   x, y;
    uint8_t p;
    uint8_t _pad[7];
} eorb_event;

/* result of the best-2 scan of ORBmatcher (ORBmatcher.cc:741-772 / :318-382) for one query */
typedef struct eorb_match {
    int32_t best_dist;    /* 256 when the database is empty */
    int32_t best_idx;     /* global row index, -1 when none; lowest index wins distance ties */
    int32_t second_dist;  /* 256 when fewer than two rows */
    int32_t accepted;     /* best_dist <= th && (float)best_dist < ratio*(float)second_dist */
} eorb_match;

/* per-shard partial result exchanged between GPUs (16 bytes): keys are (dist << 32) | global_idx */
typedef struct eorb_best2 {
    uint64_t key1;        /* smallest  (dist, idx); ~0 when none */
    uint64_t key2;        /* second smallest */
} eorb_best2;

typedef struct eorb_orb     eorb_orb;
typedef struct eorb_matcher eorb_matcher;
typedef struct eorb_evconv  eorb_evconv;

/* ---------------------------------------------------------------- library */
int         eorb_version(void);
const char* eorb_last_error(void);
int         eorb_device_count(void);                 /* 0 when no CUDA device / driver */
/* CUDA-event timing on a handle's stream, so callers never need the CUDA runtime themselves */
int         eorb_timer_create(void** timer);         /* a pair of cudaEvents */
int         eorb_timer_destroy(void* timer);
int         eorb_timer_start(void* timer, void* cuda_stream);
int         eorb_timer_stop(void* timer, void* cuda_stream);
int         eorb_timer_elapsed_ms(void* timer, float* ms);   /* synchronises on the stop event */
/* integer-pipe POPC throughput probe used as the Hamming roofline denominator */
int         eorb_probe_popc_rate(int device, double* popc32_per_sec);
/* device-vs-host evaluation of the shared scalar arithmetic (FAST arc score, fastAtan2, steered-BRIEF offsets);
 * *mismatches must come back 0.  Guards against toolchain miscompiles (see DESIGN.md "Toolchain findings"). */
int         eorb_selftest_math(int device, int* mismatches);

/* ---------------------------------------------------------------- ORB extractor
 * replaces ORB_SLAM3::ORBextractor (include/ORBextractor.h:62-136, src/ORBextractor.cc) */

/* ORBextractor::ORBextractor(const ORBxParams&)  ORBextractor.cc:420-489.
 * max_batch = frames processed per launch set by the batch entry points (>=1). */
int eorb_orb_create(const eorb_orb_params* params, int device, int max_batch, eorb_orb** out);
int eorb_orb_destroy(eorb_orb* h);
/* use a caller-owned cudaStream_t (e.g. the framework's current stream; NULL = the CUDA default stream);
 * _reset_stream returns to the handle's own non-blocking stream */
int eorb_orb_set_stream(eorb_orb* h, void* cuda_stream);
int eorb_orb_reset_stream(eorb_orb* h);
void* eorb_orb_get_stream(eorb_orb* h);
int eorb_orb_synchronize(eorb_orb* h);

/* GetLevels / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * (ORBextractor.h:83-101), mnFeaturesPerLevel and the per-instance EDGE_THRESHOLD.  Any pointer may be NULL. */
int eorb_orb_tables(const eorb_orb* h, int* nlevels, int* edge_threshold, float* scale, float* inv_scale,
                    float* sigma2, float* inv_sigma2, int* features_per_level);
/* the same tables straight from the parameters: host arithmetic only (the constructor's, ORBextractor.cc:420-489), no device and no
 * handle needed -- lets ORBextractor::ORBextractor fill its tables on any thread and create device handles lazily per calling thread */
int eorb_orb_params_tables(const eorb_orb_params* params, int* nlevels, int* edge_threshold, float* scale, float* inv_scale,
                           float* sigma2, float* inv_sigma2, int* features_per_level);
/* capacity a caller must provide per frame of w x hgt pixels: per level max(quota + 3, 4 * nIni) keypoints, nIni = round of the
 * level's aspect ratio (the octree's first pass yields up to 4 children per root, ORBextractor.cc:562-563, 620-683).
 * eorb_orb_max_keypoints: the same for the size of the last extracted frames, before the first call for the (imW, imH) of the
 * parameters. */
int eorb_orb_max_keypoints_for_size(const eorb_orb* h, int w, int hgt);
int eorb_orb_max_keypoints(const eorb_orb* h);

/* int ORBextractor::operator()(image, mask, keypoints, descriptors, vLappingArea)   ORBextractor.cc:1092-1176
 * and the keypoints-only overload :1178-1238 (want_desc = 0, desc may be NULL).
 * Host buffers.  Returns monoIndex (>= 0), EORB_EMPTY for an empty image, or an error. */
int eorb_orb_extract(eorb_orb* h, const uint8_t* img, int w, int hgt, size_t stride,
                     int lap0, int lap1, int want_desc,
                     eorb_keypoint* kps, uint8_t* desc, int cap, int* n_out);

/* Config-3 shape: nframes images (host), frame f at imgs + f*frame_stride.  Outputs are [nframes][cap]
 * (kps), [nframes][cap][32] (desc), [nframes] (n_out, mono_out).  Frames are processed max_batch at a time
 * with H2D / compute / D2H overlapped on internal streams.  Entries of kps / desc past n_out[f] are unspecified (when the
 * caller's arrays are pinned the device writes whole slabs straight into them).  On an error return every internal copy
 * has been drained: no write into the caller's arrays is still in flight. */
int eorb_orb_extract_batch(eorb_orb* h, const uint8_t* imgs, int nframes, int w, int hgt, size_t row_stride,
                           size_t frame_stride, int lap0, int lap1, int want_desc,
                           eorb_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* mono_out);

/* Same with everything resident in HBM (device pointers); nframes <= max_batch; asynchronous. */
int eorb_orb_extract_batch_device(eorb_orb* h, const uint8_t* d_imgs, int nframes, int w, int hgt, size_t row_stride,
                                  size_t frame_stride, int lap0, int lap1, int want_desc,
                                  eorb_keypoint* d_kps, uint8_t* d_desc, int cap, int* d_n_out, int* d_mono_out);

/* std::vector<cv::Mat> mvImagePyramid (ORBextractor.h:105, read by Frame.cc:876,966-985): lazy D2H of one
 * level of frame `frame` of the last call.  dst is w_l x h_l (unbordered; the reference's border is the
 * REFLECT_101 image of these pixels). */
int eorb_orb_level_size(const eorb_orb* h, int level, int* w, int* hgt);
int eorb_orb_pyramid_level(eorb_orb* h, int frame, int level, uint8_t* dst, size_t dst_stride);

/* ORBextractor::ComputeTrackedKPtsDesc (ORBextractor.cc:1316-1363) */
int eorb_orb_tracked_desc(eorb_orb* h, const uint8_t* img, int w, int hgt, size_t stride,
                          const eorb_keypoint* kps, int n, uint8_t* desc);
/* ORBextractor::AssignKPtLevelByBestDesc (ORBextractor.cc:1267-1314): updates kps[i].octave */
int eorb_orb_assign_level_by_best_desc(eorb_orb* h, const uint8_t* ref_desc, const uint8_t* img, int w, int hgt,
                                       size_t stride, eorb_keypoint* kps, int n);

/* stage taps of the last call (parity tests; not part of the reference surface) */
int eorb_orb_debug_blurred(eorb_orb* h, int frame, int level, uint8_t* dst, size_t dst_stride);
int eorb_orb_debug_candidates(eorb_orb* h, int frame, int level, int* xs, int* ys, int* scores, int cap);
int eorb_orb_debug_level_kps(eorb_orb* h, int frame, int level, int* xs, int* ys, int* scores, float* angles, int cap);
/* counts kernel launches issued by this handle since creation (bench.py's gpu_launches) */
long long eorb_orb_launch_count(const eorb_orb* h);
/* per-stage device timing for bench.py: when enabled, cudaEvents are recorded on the handle's stream around
 * each of the 6 stages (pyramid, fast, octree, index, blur, orient+desc) of every call; _stage_times
 * synchronises, returns the summed milliseconds and launch counts since the last call, and resets them. */
int eorb_orb_stage_timing(eorb_orb* h, int enable);
int eorb_orb_stage_times(eorb_orb* h, float* ms6, long long* launches6);

/* ---------------------------------------------------------------- matcher
 * replaces ORBmatcher::DescriptorDistance (ORBmatcher.cc:2360-2378) and the best/second-best scan +
 * ratio test that every ORBmatcher::Search* bottoms out in (:318-382, :741-772), in the brute-force shape
 * of Frame.cc:1228-1235 */

int eorb_descriptor_distance(const uint8_t* a, const uint8_t* b);     /* host helper, 0..256 */

int eorb_matcher_create(int device, eorb_matcher** out);
int eorb_matcher_destroy(eorb_matcher* m);
int eorb_matcher_set_stream(eorb_matcher* m, void* cuda_stream);
int eorb_matcher_reset_stream(eorb_matcher* m);
int eorb_matcher_synchronize(eorb_matcher* m);
long long eorb_matcher_launch_count(const eorb_matcher* m);
/* database shard: rows [index_offset, index_offset+ndb) of the global database.
 * _host copies into HBM; _device adopts a device pointer (not owned). */
int eorb_matcher_set_db_host(eorb_matcher* m, const uint8_t* db, int64_t ndb, int64_t index_offset);
int eorb_matcher_set_db_device(eorb_matcher* m, const uint8_t* d_db, int64_t ndb, int64_t index_offset);
/* one-call brute force, host buffers: best-2 + threshold + ratio -> out[nq] */
/* Scan engine.  POPC: 8 x popcount(xor) per pair on the integer pipe (the reference's DescriptorDistance, ORBmatcher.cc:2360-2378, as is).
 * TENSOR: the same distance as an exact +-1 int8 contraction, dist = (256 - dot) / 2, on the 5th-generation tensor cores (tcgen05.mma
 * kind::i8, accumulators in TMEM; compute capability 10.x only).  AUTO (default; environment EORB_HAMMING_ENGINE=0/1/2 overrides at
 * creation): TENSOR for at least 96 queries against at least 65536 rows, POPC otherwise.  Results are identical bit for bit. */
#define EORB_HAMMING_POPC   0
#define EORB_HAMMING_TENSOR 1
#define EORB_HAMMING_AUTO   2
int eorb_matcher_set_engine(eorb_matcher* m, int engine);
int eorb_matcher_last_engine(const eorb_matcher* m);   /* the engine the last search ran on */
int eorb_matcher_search(eorb_matcher* m, const uint8_t* q, int nq, int th, float ratio, eorb_match* out);
/* per-shard partials, device buffers, asynchronous: d_partial[nq] */
int eorb_matcher_search_device(eorb_matcher* m, const uint8_t* d_q, int nq, eorb_best2* d_partial);
/* merge nshards gathered partial arrays ([nshards][nq], e.g. the output of an NCCL all-gather) with the
 * (dist, global index) ordering, then threshold + ratio -> d_out[nq]; asynchronous */
int eorb_matcher_merge_device(eorb_matcher* m, const eorb_best2* d_gathered, int nshards, int nq,
                              int th, float ratio, eorb_match* d_out);
/* sharded search in one call (the brute-force shape of Frame.cc:1228-1235 over a database split by rows
 * across the GPUs of a box): scan of this rank's shard, ONE ncclAllGather of nq x 16 B per rank on the matcher's stream,
 * merge with the (dist, global index) ordering, threshold + ratio -> d_out[nq] on every rank; asynchronous.
 * nccl_comm is an ncclComm_t of nshards ranks.  libnccl.so.2 is resolved at run time (dlopen): no link-time dependency. */
int eorb_matcher_search_sharded(eorb_matcher* m, const uint8_t* d_q, int nq, int th, float ratio, void* nccl_comm,
                                int nshards, eorb_match* d_out);
/* thin NCCL helpers for host languages without NCCL bindings: ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy */
int eorb_nccl_unique_id(uint8_t* id128);
int eorb_nccl_comm_init_rank(void** comm, int nranks, const uint8_t* id128, int rank, int device);
int eorb_nccl_comm_destroy(void* comm);
/* convenience: one-call stateless form on host buffers (creates/destroys a matcher) */
int eorb_hamming_best2(const uint8_t* q, int nq, const uint8_t* db, int64_t ndb, int th, float ratio, eorb_match* out);
/* rotation-consistency filter: ORBmatcher.cc:784-794, 800-823 and ComputeThreeMaxima :2314-2355.
 * match12[i] = matched index or -1; entries outside the three dominant rotation bins are set to -1.
 * Returns the number of remaining matches.  Host code (the reference's is too, and it is O(matches)). */
int eorb_rotation_filter(const float* angle1, const float* angle2, int32_t* match12, int n1);

/* ---------------------------------------------------------------- event frames
 * replaces EORB_SLAM::EvImConverter::ev2im / ev2im_gauss / ev2mci_gg_f (include/Event/EventConversion.h:52-75,
 * src/Event/EventConversion.cc:173-448) and the normalisation that follows them
 * (normalizeImage :67-72; cv::normalize NORM_MINMAX at EvImBuilder.cpp:976,1055,1076,1140) */

#define EORB_EV_NEAREST 0   /* ev2im                       EventConversion.cc:173-212 */
#define EORB_EV_GAUSS   1   /* ev2im_gauss                 :215-269 */
#define EORB_EV_SE3     2   /* ev2mci_gg_f(Tcw, medDepth)  :279-360 */
#define EORB_EV_SE2     3   /* ev2mci_gg_f(params2D)       :362-448 */

#define EORB_NORM_NONE    0
#define EORB_NORM_RUNNING 1 /* normalizeImage(max,min) as inside ev2im* when normalized=true.  The reference's extremes are RUNNING
                             * extremes of every intermediate pixel value (resolveMinMaxVals after each tap, EventConversion.cc:30-38,
                             * 201, 257): with pol == 0 they equal the final frame's; with pol != 0 they depend on the event order,
                             * and the window is then replayed in order on the device (ev_ordered_kernel) to reproduce them */
#define EORB_NORM_MINMAX  2 /* cv::normalize(img,img,255,0,NORM_MINMAX,CV_8UC1) as in EvImBuilder */

typedef struct eorb_ev_params {
    int   mode;            /* EORB_EV_* */
    int   width, height;
    float sigma;           /* Gaussian splat sigma (ignored for NEAREST) */
    int   pol;             /* 1: subtract events with p == false */
    int   normalize;       /* EORB_NORM_* */
    float Tcw[16];         /* SE3: row-major 4x4 float pose (cv::Mat CV_32F) */
    float med_depth;       /* SE3 */
    float K[4];            /* Pinhole fx, fy, cx, cy (Pinhole.cpp:30-62) for SE3/SE2 */
    float se2[4];          /* SE2: omega, vx, vy, [scale] */
    int   se2_n;           /* 3 or 4 */
    int   cam_model;       /* 0: Pinhole; 1: KannalaBrandt8 (the camera of Examples/Event/EvMVSEC.yaml:50): K = fx, fy, cx, cy and kb = k1..k4,
                            * unproject / project as src/CameraModels/KannalaBrandt8.cpp:86-129, 163-190 */
    float kb[4];
} eorb_ev_params;

int eorb_ev_create(int device, int max_windows, int64_t max_events, int max_width, int max_height, eorb_evconv** out);
int eorb_ev_destroy(eorb_evconv* c);
int eorb_ev_set_stream(eorb_evconv* c, void* cuda_stream);
int eorb_ev_reset_stream(eorb_evconv* c);
int eorb_ev_synchronize(eorb_evconv* c);
long long eorb_ev_launch_count(const eorb_evconv* c);
/* one window, host buffers.  img_f32 (h*w floats) and img_u8 (h*w) may each be NULL.  minmax[2] optional.
 * Returns EORB_EMPTY (outputs zeroed) when n == 0 for the SE3/SE2 modes, like the reference. */
int eorb_ev_accumulate(eorb_evconv* c, const eorb_event* evs, int64_t n, const eorb_ev_params* p,
                       float* img_f32, uint8_t* img_u8, float* minmax);
/* nwin windows, device buffers, asynchronous.  Window i covers events [win_offsets[i], win_offsets[i+1]) of
 * d_evs (win_offsets is a HOST array of nwin+1 entries: the fixed-count slicing of
 * EvTrackManager::consumeEventsBegin, EvTrackManager.cpp:272-286).  All windows share p except the pose:
 * poses (HOST, nwin x 16 floats) may be NULL to use p->Tcw for every window.
 * d_img_f32 is [nwin][h][w]; d_img_u8 likewise (may be NULL when normalize == NONE). */
int eorb_ev_accumulate_batch_device(eorb_evconv* c, const eorb_event* d_evs, const int64_t* win_offsets, int nwin,
                                    const eorb_ev_params* p, const float* poses,
                                    float* d_img_f32, uint8_t* d_img_u8);

/* the same with HOST buffers (the fixed-count windows of one event packet, EvTrackManager.cpp:272-286, in one call): events and offsets on
 * the host, frames [nwin][h][w] back on the host; img_f32 / img_u8 may each be NULL (not both; img_u8 needs a normalisation mode).  One H2D
 * copy, the batch kernels, one D2H copy per output; synchronous. */
int eorb_ev_accumulate_batch(eorb_evconv* c, const eorb_event* evs, const int64_t* win_offsets, int nwin, const eorb_ev_params* p,
                             const float* poses, float* img_f32, uint8_t* img_u8);

/* contrast metric of event frames (second "next" row): EvImConverter::measureImageFocusLocal (what = 0),
 * measureImageFocusGlobal (1), imageMeanLocal (2) — src/Event/EventConversion.cc:79-162, 30x30 cells, cv::meanStdDev
 * per cell; avg != 0: average of the cell values (measureImageFocus), avg == 0: their median.  The reference picks the
 * motion-compensated candidate with the largest value (EvImBuilder.cpp:1206-1215); with the _device form the
 * candidates never leave the GPU.  focus_out is a host array. */
/* Jacobian of the contrast objective w.r.t. the window's SE3 motion: EvImConverter::ev2mci_gg_f_jac
 * (src/Event/EventConversion.cc:533-662, called once per optimiser iteration by src/Utils/MyOptimTypes.cpp:16).
 * Rt12 = rotation (row-major 3x3) then translation of the vertex estimate, double; K4 = fx, fy, cx, cy (Pinhole);
 * global_mean != 0: cv::mean of the product images, else the mean of their 30x30-cell means (`global` argument of the
 * reference).  jac6 (host) = d/d[wx wy wz vx vy vz].  EORB_EMPTY and a zero Jacobian when there are no events. */
int eorb_ev_mci_jac(eorb_evconv* c, const eorb_event* evs, int64_t n, int w, int hgt, float sigma, const double* Rt12, float med_depth,
                    const float* K4, int pol, int global_mean, double* jac6);
#define EORB_FOCUS_LOCAL_STD  0
#define EORB_FOCUS_GLOBAL_STD 1
#define EORB_FOCUS_LOCAL_MEAN 2
int eorb_ev_image_focus_device(eorb_evconv* c, const float* d_img_f32, int nwin, int w, int hgt, int what, int avg, float* focus_out);
int eorb_ev_image_focus(eorb_evconv* c, const float* img_f32, int w, int hgt, size_t stride_bytes, int what, int avg, float* focus_out);

/* ---------------------------------------------------------------- LK tracker (first "next" row)
 * replaces EORB_SLAM::ELK_Tracker::setRefImage / trackCurrImage (include/Event/KLT_Tracker.h,
 * src/Event/KLT_Tracker.cpp:22-91), i.e. cv::calcOpticalFlowPyrLK(ref, cur, refPts, pts, status, err,
 * Size(win, win), maxLevel, TermCriteria(COUNT+EPS, maxItr, eps) [, OPTFLOW_USE_INITIAL_FLOW]) on 8-bit frames.
 * Points are (x, y) float pairs.  The reference frame's pyramid and Scharr derivatives are built once in set_ref. */
typedef struct eorb_lk eorb_lk;
int eorb_lk_create(int device, int max_width, int max_height, int max_points, eorb_lk** out);
int eorb_lk_destroy(eorb_lk* h);
int eorb_lk_set_stream(eorb_lk* h, void* cuda_stream);
int eorb_lk_reset_stream(eorb_lk* h);
long long eorb_lk_launch_count(const eorb_lk* h);
/* EORB_EMPTY when the image or the point list is empty (the reference asserts, KLT_Tracker.cpp:24) */
int eorb_lk_set_ref(eorb_lk* h, const uint8_t* img, int w, int hgt, size_t stride, const float* pts_xy, int n, int win, int max_level);
int eorb_lk_set_ref_device(eorb_lk* h, const uint8_t* d_img, int w, int hgt, size_t stride, const float* pts_xy, int n, int win,
                           int max_level);
/* init_xy == NULL: start from the reference points; else OPTFLOW_USE_INITIAL_FLOW (KLT_Tracker.cpp:63-65, 86-88).
 * min_eig: minEigThreshold (OpenCV default 1e-4).  err may be NULL.  Returns the number of pyramid levels - 1 used
 * (OpenCV lowers maxLevel when a level would be no larger than the window). */
int eorb_lk_track(eorb_lk* h, const uint8_t* img, size_t stride, const float* init_xy, int max_iter, double eps, float min_eig,
                  float* out_xy, uint8_t* status, float* err);
int eorb_lk_track_device(eorb_lk* h, const uint8_t* d_img, size_t stride, const float* init_xy, int max_iter, double eps, float min_eig,
                         float* out_xy, uint8_t* status, float* err);

/* ELK_Tracker with its state in HBM: what follows every LK call in the reference (src/Event/KLT_Tracker.cpp)
 *   eorb_lk_set_ref_keypoints   replaces setRefImage(image, vector<KeyPoint>) :22-46 (mRefKPoints, mRefPoints, mLastTrackedPts = the points)
 *   eorb_lk_set_last_tracked    replaces setLastTrackedPts :252-262 (a list of another size switches the next call to "no initial flow", :63-70)
 *   eorb_lk_get_last_tracked    reads mLastTrackedPts back (n x 2 floats); returns n
 *   eorb_lk_track_and_match     replaces trackAndMatchCurrImage :215-234 (first_octave_only = 0) and trackAndMatchCurrImageInit :236-242
 *                               (= 1): LK from the resident last tracked points with OPTFLOW_USE_INITIAL_FLOW, then refineTrackedPts
 *                               :105-155 and refineFirstOctaveLevel :157-183 on the device.  The tracked points stay on the device as the
 *                               next call's initial flow; ONE device-to-host copy returns
 *        tracked[n]   KeyPoint(currPt, size / angle / response / octave / class_id of the reference keypoint)  (p1 / trackedKPts)
 *        matched[n]   bit 0: the point passed refineTrackedPts' test (status == 1 and inside the image, :140), i.e. the reference does
 *                     vMatches12[i] = i, vCntMatches[i]++, nMatches++ and pushes its displacement; bit 1: bit 0 and the match survives
 *                     refineFirstOctaveLevel (reference keypoint on octave 0, :171-175; equal to bit 0 without first_octave_only)
 *        px_disp[]    sqrtf(dx^2 + dy^2) of the points with bit 0, in index order (vPxDisp)
 *        counts2      {points with bit 1 = the reference's return value when the caller's vMatches12 came in empty, px_disp entries}
 *     The caller owns the vectors the reference updates in place; with vectors carried over from earlier calls the shim replays the two
 *     loops on them from bit 0 and the reference keypoints' octaves (stale entries on upper octaves are un-matched and counted, :166-176).
 *     Returns the last pyramid level used, EORB_EMPTY without reference keypoints (the reference logs and returns 0, :218-221).
 *   _device: image and outputs are device pointers, asynchronous on the tracker's stream. */
int eorb_lk_set_ref_keypoints(eorb_lk* h, const uint8_t* img, int w, int hgt, size_t stride, int img_on_device,
                              const eorb_keypoint* ref_kps, int n, int win, int max_level);
int eorb_lk_set_last_tracked(eorb_lk* h, const eorb_keypoint* kps, int n);
int eorb_lk_get_last_tracked(eorb_lk* h, float* pts_xy);
int eorb_lk_track_and_match(eorb_lk* h, const uint8_t* img, size_t stride, int img_on_device, int max_iter, double eps, float min_eig,
                            int first_octave_only, eorb_keypoint* tracked, uint8_t* matched, float* px_disp, int* counts2);
int eorb_lk_track_and_match_device(eorb_lk* h, const uint8_t* d_img, size_t stride, int max_iter, double eps, float min_eig,
                                   int first_octave_only, eorb_keypoint* d_tracked, uint8_t* d_matched, float* d_px_disp, int* d_counts2);

/* ---------------------------------------------------------------- guided matching (third "next" row)
 * The per-frame callers of DescriptorDistance in tracking:
 *   eorb_guided_frame_grid            replaces Frame::AssignFeaturesToGrid + Frame::PosInGrid (src/Frame.cc:431-460, 783-793)
 *   eorb_guided_features_in_area      replaces Frame::GetFeaturesInArea (src/Frame.cc:709-777), a batch of queries per call
 *   eorb_guided_search_for_initialization  replaces ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:714-831)
 * Keypoints are eorb_keypoint records (the undistorted keypoints the reference's grid holds), descriptors n x 32 bytes,
 * bounds4 = {mnMinX, mnMinY, mnMaxX, mnMaxY} (Frame.cc:846-866).  The grid is FRAME_GRID_COLS x FRAME_GRID_ROWS = 64 x 48
 * (include/Frame.h:45-46); cell id = col * 48 + row; lists hold keypoint indices in ascending order, like push_back does.
 * At most EORB_GUIDED_MAX_KEYPOINTS keypoints per frame (the reference extracts <= 5 * nFeatures). */
#define EORB_GUIDED_MAX_KEYPOINTS 16384
#define EORB_GRID_NCELLS (64 * 48)
typedef struct eorb_area_query {
    float x, y, r;          /* window centre and half size (factorX = factorY = r) */
    int min_level, max_level;   /* as in the reference: levels are checked iff min_level > 0 || max_level >= 0 */
} eorb_area_query;
typedef struct eorb_guided eorb_guided;
int eorb_guided_create(int device, eorb_guided** out);
int eorb_guided_destroy(eorb_guided* g);
int eorb_guided_set_stream(eorb_guided* g, void* cuda_stream);
int eorb_guided_reset_stream(eorb_guided* g);
long long eorb_guided_launch_count(const eorb_guided* g);
/* cell_start[EORB_GRID_NCELLS + 1], cell_idx[n] (host); *assigned = keypoints inside the grid */
int eorb_guided_frame_grid(eorb_guided* g, const eorb_keypoint* kps, int n, const float* bounds4, int* cell_start, int* cell_idx,
                           int* assigned);
/* counts[nq] = size of every answer; idx_out[nq * cap_per_query]: the first cap_per_query indices of every answer in the
 * reference's order (cells column by column, rows inside, list order inside a cell) */
int eorb_guided_features_in_area(eorb_guided* g, const eorb_keypoint* kps, int n, const float* bounds4, const eorb_area_query* queries,
                                 int nq, int* counts, int* idx_out, int cap_per_query);
/* prev_xy = vbPrevMatched (n1 x 2 floats, in/out), matches12[n1] = vnMatches12; returns nmatches in *nmatches.
 * window_size = the reference's int windowSize (100 at Tracking.cc's call), nnratio = mfNNratio, check_ori = mbCheckOrientation. */
int eorb_guided_search_for_initialization(eorb_guided* g, const eorb_keypoint* kps1, const uint8_t* desc1, int n1,
                                          const eorb_keypoint* kps2, const uint8_t* desc2, int n2, const float* bounds4, float* prev_xy,
                                          int window_size, float nnratio, int check_ori, int32_t* matches12, int* nmatches);
/* the same with every array resident in HBM (16-byte aligned descriptors), e.g. straight from eorb_orb_extract_batch_device;
 * d_prev_xy and d_matches12 are written on the device, *nmatches (host) after a stream synchronisation */
int eorb_guided_search_for_initialization_device(eorb_guided* g, const eorb_keypoint* d_kps1, const uint8_t* d_desc1, int n1,
                                                 const eorb_keypoint* d_kps2, const uint8_t* d_desc2, int n2, const float* bounds4,
                                                 float* d_prev_xy, int window_size, float nnratio, int check_ori, int32_t* d_matches12,
                                                 int* nmatches);

/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, float th, bool bMono = true)
 * (src/ORBmatcher.cc:1969-2150, monocular path): the tracking-rate caller of the matcher (Tracking::TrackWithMotionModel).
 * Per last-frame keypoint i: x3Dc = Rcw * x3Dw + tcw (3 floats, formed by the caller with its own cv::Mat arithmetic), valid1[i] =
 * LastFrame holds a map point at i that is not an outlier, obs1[i] = that point's Observations(), descMP = its descriptor,
 * kps1[i].octave / .angle = the last frame's keypoint.  kps2 / desc2 = the current frame (undistorted keypoints); K4 = fx, fy, cx,
 * cy of its Pinhole camera; scale_factors[nlevels] = mvScaleFactors; bounds4 = mnMinX, mnMinY, mnMaxX, mnMaxY.
 * match_cur[n2] = index of the last-frame keypoint whose map point ends up in CurrentFrame.mvpMapPoints[i2] (-1: none); the
 * return value of the reference goes to *nmatches.  (Stereo / fisheye second view and forward/backward level ranges: not on the
 * monocular path.) */
int eorb_guided_search_by_projection(eorb_guided* g, const float* x3Dc, const uint8_t* valid1, const int32_t* obs1,
                                     const eorb_keypoint* kps1, const uint8_t* descMP, int n1, const eorb_keypoint* kps2,
                                     const uint8_t* desc2, int n2, const float* bounds4, const float* K4, const float* scale_factors,
                                     int nlevels, float th, int check_ori, int32_t* match_cur, int* nmatches);
/* the same with every array resident in HBM; d_match_cur is written on the device, *nmatches after a stream synchronisation */
int eorb_guided_search_by_projection_device(eorb_guided* g, const float* d_x3Dc, const uint8_t* d_valid1, const int32_t* d_obs1,
                                            const eorb_keypoint* d_kps1, const uint8_t* d_descMP, int n1, const eorb_keypoint* d_kps2,
                                            const uint8_t* d_desc2, int n2, const float* bounds4, const float* K4,
                                            const float* scale_factors, int nlevels, float th, int check_ori, int32_t* d_match_cur,
                                            int* nmatches);
/* The same for a RECTIFIED-STEREO / RGB-D frame (Nleft == -1, mvuRight set; Tracking::TrackWithMotionModel with bMono = false):
 * level_mode = 1 when bForward, 2 when bBackward, 0 otherwise (tlc.z against CurrentFrame.mb, src/ORBmatcher.cc:1989-1990; the
 * level window becomes [octave, inf) / [0, octave], :2024-2029); u_right2 = CurrentFrame.mvuRight (n2 floats, may be NULL) and
 * mbf = CurrentFrame.mbf: a candidate with a right-image column is skipped when |(u - mbf / zc) - uRight| > radius (:2049-2055).
 * (The second pass over a fisheye rig's right camera, Nleft != -1, :2093-2160, stays with the reference body.) */
int eorb_guided_search_by_projection_stereo(eorb_guided* g, const float* x3Dc, const uint8_t* valid1, const int32_t* obs1,
                                            const eorb_keypoint* kps1, const uint8_t* descMP, int n1, const eorb_keypoint* kps2,
                                            const uint8_t* desc2, const float* u_right2, int n2, const float* bounds4, const float* K4,
                                            const float* scale_factors, int nlevels, float th, int check_ori, int level_mode, float mbf,
                                            int32_t* match_cur, int* nmatches);
int eorb_guided_search_by_projection_stereo_device(eorb_guided* g, const float* d_x3Dc, const uint8_t* d_valid1, const int32_t* d_obs1,
                                                   const eorb_keypoint* d_kps1, const uint8_t* d_descMP, int n1, const eorb_keypoint* d_kps2,
                                                   const uint8_t* d_desc2, const float* d_u_right2, int n2, const float* bounds4,
                                                   const float* K4, const float* scale_factors, int nlevels, float th, int check_ori,
                                                   int level_mode, float mbf, int32_t* d_match_cur, int* nmatches);

/* eorb_guided_search_by_projection_reloc replaces ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF,
 * const set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist) (src/ORBmatcher.cc:2189-2312; Tracking::Relocalization
 * after the PnP refinement).  Per keyframe map point i: x3Dc = Rcw * x3Dw + tcw, valid1[i] != 0 = the point exists, is not bad, is
 * not in sAlreadyFound and dist3D lies inside its distance invariance (:2205-2232); level1[i] = PredictScale(dist3D, &CurrentFrame)
 * (:2234; host arithmetic on the map point); kps1[i].angle = the keyframe's undistorted keypoint; descMP = GetDescriptor().
 * held2[i2] != 0 (may be NULL) = CurrentFrame.getMapPoint(i2) is set on entry.  No depth-sign test (:2218); window th *
 * scale[level], levels [level-1, level+1]; a slot that holds ANY point is skipped (:2253); accepted when the best distance is
 * <= orb_dist (:2266); rotation filter as in the other searches.  match_cur[i2] = keyframe index set into the frame, or -1. */
int eorb_guided_search_by_projection_reloc(eorb_guided* g, const float* x3Dc, const uint8_t* valid1, const int32_t* level1,
                                           const eorb_keypoint* kps1, const uint8_t* descMP, int n1, const eorb_keypoint* kps2,
                                           const uint8_t* desc2, const uint8_t* held2, int n2, const float* bounds4, const float* K4,
                                           const float* scale_factors, int nlevels, float th, int orb_dist, int check_ori,
                                           int32_t* match_cur, int* nmatches);
int eorb_guided_search_by_projection_reloc_device(eorb_guided* g, const float* d_x3Dc, const uint8_t* d_valid1, const int32_t* d_level1,
                                                  const eorb_keypoint* d_kps1, const uint8_t* d_descMP, int n1, const eorb_keypoint* d_kps2,
                                                  const uint8_t* d_desc2, const uint8_t* d_held2, int n2, const float* bounds4,
                                                  const float* K4, const float* scale_factors, int nlevels, float th, int orb_dist,
                                                  int check_ori, int32_t* d_match_cur, int* nmatches);

/* eorb_guided_search_by_projection_map_points replaces ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>&
 * vpMapPoints, const float th, const bool bFarPoints, const float thFarPoints) (src/ORBmatcher.cc:44-218) as called by
 * Tracking::SearchLocalPoints (src/Tracking-1.cc:2379) for a monocular frame (Nleft == -1, mvuRight < 0: the right-camera
 * blocks never run).  pts[i] = the MapPoint fields Frame::isInFrustum fills (mTrackProjX/Y, mTrackViewCos, mTrackDepth,
 * mnTrackScaleLevel, mbTrackInView) + Observations() + isBad(); descMP = GetDescriptor() (n1 x 32 bytes).  kps2 / desc2 = the
 * frame's undistorted keypoints and descriptors; held2[i2] != 0 (may be NULL) = F.getMapPoint(i2) holds a point with
 * observations on entry.  scale_factors = mvScaleFactors.  match_cur[i2] = index of the map point F.setMapPoint(i2, .)
 * received in this call (the last one when points without observations are overwritten), or -1. */
typedef struct eorb_track_point {
    float proj_x, proj_y, view_cos, depth;
    int32_t scale_level, observations;
    uint8_t in_view, bad, pad[2];
} eorb_track_point;
int eorb_guided_search_by_projection_map_points(eorb_guided* g, const eorb_track_point* pts, const uint8_t* descMP, int n1,
                                                const eorb_keypoint* kps2, const uint8_t* desc2, const uint8_t* held2, int n2,
                                                const float* bounds4, const float* scale_factors, int nlevels, float th, int far_points,
                                                float th_far, float nnratio, int32_t* match_cur, int* nmatches);
int eorb_guided_search_by_projection_map_points_device(eorb_guided* g, const eorb_track_point* d_pts, const uint8_t* d_descMP, int n1,
                                                       const eorb_keypoint* d_kps2, const uint8_t* d_desc2, const uint8_t* d_held2,
                                                       int n2, const float* bounds4, const float* scale_factors, int nlevels, float th,
                                                       int far_points, float th_far, float nnratio, int32_t* d_match_cur,
                                                       int* nmatches);
/* The same for a RECTIFIED-STEREO / RGB-D frame (Nleft == -1, mvuRight set): proj_xr[i] = mTrackProjXR of map point i, u_right2 =
 * F.mvuRight; a candidate with a right-image column is skipped when |mTrackProjXR - uRight| > r * mvScaleFactors[level]
 * (src/ORBmatcher.cc:91-96).  (The right-camera block of a fisheye rig, Nleft != -1, :149-216, stays with the reference body.) */
int eorb_guided_search_by_projection_map_points_stereo(eorb_guided* g, const eorb_track_point* pts, const float* proj_xr,
                                                       const uint8_t* descMP, int n1, const eorb_keypoint* kps2, const uint8_t* desc2,
                                                       const uint8_t* held2, const float* u_right2, int n2, const float* bounds4,
                                                       const float* scale_factors, int nlevels, float th, int far_points, float th_far,
                                                       float nnratio, int32_t* match_cur, int* nmatches);
int eorb_guided_search_by_projection_map_points_stereo_device(eorb_guided* g, const eorb_track_point* d_pts, const float* d_proj_xr,
                                                              const uint8_t* d_descMP, int n1, const eorb_keypoint* d_kps2,
                                                              const uint8_t* d_desc2, const uint8_t* d_held2, const float* d_u_right2,
                                                              int n2, const float* bounds4, const float* scale_factors, int nlevels,
                                                              float th, int far_points, float th_far, float nnratio,
                                                              int32_t* d_match_cur, int* nmatches);

/* eorb_guided_search_for_triangulation replaces ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12,
 * vector<pair<size_t, size_t>>& vMatchedPairs, bool bOnlyStereo, bool bCoarse) (src/ORBmatcher.cc:975-1214; LocalMapping::CreateNewMapPoints,
 * src/LocalMapping.cc:507, and Tracking's keyframe insertion, src/Tracking.cc:3212) for keyframes with one PINHOLE camera each (mpCamera2 ==
 * NULL; KannalaBrandt8::epipolarConstrain triangulates every candidate pair and is not taken over).  Keypoints / descriptors / FeatureVectors
 * of both keyframes as in eorb_guided_search_by_bow_kf.  flags{1,2}[i]: bit 0 = the feature takes part (GetMapPoint(i) == NULL, and
 * mvuRight[i] >= 0 when bOnlyStereo), bit 1 = mvuRight[i] >= 0 (bStereo).  F12[9] = the fundamental matrix Pinhole::epipolarConstrain forms,
 * K1.t().inv() * SkewSymmetricMatrix(t12) * R12 * K2.inv() (src/CameraModels/Pinhole.cpp:137-140; the argument F12 of the reference function
 * is not read by it), row-major, computed by the caller with the camera's own cv::Mat arithmetic; epipole2 = pKF2->mpCamera->project(R2w * Cw
 * + t2w) (:982-988); scale_factors2 / level_sigma2_2 = mvScaleFactors / mvLevelSigma2 of pKF2 (:1081, :1131).  match12[i1] = the feature
 * of pKF2 matched to feature i1 of pKF1 or -1 (vMatchedPairs = the pairs (i1, match12[i1]) in ascending i1, :1200-1208). */
int eorb_guided_search_for_triangulation(eorb_guided* g, const eorb_keypoint* kps1, const uint8_t* desc1, const uint8_t* flags1, int n1,
                                         const uint32_t* nodes1, const int32_t* start1, const uint32_t* feats1, int nn1,
                                         const eorb_keypoint* kps2, const uint8_t* desc2, const uint8_t* flags2, int n2,
                                         const uint32_t* nodes2, const int32_t* start2, const uint32_t* feats2, int nn2, const float* F12,
                                         const float* epipole2, const float* scale_factors2, const float* level_sigma2_2, int nlevels,
                                         int coarse, int check_ori, int32_t* match12, int* nmatches);

/* the same with every array resident in HBM (the FeatureVectors in the CSR form eorb_vocab_transform_resident leaves there); nentries1 =
 * start1[nn1] (the number of feature entries of the first FeatureVector); d_match12 is written on the device, *nmatches after a stream
 * synchronisation.  F12 / epipole2 / the level tables stay host arrays (they are per-call constants). */
int eorb_guided_search_for_triangulation_device(eorb_guided* g, const eorb_keypoint* d_kps1, const uint8_t* d_desc1, const uint8_t* d_flags1, int n1,
                                                const uint32_t* d_nodes1, const int32_t* d_start1, const uint32_t* d_feats1, int nn1, int nentries1,
                                                const eorb_keypoint* d_kps2, const uint8_t* d_desc2, const uint8_t* d_flags2, int n2,
                                                const uint32_t* d_nodes2, const int32_t* d_start2, const uint32_t* d_feats2, int nn2,
                                                const float* F12, const float* epipole2, const float* scale_factors2, const float* level_sigma2_2,
                                                int nlevels, int coarse, int check_ori, int32_t* d_match12, int* nmatches);

/* eorb_guided_search_windows: the matching core of the KEYFRAME-side searches of local mapping and loop closing, which all have one
 * shape -- per map point a window in a keyframe (GetFeaturesInArea(u, v, radius), src/KeyFrame.cc:873-917), keypoints of level
 * [nPredictedLevel - 1, nPredictedLevel], smallest descriptor distance, first visited among equals:
 *   ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th,
 *       float ratioHamming) (src/ORBmatcher.cc:480-593) and the overload with vpPointsKFs / vpMatchedKF (:595-712): blocking = 1 (a
 *       keypoint with vpMatched[idx] set -- on entry = held2, or by an earlier point of this call -- is skipped, :563, :678),
 *       th_high = (int)floor(TH_LOW * ratioHamming);
 *   ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, float th, bool bRight) (:1407-1617): blocking = 0,
 *       inv_level_sigma2 = mvInvLevelSigma2 (the reprojection gate :1532-1558: e2 * invSigma2[octave] > 5.99, or > 7.8 with the
 *       right-image term when u_right2[idx] >= 0, ur[i] = uv.x - bf * invz), th_high = TH_LOW;
 *   ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint)
 *       (:1619-1741): blocking = 0, no gate, th_high = TH_LOW;
 *   ORBmatcher::SearchBySim3 (:1743-1967): one call per direction, blocking = 0, th_high = TH_HIGH; the agreement test stays on the host.
 * queries[i] = the window of map point i as the reference forms it on the host from the MapPoint / KeyFrame accessors (projection
 * through the keyframe's camera, IsInImage, distance and viewing-angle gates, PredictScale): x, y = uv, r = th * scale[level],
 * min_level = level - 1, max_level = level; r <= 0 = the point failed a gate (no query).  descMP = GetDescriptor() per point.
 * kps2 / desc2 = the keyframe's undistorted keypoints and descriptors, held2 (may be NULL) = slots to skip.  bounds4 = the image bounds
 * the keyframe's grid was built with (the Frame's floats, src/KeyFrame.cc:73-110 copies F.mGrid); query_min_xy (may be NULL = bounds4's)
 * = (float)mnMinX, (float)mnMinY of the KeyFrame, which are INTS (include/KeyFrame.h:529: the window lookup subtracts the truncated
 * bounds, the two differ for distorted cameras).
 * best_idx[i] = the keypoint map point i matched (best candidate with distance <= th_high; blocking: that it claimed), else -1;
 * best_dist[i] (may be NULL) = non-blocking: distance of the best candidate whether accepted or not (256: no candidate); blocking:
 * distance of the claimed keypoint (256: no claim); match2[i2] (blocking only, may be NULL) = the point that claimed keypoint i2 or -1;
 * *nmatches = number of points with best_idx >= 0.  The map updates that follow (Replace / AddObservation / vpMatched) stay with the
 * caller, in the order of the points, as in the reference. */
int eorb_guided_search_windows(eorb_guided* g, const eorb_area_query* queries, const float* ur, const uint8_t* descMP, int n1,
                               const eorb_keypoint* kps2, const uint8_t* desc2, const uint8_t* held2, const float* u_right2, int n2,
                               const float* bounds4, const float* query_min_xy, const float* inv_level_sigma2, int nlevels, int blocking,
                               int th_high, int32_t* best_idx, int32_t* best_dist, int32_t* match2, int* nmatches);
/* the same with every array resident in HBM (results are written on the device, *nmatches after a stream synchronisation) */
int eorb_guided_search_windows_device(eorb_guided* g, const eorb_area_query* d_queries, const float* d_ur, const uint8_t* d_descMP, int n1,
                                      const eorb_keypoint* d_kps2, const uint8_t* d_desc2, const uint8_t* d_held2, const float* d_u_right2,
                                      int n2, const float* bounds4, const float* query_min_xy, const float* inv_level_sigma2, int nlevels,
                                      int blocking, int th_high, int32_t* d_best_idx, int32_t* d_best_dist, int32_t* d_match2, int* nmatches);

/* eorb_guided_search_by_bow replaces ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches)
 * (src/ORBmatcher.cc:276-478) as called by Tracking::TrackReferenceKeyFrame and Relocalization (src/Tracking-1.cc:1680, 2625)
 * for a monocular frame.  kps_kf / desc_kf = the keyframe's undistorted keypoints and descriptors, valid_kf[i] != 0 =
 * GetMapPointMatches()[i] is set and not bad; the FeatureVectors pKF->mFeatVec and F.mFeatVec arrive in the CSR form
 * eorb_vocab_transform writes (nodes ascending; features of node q = feats[start[q] .. start[q+1]) in push_back order).
 * match_f[i2] = keyframe feature index whose map point ends up in vpMapPointMatches[i2], or -1. */
int eorb_guided_search_by_bow(eorb_guided* g, const eorb_keypoint* kps_kf, const uint8_t* desc_kf, const uint8_t* valid_kf, int n1,
                              const uint32_t* kf_nodes, const int32_t* kf_start, const uint32_t* kf_feats, int nkf,
                              const eorb_keypoint* kps_f, const uint8_t* desc_f, int n2, const uint32_t* f_nodes, const int32_t* f_start,
                              const uint32_t* f_feats, int nf, float nnratio, int check_ori, int32_t* match_f, int* nmatches);
/* the same with every array resident in HBM; d_match_f is written on the device, *nmatches after a stream synchronisation */
int eorb_guided_search_by_bow_device(eorb_guided* g, const eorb_keypoint* d_kps_kf, const uint8_t* d_desc_kf, const uint8_t* d_valid_kf,
                                     int n1, const uint32_t* d_kf_nodes, const int32_t* d_kf_start, const uint32_t* d_kf_feats, int nkf,
                                     const eorb_keypoint* d_kps_f, const uint8_t* d_desc_f, int n2, const uint32_t* d_f_nodes,
                                     const int32_t* d_f_start, const uint32_t* d_f_feats, int nf, float nnratio, int check_ori,
                                     int32_t* d_match_f, int* nmatches);
/* eorb_guided_search_by_bow_kf replaces ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)
 * (src/ORBmatcher.cc:833-990; loop closing / place recognition), monocular keyframes.  Same node walk as the keyframe-frame form; the
 * candidates of the second keyframe need a good map point too (valid2, :889-893), the acceptance is bestDist1 < TH_LOW (strict, :912)
 * and the result is indexed by the FIRST keyframe's feature: match12[i1] = feature index of keyframe 2 whose map point goes to
 * vpMatches12[i1], or -1.  _device: every array resident, d_scratch_n2 = n2 ints of scratch. */
int eorb_guided_search_by_bow_kf(eorb_guided* g, const eorb_keypoint* kps1, const uint8_t* desc1, const uint8_t* valid1, int n1,
                                 const uint32_t* nodes1, const int32_t* start1, const uint32_t* feats1, int nn1, const eorb_keypoint* kps2,
                                 const uint8_t* desc2, const uint8_t* valid2, int n2, const uint32_t* nodes2, const int32_t* start2,
                                 const uint32_t* feats2, int nn2, float nnratio, int check_ori, int32_t* match12, int* nmatches);
int eorb_guided_search_by_bow_kf_device(eorb_guided* g, const eorb_keypoint* d_kps1, const uint8_t* d_desc1, const uint8_t* d_valid1, int n1,
                                        const uint32_t* d_nodes1, const int32_t* d_start1, const uint32_t* d_feats1, int nn1,
                                        const eorb_keypoint* d_kps2, const uint8_t* d_desc2, const uint8_t* d_valid2, int n2,
                                        const uint32_t* d_nodes2, const int32_t* d_start2, const uint32_t* d_feats2, int nn2, float nnratio,
                                        int check_ori, int32_t* d_match12, int32_t* d_scratch_n2, int* nmatches);

/* ---------------------------------------------------------------- bag of words + undistortion (fourth "next" row)
 * The two steps that follow extraction in the reference's Frame:
 *   eorb_vocab_transform        replaces DBoW2 TemplatedVocabulary<FORB::TDescriptor, FORB>::transform(features, BowVector,
 *                               FeatureVector, levelsup) (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1258) as called by
 *                               Frame::ComputeBoW / KeyFrame::ComputeBoW (src/Frame.cc:796-803, src/KeyFrame.cc:205-213, levelsup 4)
 *   eorb_undistort_keypoints    replaces Frame::UndistortKeyPoints (src/Frame.cc:805-840): cv::undistortPoints(K, dist, P = K)
 * A vocabulary is passed in the flat form TemplatedVocabulary::loadFromTextFile builds (:1330-1417; reading ORBvoc.txt stays
 * with the caller): node 0 = root, parent[nid] < nid (file order), is_leaf / 32-byte descriptor / weight per node; k, L,
 * scoring (L1_NORM 0, L2_NORM 1, CHI_SQUARE 2, KL 3, BHATTACHARYYA 4, DOT_PRODUCT 5) and weighting (TF_IDF 0, TF 1, IDF 2,
 * BINARY 3) are the file header's.  It stays resident in HBM.  At most EORB_BOW_MAX_FEATURES features per call. */
#define EORB_BOW_MAX_FEATURES 8192
typedef struct eorb_vocab eorb_vocab;
int eorb_vocab_create(int device, int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent, const uint8_t* is_leaf,
                      const uint8_t* desc, const double* weight, eorb_vocab** out);
int eorb_vocab_destroy(eorb_vocab* v);
int eorb_vocab_set_stream(eorb_vocab* v, void* cuda_stream);
int eorb_vocab_reset_stream(eorb_vocab* v);
long long eorb_vocab_launch_count(const eorb_vocab* v);
/* feats: n x 32 bytes.  BowVector = (bow_ids, bow_vals)[*nbow] in ascending word id (std::map order); FeatureVector in CSR form:
 * fv_nodes[*nfv] ascending, features of node q = fv_feats[fv_start[q] .. fv_start[q+1]) in ascending feature index (push_back
 * order).  Output arrays hold n entries (fv_start n + 1).  word_id / node_id (optional, n entries): the per-feature leaf word
 * and the node at level L - levelsup.  Doubles are bit-identical to DBoW2's (same summation order). */
int eorb_vocab_transform(eorb_vocab* v, const uint8_t* feats, int n, int levelsup, uint32_t* bow_ids, double* bow_vals, int* nbow,
                         uint32_t* fv_nodes, int32_t* fv_start, uint32_t* fv_feats, int* nfv, uint32_t* word_id, uint32_t* node_id);
/* the same with the descriptors resident in HBM (16-byte aligned), e.g. straight from eorb_orb_extract_batch_device */
int eorb_vocab_transform_device(eorb_vocab* v, const uint8_t* d_feats, int n, int levelsup, uint32_t* bow_ids, double* bow_vals, int* nbow,
                                uint32_t* fv_nodes, int32_t* fv_start, uint32_t* fv_feats, int* nfv, uint32_t* word_id, uint32_t* node_id);
/* the same with the results left in HBM too: BowVector and CSR FeatureVector are written to the caller's DEVICE arrays (n entries each,
 * d_fv_start n + 1; d_bow_ids / d_bow_vals may be NULL), only the counts *nbow / *nfv come back to the host.  For chaining
 * eorb_orb_extract_batch_device -> here -> eorb_guided_search_by_bow_device without a host round trip of the vectors. */
int eorb_vocab_transform_resident(eorb_vocab* v, const uint8_t* d_feats, int n, int levelsup, uint32_t* d_bow_ids, double* d_bow_vals,
                                  int* nbow, uint32_t* d_fv_nodes, int32_t* d_fv_start, uint32_t* d_fv_feats, int* nfv);
/* K4 = fx, fy, cx, cy; dist5 = k1, k2, p1, p2, k3 (mDistCoef).  Positions are replaced, the other fields copied
 * (Frame.cc:833-838); k1 == 0 copies the keypoints unchanged (:807-811).  out may alias kps. */
int eorb_undistort_keypoints(const eorb_keypoint* kps, int n, const float* K4, const float* dist5, eorb_keypoint* out);
int eorb_undistort_keypoints_device(const eorb_keypoint* d_kps, eorb_keypoint* d_out, int n, const float* K4, const float* dist5,
                                    void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* EORB_B200_H */
