/*
 * oracle/oracle.h — C API of the CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * This library is a from-scratch CPU restatement of the reference's front-end hot path
 * (src/ORBextractor.cc, src/ORBmatcher.cc:2360-2378 + best-2 loops, src/Event/EventConversion.cc)
 * including the OpenCV primitive semantics those files rely on (resize/INTER_LINEAR u8, REFLECT_101,
 * FAST-9/16 + NMS, fastAtan2, GaussianBlur 5x5 u8, cvRound, convertTo/normalize).
 *
 * It is the checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg.  NOTHING in the product (eorb_slam_b200/, include/) links, imports or calls it.
 *
 * Parity pin: the reference ships no tests/golden vectors and cannot be compiled here
 * (needs OpenCV 3.4.1 C++ headers, Eigen, glog ...).  Every OpenCV primitive restated here is pinned
 * bit-exactly against the Python cv2 4.13.0 wheel by tests/golden/make_golden.py (fixtures committed in
 * tests/golden/ as .npz files); the full-pipeline goldens come from an independent cv2-assisted restatement in
 * that same script.  Where the reference itself is not a function of its inputs (octree size ties by
 * heap pointer, FMA contraction, OOB descriptor taps for margin<19, float summation order) the pin
 * chosen is documented at the function.
 */
#ifndef EORB_ORACLE_H
#define EORB_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int   nfeatures;
    float scaleFactor;
    int   nlevels;
    int   iniThFAST;
    int   minThFAST;
    int   edgeTh;      /* <0: adaptive 19*W/752 forced odd (ORBextractor.cc:481-485) */
    int   imW, imH;    /* only used for the adaptive edge threshold */
} orc_orb_params;

/* mirrors cv::KeyPoint (28 bytes) */
typedef struct {
    float x, y, size, angle, response;
    int   octave, class_id;
} orc_keypoint;

/* mirrors EORB_SLAM::EventData (24 bytes, include/Event/EventData.h:36-58) */
typedef struct {
    double  ts;
    float   x, y;
    uint8_t p;
    uint8_t _pad[7];
} orc_event;

typedef struct {
    int32_t best_dist;   /* 256 if no candidate */
    int32_t best_idx;    /* -1 if none */
    int32_t second_dist; /* 256 if fewer than two candidates */
    int32_t accepted;    /* best<=th && best < ratio*second */
} orc_match;

typedef struct orc_orb orc_orb;

/* ---- primitives (each pinned against cv2 in tests) ---- */
int   orc_cv_round_f(float v);
int   orc_cv_round_d(double v);
void  orc_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride,
                           uint8_t* dst, int dw, int dh, size_t dstride);
void  orc_copy_make_border_reflect101(const uint8_t* src, int w, int h, size_t sstride,
                                      uint8_t* dst, int border, size_t dstride);
void  orc_gauss5x5_s2_u8(const uint8_t* src, int w, int h, size_t sstride, uint8_t* dst, size_t dstride);
/* FAST-9/16 on a ROI; returns count, writes up to cap (x,y,score) in row-major order */
int   orc_fast9_16(const uint8_t* img, int w, int h, size_t stride, int threshold, int nms,
                   int* xs, int* ys, int* scores, int cap);
float orc_fast_atan2(float y, float x);
/* DistributeOctTree on keys relative to (minX,minY); returns count; out_idx = index into input */
int   orc_distribute_octtree(const float* kx, const float* ky, const float* kresp, int n,
                             int minX, int maxX, int minY, int maxY, int N, int* out_idx, int cap);

/* ---- ORB extractor ---- */
orc_orb* orc_orb_create(const orc_orb_params* p);
void  orc_orb_destroy(orc_orb* h);
int   orc_orb_edge_threshold(const orc_orb* h);
int   orc_orb_features_per_level(const orc_orb* h, int* out /*nlevels*/);
int   orc_orb_scale_factors(const orc_orb* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2);
int   orc_orb_umax(const orc_orb* h, int* out16);
/* returns monoIndex (or -1 on empty image); *n_out = number of keypoints; desc may be NULL */
int   orc_orb_extract(orc_orb* h, const uint8_t* img, int w, int hgt, size_t stride,
                      int lap0, int lap1, int want_desc,
                      orc_keypoint* kps, uint8_t* desc, int cap, int* n_out);
/* stage taps of the LAST extract call */
int   orc_orb_level_size(const orc_orb* h, int level, int* w, int* hgt);
int   orc_orb_get_level(const orc_orb* h, int level, uint8_t* dst, size_t dstride);   /* unbordered */
int   orc_orb_get_blurred(const orc_orb* h, int level, uint8_t* dst, size_t dstride); /* valid if level had kps */
int   orc_orb_num_candidates(const orc_orb* h, int level);
int   orc_orb_get_candidates(const orc_orb* h, int level, int* xs, int* ys, int* scores, int cap);
int   orc_orb_num_level_kps(const orc_orb* h, int level);
int   orc_orb_get_level_kps(const orc_orb* h, int level, int* xs, int* ys, int* scores, float* angles, int cap);
int   orc_orb_num_fallback_cells(const orc_orb* h);
/* thread-pool batch (CPU baseline): frames are w*h contiguous; returns total keypoints */
long  orc_orb_extract_batch_mt(const orc_orb_params* p, const uint8_t* imgs, int nframes, int w, int hgt,
                               int nthreads, int want_desc, int* n_per_frame);
/* descriptors for given keypoints (ComputeTrackedKPtsDesc / AssignKPtLevelByBestDesc) */
int   orc_orb_tracked_desc(orc_orb* h, const uint8_t* img, int w, int hgt, size_t stride,
                           const orc_keypoint* kps, int n, uint8_t* desc);
int   orc_orb_assign_level_by_best_desc(orc_orb* h, const uint8_t* ref_desc, const uint8_t* img, int w, int hgt,
                                        size_t stride, orc_keypoint* kps, int n);

/* ---- matcher ---- */
int   orc_descriptor_distance(const uint8_t* a, const uint8_t* b);
void  orc_hamming_best2(const uint8_t* q, int nq, const uint8_t* db, int64_t ndb,
                        int th, float ratio, int ratio_mode, orc_match* out, int nthreads);
/* rotation-consistency filter (ORBmatcher.cc:784-823, 2314-2355). match12[i] = idx or -1; returns #kept */
int   orc_rotation_filter(const float* angle1, const float* angle2, int32_t* match12, int n1);

/* ---- guided matching (guided_oracle.cc) ---- */
/* Frame::AssignFeaturesToGrid + PosInGrid (Frame.cc:431-460, 783-793): 64 x 48 cells, cell id = col*48 + row; CSR lists
   (cell_start[3073], cell_idx[n]) with keypoint indices ascending inside a cell.  bounds4 = mnMinX, mnMinY, mnMaxX, mnMaxY.
   Returns the number of keypoints that fell inside the grid. */
int   orc_frame_grid(const orc_keypoint* kps, int n, const float* bounds4, int* cell_start, int* cell_idx);
/* Frame::GetFeaturesInArea (Frame.cc:709-777); returns the count, writes up to cap indices in the reference's order */
int   orc_features_in_area(const orc_keypoint* kps, int n, const float* bounds4, const int* cell_start, const int* cell_idx, float x,
                           float y, float r, int min_level, int max_level, int* out, int cap);
/* ORBmatcher::SearchForInitialization (ORBmatcher.cc:714-831); prev_xy = vbPrevMatched (n1 x 2, in/out); returns nmatches */
int   orc_search_for_initialization(const orc_keypoint* kps1, const uint8_t* desc1, int n1, const orc_keypoint* kps2,
                                    const uint8_t* desc2, int n2, const float* bounds4, float* prev_xy, int window_size, float nnratio,
                                    int check_ori, int32_t* matches12);

/* ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono = true) (ORBmatcher.cc:1969-2150); x3Dc = Rcw*x3Dw+tcw per
   last-frame keypoint (n1 x 3, formed by the caller), valid1 = has a non-outlier map point, obs1 = its Observations(),
   descMP = its descriptor; match_cur[n2] = last-frame index assigned to every current-frame keypoint or -1; returns nmatches */
int   orc_search_by_projection(const float* x3Dc, const uint8_t* valid1, const int32_t* obs1, const orc_keypoint* kps1,
                               const uint8_t* descMP, int n1, const orc_keypoint* kps2, const uint8_t* desc2, int n2,
                               const float* bounds4, const float* K4, const float* scale_factors, int nlevels, float th, int check_ori,
                               int32_t* match_cur);

/* the same with the rectified-stereo branches (:1989-1990, 2024-2029, 2049-2055): level_mode 0 = [oct-1, oct+1], 1 = bForward [oct, inf),
   2 = bBackward [0, oct]; u_right2 = CurrentFrame.mvuRight (n2 floats or NULL), mbf = CurrentFrame.mbf */
int   orc_search_by_projection_ex(const float* x3Dc, const uint8_t* valid1, const int32_t* obs1, const orc_keypoint* kps1,
                                  const uint8_t* descMP, int n1, const orc_keypoint* kps2, const uint8_t* desc2, int n2,
                                  const float* bounds4, const float* K4, const float* scale_factors, int nlevels, float th, int check_ori,
                                  int level_mode, float mbf, const float* u_right2, int32_t* match_cur);
/* ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:2189-2312; Tracking::Relocalization):
   valid1 = keyframe map point present, not bad, not already found, inside its distance invariance; level1 = its predicted scale level;
   held2 (may be null) = slots of the frame that hold any map point on entry */
int   orc_search_by_projection_reloc(const float* x3Dc, const uint8_t* valid1, const int32_t* level1, const orc_keypoint* kps1,
                                     const uint8_t* descMP, int n1, const orc_keypoint* kps2, const uint8_t* desc2, const uint8_t* held2, int n2,
                                     const float* bounds4, const float* K4, const float* scale_factors, int nlevels, float th, int orb_dist,
                                     int check_ori, int32_t* match_cur);

/* matching core of the keyframe-side searches (ORBmatcher.cc:480-712 SearchByProjection(KeyFrame*, Scw, ...), :1407-1741 Fuse,
   :1743-1967 SearchBySim3): the caller's window per map point, best candidate of level [min_level, max_level], optional blocking
   (vpMatched) and optional reprojection gate of Fuse; returns the number of accepted points */
/* ORBmatcher::SearchForTriangulation (ORBmatcher.cc:975-1214), pinhole keyframes; flags bit 0 = takes part, bit 1 = bStereo; F = the
   fundamental matrix of Pinhole::epipolarConstrain (row-major), ep = epipole in image 2; returns nmatches, match12[n1] */
int   orc_search_for_triangulation(const orc_keypoint* kps1, const uint8_t* desc1, const uint8_t* flags1, int n1, const uint32_t* nodes1,
                                   const int32_t* start1, const uint32_t* feats1, int nn1, const orc_keypoint* kps2, const uint8_t* desc2,
                                   const uint8_t* flags2, int n2, const uint32_t* nodes2, const int32_t* start2, const uint32_t* feats2, int nn2,
                                   const float* F, const float* ep, const float* scale2, const float* sigma2, int nlevels, int coarse,
                                   int check_ori, int32_t* match12);
typedef struct orc_area_query { float x, y, r; int32_t min_level, max_level; } orc_area_query;
int   orc_search_windows(const orc_area_query* queries, const float* ur, const uint8_t* descMP, int n1, const orc_keypoint* kps2,
                         const uint8_t* desc2, const uint8_t* held2, const float* u_right2, int n2, const float* bounds4,
                         const float* query_min_xy, const float* inv_level_sigma2, int nlevels, int blocking, int th_high, int32_t* best_idx, int32_t* best_dist,
                         int32_t* match2);

/* ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints) (ORBmatcher.cc:44-218), monocular frame.
   orc_track_point = the MapPoint fields Frame::isInFrustum fills (mTrackProjX/Y, mTrackViewCos, mTrackDepth, mnTrackScaleLevel,
   mbTrackInView) + Observations() + isBad(); held2 (may be null) = slots of F that hold a point with observations on entry;
   match_cur[n2] = index of the map point put at every keypoint of F by this call, or -1; returns nmatches */
typedef struct orc_track_point {
    float proj_x, proj_y, view_cos, depth;
    int32_t scale_level, observations;
    uint8_t in_view, bad, pad[2];
} orc_track_point;
int   orc_search_by_projection_map_points(const orc_track_point* pts, const uint8_t* descMP, int n1, const orc_keypoint* kps2,
                                          const uint8_t* desc2, const uint8_t* held2, int n2, const float* bounds4,
                                          const float* scale_factors, int nlevels, float th, int far_points, float th_far,
                                          float nnratio, int32_t* match_cur);

/* the same with the rectified-stereo test (:91-96): proj_xr = mTrackProjXR per map point, u_right2 = F.mvuRight (either may be NULL) */
int   orc_search_by_projection_map_points_ex(const orc_track_point* pts, const float* proj_xr, const uint8_t* descMP, int n1,
                                             const orc_keypoint* kps2, const uint8_t* desc2, const uint8_t* held2, const float* u_right2, int n2,
                                             const float* bounds4, const float* scale_factors, int nlevels, float th, int far_points,
                                             float th_far, float nnratio, int32_t* match_cur);

/* ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) (ORBmatcher.cc:276-478), monocular; FeatureVectors in the CSR form of
   orc_vocab_transform; valid_kf = keyframe map point present and not bad; match_f[n2] = keyframe feature index or -1 */
int   orc_search_by_bow(const orc_keypoint* kps_kf, const uint8_t* desc_kf, const uint8_t* valid_kf, const uint32_t* kf_nodes,
                        const int32_t* kf_start, const uint32_t* kf_feats, int nkf, const orc_keypoint* kps_f, const uint8_t* desc_f,
                        int n2, const uint32_t* f_nodes, const int32_t* f_start, const uint32_t* f_feats, int nf, float nnratio,
                        int check_ori, int32_t* match_f);

/* ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) (ORBmatcher.cc:833-990), monocular keyframes; match12[n1] = feature index of keyframe 2 or -1 */
int   orc_search_by_bow_kf(const orc_keypoint* kps1, const uint8_t* desc1, const uint8_t* valid1, const uint32_t* nodes1, const int32_t* start1,
                           const uint32_t* feats1, int nn1, int n1, const orc_keypoint* kps2, const uint8_t* desc2, const uint8_t* valid2,
                           const uint32_t* nodes2, const int32_t* start2, const uint32_t* feats2, int nn2, int n2, float nnratio, int check_ori,
                           int32_t* match12);

/* ---- bag of words + undistortion (bow_oracle.cc) ---- */
typedef struct orc_vocab orc_vocab;
/* flat form of what TemplatedVocabulary::loadFromTextFile builds: node 0 = root, parent[nid] < nid, children in id order,
   words numbered in order of appearance; scoring / weighting = DBoW2's enums (L1_NORM 0 ... DOT_PRODUCT 5; TF_IDF 0 ... BINARY 3) */
orc_vocab* orc_vocab_create(int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent, const uint8_t* is_leaf,
                            const uint8_t* desc, const double* weight);
void  orc_vocab_destroy(orc_vocab* v);
/* transform(features, BowVector, FeatureVector, levelsup): per-feature (word, weight, node at level L-levelsup), the
   BowVector as sorted (id, value) arrays and the FeatureVector as CSR (fv_nodes[nfv], fv_start[nfv+1], fv_feats) */
int   orc_vocab_transform(const orc_vocab* v, const uint8_t* feats, int n, int levelsup, uint32_t* word_id, double* word_w,
                          uint32_t* node_id, uint32_t* bow_ids, double* bow_vals, int* nbow, uint32_t* fv_nodes, int32_t* fv_start,
                          uint32_t* fv_feats, int* nfv);
/* cv::undistortPoints(src, dst, K, dist, noArray(), K) (Frame::UndistortKeyPoints, Frame.cc:805-840) */
void  orc_undistort_points(const float* xy, int n, const float* K4, const float* dist5, float* out_xy);

/* ---- event frames ---- */
/* mode: 0 nearest (ev2im), 1 gauss (ev2im_gauss), 2 gauss+SE3 (Tcw16,depth,K4), 3 gauss+SE2 (se2[4],K4) */
int   orc_ev_accumulate(const orc_event* evs, int64_t n, int w, int h, float sigma, int mode,
                        const float* Tcw16, float depth, const float* K4, const float* se2, int se2_n,
                        int pol, int normalize, float* img_f32, float* minmax2);
/* the same with a camera model: cam_model 0 = Pinhole (cam8[0..3] = fx, fy, cx, cy), 1 = KannalaBrandt8 (cam8[4..7] = k1..k4;
   src/CameraModels/KannalaBrandt8.cpp:86-129, 163-190) */
int   orc_ev_accumulate_cam(const orc_event* evs, int64_t n, int w, int h, float sigma, int mode,
                            const float* Tcw16, float depth, const float* cam8, int cam_model, const float* se2, int se2_n,
                            int pol, int normalize, float* img_f32, float* minmax2);
/* normalizeImage(convertTo) : u8 = sat(rint(v*alpha+beta)), alpha=255/(max-min), beta=-min*alpha */
void  orc_normalize_convert_u8(const float* img, int n, float maxVal, float minVal, uint8_t* out);
/* cv::normalize(img,out,255,0,NORM_MINMAX,CV_8UC1) */
void  orc_normalize_minmax_u8(const float* img, int n, uint8_t* out);
/* measureImageFocusLocal / Global, imageMeanLocal (EventConversion.cc:79-162); see event_oracle.cc */
float orc_image_focus(const float* img, int w, int h, int patch, int what, int avg);
/* ev2mci_gg_f_jac (EventConversion.cc:533-662); Rt12 = row-major R (9) then t (3), double */
void  orc_ev_mci_jac(const orc_event* evs, int64_t n, int w, int h, float sigma, const double* Rt12, float medDepth, const float* K4,
                     int pol, int global, double* jac6);

#ifdef __cplusplus
}
#endif
#endif
