/*
 * orb_oracle.h -- CPU ORACLE for the ORB front-end + Hamming matching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain, scalar, single-threaded restatement of the
 * reference's CPU algorithm (AaltoML/SLAM-module), used as the checker in tests/, in
 * __graft_entry__.smoke() and as the `cpu_baseline` / `--impl reference` leg of bench.py.
 * Nothing in the product path (slam-module_b200/, include/) may include, link or call it.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - PINNED against the reference's own sources compiled verbatim (oracle/_ref/libref_slam.so: static_settings.cpp,
 *     feature_search.cpp, orb_extractor.cpp, image_pyramid.cpp, feature_detector.cpp, keyframe_matcher.cpp,
 *     map_point.cpp, keyframe.cpp, bow_index.cpp, and the four openvslam/ header leaves; oracle/Makefile target `ref`,
 *     oracle/ref_slam.cpp), live and through tests/golden/golden_ref_slam.npz: scale factors and budgets, FeatureSearch
 *     order and radius queries, the chained pyramid, detectAndExtract end to end (border filter, tracker branch,
 *     ic_angle, rBRIEF, output order), matchForLoopClosures, matchForTriangulationDBoW, searchByProjection,
 *     replaceDuplication, matchMapPointsSim3, MapPoint::updateDescriptor, BowIndex (tests/test_reference_parity.py).
 *   - OpenCV arithmetic (resize, GaussianBlur, fastAtan2, FAST-9/16, cvRound): restated from OpenCV, pinned against
 *     cv2 4.13.0 outputs generated in the build container (tests/golden/golden_cv2.npz, tools/gen_golden.py).
 *   - PARITY UNPINNED: the FAST cell grid + quadtree distribution -- the reference delegates that stage to an absent
 *     parent-project class (feature_detector.cpp:89-98); the oracle follows the upstream OpenVSLAM scheme named by the
 *     north star with a documented tie-break and IS the specification.  DBoW2 itself (absent): its published tree descent,
 *     addWeight / normalize and L1Scoring::score are restated.  The summation order of Eigen's 3-term sums in the
 *     epipolar test is recalled (Eigen is not in the tree).
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LEVELS 16

typedef struct orc_params {
    int width;           /* level-0 image width  */
    int height;          /* level-0 image height */
    int levels;          /* slam.orbScaleLevels   (static_settings.cpp:31) */
    float scale_factor;  /* slam.orbScaleFactor   (static_settings.cpp:32) */
    int max_keypoints;   /* slam.maxKeypoints     (static_settings.cpp:48) */
    int ini_fast_thr;    /* upstream OpenVSLAM ini_fast_thr_ (20) */
    int min_fast_thr;    /* upstream OpenVSLAM min_fast_thr  (7)  */
} orc_params;

/* static_settings.cpp:9-15 (float products) and image_pyramid.cpp:77-78 (std::round on double). */
void orc_level_geometry(const orc_params *p, float *scales, int *widths, int *heights);
/* static_settings.cpp:39-60 */
void orc_level_budgets(const orc_params *p, int *budgets);

/* cv::resize(..., INTER_LINEAR) on 8-bit single channel (image_pyramid.cpp:79). */
void orc_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride,
                          uint8_t *dst, int dw, int dh, int dstride);
/* cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on 8-bit (image_pyramid.cpp:84). */
void orc_gaussian7_u8(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride);

/* CpuImagePyramid::update (image_pyramid.cpp:68-86).  Planes are written tightly packed
 * (stride == width) one after another, level 0 first; total = sum(w*h). */
void orc_pyramid(const orc_params *p, const uint8_t *img, int stride, uint8_t *pyr, uint8_t *blur);

/* cv::FAST(img, thr, nonmaxSuppression=true), TYPE_9_16.  Returns the number of keypoints,
 * row-major order; xs/ys/resp may be NULL to count only.  cap bounds the output arrays. */
int orc_cv_fast(const uint8_t *img, int w, int h, int stride, int thr,
                int *xs, int *ys, int *resp, int cap);
/* FAST-9/16 response (cornerScore) of one pixel: max_arc9 min |p-v| - 1 ( <0: never a corner ). */
int orc_fast_response(const uint8_t *center, int stride);

/* Upstream OpenVSLAM compute_fast_keypoints for one level: cell grid (64 px, overlap 6) over the
 * 19-px-inset working area, FAST ini->min threshold per cell, quadtree distribution to `budget`.
 * Output: level coordinates (integers), in the order of the final node list.
 * cand_*: optional dump of the candidates handed to the quadtree (working-area coordinates, in
 * cell-major / row-major order) for debugging the GPU stages; n_cand receives their number. */
int orc_detect_level(const uint8_t *img, int w, int h, int stride, int budget,
                     int ini_thr, int min_thr, int *xs, int *ys, int *resp, int cap,
                     int *cand_x, int *cand_y, int *cand_resp, int cand_cap, int *n_cand);

/* distribute_keypoints_via_tree on an explicit candidate list (working-area coords). */
int orc_distribute(const int *cx, const int *cy, const int *cresp, int n,
                   int area_w, int area_h, int budget, int *out_idx, int cap);

/* cv::fastAtan2 scalar path (degrees, [0,360)). */
float orc_fast_atan2(float y, float x);
/* ic_angle (orb_extractor.cpp:245-275). */
float orc_ic_angle(const uint8_t *img, int stride, int x, int y);
void orc_ic_moments(const uint8_t *img, int stride, int x, int y, int *m10, int *m01);
/* util::cos / util::sin (openvslam/trigonometric.h:17-46). */
float orc_util_cos(float v);
float orc_util_sin(float v);
/* compute_orb_descriptor, non-SSE branch (orb_extractor.cpp:284-352). */
void orc_descriptor(const uint8_t *blurred, int stride, int x, int y, float angle_deg, uint32_t *desc8);
/* u_max_ table (orb_extractor.cpp:174-186), 16 ints. */
void orc_umax(int *umax16);
/* rBRIEF pattern as floats, 1024 entries (orb_point_pairs.h:47). */
void orc_pattern(float *out1024);

/* compute_descriptor_distance_32 (openvslam/match_base.h:18-39). */
unsigned orc_hamming(const uint32_t *a, const uint32_t *b);

/* Full OrbExtractor::detectAndExtract (orb_extractor.cpp:73-164) for one frame, camera = all pixels
 * valid.  tracks (full-res x,y pairs; may be NULL/0) are described at level `track_level`.
 * Outputs (caller-allocated, capacity cap): x,y full-res float; angle deg; octave; desc 8 x u32;
 * track_id (-1 for detected).  lvl_x/lvl_y: integer level coordinates (for detected keypoints; for
 * tracker points the rounded level coords).  Returns the number of keypoints; level_counts[levels]
 * gets the detected count per level (tracker points excluded). */
int orc_extract(const orc_params *p, const uint8_t *img, int stride,
                const float *track_xy, const int *track_ids, int n_tracks, int track_level,
                float *x, float *y, float *angle, int *octave, uint32_t *desc, int *track_id,
                int *lvl_x, int *lvl_y, int cap, int *level_counts);

/* Brute-force degenerate case (single BoW node containing every feature, every feature owning a
 * TRIANGULATED map point) of matchForLoopClosures (keyframe_matcher.cpp:50-158).
 * matches[nA] receives idx in B or -1; returns num_matches.  ratio_is_double selects the promotion
 * used for the Lowe test (the parameter's C type lives in an absent header). */
unsigned orc_match_bruteforce(const uint32_t *descA, const float *angA, int nA,
                              const uint32_t *descB, const float *angB, int nB,
                              float ratio, unsigned thr, int check_orientation, int ratio_is_double,
                              int *matches);
/* angle_checker<int> (openvslam/match_angle_checker.h:61-134): bin index of a delta angle. */
int orc_angle_bin(float delta_angle);
/* angle_checker::get_invalid_matches on (delta, id) pairs; returns count, ids in reference order. */
int orc_angle_invalid(const float *deltas, const int *ids, int n, int *invalid_out);
/* libstdc++ std::sort order of the 30 histogram bins by size (descending), as the reference calls it. */
void orc_bin_order(const unsigned *sizes30, unsigned *order30);
/* std::partial_sort(first, last, last) order of the same bins (std::sort's heap-sort branch). */
void orc_bin_order_heap(const unsigned *sizes30, unsigned *order30);

/* Threaded drivers for the CPU baseline (frames / keyframe pairs sharded over host threads; the
 * reference itself is single-threaded).  Return wall seconds. */
/* matchForLoopClosures with the DBoW2 node buckets of the reference (keyframe_matcher.cpp:50-158): features are
 * compared only under the same vocabulary node; node[i] < 0 = in no node; elig = the caller's map-point filters. */
unsigned orc_match_bow(const uint32_t *descA, const float *angA, const int *nodeA, const unsigned char *eligA, int nA,
                       const uint32_t *descB, const float *angB, const int *nodeB, const unsigned char *eligB, int nB,
                       float ratio, unsigned thr, int check_orientation, int ratio_is_double, int *matches);

/* matchForTriangulationDBoW (keyframe_matcher.cpp:160-293) and its epipolar test (:23-44, E row-major). */
unsigned orc_match_triangulation(const uint32_t *dA, const float *aA, const int *octA, const double *bearA, const int *nodeA,
                                 const unsigned char *eligA, int nA, const uint32_t *dB, const float *aB, const double *bearB,
                                 const int *nodeB, const unsigned char *eligB, int nB, const double *E,
                                 const float *scale_factors, float residual_deg_thr, unsigned thr, int check_orientation,
                                 int *matches);
int orc_check_epipolar(const double *b1, const double *b2, const double *E, float scale, float residual_deg_thr);

/* ---- "next" rows ---------------------------------------------------------------- */
/* MapPoint::updateDescriptor (map_point.cpp:75-116): medoid index of every descriptor segment. */
void orc_medoid(const uint32_t *desc, const long long *offsets, int n_seg, int *best);
/* FeatureSearch (feature_search.cpp:22-48): the Y-sorted order and one radius query. */
void orc_feature_index(const float *x, const float *y, int n, int *order);
int orc_features_around(const float *x, const float *y, int n, float qx, float qy, float r, int *out);
/* Candidate loops of searchByProjection / replaceDuplication / findMatchesTranformedMps
 * (keyframe_matcher.cpp:356-386, 482-499, 604-627); see oracle/src/search.cpp. */
int orc_search_candidates(const float *kx, const float *ky, const int *koct, const uint32_t *kdesc, int nK,
                          unsigned char *taken, const float *qx, const float *qy, const float *qr,
                          const uint32_t *qdesc, const int *q_pred_level, int nQ, int mode, unsigned thr,
                          int *out_idx, unsigned *out_dist);

/* matchMapPointsSim3 (keyframe_matcher.cpp:633-686): both findMatchesTranformedMps directions + the agreement filter;
 * q12 / q21: per keypoint the projected query into the other keyframe (r < 0: no query); see oracle/src/search.cpp. */
int orc_match_sim3(const float *x1, const float *y1, const int *oct1, const uint32_t *d1, int n1,
                   const float *x2, const float *y2, const int *oct2, const uint32_t *d2, int n2,
                   const float *q12x, const float *q12y, const float *q12r, const uint32_t *q12desc, const int *q12lvl,
                   const float *q21x, const float *q21y, const float *q21r, const uint32_t *q21desc, const int *q21lvl,
                   int *out_pairs);

/* Bag of words behind BowIndex (bow_index.cpp:44-176); DBoW2 itself is absent: restated from its published
 * algorithm, PARITY UNPINNED.  See oracle/src/bow.cpp. */
void orc_bow_transform(const int *child_off, const int *child_ids, const uint32_t *node_desc, const double *node_weight,
                       const int *node_word, int n_nodes, int levels, const uint32_t *desc, int n, int levels_up,
                       int *out_word, double *out_weight, int *out_node);
int orc_bow_vector(const int *word, const double *weight, int n, unsigned *out_word, double *out_value);
void *orc_bowindex_create(int vocabulary_size);
void orc_bowindex_destroy(void *index);
void orc_bowindex_add(void *index, int map_id, int kf_id, const unsigned *word, const double *value, int n);
void orc_bowindex_remove(void *index, int map_id, int kf_id);
int orc_bowindex_similar(void *index, const unsigned *q_word, const double *q_value, int nq, int self_map, int self_kf,
                         float bowMinInCommonRatio, float bowScoreRatio, int *out_map, int *out_kf, float *out_score,
                         int capacity);

/* malloc settings of the timed CPU baselines: freed temporaries stay in the heap (see oracle/src/extract.cpp). */
void orc_tune_malloc(void);
double orc_bench_extract(const orc_params *p, const uint8_t *imgs, int n_frames, int threads, long *total_kp);
double orc_bench_match(const uint32_t *desc, const float *ang, int n_sets, int n_per_set,
                       const int *pairs, int n_pairs, float ratio, unsigned thr, int threads, long *total_matches);

#ifdef __cplusplus
}
#endif
#endif
