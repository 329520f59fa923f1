/*
 * slamgpu.h -- C ABI of libslamgpu.so: the B200 (sm_100a) implementation of the SLAM module's
 * data-parallel front-end hot path (image pyramid -> FAST + quadtree distribution -> intensity-
 * centroid orientation -> rBRIEF-256 -> brute-force Hamming matching with ratio / uniqueness /
 * angle-histogram checks).
 *
 * Conventions
 *   - plain C: pointers, sizes, PODs.  No C++ / torch types cross this boundary.
 *   - every entry point returns an int status (SG_OK == 0); sg_last_error() gives the text.
 *   - pointers named h_* are HOST memory, d_* are DEVICE memory of the context's GPU.
 *   - a context owns every device buffer, table and stream; the caller owns every pointer it
 *     passes in.  Sizes (image size, levels, budget, frames per batch) are fixed at sg_create,
 *     mirroring the reference, whose extractor caches a pyramid/detector sized by the first image
 *     (orb_extractor.cpp:80-81).
 *   - calls on one context must be serialised by the caller (the reference objects are stateful
 *     and single-threaded too: orb_extractor.cpp:217-219).  One context per (host thread, GPU).
 *   - there is NO CPU fallback: without a CUDA device sg_create fails with SG_ERR_CUDA.
 *
 * Each entry point cites the reference interface (file:line under the reference tree) it replaces.
 */
#ifndef SLAMGPU_H
#define SLAMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG_MAX_LEVELS 16
#define SG_ABI_VERSION 1

enum {
    SG_OK = 0,
    SG_ERR_INVALID = 1,   /* bad argument / size outside what the context was created for */
    SG_ERR_CUDA = 2,      /* CUDA runtime error (text in sg_last_error)                    */
    SG_ERR_OVERFLOW = 3,  /* an internal capacity was exceeded (candidate list, node table) */
    SG_ERR_STATE = 4      /* call order violated (e.g. detect before pyramid update)        */
};

/* Static configuration: the fields of odometry::ParametersSlam the path reads
 * (static_settings.cpp:31-32,48; orb_extractor.cpp:91) plus the batch geometry. */
typedef struct sg_params {
    int width;            /* level-0 image width  (pixels)                                     */
    int height;           /* level-0 image height                                              */
    int levels;           /* slam.orbScaleLevels                                               */
    float scale_factor;   /* slam.orbScaleFactor                                               */
    int max_keypoints;    /* slam.maxKeypoints  (budget split by static_settings.cpp:39-60)    */
    int ini_fast_thr;     /* FAST threshold tried first in every 64-px cell (upstream: 20)     */
    int min_fast_thr;     /* FAST threshold used when a cell yields nothing (upstream: 7)      */
    int max_frames;       /* frames per batch the context is sized for (>= 1)                  */
    int max_tracks;       /* tracker features per frame described besides detected ones (>= 0) */
    int track_level;      /* slam.orbLkTrackLevel                                              */
} sg_params;

typedef struct sg_ctx sg_ctx;

/* ---- life cycle ------------------------------------------------------------------------------ */
/* Replaces OrbExtractor::build / ImagePyramid::build / FeatureDetector::build
 * (orb_extractor.cpp:356-358, image_pyramid.cpp:209-219, feature_detector.cpp:138-140). */
int sg_create(int device, const sg_params *params, sg_ctx **out);
/* Objects created from a context (sg_db, sg_vocab, sg_bowdb) use its device, streams and memory pool: destroy them
 * before the context. */
void sg_destroy(sg_ctx *ctx);
const char *sg_last_error(const sg_ctx *ctx); /* ctx may be NULL: error of the failed sg_create */
int sg_abi_version(void);
/* Block until everything queued on the context's stream has finished. */
int sg_synchronize(sg_ctx *ctx);
/* The context's cudaStream_t (as void*), so a harness can bracket the launches with its own events. */
void *sg_stream(sg_ctx *ctx);
/* Number of kernels this library has launched on the context since creation. */
unsigned long long sg_launch_count(const sg_ctx *ctx);

/* StaticSettings (static_settings.cpp:9-60) + level sizes (image_pyramid.cpp:76-78).
 * Any output pointer may be NULL.  Arrays hold `levels` entries. */
int sg_get_geometry(const sg_ctx *ctx, float *scales, int *widths, int *heights, int *pitches,
                    int *budgets);
/* Capacity (per frame) of the keypoint output arrays of sg_extract*. */
int sg_keypoint_capacity(const sg_ctx *ctx);

/* ---- image pyramid: ImagePyramid::update / getLevel / getBlurredLevel (image_pyramid.hpp:20-27,
 *      CpuImagePyramid::update image_pyramid.cpp:68-86) --------------------------------------- */
/* n_frames 8-bit gray images, frame f at h_imgs + f*frame_stride, rows `pitch` bytes apart. */
int sg_pyramid_update(sg_ctx *ctx, const uint8_t *h_imgs, int pitch, size_t frame_stride, int n_frames);
/* Same, images already in device memory (pitch % 16 == 0, base 16-byte aligned). */
int sg_pyramid_update_device(sg_ctx *ctx, const uint8_t *d_imgs, int pitch, size_t frame_stride, int n_frames);
/* Copy one plane (blurred != 0: the Gaussian-blurred one) of one frame to host memory. */
int sg_pyramid_download(sg_ctx *ctx, int frame, int level, int blurred, uint8_t *h_dst, int dst_pitch);
/* Device pointer / pitch of a plane of frame 0 (frames follow each other frame_stride bytes apart):
 * the counterpart of getGpuLevel (image_pyramid.hpp:27). */
int sg_pyramid_device_plane(sg_ctx *ctx, int level, int blurred, const uint8_t **d_plane, int *pitch,
                            size_t *frame_stride);

/* ---- detection: FeatureDetector::detect (feature_detector.hpp:20-22, feature_detector.cpp:68-134)
 *      on the pyramid of the last sg_pyramid_update*.  Per level: FAST-9/16 in 64-px cells
 *      (ini -> min threshold), NMS, quadtree distribution to the level budget. ---------------- */
int sg_detect(sg_ctx *ctx);
/* Keypoints of one (frame, level) in the order of the final quadtree node list: integer level
 * coordinates and FAST response.  *n receives the count (also when it exceeds cap). */
int sg_detect_download(sg_ctx *ctx, int frame, int level, int *h_x, int *h_y, int *h_resp, int cap, int *n);
/* Debug / test hook: the candidates handed to the quadtree (working-area coords), unordered. */
int sg_detect_download_candidates(sg_ctx *ctx, int frame, int level, int *h_x, int *h_y, int *h_resp,
                                  int cap, int *n);

/* ---- full extraction: OrbExtractor::detectAndExtract (orb_extractor.hpp:16-20,
 *      orb_extractor.cpp:73-164), batched over frames ---------------------------------------- */
/* Output (SoA, host, caller-allocated, `cap` = sg_keypoint_capacity entries per frame; frame f
 * starts at index f*cap): full-resolution x,y (orb_extractor.cpp:156), angle in degrees, octave,
 * descriptor 8 x u32, track id (-1 for detected keypoints), integer level coordinates.
 * Order inside a frame: tracker keypoints, then level 0..levels-1 (orb_extractor.cpp:89-162).
 * Any array pointer may be NULL (not downloaded). */
typedef struct sg_keypoints {
    float *x, *y, *angle;
    int32_t *octave;
    uint32_t *desc;     /* 8 words per keypoint */
    int32_t *track_id;
    int32_t *lvl_x, *lvl_y;
    int32_t *count;         /* [n_frames] keypoints per frame                       */
    int32_t *level_count;   /* [n_frames * levels] detected keypoints per level     */
} sg_keypoints;

/* Tracker features to describe besides the detected ones (orb_extractor.cpp:89-124): per frame f,
 * n_tracks[f] points at h_track_xy + 2*f*max_tracks (full-res x,y pairs) with ids at
 * h_track_ids + f*max_tracks.  Pass NULL / NULL / NULL for none. */
int sg_extract(sg_ctx *ctx, const uint8_t *h_imgs, int pitch, size_t frame_stride, int n_frames,
               const float *h_track_xy, const int32_t *h_track_ids, const int32_t *n_tracks,
               sg_keypoints *h_out);
/* sg_extract pipelines the batch in chunks of `frames` frames (default 32): the host->device copy of chunk
 * c+1 and the device->host copy of chunk c-1 run under the kernels of chunk c.  Pass pinned host memory
 * (sg_host_alloc_pinned) for the copies to be asynchronous. */
int sg_set_pipeline_chunk(sg_ctx *ctx, int frames);
/* Streaming form of sg_extract (OrbExtractor::detectAndExtract, orb_extractor.hpp:16-20, without tracker points) for a
 * sequence of batches: submit queues the copies and the
 * kernels of the batch on the context's pipeline streams and returns; the host arrays of h_out (pinned) are complete
 * after sg_extract_wait(ctx, ticket).  The batch occupies the context's frame slots [base_frame, base_frame +
 * n_frames) -- a context created with max_frames = 2 x batch keeps two batches in flight on disjoint halves, so the
 * host->device copy of batch i+1 runs under the kernels and the device->host copy of batch i.  A slot range may be
 * reused once the batch that used it has been waited for; at most 8 tickets are outstanding.  h_imgs and the arrays
 * of h_out must stay valid until the wait returns. */
int sg_extract_submit(sg_ctx *ctx, const uint8_t *h_imgs, int pitch, size_t frame_stride, int n_frames, int base_frame,
                      const sg_keypoints *h_out, int *ticket);
int sg_extract_wait(sg_ctx *ctx, int ticket);
/* Device-resident variant: images in device memory, results stay on the device; fetch them with
 * sg_extract_download.  This is the call the throughput bench times. */
int sg_extract_device(sg_ctx *ctx, const uint8_t *d_imgs, int pitch, size_t frame_stride, int n_frames);
int sg_extract_download(sg_ctx *ctx, int n_frames, sg_keypoints *h_out);
/* sg_extract_device runs `parts` (1..8, default 4) independent slices of the batch on separate streams so that the
 * latency-bound stages of one slice overlap the issue-bound stages of another; 1 = one stream.  With per-stage
 * profiling switched on (sg_set_profiling) the stages run on one stream so that their event times are meaningful. */
int sg_set_overlap(sg_ctx *ctx, int parts);
/* Device views of the last extraction's results (SoA, `cap` entries per frame). */
typedef struct sg_keypoints_dev {
    const float *x, *y, *angle;
    const int32_t *octave;
    const uint32_t *desc;
    const int32_t *count;
    int cap;
} sg_keypoints_dev;
int sg_extract_device_views(sg_ctx *ctx, sg_keypoints_dev *out);

/* ---- Hamming distance: compute_descriptor_distance_32 (openvslam/match_base.h:18-39) --------
 * n pairs: h_out[i] = popcount(a[i] ^ b[i]) over 256 bits, computed on the GPU. */
int sg_hamming(sg_ctx *ctx, const uint32_t *h_a, const uint32_t *h_b, int n, uint32_t *h_out);

/* ---- brute-force matcher: the single-BoW-node case of matchForLoopClosures
 *      (keyframe_matcher.hpp:33-40, keyframe_matcher.cpp:50-158): for every A feature in index
 *      order, best and second-best Hamming distance over the still-unmatched B features, reject
 *      when thr < best or ratio*second < best, consume the B feature, then keep only matches in
 *      the three most populated 30-degree delta-angle bins (match_angle_checker.h:61-134). */
typedef struct sg_match_params {
    float ratio;            /* slam.loopClosureFeatureMatchLoweRatio (keyframe_matcher.cpp:120) */
    uint32_t thr;           /* HAMMING_DIST_THR_LOW = 50 (match_base.h:13, keyframe_matcher.cpp:115) */
    int check_orientation;  /* keyframe_matcher.cpp:48 (always true in the reference)            */
    int ratio_is_double;    /* promotion of the ratio test; the parameter's C type is in an absent header */
} sg_match_params;

/* One pair, host buffers.  h_matches[nA] receives the B index or -1; *n_matches the count. */
int sg_match_bruteforce(sg_ctx *ctx, const uint32_t *h_descA, const float *h_angA, int nA,
                        const uint32_t *h_descB, const float *h_angB, int nB,
                        const sg_match_params *mp, int32_t *h_matches, uint32_t *n_matches);

/* matchForLoopClosures with the reference's DBoW2 node buckets (keyframe_matcher.cpp:65-146): only features under
 * the same vocabulary node are compared.  h_node*[i] = node id of feature i as in the keyframe's bowFeatureVec
 * (features of a node in index order; < 0: in no node), h_elig* = the map-point filters of :79-84 / :93-96
 * evaluated by the caller (NULL: all eligible).  Every shared node runs as one pair of the batched brute-force
 * matcher; the angle histogram spans all nodes.  Outputs as sg_match_bruteforce. */
int sg_match_bow(sg_ctx *ctx, const uint32_t *h_descA, const float *h_angA, const int32_t *h_nodeA, const uint8_t *h_eligA,
                 int nA, const uint32_t *h_descB, const float *h_angB, const int32_t *h_nodeB, const uint8_t *h_eligB,
                 int nB, const sg_match_params *mp, int32_t *h_matches, uint32_t *n_matches);

/* matchForTriangulationDBoW (keyframe_matcher.hpp:54, keyframe_matcher.cpp:160-293): features WITHOUT a map point
 * (h_elig* = 1) in the same DBoW2 node; per kf1 feature the last kf2 feature with distance <= thr and <= the best so far
 * that passes check_epipolar_constraint (:23-44) and is not matched yet; angle histogram over all nodes.  The Hamming
 * candidate lists come from the GPU; the epipolar test (fp64, acos) and the uniqueness walk run on the host over those
 * short lists.  E = create_E_21(kf2.R, kf2.t, kf1.R, kf1.t) row-major; bearings are 3 doubles per keypoint. */
typedef struct sg_triangulation_params {
    double E[9];
    const float *scale_factors;   /* StaticSettings::scaleFactors, n_levels entries */
    int n_levels;
    float residual_deg_thr;       /* slam.epipolarCheckThresholdDegrees */
    uint32_t thr;                 /* HAMMING_DIST_THR_LOW = 50 */
    int check_orientation;
} sg_triangulation_params;
int sg_match_triangulation(sg_ctx *ctx, const uint32_t *h_descA, const float *h_angA, const int32_t *h_octA, const double *h_bearA,
                           const int32_t *h_nodeA, const uint8_t *h_eligA, int nA, const uint32_t *h_descB, const float *h_angB,
                           const double *h_bearB, const int32_t *h_nodeB, const uint8_t *h_eligB, int nB,
                           const sg_triangulation_params *tp, int32_t *h_matches, uint32_t *n_matches);

/* Device-resident descriptor database: n_sets keyframes, set s owns features
 * [offsets[s], offsets[s+1]) of desc (8 words each) / angle. */
typedef struct sg_db sg_db;
int sg_db_create(sg_ctx *ctx, const uint32_t *h_desc, const float *h_angle, const int64_t *h_offsets,
                 int n_sets, sg_db **out);
/* Same from device arrays (e.g. the views of sg_extract_device); data is copied. */
int sg_db_create_device(sg_ctx *ctx, const uint32_t *d_desc, const float *d_angle, const int64_t *h_offsets,
                        int n_sets, sg_db **out);
/* The descriptor sets matchForLoopClosures reads (keyframe_matcher.cpp:61-63) as a view instead of a copy: the database
 * reads d_desc / d_angle in place (they must stay valid and unchanged while
 * the view is used, and d_desc needs 32 readable bytes behind the last descriptor -- the arrays of
 * sg_extract_device_views qualify).  The extract -> match flow of one GPU then moves no descriptor at all. */
int sg_db_wrap_device(sg_ctx *ctx, const uint32_t *d_desc, const float *d_angle, const int64_t *h_offsets,
                      int n_sets, sg_db **out);
/* Releases a database.  Normally called before sg_destroy of its context; a database that outlives its context has had
 * its device memory released by sg_destroy and only the handle is freed here. */
void sg_db_destroy(sg_db *db);
/* Batched matching of keyframe pairs (h_pairs: n_pairs x {setA, setB}).  h_matches may be NULL;
 * otherwise row p holds `match_stride` ints (>= size of the largest A set), -1 padded.
 * h_n_matches[n_pairs] receives the counts. */
int sg_match_pairs(sg_ctx *ctx, const sg_db *db, const int32_t *h_pairs, int n_pairs,
                   const sg_match_params *mp, int32_t *h_matches, int match_stride, uint32_t *h_n_matches);
/* Device-resident variant used by the throughput bench: pairs already on the device, results stay
 * there (d_matches may be NULL -> only counts). */
int sg_match_pairs_device(sg_ctx *ctx, const sg_db *db, const int32_t *d_pairs, int n_pairs,
                          const sg_match_params *mp, int32_t *d_matches, int match_stride,
                          uint32_t *d_n_matches);
/* Rows of the last sg_match_* call that needed the exact full-row rescan (diagnostic). */
unsigned long long sg_match_rescans(const sg_ctx *ctx);

/* ---- candidate-list matchers: FeatureSearch radius query (feature_search.cpp:22-48) + the
 *      descriptor loops of searchByProjection (keyframe_matcher.cpp:356-386), replaceDuplication (:482-499)
 *      and findMatchesTranformedMps (:604-627).  The caller projects the map points (kf.reproject, viewing
 *      distance, predictScaleLevel stay host geometry) and passes point, radius and descriptor per query.
 *   keypoints : h_kx, h_ky (the coordinates FeatureSearch indexes), h_koct, h_kdesc (8 words each), nK <= 65535;
 *               h_order = FeatureSearch's Y-sorted keypoint indices or NULL (the library runs the same std::sort)
 *   mode 0    : best only; match when best <= thr; with h_qlevel != NULL only octaves in [level-1, level] count
 *   mode 1    : searchByProjection: keypoints with h_taken[i] != 0 are skipped (:358), best and second best with
 *               their octaves, match when best <= thr and not (same octave and best > 0.8 * second); queries
 *               are resolved IN ORDER and a matched keypoint becomes taken (h_taken is updated in place)
 *   outputs   : h_idx[q] = matched keypoint or -1, h_dist[q] = its distance (256 when unmatched), *n_matched */
int sg_search_candidates(sg_ctx *ctx, const float *h_kx, const float *h_ky, const int32_t *h_koct,
                         const uint32_t *h_kdesc, int nK, const int32_t *h_order, uint8_t *h_taken,
                         const float *h_qx, const float *h_qy, const float *h_qr, const uint32_t *h_qdesc,
                         const int32_t *h_qlevel, int nQ, int mode, uint32_t thr, int32_t *h_idx, uint32_t *h_dist,
                         uint32_t *n_matched);
/* matchMapPointsSim3 (keyframe_matcher.cpp:633-686; findMatchesTranformedMps :552-631 in both directions + the
 * agreement filter :672-685).  The Sim3 geometry stays with the caller like in sg_search_candidates: per keypoint i of
 * keyframe 1, (h_q12x, h_q12y, h_q12r)[i] is the projection of its map point into keyframe 2 with the search radius
 * margin * scaleFactors[level], h_q12desc its descriptor and h_q12level the predicted level; r < 0 marks a keypoint that
 * issues no query (no map point, seeded as already matched :643-649, not TRIANGULATED, outside the image or the
 * viewing-distance range).  h_q21* likewise per keypoint of keyframe 2.  h_order1 / h_order2: FeatureSearch order or NULL.
 * h_pairs receives 2 * n_pairs indices (keypoint of kf1, keypoint of kf2), capacity 2 * min(n1, n2), in kf1 order. */
int sg_match_sim3(sg_ctx *ctx, const float *h_x1, const float *h_y1, const int32_t *h_oct1, const uint32_t *h_desc1, int n1,
                  const int32_t *h_order1, const float *h_x2, const float *h_y2, const int32_t *h_oct2, const uint32_t *h_desc2,
                  int n2, const int32_t *h_order2, const float *h_q12x, const float *h_q12y, const float *h_q12r,
                  const uint32_t *h_q12desc, const int32_t *h_q12level, const float *h_q21x, const float *h_q21y,
                  const float *h_q21r, const uint32_t *h_q21desc, const int32_t *h_q21level, int32_t *h_pairs,
                  uint32_t *n_pairs);
/* FeatureSearch constructor (feature_search.cpp:22-31): keypoint indices sorted by y (host code, no GPU). */
int sg_feature_index(const float *h_x, const float *h_y, int n, int32_t *h_order);
/* MapPoint::updateDescriptor (map_point.cpp:75-116), batched: segment s owns descriptors
 * [offsets[s], offsets[s+1]) (at most 1024); h_best[s] = index inside the segment of the descriptor whose
 * median Hamming distance to the segment is smallest (first wins; 0 for an empty segment). */
int sg_medoid(sg_ctx *ctx, const uint32_t *h_desc, const int64_t *h_offsets, int n_seg, int32_t *h_best);

/* ---- bag of words: BowIndex::transform (bow_index.cpp:59-93), add / remove (:44-57) and
 *      getBowSimilar (:95-176) over a DBoW2 vocabulary tree.  DBoW2 is an external dependency of the reference;
 *      its published arithmetic is restated: WordValue = double, TF-IDF weighting, L1 norm, L1 scoring.
 *   vocabulary : nodes in breadth-first order, node 0 = root; children of node i are
 *                h_child_ids[h_child_off[i] .. h_child_off[i+1]) (ids larger than i); per node a 256-bit
 *                descriptor, a weight and (leaves) a word id.
 *   transform  : per feature the word id, the word weight and the feature-vector node (the node reached at level
 *                L - levels_up; the root when that is <= 0).
 *   vector     : the keyframe's BowVector from those per-feature arrays: features with weight > 0 only, the value
 *                of a word = its features' weights added in feature order, then divided by the sum of |value| taken
 *                in ascending word order.  Output words ascending (std::map order); n <= 4096. */
typedef struct sg_vocab sg_vocab;
int sg_vocab_create(sg_ctx *ctx, const int32_t *h_child_off, const int32_t *h_child_ids, const uint32_t *h_node_desc,
                    const double *h_node_weight, const int32_t *h_node_word, int n_nodes, int levels, sg_vocab **out);
void sg_vocab_destroy(sg_vocab *vocab);
int sg_bow_transform(sg_ctx *ctx, const sg_vocab *vocab, const uint32_t *h_desc, int n, int levels_up, int32_t *h_word,
                     double *h_weight, int32_t *h_node);
int sg_bow_transform_device(sg_ctx *ctx, const sg_vocab *vocab, const uint32_t *d_desc, int n, int levels_up,
                            int32_t *d_word, double *d_weight, int32_t *d_node);
int sg_bow_vector(sg_ctx *ctx, const int32_t *h_word, const double *h_weight, int n, uint32_t *h_vec_word,
                  double *h_vec_value, int *n_words);
/* Batched form (one CTA per keyframe): keyframe k owns features [h_offsets[k], h_offsets[k+1]) of h_word / h_weight (at
 * most 4096 each); its BowVector is written to h_vec_word / h_vec_value starting at h_offsets[k], n_words[k] entries. */
int sg_bow_vector_batch(sg_ctx *ctx, const int32_t *h_word, const double *h_weight, const int64_t *h_offsets, int n_keyframes,
                        uint32_t *h_vec_word, double *h_vec_value, int32_t *n_words);
/* Device-resident BowVectors of up to max_keyframes keyframes (max_words_per_keyframe each), keyed by
 * (map id, keyframe id) like the reference's MapKf.  add = BowIndex::add, remove = BowIndex::remove (unknown
 * keyframe: no-op).  sg_bow_similar = getBowSimilar: every stored keyframe except (self_map, self_kf) that shares a
 * word with the query is counted; those with more than (unsigned)(min_in_common_ratio * max count) common words are
 * scored (DBoW2 L1 score of (query, stored), narrowed to float), sorted by descending score with std::sort from the
 * (map id, keyframe id) order, and cut at the first score < best * score_ratio.  *n_out = number of results; the
 * first min(*n_out, capacity) are written. */
typedef struct sg_bowdb sg_bowdb;
int sg_bowdb_create(sg_ctx *ctx, int max_keyframes, int max_words_per_keyframe, sg_bowdb **out);
void sg_bowdb_destroy(sg_bowdb *db);
int sg_bowdb_size(const sg_bowdb *db);
int sg_bowdb_add(sg_ctx *ctx, sg_bowdb *db, int map_id, int kf_id, const uint32_t *h_vec_word, const double *h_vec_value,
                 int n_words);
int sg_bowdb_remove(sg_ctx *ctx, sg_bowdb *db, int map_id, int kf_id);
int sg_bow_similar(sg_ctx *ctx, sg_bowdb *db, const uint32_t *h_q_word, const double *h_q_value, int nq, int self_map,
                   int self_kf, float min_in_common_ratio, float score_ratio, int32_t *h_map, int32_t *h_kf, float *h_score,
                   int capacity, int *n_out);

/* ---- angle histogram: angle_checker<int> (openvslam/match_angle_checker.h:72-134) -----------
 * Test hook for the restated libstdc++ std::sort order of the 30 bins (host code, no GPU). */
void sg_angle_bin_order(const uint32_t *sizes30, uint32_t *order30);
/* Same with an explicit introsort depth limit (0 forces the heap-sort branch; std::sort uses 8). */
void sg_angle_bin_order_depth(const uint32_t *sizes30, int depth_limit, uint32_t *order30);
int sg_angle_bin(float delta_angle);

/* ---- device helpers for harnesses that must not depend on a CUDA binding of their own ------- */
int sg_device_count(void);
int sg_malloc(sg_ctx *ctx, size_t bytes, void **d_ptr);
int sg_free(sg_ctx *ctx, void *d_ptr);
int sg_memcpy_h2d(sg_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int sg_memcpy_d2h(sg_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
int sg_host_alloc_pinned(size_t bytes, void **h_ptr);
int sg_host_free_pinned(void *h_ptr);
/* CUDA-event timing on the context's stream. */
int sg_timer_start(sg_ctx *ctx);
int sg_timer_stop(sg_ctx *ctx, float *ms); /* records, synchronises, returns elapsed ms */
/* Per-stage CUDA-event timing on the context's stream.  With profiling on, the stage launchers
 * record an event after each stage of every call (the last 64 calls are remembered);
 * sg_get_stage_ms synchronises and returns the AVERAGE duration per call in ms since profiling was
 * switched on: [0] pyramid (all level kernels), [1] FAST cells, [2] quadtree distribution,
 * [3] orientation + descriptors, [4] Hamming top-K (first chunk of a call), [5] match resolve
 * (first chunk); -1 where the stage did not run.  *n_calls (may be NULL): calls averaged. */
int sg_set_profiling(sg_ctx *ctx, int on);
int sg_get_stage_ms(sg_ctx *ctx, float *ms6, int *n_calls);
/* Overwrite a buffer larger than L2 (flushes L2 between timed iterations). */
int sg_flush_l2(sg_ctx *ctx);
/* Integer-pipe micro-benchmark: dependent-free POPC.b32 issue rate of the whole chip; returns
 * popc/s and the kernel time (denominator of the matching roofline). */
int sg_microbench_popc(sg_ctx *ctx, double *popc_per_s, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* SLAMGPU_H */
