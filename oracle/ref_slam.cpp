// oracle/_ref/libref_slam.so -- the REFERENCE'S OWN hot-path sources compiled VERBATIM from /root/reference
// (static_settings.cpp, feature_search.cpp, id.cpp, map_point.cpp, keyframe.cpp, keyframe_matcher.cpp,
// orb_extractor.cpp, image_pyramid.cpp, feature_detector.cpp, bow_index.cpp, viewer_data_publisher.cpp; see
// oracle/Makefile target `ref`) behind plain C entry points.  TEST INFRASTRUCTURE ONLY: tests/ use it to pin the
// oracle restatement (oracle/src/*.cpp) against the real reference code, live in this container and through the
// golden fixtures tools/gen_golden.py writes from it.
//
// What is NOT the reference here (oracle/shim/, all absent from the reference tree):
//   * OpenCV primitives cv::resize / cv::GaussianBlur / cv::fastAtan2 / cvRound: the oracle's restatements, which are
//     pinned bit-for-bit against cv2 4.13 (tests/golden/golden_cv2.npz);
//   * tracker::FeatureDetector (the parent project's corner detector): backed by the oracle's FAST-in-cells +
//     quadtree detector -- that stage stays "parity unpinned", the reference does not contain it;
//   * Eigen, cereal, DBoW2, accelerated-arrays, tracker::Image / Camera, odometry::Parameters: minimal stand-ins;
//   * openvslam::solve::essential_solver::create_E_21 (essential_solver.cc:157-162 needs Eigen's SVD for its other
//     members): restated below; MapDB helpers of mapdb.cpp (needs ../odometry/util.hpp): getMapWithId restated below.
// This file contains scenario builders only: it fills Keyframe / MapPoint / MapDB objects from flat arrays, calls
// the reference function, and flattens the result.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <set>
#include <thread>
#include <vector>

#include "orb_oracle.h"

#include "static_settings.hpp"
#include "feature_search.hpp"
#include "orb_extractor.hpp"
#include "image_pyramid.hpp"
#include "feature_detector.hpp"
#include "keyframe.hpp"
#include "keyframe_matcher.hpp"
#include "map_point.hpp"
#include "mapdb.hpp"
#include "bow_index.hpp"
#include "openvslam/essential_solver.h"
#include "../odometry/parameters.hpp"
#include "../tracker/image.hpp"
#include "../tracker/camera.hpp"
#include "../tracker/feature_detector.hpp"
#include <accelerated-arrays/opencv_adapter.hpp>

// ---- OpenCV primitives: the oracle's cv2-pinned restatements ------------------------------------------------
namespace cv {
float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }
void resize(const Mat &src, Mat &dst, Size dsize, double, double, int interpolation) {
    assert(interpolation == INTER_LINEAR);
    (void)interpolation;
    dst.create(dsize.height, dsize.width);
    orc_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}
void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
    assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
    (void)ksize; (void)sigmaX; (void)sigmaY; (void)borderType;
    dst.create(src.rows, src.cols);
    orc_gaussian7_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, (int)dst.step);
}
void vconcat(const Mat &a, const Mat &b, Mat &dst) {
    dst.create(a.rows + b.rows, a.cols);
    a.copyTo(Mat(dst, Rect(0, 0, a.cols, a.rows)));
    b.copyTo(Mat(dst, Rect(0, a.rows, b.cols, b.rows)));
}
}  // namespace cv

// ---- the absent detector: oracle FAST-in-cells + quadtree (upstream OpenVSLAM scheme) ----------------------------
namespace tracker {
namespace {
struct OracleDetector : FeatureDetector {
    odometry::ParametersTracker params;
    accelerated::Future detect(accelerated::Image &image, std::vector<Feature::Point> &out,
                               const std::vector<Feature::Point> &, double) final {
        const int cap = params.maxTracks + 16;
        std::vector<int> xs(cap), ys(cap), rs(cap);
        int n_cand = 0;
        const int n = orc_detect_level(image.data, image.width, image.height, (int)image.stride, params.maxTracks,
                                       params.iniFastThreshold, params.minFastThreshold, xs.data(), ys.data(), rs.data(),
                                       cap, nullptr, nullptr, nullptr, 0, &n_cand);
        out.clear();
        for (int i = 0; i < n && i < cap; ++i) out.push_back(Feature::Point{(float)xs[i], (float)ys[i]});
        return accelerated::Future();
    }
};
}  // namespace
std::unique_ptr<FeatureDetector> FeatureDetector::build(int, int, accelerated::Processor &, accelerated::Image::Factory &,
                                                        accelerated::operations::StandardFactory &,
                                                        const odometry::ParametersTracker &params) {
    auto d = new OracleDetector();
    d->params = params;
    return std::unique_ptr<FeatureDetector>(d);
}
}  // namespace tracker

// ---- essential_solver.cc:139-162 (to_skew_symmetric_mat, create_E_21), restated on the Eigen stand-in ----------
namespace openvslam { namespace solve {
Mat33_t essential_solver::create_E_21(const Mat33_t &rot_1w, const Vec3_t &trans_1w, const Mat33_t &rot_2w, const Vec3_t &trans_2w) {
    const Mat33_t rot_21 = rot_2w * rot_1w.transpose();
    const Vec3_t trans_21 = -rot_21 * trans_1w + trans_2w;
    Mat33_t skew;
    skew << 0, -trans_21(2), trans_21(1), trans_21(2), 0, -trans_21(0), -trans_21(1), trans_21(0), 0;
    return skew * rot_21;
}
} }  // namespace openvslam::solve

namespace slam {
// mapdb.cpp:269-272
const MapDB &getMapWithId(MapId mapId, const MapDB &mapDB, const Atlas &atlas) {
    if (mapId == CURRENT_MAP_ID) return mapDB;
    return atlas[mapId.v];
}
// external linkage in keyframe_matcher.cpp:552, not declared in its header
std::vector<int> findMatchesTranformedMps(std::vector<MpId> mpIdsA, std::vector<bool> alreadyMatchedInA, Keyframe &kfB,
                                          const Eigen::Matrix3d &rotBAW, const Eigen::Vector3d &transBAW, MapDB &mapDB,
                                          float margin, const StaticSettings &settings);
}  // namespace slam

namespace {
using namespace slam;

struct GrayImage : tracker::CpuImage {
    cv::Mat mat;
    std::unique_ptr<accelerated::Image> acc;
    accelerated::Image::Factory factory;
    accelerated::operations::StandardFactory ops;
    accelerated::Processor proc;
    std::shared_ptr<const tracker::Camera> camera;
    GrayImage(const uint8_t *img, int w, int h, int stride) : mat(h, w, CV_8UC1, (void *)img, (size_t)stride) {
        acc = accelerated::opencv::ref(mat);
    }
    accelerated::Image &getAccImage() final { return *acc; }
    accelerated::Image::Factory &getImageFactory() final { return factory; }
    accelerated::operations::StandardFactory &getOperationsFactory() final { return ops; }
    accelerated::Processor &getProcessor() final { return proc; }
    std::shared_ptr<const tracker::Camera> getCamera() const final { return camera; }
    cv::Mat getOpenCvMat() final { return mat; }
};

odometry::Parameters make_parameters(const orc_params *p) {
    odometry::Parameters q;
    q.slam.orbScaleLevels = (unsigned)p->levels;
    q.slam.orbScaleFactor = p->scale_factor;
    q.slam.maxKeypoints = (unsigned)p->max_keypoints;
    q.tracker.iniFastThreshold = p->ini_fast_thr;
    q.tracker.minFastThreshold = p->min_fast_thr;
    return q;
}

// FeatureSearch decorator: forwards to the reference implementation and records every query it receives
struct Query { float x, y, r; };
struct RecordingSearch : FeatureSearch {
    std::unique_ptr<FeatureSearch> inner;
    mutable std::vector<Query> log;
    explicit RecordingSearch(const KeyPointVector &kps) : inner(FeatureSearch::create(kps)) {}
    void getFeaturesAround(float x, float y, float r, std::vector<size_t> &output) const final {
        log.push_back(Query{x, y, r});
        inner->getFeaturesAround(x, y, r, output);
    }
};

// a keyframe from flat arrays: identity pose, pinhole f = 1 / c = 0 camera (pixel == ray.xy / ray.z)
std::shared_ptr<Keyframe> make_keyframe(int id, const float *x, const float *y, const float *angle, const int *oct,
                                        const uint32_t *desc, const double *bearing, int n, RecordingSearch **rec = nullptr) {
    auto kf = std::make_shared<Keyframe>();
    kf->shared = std::make_shared<KeyframeShared>();
    kf->shared->camera = std::make_shared<tracker::Camera>();
    kf->id = KfId(id);
    kf->poseCW = Eigen::Matrix4d::Identity();
    kf->origPoseCW = Eigen::Matrix4d::Identity();
    kf->hasFullFeatures = true;
    kf->shared->keyPoints.resize(n);
    for (int i = 0; i < n; ++i) {
        KeyPoint &kp = kf->shared->keyPoints[i];
        kp.pt.x = x ? x[i] : 0.f;
        kp.pt.y = y ? y[i] : 0.f;
        kp.angle = angle ? angle[i] : 0.f;
        kp.octave = oct ? oct[i] : 0;
        if (bearing) kp.bearing = Eigen::Vector3d(bearing[3 * i], bearing[3 * i + 1], bearing[3 * i + 2]);
        std::memcpy(kp.descriptor.data(), desc + 8 * (size_t)i, 32);
    }
    kf->mapPoints.assign(n, MpId(-1));
    auto *r = new RecordingSearch(kf->shared->keyPoints);
    kf->shared->featureSearch.reset(r);
    if (rec) *rec = r;
    return kf;
}

void fill_feature_vector(DBoW2::FeatureVector &fv, const int *node, int n) {
    for (int i = 0; i < n; ++i) if (node[i] >= 0) fv.addFeature((DBoW2::NodeId)node[i], (unsigned)i);
}

void set_pose(Keyframe &kf, const double *pose16) {
    if (!pose16) return;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) kf.poseCW(r, c) = pose16[4 * r + c];
}

struct SettingsBox {
    odometry::Parameters params;
    StaticSettings settings;
    SettingsBox(int levels, float scale_factor) : params(make(levels, scale_factor)), settings(params) {}
    static odometry::Parameters make(int levels, float scale_factor) {
        odometry::Parameters q;
        q.slam.orbScaleLevels = (unsigned)levels;
        q.slam.orbScaleFactor = scale_factor;
        return q;
    }
};

// map points of a projection query list; ids start at id0
void add_query_points(MapDB &db, int id0, const double *pos, const float *norm, const float *min_dist, const float *max_dist,
                      const uint32_t *desc, const int *status, int n) {
    for (int q = 0; q < n; ++q) {
        MapPoint mp;
        mp.id = MpId(id0 + q);
        mp.position = Eigen::Vector3d(pos[3 * q], pos[3 * q + 1], pos[3 * q + 2]);
        mp.norm = Eigen::Vector3f(norm[3 * q], norm[3 * q + 1], norm[3 * q + 2]);
        mp.minViewingDistance = min_dist[q];
        mp.maxViewingDistance = max_dist[q];
        mp.status = status ? (MapPointStatus)status[q] : MapPointStatus::TRIANGULATED;
        std::memcpy(mp.descriptor.data(), desc + 8 * (size_t)q, 32);
        db.mapPoints.emplace(mp.id, mp);
    }
}
}  // namespace

namespace {
template <class F> double run_threads(int n_items, int threads, F fn) {
    orc_tune_malloc();
    threads = std::max(1, threads);
    std::atomic<int> next{0};
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&] { for (int i; (i = next.fetch_add(1)) < n_items;) fn(i); });
    for (auto &th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
}  // namespace

extern "C" {

// StaticSettings (static_settings.cpp:9-60)
void ref_settings(int levels, float scale_factor, int max_keypoints, float *scale_factors, float *sigma_sq, int *budgets) {
    odometry::Parameters q;
    q.slam.orbScaleLevels = (unsigned)levels;
    q.slam.orbScaleFactor = scale_factor;
    q.slam.maxKeypoints = (unsigned)max_keypoints;
    const StaticSettings s(q);
    const auto b = s.maxNumberOfKeypointsPerLevel();
    for (int l = 0; l < levels; ++l) {
        if (scale_factors) scale_factors[l] = s.scaleFactors[l];
        if (sigma_sq) sigma_sq[l] = s.levelSigmaSq[l];
        if (budgets) budgets[l] = (int)b[l];
    }
}

// FeatureSearch::create + getFeaturesAround (feature_search.cpp:22-48)
int ref_features_around(const float *x, const float *y, int n, float qx, float qy, float r, int *out) {
    KeyPointVector kps(n);
    for (int i = 0; i < n; ++i) { kps[i].pt.x = x[i]; kps[i].pt.y = y[i]; }
    const auto fs = FeatureSearch::create(kps);
    std::vector<size_t> o;
    fs->getFeaturesAround(qx, qy, r, o);
    for (size_t i = 0; i < o.size(); ++i) out[i] = (int)o[i];
    return (int)o.size();
}

// CpuImagePyramid (image_pyramid.cpp:68-86) through ImagePyramid::build; planes tightly packed like orc_pyramid
int ref_pyramid(const orc_params *p, const uint8_t *img, int stride, uint8_t *pyr, uint8_t *blur) {
    const odometry::Parameters q = make_parameters(p);
    const StaticSettings s(q);
    GrayImage im(img, p->width, p->height, stride);
    auto pyramid = ImagePyramid::build(s, im);
    pyramid->update(im);
    size_t off = 0;
    for (size_t l = 0; l < pyramid->numberOfLevels(); ++l) {
        const cv::Mat a = accelerated::opencv::ref(pyramid->getLevel(l)), b = accelerated::opencv::ref(pyramid->getBlurredLevel(l));
        for (int r = 0; r < a.rows; ++r) {
            std::memcpy(pyr + off + (size_t)r * a.cols, a.data + (size_t)r * a.step, (size_t)a.cols);
            std::memcpy(blur + off + (size_t)r * b.cols, b.data + (size_t)r * b.step, (size_t)b.cols);
        }
        off += (size_t)a.rows * a.cols;
    }
    return (int)pyramid->numberOfLevels();
}

// OrbExtractor::detectAndExtract (orb_extractor.cpp:73-164) end to end.  valid_rect: {x0, y0, x1, y1} of the pixels
// tracker::Camera::isValidPixel accepts (NULL: all).  Same outputs as orc_extract minus the level coordinates.
int ref_extract(const orc_params *p, const uint8_t *img, int stride, const float *track_xy, const int *track_ids,
                int n_tracks, int track_level, const double *valid_rect, float *x, float *y, float *angle, int *octave,
                uint32_t *desc, int *track_id, int cap) {
    odometry::Parameters q = make_parameters(p);
    q.slam.orbLkTrackLevel = (unsigned)track_level;
    const StaticSettings s(q);
    GrayImage im(img, p->width, p->height, stride);
    tracker::Camera cam;
    if (valid_rect) { cam.vx0 = valid_rect[0]; cam.vy0 = valid_rect[1]; cam.vx1 = valid_rect[2]; cam.vy1 = valid_rect[3]; }
    std::vector<tracker::Feature> tracks(n_tracks);
    for (int t = 0; t < n_tracks; ++t) {
        tracks[t].id = track_ids ? track_ids[t] : t;
        tracks[t].points[0] = tracker::Feature::Point{track_xy[2 * t], track_xy[2 * t + 1]};
    }
    auto orb = OrbExtractor::build(s);
    KeyPointVector kps;
    std::vector<int> ids;
    orb->detectAndExtract(im, cam, tracks, kps, ids);
    const int n = (int)kps.size();
    for (int i = 0; i < n && i < cap; ++i) {
        x[i] = kps[i].pt.x; y[i] = kps[i].pt.y; angle[i] = kps[i].angle; octave[i] = kps[i].octave;
        std::memcpy(desc + 8 * (size_t)i, kps[i].descriptor.data(), 32);
        if (track_id) track_id[i] = ids[i];
    }
    return n;
}

// matchForLoopClosures (keyframe_matcher.cpp:50-158).  node*: DBoW2 feature-vector node of every feature (-1: none).
// status*: 0 = no map point, 1 = TRIANGULATED map point, 2 = NOT_TRIANGULATED map point.
unsigned ref_match_loop_closures(const uint32_t *dA, const float *aA, const int *nodeA, const unsigned char *statusA, int nA,
                                 const uint32_t *dB, const float *aB, const int *nodeB, const unsigned char *statusB, int nB,
                                 float ratio, int require_triangulation, int *matches) {
    auto kf1 = make_keyframe(1, nullptr, nullptr, aA, nullptr, dA, nullptr, nA);
    auto kf2 = make_keyframe(2, nullptr, nullptr, aB, nullptr, dB, nullptr, nB);
    fill_feature_vector(kf1->shared->bowFeatureVec, nodeA, nA);
    fill_feature_vector(kf2->shared->bowFeatureVec, nodeB, nB);
    MapDB db1, db2;
    auto populate = [](Keyframe &kf, MapDB &db, const unsigned char *st, int n) {
        for (int i = 0; i < n; ++i) {
            const int s = st ? st[i] : 1;
            if (s == 0) continue;
            MapPoint mp;
            mp.id = MpId(i);
            mp.status = s == 1 ? MapPointStatus::TRIANGULATED : MapPointStatus::NOT_TRIANGULATED;
            db.mapPoints.emplace(mp.id, mp);
            kf.mapPoints[i] = mp.id;
        }
    };
    populate(*kf1, db1, statusA, nA);
    populate(*kf2, db2, statusB, nB);
    odometry::ParametersSlam ps;
    ps.loopClosureFeatureMatchLoweRatio = ratio;
    ps.requireTringulationForLoopClosures = require_triangulation != 0;
    std::vector<int> m;
    const unsigned n = matchForLoopClosures(*kf1, *kf2, db1, db2, m, ps);
    for (int i = 0; i < nA; ++i) matches[i] = m[i];
    return n;
}

// matchForTriangulationDBoW (keyframe_matcher.cpp:160-293).  has_mp*: feature already owns a map point (skipped).
// pose*: world-to-camera 4x4 row-major.  E_out receives the essential matrix the reference built (row-major).
unsigned ref_match_triangulation(const uint32_t *dA, const float *aA, const int *octA, const double *bearA, const int *nodeA,
                                 const unsigned char *has_mpA, int nA, const uint32_t *dB, const float *aB, const double *bearB,
                                 const int *nodeB, const unsigned char *has_mpB, int nB, const double *poseA, const double *poseB,
                                 int levels, float scale_factor, float residual_deg_thr, int *matches, double *E_out) {
    auto kf1 = make_keyframe(1, nullptr, nullptr, aA, octA, dA, bearA, nA);
    auto kf2 = make_keyframe(2, nullptr, nullptr, aB, nullptr, dB, bearB, nB);
    fill_feature_vector(kf1->shared->bowFeatureVec, nodeA, nA);
    fill_feature_vector(kf2->shared->bowFeatureVec, nodeB, nB);
    for (int i = 0; i < nA; ++i) if (has_mpA && has_mpA[i]) kf1->mapPoints[i] = MpId(i);
    for (int i = 0; i < nB; ++i) if (has_mpB && has_mpB[i]) kf2->mapPoints[i] = MpId(i);
    set_pose(*kf1, poseA);
    set_pose(*kf2, poseB);
    SettingsBox sb(levels, scale_factor);
    sb.params.slam.epipolarCheckThresholdDegrees = residual_deg_thr;
    const auto pairs = matchForTriangulationDBoW(*kf1, *kf2, sb.settings);
    for (int i = 0; i < nA; ++i) matches[i] = -1;
    for (const auto &pr : pairs) matches[pr.first.v] = pr.second.v;
    if (E_out) {
        const Eigen::Matrix3d E = openvslam::solve::essential_solver::create_E_21(
            kf2->poseCW.topLeftCorner<3, 3>(), kf2->poseCW.block<3, 1>(0, 3),
            kf1->poseCW.topLeftCorner<3, 3>(), kf1->poseCW.block<3, 1>(0, 3));
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) E_out[3 * r + c] = E(r, c);
    }
    return (unsigned)pairs.size();
}

// searchByProjection (keyframe_matcher.cpp:295-414) on a keyframe with identity pose and the f = 1 pinhole camera.
// taken[k] != 0: keypoint k already owns a map point with an observation (:358).  Query q is a map point (position,
// viewing normal, viewing-distance range, descriptor).  Outputs per query: matched keypoint (-1: none), the query the
// reference passed to FeatureSearch (q_x, q_y, q_r; r = -1 when the geometry tests rejected the point before the
// search) and the predicted scale level.  Returns the match count.
int ref_search_by_projection(const float *kx, const float *ky, const int *koct, const uint32_t *kdesc, int nK,
                             const unsigned char *taken, const double *pos, const float *norm, const float *min_dist,
                             const float *max_dist, const uint32_t *qdesc, int nQ, float threshold, int levels,
                             float scale_factor, int *out_idx, float *q_x, float *q_y, float *q_r, int *q_level) {
    RecordingSearch *rec = nullptr;
    auto kf = make_keyframe(7, kx, ky, nullptr, koct, kdesc, nullptr, nK, &rec);
    MapDB db;
    for (int k = 0; k < nK; ++k)
        if (taken && taken[k]) {
            MapPoint mp(MpId(k), KfId(3), KpId(0));   // observed elsewhere
            db.mapPoints.emplace(mp.id, mp);
            kf->mapPoints[k] = mp.id;
        }
    const int id0 = nK + 1000;
    add_query_points(db, id0, pos, norm, min_dist, max_dist, qdesc, nullptr, nQ);
    SettingsBox sb(levels, scale_factor);
    std::vector<MpId> mps;
    for (int q = 0; q < nQ; ++q) mps.push_back(MpId(id0 + q));
    int count = 0;
    // one call per query keeps the query log aligned with the query index; the state (observations) carries over
    for (int q = 0; q < nQ; ++q) {
        const size_t before = rec->log.size();
        count += searchByProjection(*kf, std::vector<MpId>{mps[q]}, db, nullptr, threshold, sb.settings);
        const MapPoint &mp = db.mapPoints.at(mps[q]);
        out_idx[q] = mp.observations.count(kf->id) ? mp.observations.at(kf->id).v : -1;
        if (rec->log.size() > before) { q_x[q] = rec->log.back().x; q_y[q] = rec->log.back().y; q_r[q] = rec->log.back().r; }
        else { q_x[q] = q_y[q] = 0; q_r[q] = -1; }
        const float dist = (kf->cameraCenter() - mp.position).cast<float>().norm();
        q_level[q] = mp.predictScaleLevel(dist, sb.settings);
    }
    return count;
}

// replaceDuplication<std::vector<MpId>> (keyframe_matcher.cpp:416-529), same keyframe / camera convention.
// kp_mp[k]: number of observations of the map point keypoint k already owns (0: the keypoint owns none).  q_obs[q]:
// observations the query map point has elsewhere.  Outputs: final_mp[k] = query index owning keypoint k afterwards
// (-1: none, -2: its original map point), the recorded FeatureSearch queries, and the fused count (return value).
unsigned ref_replace_duplication(const float *kx, const float *ky, const int *koct, const uint32_t *kdesc, int nK,
                                 const int *kp_mp, const double *pos, const float *norm, const float *min_dist,
                                 const float *max_dist, const uint32_t *qdesc, const int *q_obs, int nQ, float margin,
                                 int levels, float scale_factor, int *final_mp, float *q_x, float *q_y, float *q_r, int *q_level) {
    RecordingSearch *rec = nullptr;
    auto kf = make_keyframe(7, kx, ky, nullptr, koct, kdesc, nullptr, nK, &rec);
    MapDB db;
    db.keyframes.emplace(kf->id, kf);
    // keyframes standing for "elsewhere" observations
    int max_obs = 1;
    for (int k = 0; k < nK; ++k) if (kp_mp && kp_mp[k] > max_obs) max_obs = kp_mp[k];
    for (int q = 0; q < nQ; ++q) if (q_obs && q_obs[q] > max_obs) max_obs = q_obs[q];
    const uint32_t zero[8] = {0};
    std::vector<std::shared_ptr<Keyframe>> others;
    const int slots = nK + nQ;
    for (int o = 0; o < max_obs; ++o) {
        std::vector<uint32_t> dz(8 * (size_t)slots, 0u);
        others.push_back(make_keyframe(100 + o, nullptr, nullptr, nullptr, nullptr, dz.data(), nullptr, slots));
        db.keyframes.emplace(others.back()->id, others.back());
    }
    (void)zero;
    for (int k = 0; k < nK; ++k)
        if (kp_mp && kp_mp[k] > 0) {
            MapPoint mp(MpId(k), kf->id, KpId(k));
            mp.status = MapPointStatus::TRIANGULATED;
            for (int o = 0; o + 1 < kp_mp[k]; ++o) { mp.addObservation(others[o]->id, KpId(k)); others[o]->mapPoints[k] = mp.id; }
            db.mapPoints.emplace(mp.id, mp);
            kf->mapPoints[k] = mp.id;
        }
    const int id0 = nK + 1000;
    add_query_points(db, id0, pos, norm, min_dist, max_dist, qdesc, nullptr, nQ);
    for (int q = 0; q < nQ; ++q) {
        MapPoint &mp = db.mapPoints.at(MpId(id0 + q));
        for (int o = 0; q_obs && o < q_obs[q]; ++o) { mp.addObservation(others[o]->id, KpId(nK + q)); others[o]->mapPoints[nK + q] = mp.id; }
    }
    SettingsBox sb(levels, scale_factor);
    unsigned fused = 0;
    for (int q = 0; q < nQ; ++q) {
        const size_t before = rec->log.size();
        const MpId id(id0 + q);
        if (db.mapPoints.count(id)) {
            const MapPoint &mp = db.mapPoints.at(id);
            const float dist = (kf->cameraCenter() - mp.position).cast<float>().norm();
            q_level[q] = mp.predictScaleLevel(dist, sb.settings);
        } else q_level[q] = -1;
        fused += replaceDuplication(*kf, std::vector<MpId>{id}, margin, db, sb.settings);
        if (rec->log.size() > before) { q_x[q] = rec->log.back().x; q_y[q] = rec->log.back().y; q_r[q] = rec->log.back().r; }
        else { q_x[q] = q_y[q] = 0; q_r[q] = -1; }
    }
    for (int k = 0; k < nK; ++k) {
        const int v = kf->mapPoints[k].v;
        final_mp[k] = v < 0 ? -1 : (v >= id0 ? v - id0 : -2);
    }
    return fused;
}

// matchMapPointsSim3 (keyframe_matcher.cpp:633-686) with transform12 = identity and identity keyframe poses.
// mp*[k]: index into the map-point arrays (pos / desc / distances / status) of keypoint k's map point, -1: none.
// seed_pairs: n_seed (map point of kf1, map point of kf2) pairs already matched (:643-649).
// Outputs: the pairs appended by the call as (keypoint of kf1, keypoint of kf2), the queries each direction passed to
// FeatureSearch (q12: kf1's map points searched in kf2, indexed by kf1 keypoint; q21 the reverse) with r = -1 when the
// point never reached the search, and the predicted levels.  Returns the number of appended pairs.
int ref_match_sim3(const float *x1, const float *y1, const int *oct1, const uint32_t *d1, const int *mp1, int n1,
                   const float *x2, const float *y2, const int *oct2, const uint32_t *d2, const int *mp2, int n2,
                   const double *mp_pos, const float *mp_min, const float *mp_max, const uint32_t *mp_desc,
                   const int *mp_status, int n_mp, const int *seed_pairs, int n_seed, int levels, float scale_factor,
                   int *out_pairs, float *q12, int *lvl12, float *q21, int *lvl21) {
    RecordingSearch *rec1 = nullptr, *rec2 = nullptr;
    auto kf1 = make_keyframe(1, x1, y1, nullptr, oct1, d1, nullptr, n1, &rec1);
    auto kf2 = make_keyframe(2, x2, y2, nullptr, oct2, d2, nullptr, n2, &rec2);
    MapDB db;
    std::vector<float> nz(3 * (size_t)n_mp, 0.f);
    add_query_points(db, 0, mp_pos, nz.data(), mp_min, mp_max, mp_desc, mp_status, n_mp);
    for (int k = 0; k < n1; ++k) if (mp1[k] >= 0) { kf1->mapPoints[k] = MpId(mp1[k]); db.mapPoints.at(MpId(mp1[k])).addObservation(kf1->id, KpId(k)); }
    for (int k = 0; k < n2; ++k) if (mp2[k] >= 0) { kf2->mapPoints[k] = MpId(mp2[k]); db.mapPoints.at(MpId(mp2[k])).addObservation(kf2->id, KpId(k)); }
    std::vector<std::pair<MpId, MpId>> matches;
    for (int i = 0; i < n_seed; ++i) matches.emplace_back(MpId(seed_pairs[2 * i]), MpId(seed_pairs[2 * i + 1]));
    SettingsBox sb(levels, scale_factor);
    // The reference runs direction 1 -> 2 then 2 -> 1 (:651-670), each walking the keypoints in index order and
    // issuing at most one FeatureSearch query per keypoint: the logs are re-aligned by replaying the skip rules.
    matchMapPointsSim3(*kf1, *kf2, Eigen::Matrix4d::Identity(), db, matches, sb.settings);
    auto align = [&](const Keyframe &kfa, RecordingSearch *rec_b, int na, float *q, int *lvl, bool first_is_a) {
        std::vector<bool> already(na, false);
        for (int i = 0; i < n_seed; ++i) {
            const MpId id(seed_pairs[2 * i + (first_is_a ? 0 : 1)]);
            already[db.mapPoints.at(id).observations.at(kfa.id).v] = true;
        }
        size_t cursor = 0;
        for (int k = 0; k < na; ++k) {
            q[3 * k] = q[3 * k + 1] = 0; q[3 * k + 2] = -1; lvl[k] = -1;
            if (already[k] || kfa.mapPoints[k].v == -1) continue;
            const MapPoint &mp = db.mapPoints.at(kfa.mapPoints[k]);
            if (mp.status != MapPointStatus::TRIANGULATED) continue;
            // geometry tests of :570-590 on the identity transform: in front of the camera, inside the distance range
            if (!(mp.position.z() > 0)) continue;
            const double vd = mp.position.norm();
            if (vd < mp.minViewingDistance || mp.maxViewingDistance < vd) continue;
            lvl[k] = mp.predictScaleLevel(vd, sb.settings);
            const Query &rq = rec_b->log.at(cursor++);
            q[3 * k] = rq.x; q[3 * k + 1] = rq.y; q[3 * k + 2] = rq.r;
        }
        assert(cursor == rec_b->log.size());
    };
    align(*kf1, rec2, n1, q12, lvl12, true);
    align(*kf2, rec1, n2, q21, lvl21, false);
    int n_new = 0;
    for (size_t i = (size_t)n_seed; i < matches.size(); ++i, ++n_new) {
        out_pairs[2 * n_new] = db.mapPoints.at(matches[i].first).observations.at(kf1->id).v;
        out_pairs[2 * n_new + 1] = db.mapPoints.at(matches[i].second).observations.at(kf2->id).v;
    }
    return n_new;
}

// MapPoint::updateDescriptor (map_point.cpp:75-116): out_desc receives the descriptor the reference selects for
// every segment (offsets[n_seg + 1] into desc).
void ref_medoid(const uint32_t *desc, const long long *offsets, int n_seg, uint32_t *out_desc) {
    for (int s = 0; s < n_seg; ++s) {
        const int n = (int)(offsets[s + 1] - offsets[s]);
        MapDB db;
        MapPoint mp;
        mp.id = MpId(0);
        std::memset(mp.descriptor.data(), 0xff, 32);
        for (int i = 0; i < n; ++i) {
            auto kf = make_keyframe(i, nullptr, nullptr, nullptr, nullptr, desc + 8 * (size_t)(offsets[s] + i), nullptr, 1);
            db.keyframes.emplace(kf->id, kf);
            mp.observations.emplace(kf->id, KpId(0));
        }
        mp.updateDescriptor(db);
        std::memcpy(out_desc + 8 * (size_t)s, mp.descriptor.data(), 32);
    }
}

// ---- BowIndex (bow_index.cpp:31-176) over the DBoW2 stand-in; the vocabulary is loaded from a DBoW2 text file ----
struct RefBowBox { odometry::ParametersSlam *ps = nullptr; BowIndex *index = nullptr; MapDB db; Atlas atlas; };
void *ref_bow_create(const char *vocabulary_txt, float min_in_common_ratio, float score_ratio) {
    auto *ps = new odometry::ParametersSlam();
    ps->vocabularyPath = vocabulary_txt;
    ps->bowMinInCommonRatio = min_in_common_ratio;
    ps->bowScoreRatio = score_ratio;
    auto *b = new RefBowBox();
    b->ps = ps;
    b->index = new BowIndex(*ps);
    return b;
}
void ref_bow_destroy(void *h) {
    auto *b = static_cast<RefBowBox *>(h);
    delete b->index;
    delete b->ps;
    delete b;
}
// BowIndex::transform: per feature the feature-vector node (-1: dropped, weight 0) + the BowVector; returns its size
int ref_bow_transform(void *h, const uint32_t *desc, int n, int *out_node, unsigned *vec_word, double *vec_value, int cap) {
    auto *b = static_cast<RefBowBox *>(h);
    KeyPointVector kps(n);
    for (int i = 0; i < n; ++i) std::memcpy(kps[i].descriptor.data(), desc + 8 * (size_t)i, 32);
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    b->index->transform(kps, bv, fv);
    for (int i = 0; i < n; ++i) out_node[i] = -1;
    for (const auto &e : fv) for (unsigned i : e.second) out_node[i] = (int)e.first;
    int k = 0;
    for (const auto &e : bv) { if (k < cap) { vec_word[k] = e.first; vec_value[k] = e.second; } ++k; }
    return k;
}
// BowIndex::add of a keyframe described by its BowVector (current map)
void ref_bow_add(void *h, int kf_id, const unsigned *word, const double *value, int n) {
    auto *b = static_cast<RefBowBox *>(h);
    const uint32_t d[8] = {0};
    auto kf = make_keyframe(kf_id, nullptr, nullptr, nullptr, nullptr, d, nullptr, 0);
    for (int i = 0; i < n; ++i) kf->shared->bowVec[word[i]] = value[i];
    b->db.keyframes[kf->id] = kf;
    b->index->add(*kf, CURRENT_MAP_ID);
}
void ref_bow_remove(void *h, int kf_id) {
    auto *b = static_cast<RefBowBox *>(h);
    b->index->remove(MapKf{CURRENT_MAP_ID, KfId(kf_id)});
    b->db.keyframes.erase(KfId(kf_id));
}
// BowIndex::getBowSimilar for a query keyframe (id self_kf, may or may not be in the index)
int ref_bow_similar(void *h, const unsigned *q_word, const double *q_value, int nq, int self_kf, int *out_kf, float *out_score, int cap) {
    auto *b = static_cast<RefBowBox *>(h);
    const uint32_t d[8] = {0};
    auto kf = make_keyframe(self_kf, nullptr, nullptr, nullptr, nullptr, d, nullptr, 0);
    for (int i = 0; i < nq; ++i) kf->shared->bowVec[q_word[i]] = q_value[i];
    const auto sim = b->index->getBowSimilar(b->db, b->atlas, *kf);
    for (size_t i = 0; i < sim.size() && (int)i < cap; ++i) { out_kf[i] = sim[i].mapKf.kfId.v; out_score[i] = sim[i].score; }
    return (int)sim.size();
}

// ---- timed drivers for bench.py's reference arm (`"kind": "reference"`): frames / keyframe pairs sharded over host
//      threads; the reference itself is single threaded, one OrbExtractor per thread like one per Mapper (mapper.cpp:145)

double ref_bench_extract(const orc_params *p, const uint8_t *imgs, int n_frames, int threads, long *total_kp) {
    orc_tune_malloc();
    const odometry::Parameters q = make_parameters(p);
    const StaticSettings s(q);
    std::atomic<long> total{0};
    const size_t fsz = (size_t)p->width * p->height;
    threads = std::max(1, threads);
    std::atomic<int> next{0};
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&] {
            auto orb = OrbExtractor::build(s);
            tracker::Camera cam;
            KeyPointVector kps;
            std::vector<int> ids;
            for (int i; (i = next.fetch_add(1)) < n_frames;) {
                GrayImage im(imgs + fsz * (size_t)i, p->width, p->height, p->width);
                orb->detectAndExtract(im, cam, {}, kps, ids);
                total += (long)kps.size();
            }
        });
    for (auto &th : pool) th.join();
    if (total_kp) *total_kp = total;
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

double ref_bench_match(const uint32_t *desc, const float *ang, int n_sets, int n_per_set, const int *pairs, int n_pairs,
                       float ratio, int threads, long *total_matches) {
    std::vector<std::shared_ptr<Keyframe>> kfs;
    MapDB db;
    std::vector<int> node(n_per_set, 0);
    for (int i = 0; i < n_per_set; ++i) {
        MapPoint mp;
        mp.id = MpId(i);
        mp.status = MapPointStatus::TRIANGULATED;
        db.mapPoints.emplace(mp.id, mp);
    }
    for (int sidx = 0; sidx < n_sets; ++sidx) {
        auto kf = make_keyframe(sidx, nullptr, nullptr, ang + (size_t)sidx * n_per_set, nullptr, desc + (size_t)sidx * n_per_set * 8, nullptr, n_per_set);
        fill_feature_vector(kf->shared->bowFeatureVec, node.data(), n_per_set);
        for (int i = 0; i < n_per_set; ++i) kf->mapPoints[i] = MpId(i);
        kfs.push_back(kf);
    }
    odometry::ParametersSlam ps;
    ps.loopClosureFeatureMatchLoweRatio = ratio;
    std::atomic<long> total{0};
    const double secs = run_threads(n_pairs, threads, [&](int i) {
        std::vector<int> m;
        total += matchForLoopClosures(*kfs[pairs[2 * i]], *kfs[pairs[2 * i + 1]], db, db, m, ps);
    });
    if (total_matches) *total_matches = total;
    return secs;
}

}  // extern "C"
