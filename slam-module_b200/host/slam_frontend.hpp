// Host-side mirror of the reference's front-end interfaces, implemented on top of the C ABI of
// libslamgpu.so (include/slamgpu.h).  Same names, argument meaning and error behaviour as the
// reference (asserts, no exceptions, outputs cleared then filled), so that code written against
//   image_pyramid.hpp:16-30, feature_detector.hpp:15-24, orb_extractor.hpp:11-30,
//   static_settings.hpp:8-22, key_point.hpp:11-28, keyframe_matcher.hpp:33-91, feature_search.hpp:16-21,
//   bow_index.hpp:21-65, map_point.cpp:75-116
// compiles against this header after swapping the include.
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "compat.hpp"

struct sg_ctx;
struct sg_db;
struct sg_vocab;
struct sg_bowdb;

namespace slam {

// static_settings.hpp:8-22
struct StaticSettings {
    const odometry::Parameters &parameters;
    std::vector<float> scaleFactors;
    std::vector<float> levelSigmaSq;
    explicit StaticSettings(const odometry::Parameters &p);

    static constexpr unsigned ORB_PATCH_RADIUS = 19;
    static constexpr unsigned ORB_FAST_PATCH_SIZE = 31;
    static constexpr unsigned ORB_FAST_PATCH_HALF_SIZE = ORB_FAST_PATCH_SIZE / 2;

    std::vector<std::size_t> maxNumberOfKeypointsPerLevel() const;
};

using Vector2f = la::Vector2f;
using Vector2d = la::Vector2d;
using Vector3f = la::Vector3f;
using Vector3d = la::Vector3d;
using Matrix3d = la::Matrix3d;
using Matrix4d = la::Matrix4d;

// key_point.hpp:11-28 (bearing: filled later by keyframe.cpp:67)
struct KeyPoint {
    tracker::Feature::Point pt{0, 0};
    float angle = 0;
    int octave = 0;
    Vector3d bearing;
    using Descriptor = std::array<std::uint32_t, 8>;
    Descriptor descriptor{};
};
using KeyPointVector = std::vector<KeyPoint>;

// The CUDA context shared by the pyramid, the detector and the extractor built from the same
// StaticSettings + model image (the reference builds them lazily from the first image too,
// orb_extractor.cpp:80-81).
struct CudaFrontend;
std::shared_ptr<CudaFrontend> cudaFrontend(const StaticSettings &settings, int width, int height);
sg_ctx *cudaContext(const std::shared_ptr<CudaFrontend> &fe);

// image_pyramid.hpp:16-30
struct ImagePyramid {
    static std::unique_ptr<ImagePyramid> build(const StaticSettings &settings, tracker::Image &modelImage);
    virtual ~ImagePyramid();

    virtual void update(tracker::Image &image) = 0;
    virtual std::size_t numberOfLevels() const = 0;
    virtual bool isGpu() const = 0;

    virtual accelerated::Image &getLevel(std::size_t level) = 0;         // CPU (downloaded on demand)
    virtual accelerated::Image &getBlurredLevel(std::size_t level) = 0;  // CPU (downloaded on demand)
    virtual accelerated::Image &getGpuLevel(std::size_t level) = 0;      // GPU, not blurred

    // for debugging: all levels side by side in one 8-bit image (the reference renders a cv::Mat)
    virtual void debugVisualize(std::vector<std::uint8_t> &target, int &width, int &height) = 0;
};

// feature_detector.hpp:15-24
struct FeatureDetector {
    static std::unique_ptr<FeatureDetector> build(const StaticSettings &settings, tracker::Image &modelImage);
    virtual ~FeatureDetector();

    /** @return the total number of detected keypoints */
    virtual std::size_t detect(ImagePyramid &imagePyramid, std::vector<KeyPointVector> &keypointsPerLevel) = 0;
};

// orb_extractor.hpp:11-30
struct OrbExtractor {
    constexpr static int DESCRIPTOR_COLS = 32;
    virtual ~OrbExtractor() {}

    virtual void detectAndExtract(tracker::Image &img, const tracker::Camera &camera,
                                  const std::vector<tracker::Feature> &tracks, KeyPointVector &keyPoints,
                                  std::vector<int> &keyPointTrackIds) = 0;

    // Batched form (not in the reference): frames of one size, one call, results per frame.
    virtual void detectAndExtractBatch(const std::vector<tracker::Image *> &imgs, const tracker::Camera &camera,
                                       std::vector<KeyPointVector> &keyPoints) = 0;

    static std::unique_ptr<OrbExtractor> build(const StaticSettings &settings);

    enum class VisualizationMode { IMAGE_PYRAMID };
    virtual void debugVisualize(const tracker::Image &img, std::vector<std::uint8_t> &target, int &width, int &height,
                                VisualizationMode mode) const = 0;
};

// ---- matching (keyframe_matcher.hpp:10-12,33-40; openvslam/match_base.h:13-39) -------------------
constexpr unsigned int HAMMING_DIST_THR_LOW = 50;
constexpr unsigned int HAMMING_DIST_THR_HIGH = 100;
constexpr unsigned int MAX_HAMMING_DIST = 256;

// ---- the slice of the reference's data model the matchers touch (id.hpp, map_point.hpp, keyframe.hpp, mapdb.hpp) ----
struct Id { int v = -1; Id() {} explicit Id(int v) : v(v) {} };                    // id.hpp:14-22
struct KfId : Id { KfId() {} explicit KfId(int v) : Id(v) {} };
struct MpId : Id { MpId() {} explicit MpId(int v) : Id(v) {} };
struct KpId : Id { KpId() {} explicit KpId(int v) : Id(v) {} };
struct TrackId : Id { TrackId() {} explicit TrackId(int v) : Id(v) {} };
struct MapId : Id { MapId() {} explicit MapId(int v) : Id(v) {} };
inline bool operator<(const KfId &a, const KfId &b) { return a.v < b.v; }
inline bool operator<(const MpId &a, const MpId &b) { return a.v < b.v; }
inline bool operator<(const KpId &a, const KpId &b) { return a.v < b.v; }
inline bool operator<(const TrackId &a, const TrackId &b) { return a.v < b.v; }
inline bool operator==(const MpId &a, const MpId &b) { return a.v == b.v; }
inline bool operator==(const KfId &a, const KfId &b) { return a.v == b.v; }

enum class MapPointStatus { TRIANGULATED, NOT_TRIANGULATED, UNSURE, BAD };          // map_point.hpp:20
class MapDB;
class Keyframe;

// feature_search.hpp:16-21: the Y-sorted index comes from the library (sg_feature_index, the reference's std::sort)
class FeatureSearch {
public:
    static std::unique_ptr<FeatureSearch> create(const KeyPointVector &keypoints);
    virtual ~FeatureSearch() = default;
    virtual void getFeaturesAround(float x, float y, float r, std::vector<size_t> &output) const = 0;
};

// map_point.hpp:22-94 (members the matchers read or update)
class MapPoint {
public:
    MapPoint() {}
    MapPoint(MpId id, KfId keyframeId, KpId keyPointId);
    void addObservation(KfId keyframeId, KpId keyPointId);
    void eraseObservation(KfId keyframeId);
    void replaceWith(MapDB &mapDB, MapPoint &otherMp);
    int predictScaleLevel(float dist, const StaticSettings &settings) const;

    MpId id;
    TrackId trackId = TrackId(-1);
    MapPointStatus status = MapPointStatus::NOT_TRIANGULATED;
    Vector3d position;
    Vector3f norm;
    float minViewingDistance = 0;
    float maxViewingDistance = 30;
    KeyPoint::Descriptor descriptor{};
    std::map<KfId, KpId> observations;
    KfId referenceKeyframe;
    std::array<std::uint8_t, 3> color{{0, 0, 0}};   // cv::Vec3b, visualisation only (kept for the archive round trip)
};

// bowFeatureVec: DBoW2::FeatureVector = std::map<NodeId, std::vector<unsigned>> (keyframe.hpp, bow_index.cpp:59-93)
// bowVec: DBoW2::BowVector = std::map<WordId, WordValue (double)>
struct KeyframeShared {
    std::shared_ptr<const tracker::Camera> camera;
    std::string cameraModel;                          // camera->serialize() as stored in a map archive (keyframe.hpp:82)
    KeyPointVector keyPoints;
    std::unique_ptr<FeatureSearch> featureSearch;
    std::vector<std::array<std::uint8_t, 3>> colors;  // keyframe.hpp:55
    std::shared_ptr<std::vector<Vector3f>> stereoPointCloud;
    std::map<unsigned, double> bowVec;
    std::map<unsigned, std::vector<unsigned>> bowFeatureVec;
};

// keyframe.hpp:107-213
class Keyframe {
public:
    bool hasFeatureDescriptors() const { return hasFullFeatures; }
    void addObservation(MpId mapPointId, KpId keyPointId);
    void eraseObservation(MpId mapPointId);
    bool reproject(const Vector3d &point, Vector2f &reprojected) const;
    Vector3d cameraCenter() const;
    void getFeaturesAround(const Vector2f &point, float r, std::vector<size_t> &output);

    KfId id;
    KfId previousKfId, nextKfId;
    std::shared_ptr<KeyframeShared> shared;
    std::map<KpId, TrackId> keyPointToTrackId;
    std::vector<MpId> mapPoints;   // per keypoint, v == -1: none
    std::vector<float> keyPointDepth;
    Matrix4d poseCW = Matrix4d::Identity();
    Matrix4d origPoseCW = Matrix4d::Identity();
    la::Matrix<double, 3, 6> uncertainty;
    double t = 0;
    bool hasFullFeatures = true;
};

// loop_closer.hpp:35-46
struct LoopClosureEdge { KfId kfId1, kfId2; Matrix4d poseDiff = Matrix4d::Identity(); };

// mapdb.hpp:17-26
class MapDB {
public:
    std::map<KfId, std::shared_ptr<Keyframe>> keyframes;
    std::map<MpId, MapPoint> mapPoints;
    std::map<TrackId, MpId> trackIdToMapPoint;
    // serialised state the matchers never touch (mapdb.hpp:83-98), kept for the archive round trip
    std::vector<LoopClosureEdge> loopClosureEdges;
    Matrix4d prevPose = Matrix4d::Identity(), prevInputPose = Matrix4d::Identity();
    std::vector<double> discardedUncertainty = std::vector<double>(18, 0.0);   // Eigen::MatrixXd, column major
    int discardedUncertaintyRows = 3, discardedUncertaintyCols = 6;
    double firstKfTimestamp = -1.0;
    int nextMp = 0;
    KfId lastKfCandidateId, lastKfId;
};

// keyframe.hpp:216-222 / keyframe.cpp:408-424
bool reprojectToImage(const tracker::Camera &camera, const Matrix3d &rot_cw, const Vector3d &trans_cw, const Vector3d &pos_w,
                      Vector2d &reproj, float &x_right);
struct ViewerDataPublisher;   // debug hook of the reference's searchByProjection: accepted and ignored

/** keyframe_matcher.hpp:33-40.  With bowFeatureVec filled in on both keyframes: the reference's node-bucketed
 *  comparison (keyframe_matcher.cpp:65-146); with empty feature vectors every feature is in ONE node (brute force,
 *  BASELINE.json north star).
 *  `matchedMapPoints[i]` = keypoint index of kf2 matched to keypoint i of kf1, or -1.  @return match count */
unsigned int matchForLoopClosures(const Keyframe &kf1, const Keyframe &kf2, const MapDB &mapDB1, const MapDB &mapDB2,
                                  std::vector<int> &matchedMapPoints, const odometry::ParametersSlam &parameters,
                                  sg_ctx *ctx);

/** keyframe_matcher.hpp:53 (matchForTriangulationDBoW), :59-66 (searchByProjection), :72-78 (replaceDuplication, for
 *  std::vector<MpId> and std::set<MpId>), :85-91 (matchMapPointsSim3): the reference's signatures plus the context.
 *  Instantiations of the templates in slam_matchers.hpp for the types above. */
std::vector<std::pair<KpId, KpId>> matchForTriangulationDBoW(Keyframe &kf1, Keyframe &kf2, const StaticSettings &settings, sg_ctx *ctx);
int searchByProjection(Keyframe &kf, const std::vector<MpId> &mps, MapDB &mapDB, ViewerDataPublisher *dataPublisher,
                       const float threshold, const StaticSettings &settings, sg_ctx *ctx);
template <typename T>
unsigned int replaceDuplication(Keyframe &kf, const T &mapPoints, const float margin, MapDB &mapDB, const StaticSettings &settings,
                                sg_ctx *ctx);
void matchMapPointsSim3(Keyframe &kf1, Keyframe &kf2, const Matrix4d &transform12, MapDB &mapDB,
                        std::vector<std::pair<MpId, MpId>> &matches, const StaticSettings &settings, sg_ctx *ctx);
/** MapPoint::updateDescriptor (map_point.cpp:75-116) for a list of map points in one launch (the reference calls it per
 *  touched map point: mapper_helpers.cpp:122,132,312,811,1069). */
void updateDescriptors(MapDB &mapDB, const std::vector<MpId> &mapPoints, sg_ctx *ctx);

/** The same loop on bare keypoint vectors (every feature eligible). */
unsigned int bruteForceMatch(const KeyPointVector &kps1, const KeyPointVector &kps2, std::vector<int> &matches,
                             float loweRatio, bool checkOrientation, sg_ctx *ctx);

// ---- bag of words (bow_index.hpp:21-65) ---------------------------------------------------------------
namespace DBoW2 {
using BowVector = std::map<unsigned, double>;                       // WordId -> WordValue
using FeatureVector = std::map<unsigned, std::vector<unsigned>>;    // NodeId -> feature indices
}  // namespace DBoW2
using Atlas = std::vector<MapDB>;
struct MapKf { MapId mapId; KfId kfId; };   // bow_index.hpp:26-29
bool operator==(const MapKf &lhs, const MapKf &rhs);
bool operator<(const MapKf &lhs, const MapKf &rhs);
struct BowSimilar { MapKf mapKf; float score; };

struct BowVocabulary;
/** DBoW2 text vocabulary ("k L scoring weighting", then one line per node: parent is_leaf 32 descriptor bytes weight;
 *  the format bow_index.cpp:11-19 loads with loadFromTextFile) -> the flattened tree the library takes.
 *  @return false when the file cannot be read or is malformed */
bool loadVocabularyText(const std::string &path, BowVocabulary &out);

/** The loaded DBoW2 vocabulary tree, flattened: node 0 = root, the children of node i are
 *  childIds[childOff[i] .. childOff[i+1]) (breadth first, larger ids than i). */
struct BowVocabulary {
    std::vector<std::int32_t> childOff, childIds, nodeWord;
    std::vector<std::uint32_t> nodeDescriptor;   // 8 words per node
    std::vector<double> nodeWeight;
    int levels = 0;
};

/** bow_index.hpp:36-63 on the GPU: the vocabulary and every added keyframe's BowVector live on the device. */
class BowIndex {
public:
    BowIndex(const odometry::ParametersSlam &parameters, const BowVocabulary &vocabulary, sg_ctx *ctx, int maxKeyframes = 16384);
    ~BowIndex();
    BowIndex(const BowIndex &) = delete;
    BowIndex &operator=(const BowIndex &) = delete;

    void add(const Keyframe &keyframe, MapId mapId);
    void remove(MapKf mapKf);
    void transform(const KeyPointVector &keypoints, DBoW2::BowVector &bowVector, DBoW2::FeatureVector &bowFeatureVector);
    /** Keyframes similar to `kf` (its shared->bowVec), best first; `kf` itself is taken as MapKf{CURRENT_MAP_ID, kf.id}.
     *  mapDB / atlas are not read: the stored BowVectors are the device copies made by add(). */
    std::vector<BowSimilar> getBowSimilar(const MapDB &mapDB, const Atlas &atlas, const Keyframe &kf);

    static constexpr int CURRENT_MAP_ID = 0;

private:
    const odometry::ParametersSlam &parameters;
    sg_ctx *ctx;
    ::sg_vocab *vocab = nullptr;
    ::sg_bowdb *db = nullptr;
    int capacity;
};

// ---- map archives: the cereal binary files Mapper::end writes and loadMapDB reads ----------------
// (mapper.cpp:504-512, mapper_helpers.cpp:958-993; MapDB::serialize mapdb.hpp:83-98, Keyframe::serialize keyframe.hpp:199-213,
//  KeyframeShared::save / load keyframe.hpp:80-105, MapPoint::serialize map_point.hpp:78-93, KeyPoint::serialize
//  key_point.hpp:22-25 -- `octave` is written twice).  cereal itself and the parent project's Eigen / cv::Vec3b adapters
//  (../util/serialization.hpp) are not in the reference tree: the byte layout follows cereal's published binary archive
//  rules and the two adapter layouts named in MapArchiveOptions -- PARITY UNPINNED, no example archive exists offline.
struct MapArchiveOptions {
    // Eigen matrices: raw column-major coefficients; dynamic dimensions are preceded by their extent as 64-bit integers
    // (the widely used cereal adapter).  fixedSizeHeader = true additionally writes rows / cols (int32) before fixed-size
    // matrices (the other common adapter).
    bool fixedSizeHeader = false;
};
/** @return false (with `error` set) on a truncated or inconsistent archive */
bool loadMapArchive(const std::string &path, MapDB &mapDB, std::string *error = nullptr, const MapArchiveOptions &opt = MapArchiveOptions());
bool saveMapArchive(const std::string &path, const MapDB &mapDB, std::string *error = nullptr, const MapArchiveOptions &opt = MapArchiveOptions());

/** The rebuild hook next to loadMapDB's (mapper_helpers.cpp:972-988 rebuilds the BoW vectors and FeatureSearch of every
 *  loaded keyframe): a device-resident descriptor database with one set per keyframe, in std::map (KfId) order, for
 *  sg_match_pairs -- BASELINE configs[4] on a real atlas.  `keyframeIds` receives the KfId of every set.
 *  Also refreshes shared->featureSearch of every keyframe.  The caller owns the database (sg_db_destroy). */
struct ::sg_db *buildDescriptorDatabase(MapDB &mapDB, sg_ctx *ctx, std::vector<KfId> &keyframeIds);

namespace match {
/** openvslam/match_base.h:18-39, evaluated on the GPU for n descriptor pairs. */
void compute_descriptor_distance_32(const std::uint32_t *desc_1, const std::uint32_t *desc_2, int n,
                                    unsigned int *out, sg_ctx *ctx);
}
}  // namespace slam
