// Host-side mirror of the reference's front-end interfaces, implemented on top of the C ABI of
// libslamgpu.so (include/slamgpu.h).  Same names, argument meaning and error behaviour as the
// reference (asserts, no exceptions, outputs cleared then filled), so that code written against
//   image_pyramid.hpp:16-30, feature_detector.hpp:15-24, orb_extractor.hpp:11-30,
//   static_settings.hpp:8-22, key_point.hpp:11-28, keyframe_matcher.hpp:33-40
// compiles against this header after swapping the include.
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <utility>
#include <vector>

#include "compat.hpp"

struct sg_ctx;
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

// key_point.hpp:11-28 (Eigen::Vector3d bearing -> three doubles; filled later by keyframe.cpp:67)
struct KeyPoint {
    tracker::Feature::Point pt{0, 0};
    float angle = 0;
    int octave = 0;
    std::array<double, 3> bearing{{0, 0, 0}};
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

// The slice of Keyframe / MapDB that matchForLoopClosures reads (keyframe_matcher.cpp:50-158):
// keypoints with descriptors and angles, the map point id of every keypoint and its status.
enum class MapPointStatus { NOT_TRIANGULATED, TRIANGULATED, BAD };
struct MpId { int v = -1; };
struct KfId { int v = -1; };    // id.hpp:48-51
struct MapId { int v = -1; };
struct MapPoint { MapPointStatus status = MapPointStatus::NOT_TRIANGULATED; };
struct MapDB { std::vector<MapPoint> mapPoints; };   // indexed by MpId::v
// bowFeatureVec: DBoW2::FeatureVector = std::map<NodeId, std::vector<unsigned>> (keyframe.hpp, bow_index.cpp:59-93)
// bowVec: DBoW2::BowVector = std::map<WordId, WordValue (double)>
struct KeyframeShared {
    KeyPointVector keyPoints;
    std::map<unsigned, double> bowVec;
    std::map<unsigned, std::vector<unsigned>> bowFeatureVec;
};
struct Keyframe {
    KfId id;
    std::shared_ptr<KeyframeShared> shared;
    std::vector<MpId> mapPoints;   // per keypoint, v == -1: none
};

/** keyframe_matcher.hpp:33-40.  With bowFeatureVec filled in on both keyframes: the reference's node-bucketed
 *  comparison (keyframe_matcher.cpp:65-146); with empty feature vectors every feature is in ONE node (brute force,
 *  BASELINE.json north star).
 *  `matchedMapPoints[i]` = keypoint index of kf2 matched to keypoint i of kf1, or -1.  @return match count */
unsigned int matchForLoopClosures(const Keyframe &kf1, const Keyframe &kf2, const MapDB &mapDB1, const MapDB &mapDB2,
                                  std::vector<int> &matchedMapPoints, const odometry::ParametersSlam &parameters,
                                  sg_ctx *ctx);

/** The same loop on bare keypoint vectors (every feature eligible). */
unsigned int bruteForceMatch(const KeyPointVector &kps1, const KeyPointVector &kps2, std::vector<int> &matches,
                             float loweRatio, bool checkOrientation, sg_ctx *ctx);

// ---- bag of words (bow_index.hpp:21-65) ---------------------------------------------------------------
namespace DBoW2 {
using BowVector = std::map<unsigned, double>;                       // WordId -> WordValue
using FeatureVector = std::map<unsigned, std::vector<unsigned>>;    // NodeId -> feature indices
}  // namespace DBoW2
using Atlas = std::vector<MapDB>;
struct MapKf { MapId mapId; KfId kfId; };
bool operator==(const MapKf &lhs, const MapKf &rhs);
bool operator<(const MapKf &lhs, const MapKf &rhs);
struct BowSimilar { MapKf mapKf; float score; };

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

namespace match {
/** openvslam/match_base.h:18-39, evaluated on the GPU for n descriptor pairs. */
void compute_descriptor_distance_32(const std::uint32_t *desc_1, const std::uint32_t *desc_2, int n,
                                    unsigned int *out, sg_ctx *ctx);
}
}  // namespace slam
