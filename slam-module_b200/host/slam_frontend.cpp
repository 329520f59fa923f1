// Adapters: the reference's ImagePyramid / FeatureDetector / OrbExtractor / matchForLoopClosures
// interfaces on top of the C ABI of libslamgpu.so.  They own nothing but a context handle and
// host-side staging; every computation happens in the CUDA library (there is no CPU fallback:
// if the context cannot be created the adapters abort like the reference's asserts do).
#include "slam_frontend.hpp"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

#include "../../include/slamgpu.h"
#include "slam_matchers.hpp"

namespace slam {
namespace {

[[noreturn]] void die(sg_ctx *ctx, const char *what, int rc) {
    // the reference has no error codes on this path (asserts only): fail loudly
    std::fprintf(stderr, "slam-b200: %s failed (%d): %s\n", what, rc, sg_last_error(ctx));
    std::abort();
}
#define SG_CHECK(ctx, call)                          \
    do {                                             \
        const int rc_ = (call);                      \
        if (rc_ != SG_OK) die((ctx), #call, rc_);    \
    } while (0)

}  // namespace

// ---- StaticSettings (static_settings.cpp:9-60) ---------------------------------------------------------
StaticSettings::StaticSettings(const odometry::Parameters &p) : parameters(p) {
    const unsigned n = p.slam.orbScaleLevels;
    scaleFactors.assign(n, 1.0f);
    levelSigmaSq.assign(n, 1.0f);
    for (unsigned l = 1; l < n; ++l) {
        scaleFactors[l] = p.slam.orbScaleFactor * scaleFactors[l - 1];   // float products
        levelSigmaSq[l] = scaleFactors[l] * scaleFactors[l];
    }
}

std::vector<std::size_t> StaticSettings::maxNumberOfKeypointsPerLevel() const {
    const auto &s = parameters.slam;
    std::vector<std::size_t> budget(s.orbScaleLevels, 0);
    const double inv = 1.0 / s.orbScaleFactor;
    double want = s.maxKeypoints * (1.0 - inv) / (1.0 - std::pow(inv, static_cast<double>(s.orbScaleLevels)));
    unsigned total = 0;
    for (unsigned l = 0; l + 1 < s.orbScaleLevels; ++l, want *= inv) {
        budget[l] = static_cast<std::size_t>(std::round(want));
        total += static_cast<unsigned>(budget[l]);
    }
    budget[s.orbScaleLevels - 1] = static_cast<std::size_t>(std::max(static_cast<int>(s.maxKeypoints) - static_cast<int>(total), 0));
    return budget;
}

// ---- shared CUDA context ----------------------------------------------------------------------------
struct CudaFrontend {
    sg_ctx *ctx = nullptr;
    sg_params params{};
    int levels = 0, cap = 0;
    std::vector<float> scales;
    std::vector<int> widths, heights;
    ~CudaFrontend() { if (ctx) sg_destroy(ctx); }
};

std::shared_ptr<CudaFrontend> cudaFrontend(const StaticSettings &settings, int width, int height) {
    static std::mutex mu;
    static std::map<std::tuple<const StaticSettings *, int, int>, std::weak_ptr<CudaFrontend>> live;
    std::lock_guard<std::mutex> lock(mu);
    auto &slot = live[std::make_tuple(&settings, width, height)];
    if (auto fe = slot.lock()) return fe;
    const auto &s = settings.parameters.slam;
    auto fe = std::make_shared<CudaFrontend>();
    sg_params &p = fe->params;
    p.width = width; p.height = height;
    p.levels = static_cast<int>(s.orbScaleLevels);
    p.scale_factor = s.orbScaleFactor;
    p.max_keypoints = static_cast<int>(s.maxKeypoints);
    p.ini_fast_thr = s.orbIniFastThreshold; p.min_fast_thr = s.orbMinFastThreshold;
    p.max_frames = std::max(1, s.cudaMaxFrames);
    p.max_tracks = std::max(0, s.cudaMaxTracks);
    p.track_level = static_cast<int>(s.orbLkTrackLevel);
    const int rc = sg_create(s.cudaDevice, &p, &fe->ctx);
    if (rc != SG_OK) die(nullptr, "sg_create", rc);
    fe->levels = p.levels;
    fe->cap = sg_keypoint_capacity(fe->ctx);
    fe->scales.resize(p.levels); fe->widths.resize(p.levels); fe->heights.resize(p.levels);
    SG_CHECK(fe->ctx, sg_get_geometry(fe->ctx, fe->scales.data(), fe->widths.data(), fe->heights.data(), nullptr, nullptr));
    // the adapter's StaticSettings and the library's tables must agree bit for bit
    for (int l = 0; l < p.levels; ++l) assert(fe->scales[l] == settings.scaleFactors[l]);
    slot = fe;
    return fe;
}

sg_ctx *cudaContext(const std::shared_ptr<CudaFrontend> &fe) { return fe ? fe->ctx : nullptr; }

namespace {

// ---- ImagePyramid -----------------------------------------------------------------------------------
class CudaImagePyramid : public ImagePyramid {
public:
    CudaImagePyramid(const StaticSettings &settings, tracker::Image &model)
        : fe(cudaFrontend(settings, model.width, model.height)), cpu(fe->levels), blurred(fe->levels), gpu(fe->levels),
          cpuData(fe->levels), blurData(fe->levels), cpuValid(fe->levels, false), blurValid(fe->levels, false) {}

    void update(tracker::Image &image) final {
        assert(image.width == fe->params.width && image.height == fe->params.height);
        auto &acc = image.getAccImage();
        if (acc.storageType == accelerated::Image::StorageType::GPU)
            SG_CHECK(fe->ctx, sg_pyramid_update_device(fe->ctx, acc.data, acc.stride, (size_t)acc.stride * acc.height, 1));
        else
            SG_CHECK(fe->ctx, sg_pyramid_update(fe->ctx, image.gray, image.stride, 0, 1));
        std::fill(cpuValid.begin(), cpuValid.end(), false);
        std::fill(blurValid.begin(), blurValid.end(), false);
        for (int l = 0; l < fe->levels; ++l) {
            const std::uint8_t *d = nullptr;
            int pitch = 0;
            SG_CHECK(fe->ctx, sg_pyramid_device_plane(fe->ctx, l, 0, &d, &pitch, nullptr));
            gpu[l].width = fe->widths[l]; gpu[l].height = fe->heights[l];
            gpu[l].storageType = accelerated::Image::StorageType::GPU;
            gpu[l].data = d; gpu[l].stride = pitch;
        }
        updated = true;
    }
    std::size_t numberOfLevels() const final { return fe->levels; }
    bool isGpu() const final { return true; }

    accelerated::Image &getLevel(std::size_t level) final { return fetch(level, false); }
    accelerated::Image &getBlurredLevel(std::size_t level) final { return fetch(level, true); }
    accelerated::Image &getGpuLevel(std::size_t level) final {
        assert(updated && level < gpu.size());
        return gpu[level];
    }

    void debugVisualize(std::vector<std::uint8_t> &target, int &width, int &height) final {
        // levels side by side, top aligned (the reference draws the same layout into a cv::Mat)
        width = 0; height = fe->heights[0];
        for (int l = 0; l < fe->levels; ++l) width += fe->widths[l];
        target.assign((size_t)width * height, 0);
        int x0 = 0;
        for (int l = 0; l < fe->levels; ++l) {
            const accelerated::Image &im = getLevel(l);
            for (int y = 0; y < im.height; ++y)
                std::memcpy(&target[(size_t)y * width + x0], im.data + (size_t)y * im.stride, im.width);
            x0 += im.width;
        }
    }

    std::shared_ptr<CudaFrontend> fe;
    bool updated = false;

private:
    accelerated::Image &fetch(std::size_t level, bool blur) {
        assert(updated && level < cpu.size());
        auto &img = blur ? blurred[level] : cpu[level];
        auto &buf = blur ? blurData[level] : cpuData[level];
        auto valid = blur ? blurValid[level] : cpuValid[level];
        if (!valid) {
            const int w = fe->widths[level], h = fe->heights[level];
            buf.resize((size_t)w * h);
            SG_CHECK(fe->ctx, sg_pyramid_download(fe->ctx, 0, (int)level, blur ? 1 : 0, buf.data(), w));
            img.width = w; img.height = h; img.storageType = accelerated::Image::StorageType::CPU;
            img.data = buf.data(); img.stride = w;
            (blur ? blurValid : cpuValid)[level] = true;
        }
        return img;   // valid until the next update(), like the reference's lazily created refs
    }
    std::vector<accelerated::Image> cpu, blurred, gpu;
    std::vector<std::vector<std::uint8_t>> cpuData, blurData;
    std::vector<bool> cpuValid, blurValid;
};

// ---- FeatureDetector --------------------------------------------------------------------------------
class CudaFeatureDetector : public FeatureDetector {
public:
    CudaFeatureDetector(const StaticSettings &settings, tracker::Image &model)
        : fe(cudaFrontend(settings, model.width, model.height)) {}

    std::size_t detect(ImagePyramid &imagePyramid, std::vector<KeyPointVector> &keypointsPerLevel) final {
        auto *pyr = dynamic_cast<CudaImagePyramid *>(&imagePyramid);
        assert(pyr && pyr->fe == fe && "the CUDA detector works on the CUDA pyramid of the same settings");
        (void)pyr;
        SG_CHECK(fe->ctx, sg_detect(fe->ctx));
        keypointsPerLevel.resize(fe->levels);
        std::size_t total = 0;
        std::vector<int> xs(fe->cap), ys(fe->cap);
        for (int l = 0; l < fe->levels; ++l) {
            int n = 0;
            SG_CHECK(fe->ctx, sg_detect_download(fe->ctx, 0, l, xs.data(), ys.data(), nullptr, fe->cap, &n));
            auto &out = keypointsPerLevel[l];
            out.clear();
            out.reserve(n);
            for (int i = 0; i < n; ++i) {
                KeyPoint kp;
                kp.pt = {static_cast<float>(xs[i]), static_cast<float>(ys[i])};   // level coordinates
                kp.angle = 0;                                                     // computed elsewhere
                kp.octave = l;
                out.push_back(kp);
            }
            total += out.size();
        }
        return total;
    }

private:
    std::shared_ptr<CudaFrontend> fe;
};

// ---- OrbExtractor -----------------------------------------------------------------------------------
class CudaOrbExtractor : public OrbExtractor {
public:
    explicit CudaOrbExtractor(const StaticSettings &s) : settings(s), parameters(s.parameters.slam) {}

    void detectAndExtract(tracker::Image &img, const tracker::Camera &camera, const std::vector<tracker::Feature> &tracks,
                          KeyPointVector &keypts, std::vector<int> &keyptTrackIds) final {
        ensure(img);
        keypts.clear();
        keyptTrackIds.clear();
        // tracker points: the camera test of orb_extractor.cpp:101 is applied here, the margin test in the library
        // the context's track capacity is fixed at creation: grow it (a new context) rather than drop tracker points
        if ((int)tracks.size() > fe->params.max_tracks) regrow((int)tracks.size(), img);
        const int T = fe->params.max_tracks;
        int nTracks = 0;
        trackXy.clear(); trackIds.clear();
        for (const auto &track : tracks) {
            const auto &pt = track.points[0];
            if (nTracks < T && camera.isValidPixel(pt.x, pt.y)) {
                trackXy.push_back(pt.x); trackXy.push_back(pt.y);
                trackIds.push_back(track.id);
                ++nTracks;
            }
        }
        trackXy.resize(2 * (size_t)std::max(T, 1));
        trackIds.resize((size_t)std::max(T, 1));
        sg_keypoints out = outputs(1);
        SG_CHECK(fe->ctx, sg_extract(fe->ctx, img.gray, img.stride, 0, 1, nTracks ? trackXy.data() : nullptr,
                                     nTracks ? trackIds.data() : nullptr, nTracks ? &nTracks : nullptr, &out));
        assemble(0, camera, keypts, &keyptTrackIds);
    }

    void detectAndExtractBatch(const std::vector<tracker::Image *> &imgs, const tracker::Camera &camera,
                               std::vector<KeyPointVector> &keyPoints) final {
        keyPoints.clear();
        if (imgs.empty()) return;
        ensure(*imgs[0]);
        const int n = (int)imgs.size(), w = fe->params.width, h = fe->params.height;
        assert(n <= fe->params.max_frames);
        stage.resize((size_t)n * w * h);
        for (int f = 0; f < n; ++f) {
            assert(imgs[f]->width == w && imgs[f]->height == h);
            for (int y = 0; y < h; ++y)
                std::memcpy(&stage[((size_t)f * h + y) * w], imgs[f]->gray + (size_t)y * imgs[f]->stride, w);
        }
        sg_keypoints out = outputs(n);
        SG_CHECK(fe->ctx, sg_extract(fe->ctx, stage.data(), w, (size_t)w * h, n, nullptr, nullptr, nullptr, &out));
        keyPoints.resize(n);
        for (int f = 0; f < n; ++f) assemble(f, camera, keyPoints[f], nullptr);
    }

    void debugVisualize(const tracker::Image &img, std::vector<std::uint8_t> &target, int &width, int &height,
                        VisualizationMode mode) const final {
        assert(mode == VisualizationMode::IMAGE_PYRAMID);
        (void)img; (void)mode;
        assert(imagePyramid);
        imagePyramid->debugVisualize(target, width, height);
    }

private:
    void ensure(tracker::Image &img) {
        if (!fe) {
            fe = cudaFrontend(settings, img.width, img.height);
            imagePyramid = ImagePyramid::build(settings, img);
        }
        assert(img.width == fe->params.width && img.height == fe->params.height);
    }
    // more tracker points than the context was sized for (the reference keeps every valid track, orb_extractor.cpp:89-124):
    // a private context with room for them replaces the shared one for this extractor
    void regrow(int nTracks, tracker::Image &img) {
        int cap = std::max(64, fe->params.max_tracks);
        while (cap < nTracks) cap *= 2;
        std::fprintf(stderr, "slam-b200: %d tracker points exceed cudaMaxTracks = %d: re-creating the extractor's context for %d\n",
                     nTracks, fe->params.max_tracks, cap);
        auto grown = std::make_shared<CudaFrontend>();
        grown->params = fe->params;
        grown->params.max_tracks = cap;
        const int rc = sg_create(settings.parameters.slam.cudaDevice, &grown->params, &grown->ctx);
        if (rc != SG_OK) die(nullptr, "sg_create", rc);
        grown->levels = fe->levels;
        grown->cap = sg_keypoint_capacity(grown->ctx);
        grown->scales = fe->scales; grown->widths = fe->widths; grown->heights = fe->heights;
        fe = grown;
        (void)img;
    }
    sg_keypoints outputs(int n) {
        const size_t m = (size_t)n * fe->cap;
        x.resize(m); y.resize(m); angle.resize(m); octave.resize(m); desc.resize(8 * m); trackId.resize(m);
        count.resize(n);
        sg_keypoints o{};
        o.x = x.data(); o.y = y.data(); o.angle = angle.data(); o.octave = octave.data(); o.desc = desc.data();
        o.track_id = trackId.data(); o.count = count.data();
        return o;
    }
    // output assembly of orb_extractor.cpp:120-124,153-162 + dropInvalidKeypoints (:221-237)
    void assemble(int f, const tracker::Camera &camera, KeyPointVector &keypts, std::vector<int> *ids) const {
        const size_t base = (size_t)f * fe->cap;
        keypts.reserve(count[f]);
        for (int i = 0; i < count[f]; ++i) {
            const size_t k = base + i;
            if (trackId[k] < 0 && !camera.isValidPixel(x[k], y[k])) continue;
            KeyPoint kp;
            kp.pt = {x[k], y[k]};
            kp.angle = angle[k];
            kp.octave = octave[k];
            std::memcpy(kp.descriptor.data(), &desc[8 * k], 32);
            keypts.push_back(kp);
            if (ids) ids->push_back(trackId[k]);
        }
    }

    const StaticSettings &settings;
    const odometry::ParametersSlam &parameters;
    std::shared_ptr<CudaFrontend> fe;
    std::unique_ptr<ImagePyramid> imagePyramid;
    std::vector<float> x, y, angle, trackXy;
    std::vector<std::int32_t> octave, trackId, count, trackIds;
    std::vector<std::uint32_t> desc;
    std::vector<std::uint8_t> stage;
};

}  // namespace

std::unique_ptr<ImagePyramid> ImagePyramid::build(const StaticSettings &s, tracker::Image &img) {
    // image_pyramid.cpp:209-219 picks CPU or GPU here; this build has one implementation and no CPU fallback
    return std::unique_ptr<ImagePyramid>(new CudaImagePyramid(s, img));
}
ImagePyramid::~ImagePyramid() = default;

std::unique_ptr<FeatureDetector> FeatureDetector::build(const StaticSettings &s, tracker::Image &img) {
    return std::unique_ptr<FeatureDetector>(new CudaFeatureDetector(s, img));
}
FeatureDetector::~FeatureDetector() = default;

std::unique_ptr<OrbExtractor> OrbExtractor::build(const StaticSettings &s) {
    return std::unique_ptr<OrbExtractor>(new CudaOrbExtractor(s));
}

// ---- matching ---------------------------------------------------------------------------------------
namespace {
unsigned int matchSubset(const KeyPointVector &kps1, const std::vector<int> &idx1, const KeyPointVector &kps2,
                         const std::vector<int> &idx2, std::vector<int> &matched, float ratio, bool checkOrientation,
                         sg_ctx *ctx) {
    const int nA = (int)idx1.size(), nB = (int)idx2.size();
    if (nA == 0 || nB == 0) return 0;
    std::vector<std::uint32_t> dA(8 * (size_t)nA), dB(8 * (size_t)nB);
    std::vector<float> aA(nA), aB(nB);
    for (int i = 0; i < nA; ++i) { std::memcpy(&dA[8 * (size_t)i], kps1[idx1[i]].descriptor.data(), 32); aA[i] = kps1[idx1[i]].angle; }
    for (int i = 0; i < nB; ++i) { std::memcpy(&dB[8 * (size_t)i], kps2[idx2[i]].descriptor.data(), 32); aB[i] = kps2[idx2[i]].angle; }
    sg_match_params mp{};
    mp.ratio = ratio; mp.thr = HAMMING_DIST_THR_LOW; mp.check_orientation = checkOrientation ? 1 : 0; mp.ratio_is_double = 0;
    std::vector<std::int32_t> m(nA);
    std::uint32_t n = 0;
    SG_CHECK(ctx, sg_match_bruteforce(ctx, dA.data(), aA.data(), nA, dB.data(), aB.data(), nB, &mp, m.data(), &n));
    for (int i = 0; i < nA; ++i)
        if (m[i] >= 0) matched[idx1[i]] = idx2[m[i]];
    return n;
}
}  // namespace

unsigned int matchForLoopClosures(const Keyframe &kf1, const Keyframe &kf2, const MapDB &mapDB1, const MapDB &mapDB2,
                                  std::vector<int> &matchedMapPoints, const odometry::ParametersSlam &parameters,
                                  sg_ctx *ctx) {
    const auto &kps1 = kf1.shared->keyPoints, &kps2 = kf2.shared->keyPoints;
    matchedMapPoints.resize(kps1.size(), -1);                       // keyframe_matcher.cpp:61
    // the map-point filters of :79-84 (kf1) and :93-96 (kf2)
    std::vector<std::uint8_t> elig1(kps1.size(), 0), elig2(kps2.size(), 0);
    for (size_t i = 0; i < kps1.size(); ++i) {
        const MpId id = kf1.mapPoints.at(i);
        if (id.v == -1) continue;
        if (parameters.requireTringulationForLoopClosures && mapDB1.mapPoints.at(id).status != MapPointStatus::TRIANGULATED) continue;
        elig1[i] = 1;
    }
    for (size_t i = 0; i < kps2.size(); ++i) {
        const MpId id = kf2.mapPoints.at(i);
        if (id.v == -1 || mapDB2.mapPoints.at(id).status != MapPointStatus::TRIANGULATED) continue;
        elig2[i] = 1;
    }
    const auto &fv1 = kf1.shared->bowFeatureVec, &fv2 = kf2.shared->bowFeatureVec;
    // keyframe_matcher.cpp:65-147 walks the two feature vectors: with empty vectors the reference finds nothing.  (The
    // single-node brute-force form of the north star is bruteForceMatch below, an explicit call.)
    if (fv1.empty() || fv2.empty()) return 0;
    // node of every feature; DBoW2 lists the features of a node in index order (bow_index.cpp:59-93), which is
    // the order the library assumes inside a node
    auto nodes = [](const std::map<unsigned, std::vector<unsigned>> &fv, size_t n) {
        std::vector<std::int32_t> node(n, -1);
        for (const auto &kv : fv) {
            assert(std::is_sorted(kv.second.begin(), kv.second.end()));
            for (unsigned i : kv.second) node.at(i) = (std::int32_t)kv.first;
        }
        return node;
    };
    const auto node1 = nodes(fv1, kps1.size()), node2 = nodes(fv2, kps2.size());
    std::vector<std::uint32_t> d1(8 * kps1.size()), d2(8 * kps2.size());
    std::vector<float> a1(kps1.size()), a2(kps2.size());
    for (size_t i = 0; i < kps1.size(); ++i) { std::memcpy(&d1[8 * i], kps1[i].descriptor.data(), 32); a1[i] = kps1[i].angle; }
    for (size_t i = 0; i < kps2.size(); ++i) { std::memcpy(&d2[8 * i], kps2[i].descriptor.data(), 32); a2[i] = kps2[i].angle; }
    sg_match_params mp{};
    mp.ratio = parameters.loopClosureFeatureMatchLoweRatio; mp.thr = HAMMING_DIST_THR_LOW; mp.check_orientation = 1;
    std::vector<std::int32_t> m(std::max<size_t>(kps1.size(), 1));
    std::uint32_t n = 0;
    SG_CHECK(ctx, sg_match_bow(ctx, d1.data(), a1.data(), node1.data(), elig1.data(), (int)kps1.size(), d2.data(), a2.data(),
                               node2.data(), elig2.data(), (int)kps2.size(), &mp, m.data(), &n));
    for (size_t i = 0; i < kps1.size(); ++i) matchedMapPoints[i] = m[i];
    return n;
}

unsigned int bruteForceMatch(const KeyPointVector &kps1, const KeyPointVector &kps2, std::vector<int> &matches,
                             float loweRatio, bool checkOrientation, sg_ctx *ctx) {
    matches.assign(kps1.size(), -1);
    std::vector<int> idx1(kps1.size()), idx2(kps2.size());
    for (size_t i = 0; i < idx1.size(); ++i) idx1[i] = (int)i;
    for (size_t i = 0; i < idx2.size(); ++i) idx2[i] = (int)i;
    return matchSubset(kps1, idx1, kps2, idx2, matches, loweRatio, checkOrientation, ctx);
}

// ---- data-model slice (map_point.cpp:11-17,65-73,117-184; keyframe.cpp:248-307,408-424) -------------------------
MapPoint::MapPoint(MpId id, KfId keyframeId, KpId keyPointId) : id(id), referenceKeyframe(keyframeId) {
    assert(keyframeId.v != -1);
    addObservation(keyframeId, keyPointId);
}
void MapPoint::addObservation(KfId keyframeId, KpId keyPointId) {
    assert(!observations.count(keyframeId));
    observations.emplace(keyframeId, keyPointId);
}
void MapPoint::eraseObservation(KfId keyframeId) {
    assert(observations.count(keyframeId));
    observations.erase(keyframeId);
}
// map_point.cpp:117-158: every observation of this point moves to `otherMp` (or is dropped where otherMp is seen too)
void MapPoint::replaceWith(MapDB &mapDB, MapPoint &otherMp) {
    assert(id.v != -1 && otherMp.id.v != -1 && mapDB.mapPoints.count(id) && mapDB.mapPoints.count(otherMp.id));
    if (otherMp.id == id) return;
    if (trackId.v != -1) {
        if (otherMp.trackId.v == -1) {
            mapDB.trackIdToMapPoint.at(trackId) = otherMp.id;
            otherMp.trackId = trackId;
        } else {
            mapDB.trackIdToMapPoint.erase(trackId);
        }
    }
    for (const auto &obs : observations) {
        Keyframe &kf = *mapDB.keyframes.at(obs.first);
        kf.keyPointToTrackId.erase(obs.second);
        if (!otherMp.observations.count(obs.first)) {
            kf.mapPoints[obs.second.v] = otherMp.id;
            otherMp.addObservation(obs.first, obs.second);
        } else {
            kf.mapPoints[obs.second.v] = MpId(-1);
        }
    }
    status = MapPointStatus::BAD;
    mapDB.mapPoints.erase(id);     // `this` dies here, exactly as in the reference
}
// map_point.cpp:174-183
int MapPoint::predictScaleLevel(float dist, const StaticSettings &settings) const {
    const float ratio = maxViewingDistance / dist;
    const int scale = std::ceil(std::log(ratio) / std::log(settings.parameters.slam.orbScaleFactor));
    return std::min(std::max(scale, 0), static_cast<int>(settings.scaleFactors.size() - 1));
}

void Keyframe::addObservation(MpId mapPointId, KpId keyPointId) {
    assert(mapPoints[keyPointId.v].v == -1);
    mapPoints[keyPointId.v] = mapPointId;
}
void Keyframe::eraseObservation(MpId mapPointId) {
    const auto it = std::find(mapPoints.begin(), mapPoints.end(), mapPointId);
    assert(it != mapPoints.end());
    it->v = -1;
    keyPointToTrackId.erase(KpId((int)std::distance(mapPoints.begin(), it)));
}
Vector3d Keyframe::cameraCenter() const {   // keyframe.hpp:23-25: -R^T t
    return -poseCW.topLeftCorner<3, 3>().transpose() * poseCW.block<3, 1>(0, 3);
}
bool Keyframe::reproject(const Vector3d &pointW, Vector2f &reprojected) const {
    float unused = 0;
    Vector2d pix;
    const bool visible = reprojectToImage(*shared->camera, poseCW.topLeftCorner<3, 3>(), poseCW.block<3, 1>(0, 3), pointW, pix, unused);
    reprojected = Vector2f((float)pix(0), (float)pix(1));
    return visible;
}
void Keyframe::getFeaturesAround(const Vector2f &point, float r, std::vector<size_t> &output) {
    assert(shared->featureSearch);
    shared->featureSearch->getFeaturesAround(point(0), point(1), r, output);
}
bool reprojectToImage(const tracker::Camera &camera, const Matrix3d &rot_cw, const Vector3d &trans_cw, const Vector3d &pos_w,
                      Vector2d &reproj, float &x_right) {
    const Vector3d pos_c = rot_cw * pos_w + trans_cw;
    x_right = 0.0;
    if (!camera.rayToPixel(pos_c, reproj)) return false;
    if (!camera.isValidPixel(reproj)) return false;
    x_right = (float)reproj(0);
    return true;
}

// ---- FeatureSearch (feature_search.cpp:22-48): index order from the library, the radius walk over it ---------------
namespace {
class IndexedFeatureSearch : public FeatureSearch {
public:
    explicit IndexedFeatureSearch(const KeyPointVector &kps) : x(kps.size()), y(kps.size()), order(kps.size()) {
        for (size_t i = 0; i < kps.size(); ++i) { x[i] = kps[i].pt.x; y[i] = kps[i].pt.y; }
        if (!kps.empty()) sg_feature_index(x.data(), y.data(), (int)kps.size(), order.data());
        sortedY.resize(kps.size());
        for (size_t p = 0; p < order.size(); ++p) sortedY[p] = y[(size_t)order[p]];
    }
    void getFeaturesAround(float qx, float qy, float r, std::vector<size_t> &output) const final {
        output.clear();
        for (auto it = std::lower_bound(sortedY.begin(), sortedY.end(), qy - r); it != sortedY.end() && *it <= qy + r; ++it) {
            const size_t i = (size_t)order[(size_t)(it - sortedY.begin())];
            const float dx = qx - x[i], dy = qy - y[i];
            if (dx * dx + dy * dy < r * r) output.push_back(i);
        }
    }
private:
    std::vector<float> x, y, sortedY;
    std::vector<std::int32_t> order;
};

// essential_solver.cc:139-162 (create_E_21) on the adapter's matrix types
struct AdapterTypes {
    using Vector2f = slam::Vector2f;
    using Vector2d = slam::Vector2d;
    using Vector3d = slam::Vector3d;
    using Matrix3d = slam::Matrix3d;
    using Matrix4d = slam::Matrix4d;
    static Matrix3d createE21(const Matrix3d &rot_1w, const Vector3d &trans_1w, const Matrix3d &rot_2w, const Vector3d &trans_2w) {
        const Matrix3d rot_21 = rot_2w * rot_1w.transpose();
        const Vector3d t = -rot_21 * trans_1w + trans_2w;
        Matrix3d skew;
        skew << 0, -t(2), t(1), t(2), 0, -t(0), -t(1), t(0), 0;
        return skew * rot_21;
    }
};
}  // namespace

std::unique_ptr<FeatureSearch> FeatureSearch::create(const KeyPointVector &kps) {
    return std::unique_ptr<FeatureSearch>(new IndexedFeatureSearch(kps));
}

// ---- the candidate-list matchers: slam_matchers.hpp instantiated for the types of this header -----------------------
std::vector<std::pair<KpId, KpId>> matchForTriangulationDBoW(Keyframe &kf1, Keyframe &kf2, const StaticSettings &settings, sg_ctx *ctx) {
    return cuda_matchers::matchForTriangulationDBoW<AdapterTypes, KpId>(kf1, kf2, settings, ctx);
}
int searchByProjection(Keyframe &kf, const std::vector<MpId> &mps, MapDB &mapDB, ViewerDataPublisher *, const float threshold,
                       const StaticSettings &settings, sg_ctx *ctx) {
    return cuda_matchers::searchByProjection<AdapterTypes>(kf, mps, mapDB, threshold, settings, ctx);
}
template <typename T>
unsigned int replaceDuplication(Keyframe &kf, const T &mapPoints, const float margin, MapDB &mapDB, const StaticSettings &settings,
                                sg_ctx *ctx) {
    return cuda_matchers::replaceDuplication<AdapterTypes>(kf, mapPoints, margin, mapDB, settings, ctx);
}
template unsigned int replaceDuplication<std::vector<MpId>>(Keyframe &, const std::vector<MpId> &, const float, MapDB &,
                                                            const StaticSettings &, sg_ctx *);
template unsigned int replaceDuplication<std::set<MpId>>(Keyframe &, const std::set<MpId> &, const float, MapDB &,
                                                         const StaticSettings &, sg_ctx *);
void matchMapPointsSim3(Keyframe &kf1, Keyframe &kf2, const Matrix4d &transform12, MapDB &mapDB,
                        std::vector<std::pair<MpId, MpId>> &matches, const StaticSettings &settings, sg_ctx *ctx) {
    cuda_matchers::matchMapPointsSim3<AdapterTypes>(kf1, kf2, transform12, mapDB, matches, settings, ctx);
}
void updateDescriptors(MapDB &mapDB, const std::vector<MpId> &mapPoints, sg_ctx *ctx) {
    cuda_matchers::updateDescriptors(mapDB, mapPoints, ctx);
}

// ---- DBoW2 text vocabulary (bow_index.cpp:11-19 -> TemplatedVocabulary::loadFromTextFile) ---------------------------
bool loadVocabularyText(const std::string &path, BowVocabulary &out) {
    std::FILE *f = std::fopen(path.c_str(), "r");
    if (!f) return false;
    int k = 0, L = 0, scoring = 0, weighting = 0;
    if (std::fscanf(f, "%d %d %d %d", &k, &L, &scoring, &weighting) != 4 || k < 0 || k > 20 || L < 1 || L > 10 || scoring < 0
        || scoring > 5 || weighting < 0 || weighting > 3) { std::fclose(f); return false; }
    // node ids are the line numbers (root = 0); children keep file order, which is DBoW2's tie order in the descent
    std::vector<std::vector<std::int32_t>> children(1);
    std::vector<std::uint32_t> desc(8, 0);
    std::vector<double> weight(1, 0.0);
    std::vector<std::int32_t> word(1, -1);
    int pid = 0, leaf = 0, nWords = 0;
    while (std::fscanf(f, "%d %d", &pid, &leaf) == 2) {
        const std::int32_t nid = (std::int32_t)children.size();
        if (pid < 0 || pid >= nid) { std::fclose(f); return false; }
        children.emplace_back();
        children[(size_t)pid].push_back(nid);
        std::uint8_t bytes[32];
        for (int i = 0; i < 32; ++i) { int v = 0; if (std::fscanf(f, "%d", &v) != 1) { std::fclose(f); return false; } bytes[i] = (std::uint8_t)v; }
        std::uint32_t w8[8];
        std::memcpy(w8, bytes, 32);
        desc.insert(desc.end(), w8, w8 + 8);
        double w = 0;
        if (std::fscanf(f, "%lf", &w) != 1) { std::fclose(f); return false; }
        weight.push_back(w);
        word.push_back(leaf > 0 ? nWords++ : -1);
    }
    std::fclose(f);
    out = BowVocabulary();
    out.levels = L;
    out.childOff.push_back(0);
    for (const auto &c : children) {
        out.childIds.insert(out.childIds.end(), c.begin(), c.end());
        out.childOff.push_back((std::int32_t)out.childIds.size());
    }
    out.nodeWord = word;
    out.nodeDescriptor = desc;
    out.nodeWeight = weight;
    return nWords > 0;
}

// ---- BowIndex (bow_index.cpp:31-188) ----------------------------------------------------------------------
bool operator==(const MapKf &lhs, const MapKf &rhs) { return lhs.mapId.v == rhs.mapId.v && lhs.kfId.v == rhs.kfId.v; }
bool operator<(const MapKf &lhs, const MapKf &rhs) {
    if (lhs.mapId.v == rhs.mapId.v) return lhs.kfId.v < rhs.kfId.v;
    return lhs.mapId.v < rhs.mapId.v;
}

BowIndex::BowIndex(const odometry::ParametersSlam &p, const BowVocabulary &v, sg_ctx *c, int maxKeyframes)
    : parameters(p), ctx(c), capacity(maxKeyframes) {
    assert(!v.nodeWord.empty() && v.childOff.size() == v.nodeWord.size() + 1);
    SG_CHECK(ctx, sg_vocab_create(ctx, v.childOff.data(), v.childIds.data(), v.nodeDescriptor.data(), v.nodeWeight.data(),
                                  v.nodeWord.data(), (int)v.nodeWord.size(), v.levels, &vocab));
    SG_CHECK(ctx, sg_bowdb_create(ctx, maxKeyframes, 4096, &db));
}

BowIndex::~BowIndex() {
    sg_bowdb_destroy(db);
    sg_vocab_destroy(vocab);
}

namespace {
void flatten(const DBoW2::BowVector &v, std::vector<std::uint32_t> &words, std::vector<double> &values) {
    words.clear(); values.clear();
    for (const auto &e : v) { words.push_back(e.first); values.push_back(e.second); }
}
}  // namespace

void BowIndex::add(const Keyframe &keyframe, MapId mapId) {
    std::vector<std::uint32_t> words;
    std::vector<double> values;
    flatten(keyframe.shared->bowVec, words, values);
    SG_CHECK(ctx, sg_bowdb_add(ctx, db, mapId.v, keyframe.id.v, words.data(), values.data(), (int)words.size()));
}

void BowIndex::remove(MapKf mapKf) { SG_CHECK(ctx, sg_bowdb_remove(ctx, db, mapKf.mapId.v, mapKf.kfId.v)); }

void BowIndex::transform(const KeyPointVector &keypoints, DBoW2::BowVector &bowVector, DBoW2::FeatureVector &bowFeatureVector) {
    bowVector.clear();
    bowFeatureVector.clear();
    const int n = (int)keypoints.size();
    if (n == 0) return;
    std::vector<std::uint32_t> desc(8 * (size_t)n), vecWord(n);
    std::vector<std::int32_t> word(n), node(n);
    std::vector<double> weight(n), vecValue(n);
    for (int i = 0; i < n; ++i) std::memcpy(&desc[8 * (size_t)i], keypoints[i].descriptor.data(), 32);
    const int levelsUp = 4;                                                         // bow_index.cpp:85
    SG_CHECK(ctx, sg_bow_transform(ctx, vocab, desc.data(), n, levelsUp, word.data(), weight.data(), node.data()));
    int nWords = 0;
    SG_CHECK(ctx, sg_bow_vector(ctx, word.data(), weight.data(), n, vecWord.data(), vecValue.data(), &nWords));
    for (int i = 0; i < nWords; ++i) bowVector.emplace_hint(bowVector.end(), vecWord[i], vecValue[i]);
    for (int i = 0; i < n; ++i)                                                      // fv.addFeature(nid, i_feature), w > 0 only
        if (weight[i] > 0) bowFeatureVector[(unsigned)node[i]].push_back((unsigned)i);
}

std::vector<BowSimilar> BowIndex::getBowSimilar(const MapDB &, const Atlas &, const Keyframe &kf) {
    std::vector<std::uint32_t> words;
    std::vector<double> values;
    flatten(kf.shared->bowVec, words, values);
    const int cap = std::max(sg_bowdb_size(db), 1);
    std::vector<std::int32_t> maps(cap), kfs(cap);
    std::vector<float> scores(cap);
    int n = 0;
    SG_CHECK(ctx, sg_bow_similar(ctx, db, words.data(), values.data(), (int)words.size(), CURRENT_MAP_ID, kf.id.v,
                                 parameters.bowMinInCommonRatio, parameters.bowScoreRatio, maps.data(), kfs.data(), scores.data(),
                                 cap, &n));
    std::vector<BowSimilar> similar((size_t)n);
    for (int i = 0; i < n; ++i) { similar[i].mapKf.mapId.v = maps[i]; similar[i].mapKf.kfId.v = kfs[i]; similar[i].score = scores[i]; }
    return similar;
}

namespace match {
void compute_descriptor_distance_32(const std::uint32_t *desc_1, const std::uint32_t *desc_2, int n, unsigned int *out,
                                    sg_ctx *ctx) {
    SG_CHECK(ctx, sg_hamming(ctx, desc_1, desc_2, n, out));
}
}  // namespace match

}  // namespace slam
