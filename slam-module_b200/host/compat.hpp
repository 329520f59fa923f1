// Minimal stand-ins for the parent-project types the reference's hot-path interfaces mention but
// that are NOT part of the reference tree (../tracker/*.hpp, ../odometry/parameters.hpp,
// accelerated-arrays).  Only the members the hot path reads are present, with the reference's names, so the
// adapters in slam_frontend.hpp keep the reference's signatures.  A maintainer integrating libslamgpu into
// the real tree deletes this header and includes the real ones.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "linalg.hpp"

namespace accelerated {
// accelerated::Image as the hot path uses it: width / height / storageType (image_pyramid.cpp:211,
// feature_detector.cpp:79) plus, for CPU images, what accelerated::opencv::ref() would expose.
struct Image {
    enum class StorageType { CPU, GPU };
    int width = 0, height = 0;
    StorageType storageType = StorageType::CPU;
    // CPU: host pointer; GPU: device pointer of the context's GPU
    const std::uint8_t *data = nullptr;
    int stride = 0;   // bytes between rows
};
}  // namespace accelerated

namespace tracker {
// tracker::Feature: id, points[0] (orb_extractor.cpp:90,123), Point{x, y} (key_point.hpp:14)
struct Feature {
    struct Point { float x, y; };
    int id = -1;
    std::array<Point, 2> points{};
    float depth = -1;
};

// tracker::Image: an 8-bit gray frame (orb_extractor.cpp:74, image_pyramid.cpp:69-73)
struct Image {
    int width = 0, height = 0;
    const std::uint8_t *gray = nullptr;   // row-major
    int stride = 0;
    accelerated::Image acc;
    Image() = default;
    Image(const std::uint8_t *data, int w, int h, int rowStride) : width(w), height(h), gray(data), stride(rowStride) {
        acc.width = w; acc.height = h; acc.data = data; acc.stride = rowStride;
    }
    accelerated::Image &getAccImage() { return acc; }
};

// tracker::Camera: isValidPixel on the extraction path (orb_extractor.cpp:101,231); rayToPixel / pixelToRay for the
// projection matchers (keyframe.cpp:408-444).  The default is a pinhole model over all pixels.
class Camera {
public:
    double fx = 1, fy = 1, cx = 0, cy = 0;
    virtual ~Camera() = default;
    virtual bool isValidPixel(double x, double y) const { (void)x; (void)y; return true; }
    bool isValidPixel(const slam::la::Vector2d &p) const { return isValidPixel(p(0), p(1)); }
    virtual bool rayToPixel(const slam::la::Vector3d &ray, slam::la::Vector2d &pix) const {
        if (!(ray(2) > 0)) return false;
        pix = slam::la::Vector2d(fx * (ray(0) / ray(2)) + cx, fy * (ray(1) / ray(2)) + cy);
        return true;
    }
    virtual bool pixelToRay(const slam::la::Vector2d &pix, slam::la::Vector3d &ray) const {
        ray = slam::la::Vector3d((pix(0) - cx) / fx, (pix(1) - cy) / fy, 1.0).normalized();
        return true;
    }
};
}  // namespace tracker

namespace odometry {
// The fields of odometry::ParametersSlam the path reads.
struct ParametersSlam {
    unsigned orbScaleLevels = 8;              // static_settings.cpp:31
    float orbScaleFactor = 1.2f;              // static_settings.cpp:32
    unsigned maxKeypoints = 1000;             // static_settings.cpp:48
    unsigned orbLkTrackLevel = 0;             // orb_extractor.cpp:91
    bool useGpuImagePyramid = true;           // image_pyramid.cpp:211
    std::string slamFeatureDetector = "FAST"; // feature_detector.cpp:38-41
    float loopClosureFeatureMatchLoweRatio = 0.8f;     // keyframe_matcher.cpp:120
    bool requireTringulationForLoopClosures = true;    // keyframe_matcher.cpp:82 (sic)
    float epipolarCheckThresholdDegrees = 0.2f;        // keyframe_matcher.cpp:168
    std::string vocabularyPath;                        // bow_index.cpp:35 (DBoW2 text vocabulary, see loadVocabularyText)
    float bowMinInCommonRatio = 0.8f;         // bow_index.cpp:144
    float bowScoreRatio = 0.75f;              // bow_index.cpp:170
    // upstream OpenVSLAM FAST thresholds of the detector the north star names
    int orbIniFastThreshold = 20, orbMinFastThreshold = 7;
    // batch geometry of the CUDA context (not in the reference: it handles one frame per call)
    int cudaDevice = 0, cudaMaxFrames = 1, cudaMaxTracks = 512;
};
struct ParametersTracker {
    int maxTracks = 200;
    std::string featureDetector = "FAST";
    double gfttMinDistance = 15;
};
struct Parameters {
    ParametersSlam slam;
    ParametersTracker tracker;
};
}  // namespace odometry
