// Stand-in for the parent project's generated parameter struct (../codegen/output/cmd_parameters.hpp is absent from the
// reference tree).  Only the fields the compiled reference files read; C types are this shim's ASSUMPTION (float for
// ratios / factors, unsigned for counts) -- the real header is not available.
#pragma once
#include <string>
namespace cmd {
struct ParametersSlam {
    unsigned orbScaleLevels = 8;                       // static_settings.cpp:31
    float orbScaleFactor = 1.2f;                       // static_settings.cpp:32
    unsigned maxKeypoints = 1000;                      // static_settings.cpp:48
    unsigned orbLkTrackLevel = 0;                      // orb_extractor.cpp:91
    bool useGpuImagePyramid = false;                   // image_pyramid.cpp:211
    std::string slamFeatureDetector;                   // feature_detector.cpp:38
    float loopClosureFeatureMatchLoweRatio = 0.8f;     // keyframe_matcher.cpp:120
    bool requireTringulationForLoopClosures = true;    // keyframe_matcher.cpp:82
    float epipolarCheckThresholdDegrees = 0.2f;        // keyframe_matcher.cpp:168
    std::string vocabularyPath;                        // bow_index.cpp:35
    float bowMinInCommonRatio = 0.8f;                  // bow_index.cpp:141
    float bowScoreRatio = 0.75f;                       // bow_index.cpp:168
    bool visualizeMapPointSearch = false;              // keyframe_matcher.cpp:307
};
struct ParametersTracker {
    std::string featureDetector = "FAST";              // feature_detector.cpp:39
    int maxTracks = 0;                                 // feature_detector.cpp:37
    double gfttMinDistance = 0;                        // feature_detector.cpp:81
    int iniFastThreshold = 20, minFastThreshold = 7;   // upstream OpenVSLAM ini_fast_thr_ / min_fast_thr (shim detector)
};
}  // namespace cmd
