// stand-in for ../tracker/feature_detector.hpp (absent).  The real class is the parent project's corner detector;
// the reference only fixes its contract (feature_detector.cpp:73-101).  oracle/ref_slam.cpp backs build() with the
// oracle's FAST-in-cells + quadtree detector (oracle/src/fast_detect.cpp), the scheme BASELINE.json's north star names.
#pragma once
#include <memory>
#include <vector>
#include <accelerated-arrays/image.hpp>
#include <accelerated-arrays/standard_ops.hpp>
#include <accelerated-arrays/future.hpp>
#include "track.hpp"
#include "../odometry/parameters.hpp"
namespace tracker {
struct FeatureDetector {
    virtual ~FeatureDetector() = default;
    virtual accelerated::Future detect(accelerated::Image &image, std::vector<Feature::Point> &out,
                                       const std::vector<Feature::Point> &previous, double minDistance) = 0;
    static std::unique_ptr<FeatureDetector> build(int width, int height, accelerated::Processor &processor,
                                                  accelerated::Image::Factory &imgFactory,
                                                  accelerated::operations::StandardFactory &opFactory,
                                                  const odometry::ParametersTracker &params);
};
}  // namespace tracker
