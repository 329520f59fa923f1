// stand-in for ../tracker/image.hpp (absent): an 8-bit gray CPU frame
#pragma once
#include <memory>
#include <vector>
#include <Eigen/Dense>
#include <opencv2/core.hpp>
#include <accelerated-arrays/image.hpp>
#include <accelerated-arrays/standard_ops.hpp>
#include "camera.hpp"
namespace tracker {
struct Image {
    virtual ~Image() = default;
    virtual accelerated::Image &getAccImage() = 0;
    virtual accelerated::Image::Factory &getImageFactory() = 0;
    virtual accelerated::operations::StandardFactory &getOperationsFactory() = 0;
    virtual accelerated::Processor &getProcessor() = 0;
    virtual std::shared_ptr<const Camera> getCamera() const = 0;
    virtual float getDepth(const Eigen::Vector2f &) const { return -1; }
    virtual bool hasStereoPointCloud() const { return false; }
    virtual const std::vector<Eigen::Vector3f> &getStereoPointCloud() const { static std::vector<Eigen::Vector3f> e; return e; }
};
struct CpuImage : Image {
    virtual cv::Mat getOpenCvMat() = 0;
};
}  // namespace tracker
