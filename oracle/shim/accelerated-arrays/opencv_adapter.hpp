// stand-in for accelerated-arrays/opencv_adapter.hpp: views between cv::Mat and accelerated::Image (no copies)
#pragma once
#include <opencv2/core.hpp>
#include "image.hpp"
namespace accelerated { namespace opencv {
inline std::unique_ptr<Image> ref(cv::Mat &m) {
    std::unique_ptr<Image> i(new Image());
    i->width = m.cols; i->height = m.rows; i->data = m.data; i->stride = m.step;
    return i;
}
inline cv::Mat ref(Image &i) { return cv::Mat(i.height, i.width, CV_8UC1, i.data, i.stride); }
inline cv::Mat emptyLike(const Image &i) { return cv::Mat(i.height, i.width, CV_8UC1); }
inline Future copy(Image &src, cv::Mat &dst) { ref(src).copyTo(dst); return Future(); }
} }  // namespace accelerated::opencv
