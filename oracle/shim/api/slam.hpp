// stand-in for ../api/slam.hpp (absent): the types keyframe.hpp / mapdb.hpp mention
#pragma once
#include <Eigen/Dense>
#include "../tracker/track.hpp"
#include "../tracker/image.hpp"
namespace slam {
using Feature = tracker::Feature;
struct Pose {
    Eigen::Matrix4d pose = Eigen::Matrix4d::Identity();
    Eigen::Matrix<double, 3, 6> uncertainty;
    double t = 0;
    int frameNumber = 0;
};
}  // namespace slam
